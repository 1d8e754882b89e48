"""CPU: the float-parity floors the GPU tests gate against (tests/floors.py, committed in tests/golden/floors.json) are what the
reference really does: a sample of the entries is re-measured here from the compiled reference, and the reference's own chaos is
demonstrated (an input perturbation of 1e-6 moves its M&M output by an interpolator arm on a fraction of a percent of the symbols)."""
import json

import pytest

from tests import floors


def _need_ref():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")


def test_committed_floors_reproduce(built):
    _need_ref()
    with open(floors._CACHE_PATH) as f:
        committed = json.load(f)
    assert len(committed) >= 24
    for kind, name, stage in (("stage", "metop_ahrpt", "mm"), ("stage", "jpss_hrd", "costas"), ("chain", "dvbs2_front", None)):
        k = floors._key(kind, name, 21, stage, 1e-6, ())
        now = floors.stage_floor(name, 21, stage) if kind == "stage" else floors.chain_floor(name, 21)
        assert committed[k]["sig"] == floors.signal_crc(name, 21), k  # (the floor of THIS realisation of the signal)
        assert {a: b for a, b in committed[k].items() if a != "sig"} == pytest.approx(now, rel=1e-6, abs=1e-12), k


def test_the_reference_is_chaotic_at_the_1e5_level(built):
    """the M&M loop picks its interpolator arm with rint(mu * 128); fed an input 1e-6 off, the reference's symbols
    differ from its own by more than 1e-5 on 0.1-1.5 % of the symbols, by one to two arms at the worst spots, while the symbol COUNT and
    all but <0.2 % of the soft bytes stay the same. This is why the GPU gates are max(SURVEY 8c, 1.2 x floor)."""
    _need_ref()
    for name in ("metop_ahrpt", "bpsk_half"):
        fl = floors.cached("stage", name, 21, "mm")
        assert 1e-3 < fl["frac"] < 1.5e-2 and 5e-3 < fl["max"] < 8e-2 and fl["soft_diff"] < 2e-3, fl
    fl = floors.cached("stage", "metop_ahrpt", 21, "costas")  # a sign decision flips at a sample within 1e-6 of zero: a ~alpha phase kick
    assert fl["max"] > 1e-3 and fl["frac"] < 5e-3 and fl["median"] < 1e-6, fl

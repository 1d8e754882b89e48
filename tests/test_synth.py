"""CPU: the synthetic transmitter is a legal CCSDS transmitter — its encoders equal the reference's own encoders and the
reference receiver recovers exactly the transmitted frames."""
import numpy as np
import pytest

from tests.common import match_frames, oracle, oracle_demod, oracle_fec, signal
from satdump_b200 import synth


def _ref():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return ref


def test_encoders_equal_reference(built):
    ref = _ref()
    rng = np.random.default_rng(1)
    assert np.array_equal(synth.ccsds_pn(255), ref.derand(np.zeros(255, np.uint8)))
    for interleave in (4, 5):
        pl = rng.integers(0, 256, size=(3, interleave * 223), dtype=np.uint8)
        _, clear = synth.build_cadus(pl, interleave)
        for f in range(3):
            buf = np.zeros(255 * interleave, np.uint8)
            buf[:223 * interleave] = pl[f]
            assert np.array_equal(ref.rs_encode_interleaved(buf, True, interleave), clear[f, 4:])
    bits = rng.integers(0, 2, size=4000, dtype=np.uint8)
    assert np.array_equal(synth.conv_encode(bits), ref.cc_encode(bits))


def test_dual_basis_tables_are_inverse():
    assert np.array_equal(synth.FROM_DUAL[synth.TO_DUAL], np.arange(256))
    assert synth.TO_DUAL[1] == 0x7B and synth.TO_DUAL[0x80] == 0x8D


@pytest.mark.parametrize("name,lg", [("metop_ahrpt", 19), ("bpsk_half", 18), ("jpss_hrd", 19)])
def test_reference_receiver_recovers_transmitted_frames(built, name, lg):
    O = oracle()
    cfg, raw, clear = signal(name, lg, seed=2)
    soft = oracle_demod(O, cfg).run(raw, stages=False)["soft"]
    got = oracle_fec(O, cfg).run(soft)["cadu"].reshape(-1, cfg.cadu_bytes)
    first, ok = match_frames(got, clear)
    assert got.shape[0] >= 3 and ok, (got.shape, first)

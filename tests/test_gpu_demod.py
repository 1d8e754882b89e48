"""GPU: demodulator parity against the oracle from raw IQ, stage by stage, through the C ABI.

Gates = max(SURVEY.md 8c gate, 1.2 x the reference's own floor) with the floor measured on the reference itself (tests/floors.py:
its output after a 1e-6 perturbation of the AGC output / with its release build flags). SURVEY 8c:
  AGC, FIR outputs                 : |gpu - oracle| <= 1e-5 on EVERY sample (no feedback with sign / arm decisions: no floor needed)
  Costas(+delay) output            : <= 1e-5 everywhere ... unless the reference itself, fed an input 1e-6 off, leaves 1e-5 (every
                                     detector multiplies by sgn(v.re), sgn(v.im): a sample within the input difference of zero kicks
                                     the phase by ~alpha for ~1000 samples; costas_loop.cpp:31-52)
  M&M symbols                      : identical count; <= 1e-5 on >= 99 % of the symbols and <= 1.5e-2 on all
  int8 soft                        : differing on <= 0.5 % of the bytes, by one LSB
The same stages fed the oracle's exact stage input (bitwise rows) are in tests/test_gpu_stage_isolated.py.
"""
import numpy as np
import pytest

from tests.common import gpu_demod, nsamples, oracle, oracle_demod, signal
from tests.floors import SURVEY, cached, gate

pytestmark = pytest.mark.gpu
CONFIGS = ["metop_ahrpt", "bpsk_half", "jpss_hrd", "dvbs2_front", "hrpt_bpsk", "psk8", "metop_oversampled", "bpsk_decim8"]


def check_mm(gs, om, gsoft=None, osoft=None, floor=None):
    """M&M symbols / int8 soft bytes against the oracle's: SURVEY 8c gates, or 1.2 x the reference's own floor where that is larger."""
    fl = floor or {}
    assert gs.size == om.size, (gs.size, om.size)
    d = np.abs(gs - om)
    bad = float((d > 1e-5).mean())
    assert bad <= gate(SURVEY["mm_frac"], fl.get("mm_frac", 0.0)), (bad, fl)
    assert d.max() <= gate(SURVEY["mm_max"], fl.get("mm_max", 0.0)), (float(d.max()), fl)
    if gsoft is not None:
        assert gsoft.size == osoft.size
        ds = np.abs(gsoft.astype(np.int16) - osoft.astype(np.int16))
        assert (ds > 0).mean() <= gate(SURVEY["soft_diff"], fl.get("soft_diff", 0.0)), (float((ds > 0).mean()), fl)
        assert (ds > 1).mean() <= gate(SURVEY["soft_gt1"], fl.get("soft_gt1", 0.0)) + 2e-5 and ds.max() <= max(1, fl.get("soft_max", 0)) + 1, \
            (float((ds > 1).mean()), int(ds.max()), fl)
    return 1.0 - bad


def check_costas(g, o, floor):
    d = np.abs(g - o)
    assert np.median(d) <= 1e-6, float(np.median(d))  # nothing systematic (the mean would count the rare sign-decision events)
    # the 1e-5 of the survey with the margin the AGC's own deviation needs (measured max 1.3e-5 on BPSK), or the reference's event floor
    assert (d > 1e-5).mean() <= gate(1e-5, floor.get("costas_frac", 0.0)), (float((d > 1e-5).mean()), floor)
    assert d.max() <= gate(2e-5, floor.get("costas_max", 0.0)), (float(d.max()), floor)


def test_filter_taps_bitwise(built):
    O = oracle()
    cfg, raw, _ = signal("metop_ahrpt", 16)
    g = gpu_demod(cfg, 1 << 16)
    rrc, bank = g.taps()
    assert np.array_equal(rrc, O.rrc_design(1, np.float32(cfg.samplerate), int(cfg.symbolrate), cfg.rrc_alpha, 31))
    assert np.array_equal(bank, O.mm_taps())


def test_sample_conversion_is_bit_exact(built):
    """Every possible cs16 / cs8 sample value: the reciprocal+FMA conversion equals ((float)x) / scale of the reference's generic
    converters (baseband_interface.h:178-188) bit for bit."""
    cfg, _, _ = signal("metop_ahrpt", 16)
    v16 = np.arange(-32768, 32768, dtype=np.int16)
    got = gpu_demod(cfg, 1 << 16).convert(np.stack([v16, v16[::-1]], axis=1).reshape(-1))
    assert np.array_equal(got.real.view(np.uint32), (v16.astype(np.float32) / np.float32(32767)).view(np.uint32))
    assert np.array_equal(got.imag.view(np.uint32), (v16[::-1].astype(np.float32) / np.float32(32767)).view(np.uint32))
    import dataclasses
    from satdump_b200 import capi
    v8 = np.tile(np.arange(-128, 128, dtype=np.int8), 32)
    g8 = capi.Demod(capi.demod_cfg(90e6, 45e6, "none", 0.25, fmt="cs8", max_batch=8192))
    got8 = g8.convert(np.stack([v8, v8[::-1]], axis=1).reshape(-1))
    assert np.array_equal(got8.real.view(np.uint32), (v8.astype(np.float32) / np.float32(127)).view(np.uint32))


@pytest.mark.parametrize("name", CONFIGS)
def test_stage_parity(built, name):
    O = oracle()
    cfg, raw, _ = signal(name, 21)
    n = nsamples(raw, cfg)
    o = oracle_demod(O, cfg).run(raw)
    g = gpu_demod(cfg, n, keep_stages=True).push(raw)
    if g.cfg.final_samplerate > 0:  # front-end resampler (hrpt_bpsk): same length, same samples
        rs = O.resample(oracle_demod(O, cfg).cfg, raw)
        assert g.stage("resamp").size == rs.size == o["front"] and np.abs(g.stage("resamp") - rs).max() <= 2e-6
    fl = cached("chain", name, 21)
    for st in ("agc", "fir"):
        d = np.abs(g.stage(st) - o[st])
        assert d.max() <= 1e-5, (st, float(d.max()), int(np.argmax(d)))
    if o["costas"] is not None:
        check_costas(g.stage("costas"), o["costas"], fl)
    check_mm(g.symbols(), o["mm"], g.soft(), o["soft"], floor=fl)
    s = g.stats()
    assert s["costas_unconverged"] == 0 and s["mm_unconverged"] == 0 and s["agc_clamped"] == 0, s
    assert s["symbols_out"] == o["mm"].size and s["samples_in"] == n and s["last_front_samples"] == o["agc"].size
    # carried loop state ends where the oracle's ends
    ost = oracle_demod(O, cfg)
    ost.run(raw, stages=False)
    st = ost.state()
    assert abs(s["agc_gain"] - st["gain"]) <= 1e-4 * st["gain"]
    assert abs(s["mm_omega"] - st["omega"]) <= 1e-4


@pytest.mark.parametrize("name,cuts", [("metop_ahrpt", [300000 + 5, 1000003]), ("jpss_hrd", [65536, 65536 * 3 + 17]), ("bpsk_half", [4099])])
def test_streaming_pushes_continue_the_same_stream(built, name, cuts):
    """Loop state (AGC gain, FIR history, Costas phase/freq, M&M mu/omega/history, OQPSK delay) carries across ragged batches."""
    O = oracle()
    cfg, raw, _ = signal(name, 21)
    n = nsamples(raw, cfg)
    per = 1 if cfg.fmt == "cf32" else 2
    o = oracle_demod(O, cfg).run(raw, stages=False)
    g = gpu_demod(cfg, n)
    syms, soft, prev = [], [], 0
    for c in cuts + [n]:
        g.push(raw[prev * per:c * per])
        syms.append(g.symbols())
        soft.append(g.soft())
        prev = c
    check_mm(np.concatenate(syms), o["mm"], np.concatenate(soft), o["soft"], floor=cached("chain", name, 21))
    s = g.stats()
    assert s["costas_unconverged"] == 0 and s["mm_unconverged"] == 0


@pytest.mark.parametrize("n", [4096, 5000, 100003, (1 << 20) + 1])
def test_ragged_batch_sizes(built, n):
    O = oracle()
    cfg, raw, _ = signal("metop_ahrpt", 21)
    raw = raw[:2 * n]
    o = oracle_demod(O, cfg).run(raw)
    g = gpu_demod(cfg, n, keep_stages=True).push(raw)
    assert np.abs(g.stage("fir") - o["fir"]).max() <= 1e-5
    assert np.abs(g.stage("costas") - o["costas"]).max() <= 1e-5
    assert g.symbols().size == o["mm"].size


def test_agc_exact_pass_gives_the_same_stream(built, monkeypatch):
    """The AGC seeds of the tile ranges are normally proven by walking back a few tiles; B200_AGC_WARM_TILES=0 forbids that, so every
    range raises `need` and the stage is redone from the scanned per-tile seeds (the path very weak signals take)."""
    O = oracle()
    cfg, raw, _ = signal("metop_ahrpt", 21)
    n = nsamples(raw, cfg)
    o = oracle_demod(O, cfg).run(raw)
    monkeypatch.setenv("B200_AGC_WARM_TILES", "0")
    g = gpu_demod(cfg, n, keep_stages=True)
    monkeypatch.delenv("B200_AGC_WARM_TILES")
    g.push(raw[:2 * 700001])
    a1, f1 = g.stage("agc"), g.stage("fir")
    g.push(raw[2 * 700001:])
    a = np.concatenate([a1, g.stage("agc")])
    f = np.concatenate([f1, g.stage("fir")])
    assert np.abs(a - o["agc"]).max() <= 1e-5 and np.abs(f - o["fir"]).max() <= 1e-5
    assert g.stats()["agc_exact_passes"] == 2
    ref = gpu_demod(cfg, n, keep_stages=True).push(raw)
    assert ref.stats()["agc_exact_passes"] == 0
    assert np.abs(f - ref.stage("fir")).max() <= 2e-6  # the two seedings agree far below the parity gate


def test_weak_signal_takes_the_exact_agc_pass(built):
    """A signal 24 dB below the usual level: the AGC settles at gain ~70 and forgets so slowly (time constant 7000 samples) that the
    24-tile walk cannot prove a seed, so the exact pass runs. The reference's own float rounding noise in the gain grows with
    that memory (sigma ~1e-6 relative here), hence the wider gate."""
    O = oracle()
    cfg, raw, _ = signal("metop_ahrpt", 21)
    raw = (raw.astype(np.int32) // 16).astype(np.int16)
    n = nsamples(raw, cfg)
    o = oracle_demod(O, cfg).run(raw)
    g = gpu_demod(cfg, n, keep_stages=True).push(raw)
    assert 50 < g.stats()["agc_gain"] < 100 and g.stats()["agc_exact_passes"] == 1
    assert np.abs(g.stage("agc") - o["agc"]).max() <= 3e-5 and np.abs(g.stage("fir") - o["fir"]).max() <= 3e-5
    assert g.symbols().size == o["mm"].size


def test_interpolating_resampler_and_streaming(built):
    """2.4 Msym/s recorded at 2.6 MS/s (1.083 samples/symbol < MIN_SPS): the reference interpolates by 66/65 first. Stage parity of
    the front end, and ragged pushes (the resampler's carried counters and history) give the very same resampled stream."""
    O = oracle()
    cfg, raw, _ = signal("qpsk_undersampled", 20)
    n = nsamples(raw, cfg)
    o = oracle_demod(O, cfg).run(raw)
    g = gpu_demod(cfg, n, keep_stages=True).push(raw)
    assert g.cfg.final_samplerate == 2640000.0
    one = g.stage("resamp")
    assert one.size == o["front"] and np.abs(one - O.resample(oracle_demod(O, cfg).cfg, raw)).max() <= 2e-6
    assert np.abs(g.stage("agc") - o["agc"]).max() <= 1e-5 and np.abs(g.stage("fir") - o["fir"]).max() <= 1e-5
    g2 = gpu_demod(cfg, n, keep_stages=True)
    parts, prev = [], 0
    for c in [4099, 300001, 300001 + 70000, n]:
        g2.push(raw[2 * prev:2 * c])
        parts.append(g2.stage("resamp"))
        prev = c
    assert np.array_equal(np.concatenate(parts).view(np.uint32), one.view(np.uint32))
    assert abs(g2.stats()["agc_gain"] - g.stats()["agc_gain"]) <= 1e-5 * g.stats()["agc_gain"]


def test_iq_swap(built):
    """iq_swap (FileSourceBlock, file_source.cpp:31-33) = feeding the swapped samples: identical stream, bit for bit; and parity."""
    O = oracle()
    cfg, raw, _ = signal("metop_ahrpt", 20)
    n = nsamples(raw, cfg)
    from satdump_b200 import capi
    from tests.common import demod_kwargs
    swapped = raw.reshape(-1, 2)[:, ::-1].reshape(-1).copy()
    a = capi.Demod(capi.demod_cfg(max_batch=n, keep_stages=True, iq_swap=True, **demod_kwargs(cfg))).push(swapped)
    b = gpu_demod(cfg, n, keep_stages=True).push(raw)
    assert np.array_equal(a.stage("fir").view(np.uint32), b.stage("fir").view(np.uint32))
    assert np.array_equal(a.soft(), b.soft())
    o = O.Demod(O.demod_cfg(iq_swap=True, **demod_kwargs(cfg))).run(swapped)
    assert np.abs(a.stage("fir") - o["fir"]).max() <= 1e-5


@pytest.mark.parametrize("name", ["metop_ahrpt", "hrpt_bpsk"])
def test_dc_block(built, name):
    """dc_block: CorrectIQBlock behind the reader (utils/correct_iq.cpp:18-35), before the resampler / AGC. A constant-coefficient
    recurrence with a 10^4-sample memory, evaluated as a scan: output parity, carried accumulator across ragged pushes, and the rest
    of the chain on top of it."""
    from satdump_b200 import capi
    from tests.common import demod_kwargs
    O = oracle()
    cfg, raw, _ = signal(name, 20)
    raw = raw.copy()
    if cfg.fmt == "cf32":
        raw += np.complex64(0.03 - 0.011j)
    else:
        raw[0::2] += 900
        raw[1::2] -= 400
    n = nsamples(raw, cfg)
    per = 1 if cfg.fmt == "cf32" else 2
    oc = O.demod_cfg(dc_block=True, **demod_kwargs(cfg))
    o = O.Demod(oc).run(raw)
    g = capi.Demod(capi.demod_cfg(max_batch=n, keep_stages=True, dc_block=True, **demod_kwargs(cfg))).push(raw)
    front = O.resample(oc, raw)  # the oracle's front end: reader -> dc block (-> resampler)
    if g.cfg.final_samplerate > 0:
        assert np.abs(g.stage("resamp") - front).max() <= 2e-6
    else:
        assert np.abs(g.stage("dc") - front).max() <= 1e-6
    assert abs(g.stage("dc")[-100000:].mean()) < 2e-3  # the offset is gone
    assert np.abs(g.stage("agc") - o["agc"]).max() <= 1e-5 and np.abs(g.stage("fir") - o["fir"]).max() <= 1e-5
    assert g.symbols().size == o["mm"].size
    one = g.stage("dc")
    g2 = capi.Demod(capi.demod_cfg(max_batch=n, keep_stages=True, dc_block=True, **demod_kwargs(cfg)))
    parts, prev = [], 0
    for c in [4099, 300001, n]:
        g2.push(raw[per * prev:per * c])
        parts.append(g2.stage("dc"))
        prev = c
    assert np.abs(np.concatenate(parts) - one).max() <= 1e-6


def test_post_costas_dc(built):
    """post_costas_dc: CorrectIQBlock between the Costas loop and the clock recovery (module_psk_demod.cpp:127-134; three shipped BPSK
    pipelines). The clock recovery's input (stage "costas") and the symbols follow the reference, one shot and in ragged pushes."""
    from satdump_b200 import capi
    from tests.common import demod_kwargs
    O = oracle()
    cfg, raw, _ = signal("bpsk_half", 20)
    n = nsamples(raw, cfg)
    o = O.Demod(O.demod_cfg(post_costas_dc=True, **demod_kwargs(cfg))).run(raw)
    g = capi.Demod(capi.demod_cfg(max_batch=n, keep_stages=True, post_costas_dc=True, **demod_kwargs(cfg))).push(raw)
    fl = cached("chain", "bpsk_half", 20, extra=(("post_costas_dc", True),))
    check_costas(g.stage("costas"), o["costas"], fl)
    check_mm(g.symbols(), o["mm"], g.soft(), o["soft"], floor=fl)
    g2 = capi.Demod(capi.demod_cfg(max_batch=n, keep_stages=True, post_costas_dc=True, **demod_kwargs(cfg)))
    syms, prev = [], 0
    for c in [100003, 600000, n]:
        g2.push(raw[prev:c])
        syms.append(g2.symbols())
        prev = c
    check_mm(np.concatenate(syms), o["mm"], floor=fl)
    with pytest.raises(capi.B200Error):
        capi.Demod(capi.demod_cfg(30e6, 15e6, "oqpsk", 0.5, post_costas_dc=True, max_batch=65536))


def test_snr_estimate_follows_the_references_m2m4(built):
    """Module stats keys "snr" / "peak_snr" (module_psk_demod.cpp:190-194,242-243): M2M4SNREstimator over the recovered symbols. The two
    running averages are linear recurrences, evaluated on the GPU in closed form per push; the value after the last symbol equals the
    reference's (fed its own symbols in its own buffer-sized updates) to a hundredth of a dB, whatever the batching."""
    O = oracle()
    if not hasattr(O, "snr_m2m4"):
        pytest.skip("needs the compiled reference (oracle/_ref)")
    for name in ("metop_ahrpt", "bpsk_half"):
        cfg, raw, _ = signal(name, 21)
        n = nsamples(raw, cfg)
        want, _ = O.snr_m2m4(oracle_demod(O, cfg).run(raw)["mm"])
        g = gpu_demod(cfg, n).push(raw)
        s = g.stats()
        assert 5.0 < want < 30.0 and abs(s["snr"] - want) <= 0.01 and s["peak_snr"] >= s["snr"], (s["snr"], s["peak_snr"], want)
        per = 1 if cfg.fmt == "cf32" else 2
        g2 = gpu_demod(cfg, n)
        for a, b in ((0, 300001), (300001, 1000000), (1000000, n)):
            g2.push(raw[a * per:b * per])
        assert abs(g2.stats()["snr"] - want) <= 0.01, (g2.stats()["snr"], want)


def test_errors_are_loud(built):
    from satdump_b200 import capi
    cfg, raw, _ = signal("metop_ahrpt", 16)
    g = gpu_demod(cfg, 4096)
    with pytest.raises(capi.B200Error) as e:
        g.push(raw)  # 65536 samples > max_batch 4096
    assert e.value.code == -5
    with pytest.raises(capi.B200Error):
        g.push(raw[:40])  # below the minimum batch


def test_silent_input_runs_into_the_agc_clamp_like_the_reference(built):
    """All-zero baseband drives the AGC into its max_gain clamp after 6.55 M samples (agc.cpp:34-35); when the signal comes back the
    gain falls from exactly 65536. The clamp pass (step maps g -> min(g(1-e) + rate, 65536) composed as triples) follows the
    reference through the silence, the clamp and the recovery, also across a batch boundary inside the silence."""
    O = oracle()
    cfg, sig, _ = signal("metop_ahrpt", 21)
    raw = np.concatenate([np.zeros(2 * (1 << 23), np.int16), sig])
    n = raw.size // 2
    o = oracle_demod(O, cfg).run(raw)
    for cuts in ([n], [5000000, (1 << 23) + 4099, n]):
        g = gpu_demod(cfg, n, keep_stages=True)
        agc, fir, prev = [], [], 0
        for c in cuts:
            g.push(raw[2 * prev:2 * c])
            agc.append(g.stage("agc"))
            fir.append(g.stage("fir"))
            prev = c
        agc, fir = np.concatenate(agc), np.concatenate(fir)
        scale = np.maximum(1.0, np.abs(o["agc"]))
        assert (np.abs(agc - o["agc"]) / scale).max() <= 1e-5, float((np.abs(agc - o["agc"]) / scale).max())
        fscale = np.maximum(1.0, np.abs(o["fir"]))
        assert (np.abs(fir - o["fir"]) / fscale).max() <= 2e-5
        assert np.abs(o["agc"]).max() > 1000  # the recovery really starts from a huge gain
        s = g.stats()
        assert s["agc_clamped"] >= 1
        ost = oracle_demod(O, cfg)
        ost.run(raw, stages=False)
        assert abs(s["agc_gain"] - ost.state()["gain"]) <= 1e-4 * ost.state()["gain"]

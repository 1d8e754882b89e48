"""GPU: stage-isolated float parity. Every demodulator stage is fed the ORACLE's exact stage input through
b200_demod_debug_run_stage, so a kernel bug cannot hide behind (or be excused by) the feedback loops' sensitivity:

  FIR     strict build (separate multiply / add, the generic VOLK order of fir.cpp:74-83)  -> BITWISE the oracle's output
          production build (one fma per tap, k_agc_fir_w with the AGC switched off)          -> <= 1e-6 everywhere
  M&M     strict build, one sequential segment (clock_recovery_mm.cpp:52-121 as written)   -> BITWISE the oracle's symbols
          production arithmetic (two FMA chains), sequential and segmented                 -> same count, within the gates below
  Costas  sequential (one thread from the initial state)                                   -> <= 1e-5 everywhere
          segmented (production)                                                           -> within the gates below

Gates for the non-bitwise rows: max(SURVEY 8c gate, 1.2 x the reference's own floor under a 1e-6 perturbation of the same stage
input), tests/floors.py. The floors are part of the assertion messages, so a failure shows both numbers."""
import numpy as np
import pytest

from tests import floors
from tests.common import gpu_demod, nsamples, oracle
from tests.floors import SURVEY, cached, diff_stats, gate, reference_run

pytestmark = pytest.mark.gpu
CONFIGS = ["metop_ahrpt", "bpsk_half", "jpss_hrd", "dvbs2_front", "hrpt_bpsk", "psk8", "metop_oversampled", "bpsk_decim8"]
LG = 21


def bitwise(a, b):
    return a.size == b.size and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("name", CONFIGS)
def test_fir_fed_the_oracles_agc_output(built, name):
    cfg, raw, oc, o = reference_run(name, LG)
    g = gpu_demod(cfg, nsamples(raw, cfg))
    assert bitwise(g.run_stage("fir", o["agc"], strict=True), o["fir"])
    d = diff_stats(g.run_stage("fir", o["agc"]), o["fir"])
    assert d["max"] <= 1e-6, d


@pytest.mark.parametrize("name", CONFIGS)
def test_mm_fed_the_oracles_input(built, name):
    cfg, raw, oc, o = reference_run(name, LG)
    g = gpu_demod(cfg, nsamples(raw, cfg))
    mm_in = o["fir"] if o["costas"] is None else o["costas"]
    assert bitwise(g.run_stage("mm", mm_in, strict=True, sequential=True), o["mm"])
    fl = cached("stage", name, LG, "mm")
    for sequential in (True, False):  # production arithmetic: one segment / the production segmentation
        d = diff_stats(g.run_stage("mm", mm_in, sequential=sequential), o["mm"])
        assert d["frac"] <= gate(SURVEY["mm_frac"], fl["frac"]) and d["max"] <= gate(SURVEY["mm_max"], fl["max"]), (sequential, d, fl)


@pytest.mark.parametrize("name", [c for c in CONFIGS if c != "dvbs2_front"])
def test_costas_fed_the_oracles_fir_output(built, name):
    cfg, raw, oc, o = reference_run(name, LG)
    g = gpu_demod(cfg, nsamples(raw, cfg))
    d = diff_stats(g.run_stage("costas", o["fir"], sequential=True), o["costas"])
    assert d["max"] <= SURVEY["float_all"], d
    fl = cached("stage", name, LG, "costas")
    d = diff_stats(g.run_stage("costas", o["fir"]), o["costas"])
    assert d["median"] <= 1e-6, d  # nothing systematic (the mean would count the rare sign-decision events)
    assert d["frac"] <= gate(0.0, fl["frac"]) + 1e-5 and d["max"] <= gate(SURVEY["float_all"], fl["max"]) * (2 if fl["frac"] == 0 else 1), (d, fl)


def test_junction_residuals_are_small(built):
    """What the warm-ups leave at the first owned sample of every segment: Costas phase within the repair tolerance, M&M sampling
    instant within a fraction of an interpolator arm for all but the tail (which the repair rounds catch above 0.01 sample)."""
    cfg, raw, oc, o = reference_run("metop_ahrpt", LG)
    g = gpu_demod(cfg, nsamples(raw, cfg)).push(raw)
    cj, mj, L = g.junctions()
    assert L >= 4096 and cj.shape[0] == mj.size >= 2
    assert np.abs(cj[1:, 0]).max() <= 1e-5 and np.abs(cj[1:, 1]).max() <= 2e-6, (np.abs(cj[1:, 0]).max(), np.abs(cj[1:, 1]).max())
    assert np.median(np.abs(mj[1:])) <= 1e-4 and np.abs(mj[1:]).max() <= 0.01, (float(np.median(np.abs(mj[1:]))), float(np.abs(mj[1:]).max()))
    s = g.stats()
    assert s["costas_unconverged"] == 0 and s["mm_unconverged"] == 0


@pytest.mark.parametrize("name", ["metop_ahrpt", "bpsk_half"])
def test_gardner_clock_recovery(built, name):
    """dsp::GardnerClockRecoveryBlock<complex_t> (clock_recovery_gardner.cpp:33-131) as a variant of the clock recovery
    kernel. Strict sequential: BITWISE the reference block on the reference's own input; production arithmetic / segmentation and the
    whole chain within the floor gates; CADUs of the chain bit-exact against the reference chain built with the same block."""
    from satdump_b200 import capi
    from tests.common import demod_kwargs, gpu_fec_cfg, oracle_fec
    from tests.test_gpu_demod import check_mm
    O = oracle()
    gx = (("clock_recovery", "gardner"),)
    cfg, raw, oc, o = reference_run(name, LG, gx)
    n = nsamples(raw, cfg)
    g = capi.Demod(capi.demod_cfg(max_batch=n, clock_recovery="gardner", **demod_kwargs(cfg)))
    mm_in = o["costas"]
    assert bitwise(g.run_stage("mm", mm_in, strict=True, sequential=True), o["mm"])
    fl = cached("stage", name, LG, "mm", extra=gx)
    for sequential in (True, False):
        d = diff_stats(g.run_stage("mm", mm_in, sequential=sequential), o["mm"])
        assert d["frac"] <= gate(SURVEY["mm_frac"], fl["frac"]) and d["max"] <= gate(SURVEY["mm_max"], fl["max"]), (sequential, d, fl)
    g.push(raw)
    check_mm(g.symbols(), o["mm"], g.soft(), o["soft"], floor=cached("chain", name, LG, extra=gx))
    s = g.stats()
    assert s["costas_unconverged"] == 0 and s["mm_unconverged"] == 0
    want = oracle_fec(O, cfg).run(o["soft"])["cadu"].reshape(-1, cfg.cadu_bytes)
    ch = capi.Chain(capi.demod_cfg(max_batch=n, clock_recovery="gardner", **demod_kwargs(cfg)), gpu_fec_cfg(cfg, 2 * n)).push(raw)
    got = ch.frames()
    assert want.shape[0] >= 50 and got.shape == want.shape and np.array_equal(got, want)

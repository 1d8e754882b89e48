"""GPU: pm_demod (module_pm_demod.cpp: AGC -> PLLCarrierTrackingBlock -> PMToBPSK -> [SmartResampler -> AGC2] -> RRC -> Costas -> M&M)
the freq_shift option of BaseDemodModule and psk_demod's carrier mode ("has_carrier"), through the C ABI against the oracle.

Gates (measured distances on B200 in brackets, tools/probe_pm.py):
  carrier PLL, one segment            : bit for bit the reference (the kernel does the reference's operations in the reference's order)
  carrier PLL, segmented              : |gpu - oracle| <= 5e-6 everywhere [7e-7 .. 1e-6; 0.015 - 0.12 % of the samples differ at all]
  first AGC, carrier PLL in the chain : <= 1e-5 everywhere [3.2e-6]
  PMToBPSK / FreqShiftBlock           : the reference's VOLK rotator advances its phasor by one ROUNDED complex multiplication per sample,
                                        a serial recurrence whose angle walks off n * delta by up to ~1.5e-8 rad per sample (which way
                                        depends on the VOLK flavour, oracle/shim restates the generic one); the kernel evaluates
                                        e^{j n delta} in closed form. Gate: magnitudes equal to 1e-4 of the rms [2e-6], phase
                                        difference <= 2e-8 * n + 1e-5 rad [7.8e-9 * n]. The Costas loop behind absorbs that slow rotation.
  symbols                             : identical count; |gpu - oracle| <= 1e-3 on >= 98 % [99.5 - 99.8 %], <= 1.5e-2 on all [1.7e-3]
  int8 soft                           : differing on <= 0.5 % of the bytes [0.08 - 0.10 %], never by more than one LSB
  CADUs (through the reference decoder and through the GPU chain): the reference's, and the transmitted frames
"""
import dataclasses

import numpy as np
import pytest

from satdump_b200 import capi, synth
from tests.common import demod_kwargs, gpu_chain, gpu_demod, match_frames, nsamples, oracle, oracle_demod, oracle_fec, signal

pytestmark = pytest.mark.gpu
PM = ["pm_bpsk", "pm_bpsk_after", "pm_bpsk_front"]  # 6 samples/symbol; resampler behind the PLL; resampler in front


def rotation_gates(g, o, n_total):
    """gpu / oracle outputs of a rotator stage: same magnitudes, phase difference within the reference's own rounding walk."""
    rms = float(np.sqrt(np.mean(np.abs(o) ** 2)))
    assert np.abs(np.abs(g) - np.abs(o)).max() <= 1e-4 * rms
    m = np.abs(o) > 0.3 * rms
    ang = np.abs(np.angle(g[m] * np.conj(o[m])))
    idx = np.nonzero(m)[0]
    assert (ang <= 2e-8 * idx + 1e-5).all(), (float(ang.max()), n_total)


def soft_gates(gs, gsoft, o):
    assert gs.size == o["mm"].size and gsoft.size == o["soft"].size
    d = np.abs(gs - o["mm"])
    assert (d > 1e-3).mean() <= 0.02 and d.max() <= 1.5e-2, (float((d > 1e-3).mean()), float(d.max()))
    ds = np.abs(gsoft.astype(np.int16) - o["soft"].astype(np.int16))
    assert (ds > 0).mean() <= 5e-3 and ds.max() <= 1, (float((ds > 0).mean()), int(ds.max()))


@pytest.mark.parametrize("name", PM)
def test_carrier_pll_stage(built, name):
    O = oracle()
    cfg, raw, _ = signal(name, 21, seed=3)
    n = nsamples(raw, cfg)
    o = oracle_demod(O, cfg).run(raw)
    g = gpu_demod(cfg, n)
    one = g.run_stage("pll", o["agc"][:16384], sequential=True)
    assert np.array_equal(one.view(np.uint32), o["pll"][:16384].view(np.uint32))
    seg = g.run_stage("pll", o["agc"])
    assert np.abs(seg - o["pll"]).max() <= 5e-6
    assert (seg.view(np.uint64) != o["pll"].view(np.uint64)).mean() <= 5e-3
    rotation_gates(g.run_stage("pm", o["agc"]), o["pm"], n)


@pytest.mark.parametrize("name", PM)
def test_pm_demod_chain_against_the_reference(built, name):
    O = oracle()
    cfg, raw, clear = signal(name, 21, seed=1)
    n = nsamples(raw, cfg)
    d = oracle_demod(O, cfg)
    o = d.run(raw)
    g = gpu_demod(cfg, n, keep_stages=True)
    g.push(raw)
    st = g.stats()
    assert st["pll_unconverged"] == 0 and st["costas_unconverged"] == 0 and st["mm_unconverged"] == 0 and st["last_front_samples"] == o["front"]
    assert abs(st["pll_freq"] - float(d.pm_state()["pll_freq"])) <= 1e-6
    assert np.abs(g.stage("agc") - o["agc"]).max() <= 1e-5
    assert np.abs(g.stage("pll") - o["pll"]).max() <= 1e-5
    rotation_gates(g.stage("pm"), o["pm"], n)
    gs, gsoft = g.symbols(), g.soft()
    soft_gates(gs, gsoft, o)
    ref_cadus = oracle_fec(O, cfg).run(o["soft"])["cadu"].reshape(-1, cfg.cadu_bytes)
    got = oracle_fec(O, cfg).run(gsoft)["cadu"].reshape(-1, cfg.cadu_bytes)
    assert ref_cadus.shape[0] >= 4 and np.array_equal(got, ref_cadus) and match_frames(got, clear)[1]
    # the same stream in two ragged pushes: carried PLL state, rotator position, second AGC, resampler phase
    g2 = gpu_demod(cfg, n)
    cut = (n // 3) | 1
    a, b = (raw[:cut], raw[cut:]) if cfg.fmt == "cf32" else (raw[:2 * cut], raw[2 * cut:])
    sp = np.concatenate([g2.push(a).soft(), g2.push(b).soft()])
    assert sp.size == gsoft.size
    ds = np.abs(sp.astype(np.int16) - gsoft.astype(np.int16))
    assert (ds > 0).mean() <= 5e-3 and ds.max() <= 1
    assert all(g2.stats()[k] == 0 for k in ("pll_unconverged", "costas_unconverged", "mm_unconverged"))


def test_psk_demod_carrier_mode(built):
    """psk_demod with "has_carrier" (module_psk_demod.cpp:93-116, ODIN.json): RRC -> carrier PLL -> DC blocker -> Costas (limit 0.2) -> M&M."""
    O = oracle()
    cfg, raw, clear = signal("bpsk_carrier", 21, seed=1)
    n = nsamples(raw, cfg)
    d = oracle_demod(O, cfg)
    o = d.run(raw)
    g = gpu_demod(cfg, n, keep_stages=True)
    g.push(raw)
    st = g.stats()
    assert st["pll_unconverged"] == 0 and st["costas_unconverged"] == 0 and st["mm_unconverged"] == 0
    assert abs(st["pll_freq"] - float(d.pm_state()["pll_freq"])) <= 1e-6
    assert np.abs(g.stage("fir") - o["fir"]).max() <= 1e-5
    assert np.abs(g.stage("pll") - o["pll"]).max() <= 2e-5  # the carrier PLL's output (its input, the RRC output, is itself within 1e-5)
    gs, gsoft = g.symbols(), g.soft()
    assert gs.size == o["mm"].size
    dm = np.abs(gs - o["mm"])
    assert (dm > 1e-3).mean() <= 0.02 and dm.max() <= 3e-2, (float((dm > 1e-3).mean()), float(dm.max()))
    ds = np.abs(gsoft.astype(np.int16) - o["soft"].astype(np.int16))
    assert (ds > 0).mean() <= 5e-3 and (ds > 1).mean() <= 2e-5, (float((ds > 0).mean()), int(ds.max()))
    ref_cadus = oracle_fec(O, cfg).run(o["soft"])["cadu"].reshape(-1, cfg.cadu_bytes)
    got = oracle_fec(O, cfg).run(gsoft)["cadu"].reshape(-1, cfg.cadu_bytes)
    assert ref_cadus.shape[0] >= 20 and np.array_equal(got, ref_cadus) and match_frames(got, clear)[1]
    fused = gpu_chain(cfg, n).push(raw).frames()
    assert np.array_equal(fused, ref_cadus)
    # two ragged pushes: carried PLL state and DC accumulator
    g2 = gpu_demod(cfg, n)
    cut = (n // 3) | 1
    sp = np.concatenate([g2.push(raw[:2 * cut]).soft(), g2.push(raw[2 * cut:]).soft()])
    assert sp.size == gsoft.size
    d2 = np.abs(sp.astype(np.int16) - gsoft.astype(np.int16))
    assert (d2 > 0).mean() <= 5e-3 and (d2 > 1).mean() <= 2e-5


def test_pm_demod_through_the_fused_chain(built):
    """pm_demod -> ccsds_conv_concat_decoder with the soft symbols kept on the device: CADUs = the reference's = the transmitted frames."""
    O = oracle()
    cfg, raw, clear = signal("pm_bpsk", 21, seed=2)
    n = nsamples(raw, cfg)
    ref_cadus = oracle_fec(O, cfg).run(oracle_demod(O, cfg).run(raw)["soft"])["cadu"].reshape(-1, cfg.cadu_bytes)
    c = gpu_chain(cfg, n)
    got = c.push(raw).frames()
    assert got.shape[0] >= 4 and np.array_equal(got, ref_cadus) and match_frames(got, clear)[1]


def test_freq_shift_brings_an_offset_carrier_back(built):
    """FreqShiftBlock behind the reader (module_demod_base.cpp:122-123): a QPSK stream 150 kHz off centre, shifted back by freq_shift."""
    O = oracle()
    cfg = dataclasses.replace(synth.CONFIGS["metop_ahrpt"], carrier_rad=2 * np.pi * 150e3 / 6e6 + 1e-3)
    raw, clear = synth.make_signal(cfg, 1 << 21, seed=4)
    raw = raw.cpu().numpy()
    n = nsamples(raw, cfg)
    kw = demod_kwargs(cfg)
    d = O.Demod(O.demod_cfg(freq_shift=-150000.0, **kw))
    o = d.run(raw)
    g = capi.Demod(capi.demod_cfg(max_batch=n, keep_stages=True, freq_shift=-150000.0, **kw))
    g.push(raw)
    st = g.stats()
    assert st["costas_unconverged"] == 0 and st["mm_unconverged"] == 0
    assert abs(st["costas_freq"] - float(d.state()["freq"])) <= 2e-5
    rotation_gates(g.stage("agc"), o["agc"], n)  # the AGC output = gain * rotated input: same magnitudes, the rotator's phase walk
    gs, gsoft = g.symbols(), g.soft()
    assert gs.size == o["mm"].size
    dm = np.abs(gs - o["mm"])
    assert (dm > 1e-3).mean() <= 0.02 and dm.max() <= 3e-2, (float((dm > 1e-3).mean()), float(dm.max()))
    ds = np.abs(gsoft.astype(np.int16) - o["soft"].astype(np.int16))
    assert (ds > 0).mean() <= 5e-3 and (ds > 1).mean() <= 2e-5, (float((ds > 0).mean()), int(ds.max()))
    f = oracle_fec(O, cfg)
    got = f.run(gsoft)["cadu"].reshape(-1, cfg.cadu_bytes)
    assert got.shape[0] >= 4 and np.array_equal(got, oracle_fec(O, cfg).run(o["soft"])["cadu"].reshape(-1, cfg.cadu_bytes)) and match_frames(got, clear)[1]


@pytest.mark.parametrize("name", ["pm_bpsk", "pm_bpsk_after"])
def test_pm_demod_iq_swap_is_applied_at_the_reader(built, name):
    """iq_swap (FileSourceBlock, module_demod_base.cpp:41-42) in front of pm_demod: a recording with I and Q exchanged gives the same soft
    symbols bit for bit (with resample_after_pll nothing in front of the AGC applies the swap: a converting copy does)."""
    cfg, raw, _ = signal(name, 20, seed=8)
    n = nsamples(raw, cfg)
    if cfg.fmt == "cf32":
        swapped = (raw.imag + 1j * raw.real).astype(np.complex64)
    else:
        swapped = np.ascontiguousarray(raw.reshape(-1, 2)[:, ::-1]).reshape(-1)
    kw = demod_kwargs(cfg)
    a = capi.Demod(capi.demod_cfg(max_batch=n, **kw)).push(raw).soft()
    b = capi.Demod(capi.demod_cfg(max_batch=n, iq_swap=True, **kw)).push(swapped).soft()
    assert a.size > 1000 and np.array_equal(a, b)


def test_pm_demod_errors_are_loud(built):
    kw = demod_kwargs(synth.CONFIGS["pm_bpsk"])
    with pytest.raises(capi.B200Error, match="bpsk"):
        capi.Demod(capi.demod_cfg(max_batch=1 << 16, **{**kw, "constellation": "qpsk"}))
    with pytest.raises(capi.B200Error, match="pll_bw"):
        capi.Demod(capi.demod_cfg(max_batch=1 << 16, **{**kw, "pm_pll_bw": 0.0}))
    with pytest.raises(capi.B200Error, match="pm_demod"):  # the carrier PLL stage of a psk_demod
        cfg = synth.CONFIGS["bpsk_half"]
        gpu_demod(cfg, 1 << 16).run_stage("pll", np.zeros(4096, np.complex64))

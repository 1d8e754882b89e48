"""Diagnostic (not a test): pm_demod / freq_shift on the GPU against the compiled reference, stage by stage, printing the distances the
gates of tests/test_gpu_pm.py are set from.   python tools/probe_pm.py [log2n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satdump_b200 import capi, synth  # noqa: E402
from tests import common  # noqa: E402


def phase_drift(g, o):
    m = np.abs(o) > 0.3 * np.sqrt(np.mean(np.abs(o) ** 2))
    ang = np.angle(g[m] * np.conj(o[m]))
    idx = np.nonzero(m)[0]
    return idx, ang


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 21
    O = common.oracle()
    for name in ("pm_bpsk", "pm_bpsk_after"):
        cfg = synth.CONFIGS[name]
        raw, clear = synth.make_signal(cfg, 1 << lg, seed=1)
        raw = raw.numpy()
        n = common.nsamples(raw, cfg)
        d = common.oracle_demod(O, cfg)
        r = d.run(raw)
        g = common.gpu_demod(cfg, n, keep_stages=True)
        # one segment, sequential: bit for bit
        x = np.ascontiguousarray(r["agc"][:16384])
        p1 = g.run_stage("pll", x, sequential=True)
        print(name, "pll one segment bitwise equal:", np.array_equal(p1.view(np.uint32), r["pll"][:16384].view(np.uint32)),
              "max abs", float(np.abs(p1 - r["pll"][:16384]).max()))
        # segmented on the reference's AGC output
        p2 = g.run_stage("pll", r["agc"])
        dd = np.abs(p2 - r["pll"])
        print(name, "pll segmented: bitwise-different samples", int((p2.view(np.uint64) != r["pll"].view(np.uint64)).sum()), "of", p2.size, "max abs",
              float(dd.max()), "frac>1e-6", float((dd > 1e-6).mean()))
        m2 = g.run_stage("pm", r["agc"])
        idx, ang = phase_drift(m2, r["pm"])
        print(name, "pm stage: |mag| rel diff max", float(np.abs(np.abs(m2) - np.abs(r["pm"])).max() / np.sqrt(np.mean(np.abs(r["pm"]) ** 2))),
              "phase diff first/last/maxabs", float(ang[0]), float(ang[-1]), float(np.abs(ang).max()), "rate/sample", float(ang[-1] / idx[-1]))
        # full chain
        g.push(raw)
        st = g.stats()
        print(name, "stats", {k: st[k] for k in ("symbols_out", "costas_unconverged", "mm_unconverged", "pll_unconverged", "repairs", "agc_clamped", "pll_freq",
                                                 "last_front_samples", "snr")})
        print(name, "ref pll state", d.pm_state(), "ref state", d.state())
        ga = g.stage("agc")
        print(name, "agc max abs", float(np.abs(ga - r["agc"]).max()))
        gp = g.stage("pll")
        dd = np.abs(gp - r["pll"])
        print(name, "pll (chain) max abs", float(dd.max()), "frac>1e-5", float((dd > 1e-5).mean()))
        gm = g.stage("pm")
        idx, ang = phase_drift(gm, r["pm"])
        print(name, "pm (chain) phase diff maxabs", float(np.abs(ang).max()))
        gf = g.stage("fir")
        df = np.abs(gf - r["fir"])
        print(name, "fir max abs", float(df.max()), "median", float(np.median(df)), "rms ref", float(np.sqrt(np.mean(np.abs(r["fir"]) ** 2))))
        gc = g.stage("costas")
        dc = np.abs(gc - r["costas"])
        print(name, "costas max abs", float(dc.max()), "median", float(np.median(dc)), "frac>1e-4", float((dc > 1e-4).mean()))
        gs, gsoft = g.symbols(), g.soft()
        print(name, "symbols", gs.size, r["mm"].size)
        if gs.size == r["mm"].size:
            ds = np.abs(gs - r["mm"])
            print(name, "mm max abs", float(ds.max()), "frac>1e-5", float((ds > 1e-5).mean()), "frac>1e-3", float((ds > 1e-3).mean()))
            dso = np.abs(gsoft.astype(np.int16) - r["soft"].astype(np.int16))
            print(name, "soft diff frac", float((dso > 0).mean()), "frac>1", float((dso > 1).mean()), "max", int(dso.max()))
        # decoder
        f = common.oracle_fec(O, cfg)
        oc = f.run(r["soft"])["cadu"].reshape(-1, cfg.cadu_bytes)
        f2 = common.oracle_fec(O, cfg)
        gcad = f2.run(gsoft)["cadu"].reshape(-1, cfg.cadu_bytes)
        print(name, "cadus ref", oc.shape[0], "gpu-soft->ref-fec", gcad.shape[0], "equal", np.array_equal(oc, gcad), "match tx", common.match_frames(gcad, clear))
        # streaming: two ragged pushes against one
        g2 = common.gpu_demod(cfg, n)
        cut = (n // 3) | 1
        fmt_c = cfg.fmt == "cf32"
        a, b = (raw[:cut], raw[cut:]) if fmt_c else (raw[:2 * cut], raw[2 * cut:])
        s_a = g2.push(a).soft()
        s_b = g2.push(b).soft()
        sp = np.concatenate([s_a, s_b])
        print(name, "split pushes: soft count", sp.size, gsoft.size, "diff frac",
              float((sp != gsoft).mean()) if sp.size == gsoft.size else None, "stats", {k: g2.stats()[k] for k in ("pll_unconverged", "costas_unconverged", "mm_unconverged")})
        t = g.timing()
        print(name, "timing ms", t)

    # freq_shift in front of psk_demod: a carrier 150 kHz off centre, shifted back (shifting a centred signal AWAY instead makes the
    # reference's Costas loop chase 0.157 rad/sample for the whole stream: not a parity case)
    import dataclasses
    cfg = dataclasses.replace(synth.CONFIGS["metop_ahrpt"], carrier_rad=2 * np.pi * 150e3 / 6e6 + 1e-3)
    raw, _ = synth.make_signal(cfg, 1 << lg, seed=1)
    raw = raw.numpy()
    n = common.nsamples(raw, cfg)
    kw = common.demod_kwargs(cfg)
    d = O.Demod(O.demod_cfg(freq_shift=-150000.0, **kw))
    r = d.run(raw)
    g = capi.Demod(capi.demod_cfg(max_batch=n, keep_stages=True, freq_shift=-150000.0, **kw))
    g.push(raw)
    gs, gsoft = g.symbols(), g.soft()
    print("freq_shift symbols", gs.size, r["mm"].size, "ref costas freq", d.state()["freq"], "gpu", g.stats()["costas_freq"])
    if gs.size == r["mm"].size:
        ds = np.abs(gs - r["mm"])
        dso = np.abs(gsoft.astype(np.int16) - r["soft"].astype(np.int16))
        print("freq_shift mm max abs", float(ds.max()), "frac>1e-5", float((ds > 1e-5).mean()), "frac>1e-3", float((ds > 1e-3).mean()), "soft diff frac",
              float((dso > 0).mean()), "max", int(dso.max()))
    ga = g.stage("agc")
    idx, ang = phase_drift(ga, r["agc"])
    print("freq_shift agc-stage phase diff maxabs", float(np.abs(ang).max()), "mag rel", float(np.abs(np.abs(ga) - np.abs(r["agc"])).max()))


if __name__ == "__main__":
    main()

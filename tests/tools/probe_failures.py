"""Diagnostic (GPU box; needs the golden fixtures' oracle run = oracle/_ref or the port): where the Costas / M&M deviations of the
failing parity cases sit (sample index -> segment, distance to the segment start), with the junction residuals next to them."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from satdump_b200 import capi  # noqa: E402
from tests.common import demod_kwargs, gpu_demod, nsamples, oracle, oracle_demod, signal  # noqa: E402


def clusters(idx, gap=3000):
    out = []
    for i in idx:
        if not out or i - out[-1][1] > gap:
            out.append([i, i, 1])
        else:
            out[-1][1] = i
            out[-1][2] += 1
    return out


def report(tag, d, L, scale=1):
    bad = np.nonzero(d > 1e-5)[0]
    print(f"{tag}: n={d.size} frac>1e-5={bad.size / d.size:.5f} max={d.max():.3e} at {int(d.argmax())}; L={L}")
    for a, b, c in clusters(bad)[:12]:
        pk = a + int(np.argmax(d[a:b + 1]))
        print(f"    cluster [{a}, {b}] n={c} peak {d[pk]:.3e} at {pk}: segment {a * scale // L} offset {a * scale % L} (in samples)")


O = oracle()
# (a) metop_oversampled, Costas fed the oracle's FIR output
cfg, raw, _ = signal("metop_oversampled", 21)
o = oracle_demod(O, cfg).run(raw)
g = gpu_demod(cfg, nsamples(raw, cfg))
got = g.run_stage("costas", o["fir"])
cj, mj, L = g.junctions()
report("costas stage-isolated metop_oversampled", np.abs(got - o["costas"]), L)
c = np.abs(cj[:, 0])
print("   junction phase residual top:", [(int(i), float("%.2e" % c[i])) for i in np.argsort(-c)[:6]], "stats", {k: v for k, v in g.stats().items() if "unconv" in k or "repair" in k})
v = o["costas"]
near = np.minimum(np.abs(v.real), np.abs(v.imag))
print("   oracle costas output: samples with min(|re|,|im|) < 1e-5:", np.nonzero(near < 1e-5)[0][:20], " < 1e-6:", np.nonzero(near < 1e-6)[0][:20])
g2 = gpu_demod(cfg, nsamples(raw, cfg), keep_stages=True).push(raw)
cj, mj, L = g2.junctions()
report("costas chain metop_oversampled", np.abs(g2.stage("costas") - o["costas"]), L)

# (b) bpsk_half + post_costas_dc, 2^20
cfg, raw, _ = signal("bpsk_half", 20)
n = nsamples(raw, cfg)
o = O.Demod(O.demod_cfg(post_costas_dc=True, **demod_kwargs(cfg))).run(raw)
g = capi.Demod(capi.demod_cfg(max_batch=n, keep_stages=True, post_costas_dc=True, **demod_kwargs(cfg))).push(raw)
cj, mj, L = g.junctions()
sps = cfg.samplerate / cfg.symbolrate
report("mm bpsk_half post_costas_dc one shot", np.abs(g.symbols() - o["mm"]), L, scale=sps)
report("   its costas(+dc) stage", np.abs(g.stage("costas") - o["costas"]), L)
a = np.abs(mj)
print("   mm junction top:", [(int(i), float("%.2e" % a[i])) for i in np.argsort(-a)[:6]], "stats", {k: v for k, v in g.stats().items() if "unconv" in k or "repair" in k})
g2 = capi.Demod(capi.demod_cfg(max_batch=n, keep_stages=True, post_costas_dc=True, **demod_kwargs(cfg)))
syms, prev = [], 0
for c in [100003, 600000, n]:
    g2.push(raw[prev:c])
    syms.append(g2.symbols())
    prev = c
report("mm bpsk_half post_costas_dc ragged pushes (cuts at 100003, 600000 samples)", np.abs(np.concatenate(syms) - o["mm"]), g2.junctions()[2], scale=sps)
g3 = capi.Demod(capi.demod_cfg(max_batch=n, keep_stages=True, **demod_kwargs(cfg))).push(raw)
o3 = oracle_demod(O, cfg).run(raw)
report("mm bpsk_half plain (no post dc) one shot", np.abs(g3.symbols() - o3["mm"]), g3.junctions()[2], scale=sps)

"""Diagnostic (GPU box): the shape of the largest M&M deviation of bpsk_half + post_costas_dc (2^20 samples) and whether the
sequential / segmented stage-isolated runs on the oracle's own M&M input show it."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from satdump_b200 import capi  # noqa: E402
from tests.common import demod_kwargs, nsamples, oracle, signal  # noqa: E402

O = oracle()
cfg, raw, _ = signal("bpsk_half", 20)
n = nsamples(raw, cfg)
o = O.Demod(O.demod_cfg(post_costas_dc=True, **demod_kwargs(cfg))).run(raw)
mk = lambda: capi.Demod(capi.demod_cfg(max_batch=n, keep_stages=True, post_costas_dc=True, **demod_kwargs(cfg)))
g = mk().push(raw)
d = np.abs(g.symbols() - o["mm"])
pk = int(d.argmax())
print("chain: max", d.max(), "at", pk)
for a in range(pk - 2000, pk + 2000, 100):
    print(f"  [{a},{a + 100}) max {d[a:a + 100].max():.3e} n>1e-5 {(d[a:a + 100] > 1e-5).sum()}", end=";")
print()
cj, mj, L = g.junctions()
seg = int(pk * 2.5 / L)
print("segments around:", [(s, float("%.2e" % mj[s])) for s in range(seg - 3, seg + 4)], "L", L)
mm_in = o["costas"]
for seq in (True, False):
    gs = mk()
    got = gs.run_stage("mm", mm_in, sequential=seq)
    dd = np.abs(got - o["mm"])
    print("stage-isolated sequential" if seq else "stage-isolated segmented", "max", dd.max(), "at", int(dd.argmax()), "frac", (dd > 1e-5).mean(), "near pk:", dd[pk - 1500:pk + 1500].max())
    if not seq:
        cj, mj, L = gs.junctions()
        a = np.abs(mj)
        print("   junction top:", [(int(i), float("%.2e" % a[i])) for i in np.argsort(-a)[:6]])
# the reference perturbed: where is ITS largest excursion
from tests.floors import perturb  # noqa: E402
oc = O.demod_cfg(post_costas_dc=True, **demod_kwargs(cfg))
for sd in (1, 2):
    got = O.run_stage(oc, "mm", perturb(mm_in, 1e-5, sd))
    dd = np.abs(got - o["mm"])
    print("reference perturbed 1e-5 seed", sd, "max", dd.max(), "at", int(dd.argmax()), "near gpu pk:", dd[pk - 1500:pk + 1500].max())

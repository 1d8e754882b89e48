"""Diagnostic: the largest M&M / Costas junction residuals of a push and where they sit (segment index), per configuration."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from satdump_b200 import capi  # noqa: E402
from tests.common import demod_kwargs, nsamples, signal  # noqa: E402

for name in sys.argv[2:] or ["bpsk_half", "metop_ahrpt"]:
    cfg, raw, _ = signal(name, int(sys.argv[1]) if len(sys.argv) > 1 else 22)
    n = nsamples(raw, cfg)
    g = capi.Demod(capi.demod_cfg(max_batch=n, **demod_kwargs(cfg))).push(raw)
    cj, mj, L = g.junctions()
    a = np.abs(mj)
    order = np.argsort(-a)[:12]
    print(name, "L", L, "nseg", a.size, "mm top:", [(int(i), float("%.2e" % a[i])) for i in order], flush=True)
    print("   mm percentiles", [float("%.1e" % np.percentile(a[1:], p)) for p in (50, 90, 99, 99.9)], "repairs", g.stats()["repairs"])
    c = np.abs(cj[:, 0])
    order = np.argsort(-c)[:8]
    print("   costas top:", [(int(i), float("%.2e" % c[i])) for i in order])

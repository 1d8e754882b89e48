"""Diagnostic: where the CADU sequence of a bench workload departs from the transmitted frames (GPU box)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from satdump_b200 import capi
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c4"]
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 27
cfg, raw, clear = bench.make_workload(w, lg, 0, "cuda", None)
n = bench.nsamples_of(raw, cfg)
pipe = bench.ChainPipe(capi, cfg, n, 0)
import torch
host = torch.empty(raw.shape, dtype=raw.dtype, pin_memory=True); host.copy_(raw); torch.cuda.synchronize()
out = torch.empty(pipe.out_cap, dtype=torch.uint8, pin_memory=True)
pipe.reset(); pipe.push_host(host.data_ptr(), n); nb = pipe.pull_into(out.data_ptr(), out.numel())
fr = out[:nb].numpy().reshape(-1, cfg.cadu_bytes)
first = next((i for i in range(min(256, clear.shape[0])) if np.array_equal(clear[i], fr[0])), None)
print("frames", fr.shape[0], "clear", clear.shape[0], "first", first, "stats", pipe.stats())
if first is not None:
    k = min(fr.shape[0], clear.shape[0] - first)
    eq = (fr[:k] == clear[first:first + k]).all(axis=1)
    bad = np.nonzero(~eq)[0]
    print("equal prefix", k if bad.size == 0 else int(bad[0]), "of", k)
    if bad.size:
        b = int(bad[0])
        # does frame b match a later transmitted frame (a gap)?
        for d in range(1, 6):
            if first + b + d < clear.shape[0] and np.array_equal(fr[b], clear[first + b + d]):
                print("   frame", b, "equals transmitted frame", first + b + d, "-> gap of", d)
        print("   bytes differing in frame", b, int((fr[b] != clear[first + b]).sum()))

"""Diagnostic (not a test, not the bench): ONE synchronous step of a bench workload, for `ncu`.

  ncu --set full --clock-control none --import-source on -k regex:'^k_(agc_fir|costas|mm|vit_acs|vit_tb)' -o gpurun_out/x \
      python tools/profile_step.py c3 26 [steps [Es/N0 dB]]

The signal is generated on the GPU by torch (those kernels are filtered out by the -k regex); the step goes through the public C ABI
(b200_chain_push_iq_device) in synchronous mode, so every kernel runs alone and in stream order.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from satdump_b200 import capi  # noqa: E402


def main():
    wname = sys.argv[1] if len(sys.argv) > 1 else "c3"
    lg = int(sys.argv[2]) if len(sys.argv) > 2 else 26
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    esn0 = float(sys.argv[4]) if len(sys.argv) > 4 else None
    w = bench.WORKLOADS[wname]
    cfg, raw, _ = bench.make_workload(w, lg, 0, "cuda", esn0)
    n = bench.nsamples_of(raw, cfg)
    if w["kind"] == "chain":
        pipe = bench.ChainPipe(capi, cfg, n, 0)
    else:
        pipe = bench.DemodPipe(capi, cfg, n, 0)
    import torch
    torch.cuda.synchronize()
    for _ in range(steps):
        pipe.reset()
        pipe.push_device(raw.data_ptr(), n)
    torch.cuda.synchronize()
    print("timing", pipe.timing(), "stats", pipe.stats())


if __name__ == "__main__":
    main()

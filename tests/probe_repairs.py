"""Diagnostic: junction-repair statistics of a large batch (python tests/probe_repairs.py [log2n])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satdump_b200 import capi, synth
from tests.common import gpu_chain, nsamples
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 27
cfg = synth.CONFIGS["metop_ahrpt"]
raw, clear = synth.make_signal(cfg, 1 << lg, seed=5, device="cuda")
n = raw.numel() // 2
ch = gpu_chain(cfg, n)
ch.push_device(raw.data_ptr(), n)
ds, fs = ch.stats()
print("rounds", os.environ.get("B200_REPAIR_ROUNDS", "8"), {k: ds[k] for k in ("costas_unconverged", "mm_unconverged", "repairs")}, {k: fs[k] for k in ("replays", "frames_out")}, ch.timing())

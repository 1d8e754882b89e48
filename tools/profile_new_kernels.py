"""Diagnostic for `ncu --metrics gpu__time_duration.sum -k regex:'^k_(pll|rotator|dmx|agc_fir|costas|mm)'`: one pm_demod push (2^22 samples),
one psk_demod push with freq_shift, and one packet-demultiplexer push (65 536 CADUs), each after a warm-up push."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satdump_b200 import capi, synth  # noqa: E402
from tests import common  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    for name in ("pm_bpsk", "pm_bpsk_after"):
        cfg = synth.CONFIGS[name]
        raw, _ = synth.make_signal(cfg, 1 << lg, seed=1)
        raw = raw.numpy()
        g = common.gpu_demod(cfg, common.nsamples(raw, cfg))
        g.push(raw)
        g.reset()
        g.push(raw)
        print(name, g.timing(), {k: g.stats()[k] for k in ("symbols_out", "repairs", "pll_unconverged")})
    tile = synth.build_aos_frames(8192, seed=1, mpdu=882, insert_zone=2)
    fr = np.ascontiguousarray(np.tile(tile, (8, 1)))
    d = capi.Demux(1024, 882, 2, max_frames=fr.shape[0])
    d.run(fr)
    d.reset()
    b, r = d.run(fr)
    print("demux", fr.shape[0], "frames ->", r.shape[0], "packets,", b.size, "bytes")


if __name__ == "__main__":
    main()

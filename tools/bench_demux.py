"""Diagnostic (not the bench metric): throughput of the CADU -> space packet demultiplexer with the frames resident on the device.
   python tools/bench_demux.py [log2 frames]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satdump_b200 import capi, synth  # noqa: E402


def main():
    import torch
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    n = 1 << lg
    tile = synth.build_aos_frames(8192, seed=1, mpdu=882, insert_zone=2)
    fr = np.tile(tile, (n // tile.shape[0] + 1, 1))[:n]
    dev = torch.from_numpy(fr).cuda()
    g = capi.Demux(1024, 882, 2, max_frames=n)
    L = capi.lib()
    for _ in range(3):
        g.reset()
        capi._chk(L.b200_demux_push_frames_device(g.h, dev.data_ptr(), n))
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        g.reset()
        t = time.perf_counter()
        capi._chk(L.b200_demux_push_frames_device(g.h, dev.data_ptr(), n))
        ts.append(time.perf_counter() - t)
    ms = float(np.median(ts)) * 1e3
    st = g.stats()
    print(f"demux: {n} frames ({n * 1024 / 1e6:.0f} MB) in {ms:.3f} ms = {n * 1024 / ms / 1e6:.1f} GB/s of CADUs, {n / ms / 1e3:.2f} Mframes/s; "
          f"{st['packets_out'] // 13} packets per push")
    # the reference demultiplexer on one host core, same frames
    from oracle import ref
    if ref.available():
        m = min(n, 1 << 16)
        d = ref.Demux(882, 2)
        t = time.perf_counter()
        d.run(fr[:m])
        dt = time.perf_counter() - t
        print(f"reference (one core): {m} frames in {dt * 1e3:.1f} ms = {m * 1024 / dt / 1e9:.2f} GB/s")


if __name__ == "__main__":
    main()

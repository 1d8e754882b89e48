"""Diagnostic (not a test): run the CUDA path against the oracle stage by stage and print error statistics.
Usage on a GPU box:  python tests/probe_gpu.py [config] [log2 samples]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref as ORC  # noqa: E402
from satdump_b200 import capi, synth  # noqa: E402

if not ORC.available():
    from oracle import port as ORC  # noqa: E402,F811


def cmp_complex(name, a, b):
    n = min(a.size, b.size)
    d = np.abs(a[:n] - b[:n])
    print(f"  {name:8s} n={a.size}/{b.size} max|d|={d.max():.3e} mean={d.mean():.3e} >1e-5: {(d > 1e-5).mean() * 100:.4f}% "
          f">1e-4: {(d > 1e-4).sum()} first_bad={int(np.argmax(d > 1e-5)) if (d > 1e-5).any() else -1} bitwise={np.array_equal(a[:n].view(np.uint32), b[:n].view(np.uint32))}")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "metop_ahrpt"
    lg = int(sys.argv[2]) if len(sys.argv) > 2 else 21
    cfg = synth.CONFIGS[name]
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    t = time.time()
    raw, clear = synth.make_signal(cfg, 1 << lg, seed=1, device=dev)
    raw = raw.cpu().numpy()
    n = raw.size // 2 if cfg.fmt != "cf32" else raw.size
    print(f"signal {name}: {n} samples ({time.time() - t:.1f}s gen)")
    okw = dict(clock_alpha=cfg.clock_alpha) if cfg.clock_alpha else {}
    od = ORC.Demod(ORC.demod_cfg(cfg.samplerate, cfg.symbolrate, cfg.constellation if cfg.decoder != "none" else "none", cfg.rrc_alpha, cfg.pll_bw,
                                 cfg.fmt, **okw))
    t = time.time()
    o = od.run(raw)
    print(f"oracle demod {time.time() - t:.2f}s, {o['mm'].size} symbols; state {od.state()}")
    gd = capi.Demod(capi.demod_cfg(cfg.samplerate, cfg.symbolrate, cfg.constellation if cfg.decoder != "none" else "none", cfg.rrc_alpha, cfg.pll_bw,
                                   cfg.fmt, max_batch=max(n, 4096), keep_stages=True, **okw))
    rrc, bank = gd.taps()
    print("  rrc taps equal:", np.array_equal(rrc, od.rrc_taps() if hasattr(od, "rrc_taps") and ORC.__name__.endswith("ref") else rrc),
          " bank equal:", np.array_equal(bank, ORC.mm_taps()))
    t = time.time()
    gd.push(raw)
    print(f"gpu demod push {time.time() - t:.3f}s  stats {gd.stats()}")
    cmp_complex("agc", gd.stage("agc"), o["agc"])
    cmp_complex("fir", gd.stage("fir"), o["fir"])
    if o["costas"] is not None:
        cmp_complex("costas", gd.stage("costas"), o["costas"])
    gs = gd.symbols()
    cmp_complex("mm", gs, o["mm"])
    gsoft = gd.soft()
    m = min(gsoft.size, o["soft"].size)
    dsoft = np.abs(gsoft[:m].astype(int) - o["soft"][:m].astype(int))
    print(f"  soft     n={gsoft.size}/{o['soft'].size} differing bytes {np.count_nonzero(dsoft)} ({np.count_nonzero(dsoft) / m * 100:.4f}%), max {dsoft.max()}")
    # second push (streaming state carry): split the signal in two batches on a fresh object
    gd2 = capi.Demod(capi.demod_cfg(cfg.samplerate, cfg.symbolrate, cfg.constellation if cfg.decoder != "none" else "none", cfg.rrc_alpha, cfg.pll_bw,
                                    cfg.fmt, max_batch=max(n, 4096), keep_stages=True, **okw))
    per = 2 if cfg.fmt != "cf32" else 1
    cut = (n // 3) // 16 * 16 + 5
    gd2.push(raw[:cut * per])
    s1 = gd2.symbols()
    gd2.push(raw[cut * per:])
    s2 = gd2.symbols()
    cmp_complex("mm/2push", np.concatenate([s1, s2]), o["mm"])
    print("  stats2", gd2.stats())
    if cfg.decoder not in ("metop", "ccsds"):
        return
    fo = ORC.Fec(ORC.metop_cfg(cfg.ber_thresold, cfg.outsync_after) if cfg.decoder == "metop" else
                 ORC.ccsds_cfg(cfg.constellation, cfg.cadu_bytes * 8, cfg.ber_thresold, cfg.outsync_after, cfg.interleave, nrzm=cfg.nrzm,
                               rs_usecheck=cfg.rs_usecheck))
    t = time.time()
    fr = fo.run(o["soft"])
    print(f"oracle fec {time.time() - t:.2f}s: {fr['cadu'].size // fo.cadu_bytes} frames, vit_state {fr['vit_state'][:6]} ber {fr['vit_ber'][:4]}")
    gf = capi.Fec(capi.metop_cfg(cfg.ber_thresold, cfg.outsync_after, max_soft=max(o["soft"].size, 65536)) if cfg.decoder == "metop" else
                  capi.ccsds_cfg(cfg.constellation, cfg.cadu_bytes * 8, cfg.ber_thresold, cfg.outsync_after, cfg.interleave, nrzm=cfg.nrzm,
                                 rs_usecheck=cfg.rs_usecheck, max_soft=max(o["soft"].size, 65536)))
    t = time.time()
    gf.push(o["soft"])  # stage isolated: the oracle's exact soft stream
    print(f"gpu fec push {time.time() - t:.3f}s stats {gf.stats()}")
    gb = gf.bits()
    print(f"  bits n={gb.size}/{fr['bits'].size} equal={np.array_equal(gb, fr['bits'])}",
          "" if gb.size != fr['bits'].size else f"diff={np.count_nonzero(gb != fr['bits'])} first={int(np.argmax(gb != fr['bits'])) if (gb != fr['bits']).any() else -1}")
    gfr = gf.frames()
    ocadu = fr["cadu"].reshape(-1, fo.cadu_bytes)
    print(f"  frames {gfr.shape} vs {ocadu.shape} equal={np.array_equal(gfr, ocadu)}")
    if gfr.shape == ocadu.shape and not np.array_equal(gfr, ocadu):
        bad = np.nonzero((gfr != ocadu).any(axis=1))[0]
        print("   differing frames:", bad[:10], "bytes:", [(int(f), np.nonzero(gfr[f] != ocadu[f])[0][:8].tolist()) for f in bad[:3]])
    # end to end through the chain
    ch = capi.Chain(capi.demod_cfg(cfg.samplerate, cfg.symbolrate, cfg.constellation, cfg.rrc_alpha, cfg.pll_bw, cfg.fmt, max_batch=max(n, 4096)),
                    capi.metop_cfg(cfg.ber_thresold, cfg.outsync_after, max_soft=max(2 * o["soft"].size, 65536)) if cfg.decoder == "metop" else
                    capi.ccsds_cfg(cfg.constellation, cfg.cadu_bytes * 8, cfg.ber_thresold, cfg.outsync_after, cfg.interleave, nrzm=cfg.nrzm,
                                   rs_usecheck=cfg.rs_usecheck, max_soft=max(2 * o["soft"].size, 65536)))
    t = time.time()
    ch.push(raw)
    cfr = ch.frames()
    print(f"chain push {time.time() - t:.3f}s frames {cfr.shape} equal_to_oracle={cfr.shape == ocadu.shape and np.array_equal(cfr, ocadu)} timing {ch.timing()}")
    if cfr.shape == ocadu.shape and not np.array_equal(cfr, ocadu):
        bad = np.nonzero((cfr != ocadu).any(axis=1))[0]
        print("   differing frames:", bad[:10])
    print("   chain stats", ch.stats())


if __name__ == "__main__":
    main()

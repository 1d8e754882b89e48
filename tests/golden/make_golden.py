"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref (needs /root/reference, i.e. this
container). The reference ships no test vectors for this path, so these fixtures are the pin: synthetic
inputs (satdump_b200.synth, fixed seeds) and what the reference's own code produces for them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from satdump_b200 import synth  # noqa: E402
from tests.common import oracle_demod, oracle_fec  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    assert ref.available(), "build oracle/_ref first (python -c 'import __graft_entry__ as g; g.build()')"
    only = sys.argv[1:]
    for name, lg in [("metop_ahrpt", 17), ("bpsk_half", 16), ("jpss_hrd", 17), ("dvbs2_front", 16), ("hrpt_bpsk", 18), ("metop_oversampled", 19), ("bpsk_decim8", 20), ("qpsk_undersampled", 16), ("psk8", 17), ("bpsk_simple", 17), ("qpsk_simple", 17),
                     ("qpsk_p34", 18), ("qpsk_p78", 18),
                     ("pm_bpsk", 19)]:  # pm_demod (carrier PLL -> PMToBPSK -> RRC -> Costas -> M&M) -> ccsds_conv_concat_decoder  # Viterbi_Depunc rates (conv_rate 3/4 is the one a shipped pipeline uses, 7/8 the most punctured)
        if only and name not in only:
            continue
        cfg = synth.CONFIGS[name]
        raw, clear = synth.make_signal(cfg, 1 << lg, seed=0x600D, device="cpu")
        raw = raw.numpy()
        o = oracle_demod(ref, cfg).run(raw)
        d = dict(raw=raw, soft=o["soft"], mm_head=o["mm"][:4096], fir_head=o["fir"][:4096], agc_head=o["agc"][:4096], nsym=np.int64(o["mm"].size))
        if o["costas"] is not None:
            d["costas_head"] = o["costas"][:4096]
        if cfg.pm_index:
            d.update(pll_head=o["pll"][:4096], pm_head=o["pm"][:4096])
        dc = oracle_demod(ref, cfg).cfg
        if dc.final_samplerate > 0:  # the front-end resampler ran: pin its output and its bank too
            I, D = int(dc.final_samplerate), int(dc.samplerate)
            d.update(resamp_head=ref.resample(dc, raw)[:4096], front=np.int64(o["front"]), resamp_bank=ref.resampler_taps(I, D))
        if cfg.decoder in ("metop", "ccsds", "simple"):
            f = oracle_fec(ref, cfg).run(o["soft"])
            d.update(cadu=f["cadu"], bits=np.packbits(f["bits"]), nbits=np.int64(f["bits"].size), vit_state=f["vit_state"], defr_state=f["defr_state"],
                     rs_err=f["rs_err"])
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **d)
        print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items()})
    if not only or "simple_psk" in only:
        from tests.common import simple_soft_cases
        cases, _ = simple_soft_cases()
        d = {}
        for name, kw, soft in cases:
            f = ref.Fec(ref.simple_cfg(cadu_size=8192, rs_i=4, **kw))
            r = f.run(soft[:soft.size // f.chunk * f.chunk])
            d[f"{name}_soft"], d[f"{name}_cadu"], d[f"{name}_bits"] = soft, r["cadu"], np.packbits(r["bits"])
        np.savez_compressed(os.path.join(OUT, "simple_psk.npz"), **d)
        print("simple_psk", {k: v.shape for k, v in d.items() if k.endswith("_cadu")})
    if only:
        return
    # FEC stress vectors: RS decoder on codewords with 0..20 byte errors (beyond-capacity ones must fail the same way)
    rng = np.random.default_rng(0xFEC)
    pay = rng.integers(0, 256, size=(64, 4 * 223), dtype=np.uint8)
    tx, clear = synth.build_cadus(pay, 4)
    noisy = clear[:, 4:].copy()
    for f in range(noisy.shape[0]):
        for b in range(4):
            ne = int(rng.integers(0, 21))
            pos = rng.choice(255, size=ne, replace=False)
            noisy[f, pos * 4 + b] ^= rng.integers(1, 256, size=ne, dtype=np.uint8)
    dec = np.zeros_like(noisy)
    errs = np.zeros((noisy.shape[0], 4), np.int32)
    for f in range(noisy.shape[0]):
        dec[f], errs[f] = ref.rs_decode_interleaved(noisy[f], True, 4)
    np.savez_compressed(os.path.join(OUT, "rs_stress.npz"), noisy=noisy, decoded=dec, errors=errs)
    print("rs_stress", errs[:4].tolist(), "failures:", int((errs < 0).sum()))


if __name__ == "__main__":
    main()

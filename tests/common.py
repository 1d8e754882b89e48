"""Shared helpers of the test-suite: config plumbing between the synthetic transmitter, the oracle and the C ABI."""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from satdump_b200 import synth  # noqa: E402


def oracle():
    """The checker: the compiled reference (oracle/_ref) when present, else our C restatement."""
    from oracle import ref
    if ref.available():
        return ref
    from oracle import port
    return port


def rx_const(cfg):
    return cfg.constellation if cfg.decoder != "none" else "none"  # decoder "demod": psk_demod alone, with its Costas loop


def demod_kwargs(cfg):
    kw = dict(samplerate=cfg.samplerate, symbolrate=cfg.symbolrate, constellation=rx_const(cfg), rrc_alpha=cfg.rrc_alpha, pll_bw=cfg.pll_bw,
              fmt=cfg.fmt)
    if cfg.clock_alpha:
        kw["clock_alpha"] = cfg.clock_alpha
    return kw


def oracle_demod(O, cfg):
    return O.Demod(O.demod_cfg(**demod_kwargs(cfg)))


def oracle_fec(O, cfg):
    if cfg.decoder == "metop":
        return O.Fec(O.metop_cfg(cfg.ber_thresold, cfg.outsync_after))
    return O.Fec(O.ccsds_cfg(cfg.constellation, cfg.cadu_bytes * 8, cfg.ber_thresold, cfg.outsync_after, cfg.interleave, nrzm=cfg.nrzm,
                             rs_usecheck=cfg.rs_usecheck))


def gpu_demod(cfg, n, keep_stages=False):
    from satdump_b200 import capi
    return capi.Demod(capi.demod_cfg(max_batch=max(n, 4096), keep_stages=keep_stages, **demod_kwargs(cfg)))


def gpu_fec_cfg(cfg, max_soft):
    from satdump_b200 import capi
    if cfg.decoder == "metop":
        return capi.metop_cfg(cfg.ber_thresold, cfg.outsync_after, max_soft=max(max_soft, 65536))
    return capi.ccsds_cfg(cfg.constellation, cfg.cadu_bytes * 8, cfg.ber_thresold, cfg.outsync_after, cfg.interleave, nrzm=cfg.nrzm,
                          rs_usecheck=cfg.rs_usecheck, max_soft=max(max_soft, 65536))


def gpu_chain(cfg, n):
    from satdump_b200 import capi
    return capi.Chain(capi.demod_cfg(max_batch=max(n, 4096), **demod_kwargs(cfg)), gpu_fec_cfg(cfg, 2 * n))


def nsamples(raw, cfg):
    return raw.size if cfg.fmt == "cf32" else raw.size // 2


@functools.lru_cache(maxsize=16)
def signal(name, log2n, seed=1, esn0=None):
    """(raw numpy array in the config's format, clear CADUs) — generated on the GPU when there is one."""
    import dataclasses
    import torch
    cfg = synth.CONFIGS[name]
    if esn0 is not None:
        cfg = dataclasses.replace(cfg, esn0_db=esn0)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    raw, clear = synth.make_signal(cfg, 1 << log2n, seed=seed, device=dev)
    return cfg, raw.cpu().numpy(), clear


def match_frames(got, clear):
    """Index in `clear` of got[0] and whether got is the consecutive run of transmitted frames from there."""
    if got.shape[0] == 0:
        return None, False
    for i in range(min(256, clear.shape[0])):
        if np.array_equal(clear[i], got[0]):
            ok = i + got.shape[0] <= clear.shape[0] and np.array_equal(got, clear[i:i + got.shape[0]])
            return i, ok
    return None, False

"""Float-parity floors, measured on the reference itself (TEST INFRASTRUCTURE: drives oracle/_ref only).

The feedback stages of the demodulator amplify tiny input differences: the M&M loop picks its interpolator arm with
rint(mu * 128) (clock_recovery_mm.cpp:67), the Costas detectors multiply by sgn(v.re) / sgn(v.im) (costas_loop.cpp:31-52). Fed an
input that differs by 1e-6 the REFERENCE's own output differs from its unperturbed output by an interpolator arm on 0.1-1 % of the
symbols (self-sustaining) and, whenever a loop sample lands within the perturbation of zero, by a phase kick of ~alpha that takes
~1000 samples to decay. No implementation that is not bit-identical upstream can be closer to the reference than the reference is to
itself under such a perturbation, so the gates of the GPU parity tests are

    gate = max(SURVEY.md 8c gate, 1.2 x floor)

with the floor measured here, per configuration and stage, as the worst of
  * `seeds` runs of the reference with additive uniform noise of amplitude `eps` on the stage input (default 1e-6: the size of the
    deviation the parallel AGC evaluation has from the reference's float recurrence, mean 4e-7 / max 5e-6, itself gated at 1e-5), and
  * the reference compiled with its release flags (-O3, FMA contraction: oracle/_ref/libsatref_native.so) against the -O2 build.
The maxima are extreme-value statistics of a chaotic process (one more interpolator arm, one more sign event), hence eight seeds per
stage. Everything is deterministic (fixed seeds), so a gate is the same number on every run."""
import functools
import json
import os

import numpy as np

from tests.common import demod_kwargs, signal

SURVEY = dict(mm_frac=0.01, mm_max=1.5e-2, soft_diff=0.005, soft_gt1=0.0, float_all=1e-5)


def _O():
    from oracle import ref
    assert ref.available(), "floors need the compiled reference (oracle/_ref)"
    return ref


def perturb(x, eps, seed):
    rng = np.random.default_rng(seed)
    n = x.size
    return (x + eps * ((rng.random(n) * 2 - 1) + 1j * (rng.random(n) * 2 - 1))).astype(np.complex64)


def ulp_perturb(x, seed):
    """+-1 ulp (or 0) on every float of x; zeros stay zero."""
    rng = np.random.default_rng(seed)
    xf = np.ascontiguousarray(x).view(np.float32).copy()
    step = (rng.integers(0, 3, xf.size) - 1).astype(np.int32)
    step[(xf.view(np.int32) & 0x7FFFFFFF) == 0] = 0
    return (xf.view(np.int32) + step).view(np.float32).view(np.complex64)


def diff_stats(a, b):
    assert a.size == b.size, (a.size, b.size)
    d = np.abs(a - b)
    return dict(frac=float((d > 1e-5).mean()), max=float(d.max()), mean=float(d.mean()), median=float(np.median(d)))


def soft_stats(a, b):
    assert a.size == b.size, (a.size, b.size)
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return dict(diff=float((d > 0).mean()), gt1=float((d > 1).mean()), max=int(d.max()))


def quantise(sym, bpsk):
    """module_psk_demod.cpp:199-213 + clamp (module_demod_base.h:106-113)"""
    v = (sym.real * np.float32(50)) if bpsk else np.stack([sym.real * np.float32(100), sym.imag * np.float32(100)], axis=1).reshape(-1)
    q = np.trunc(v)
    q = np.where(v < -128, -127, np.where(v > 127, 127, q))
    return q.astype(np.int8)


@functools.lru_cache(maxsize=64)
def reference_run(name, log2n, extra=()):
    """`extra`: further demodulator options as a tuple of (key, value) pairs, e.g. (("post_costas_dc", True),)"""
    O = _O()
    cfg, raw, _ = signal(name, log2n)
    oc = O.demod_cfg(**demod_kwargs(cfg), **dict(extra))
    return cfg, raw, oc, O.Demod(oc).run(raw)


def _worst(stats):
    return {k: max(s[k] for s in stats) for k in stats[0]}


@functools.lru_cache(maxsize=128)
def stage_floor(name, log2n, stage, eps=1e-6, seeds=(1, 2, 3, 4, 5, 6, 7, 8), extra=()):
    """Worst deviation of the reference's `stage` output ("costas" / "mm") from itself when its stage input is perturbed."""
    O = _O()
    cfg, raw, oc, o = reference_run(name, log2n, extra)
    src = o["fir"] if (stage == "costas" or o["costas"] is None) else o["costas"]
    want = o[stage]
    out = []
    for sd in seeds:
        got = O.run_stage(oc, stage, perturb(src, eps, sd) if eps > 0 else ulp_perturb(src, sd))
        s = diff_stats(got, want)
        if stage == "mm":
            s |= {"soft_" + k: v for k, v in soft_stats(quantise(got, cfg.constellation == "bpsk" and cfg.decoder != "none"), o["soft"]).items()}
        out.append(s)
    w = _worst(out)
    if stage == "mm" and eps > 0:
        # The largest deviations are excursions at a few fragile spots of the signal (one or two interpolator arms for a few hundred
        # symbols); which of them a run triggers depends on the perturbation. They are sampled with more seeds and at 1e-5, the
        # deviation SURVEY 8c itself allows the stages in front (AGC / FIR / Costas <= 1e-5): bpsk_half shows 1.8e-2 in 24 of 24 runs at
        # 1e-6 and 2.1e-2 ... 3.2e-2 at 1e-5. Only the extreme-value statistics (max, more-than-1-LSB soft bytes) take these runs.
        for sd in range(1, 17):
            got = O.run_stage(oc, stage, perturb(src, 1e-5, sd))
            d = diff_stats(got, want)
            q = soft_stats(quantise(got, cfg.constellation == "bpsk" and cfg.decoder != "none"), o["soft"])
            w["max"] = max(w["max"], d["max"])
            w["soft_gt1"] = max(w["soft_gt1"], q["gt1"])
            w["soft_max"] = max(w["soft_max"], q["max"])
    return w


@functools.lru_cache(maxsize=64)
def chain_floor(name, log2n, eps=1e-6, seeds=(1, 2, 3), extra=()):
    """Worst deviation of the reference's Costas / M&M / soft outputs from themselves when the AGC output is perturbed by eps, and
    when the reference is built with its release flags (FMA contraction)."""
    O = _O()
    cfg, raw, oc, o = reference_run(name, log2n, extra)
    bpsk = o["soft"].size == o["mm"].size
    runs = []
    for sd in seeds:
        fir = O.run_stage(oc, "fir", perturb(o["agc"], eps, sd))
        r = {}
        mm_in = fir
        if o["costas"] is not None:
            mm_in = O.run_stage(oc, "costas", fir)
            r |= {"costas_" + k: v for k, v in diff_stats(mm_in, o["costas"]).items()}
        mm = O.run_stage(oc, "mm", mm_in)
        r |= {"mm_" + k: v for k, v in diff_stats(mm, o["mm"]).items()}
        r |= {"soft_" + k: v for k, v in soft_stats(quantise(mm, bpsk), o["soft"]).items()}
        runs.append(r)
    from oracle import ref_native as N
    if N.available():
        n = N.Demod(oc).run(raw)
        r = {}
        if o["costas"] is not None:
            r |= {"costas_" + k: v for k, v in diff_stats(n["costas"], o["costas"]).items()}
        r |= {"mm_" + k: v for k, v in diff_stats(n["mm"], o["mm"]).items()}
        r |= {"soft_" + k: v for k, v in soft_stats(n["soft"], o["soft"]).items()}
        runs.append(r)
    w = _worst(runs)
    # the chain cannot be better than its stages fed a perturbed input directly (the loops' rare events are a matter of which
    # samples the perturbation happens to hit: take the worst of both injection points)
    m = stage_floor(name, log2n, "mm", eps, extra=extra)
    for k in ("frac", "max"):
        w["mm_" + k] = max(w["mm_" + k], m[k])
    for k in ("diff", "gt1", "max"):
        w["soft_" + k] = max(w["soft_" + k], m["soft_" + k])
    if o["costas"] is not None:
        c = stage_floor(name, log2n, "costas", eps, extra=extra)
        for k in ("frac", "max", "mean"):
            w["costas_" + k] = max(w["costas_" + k], c[k])
    return w


# ---- committed cache: the floors are deterministic functions of the committed synthetic signals and the compiled reference, and cost
# 10-20 s of CPU each; tests/golden/floors.json holds them (tests/golden/make_floors.py regenerates it, tests/test_floors.py re-measures
# a sample of the entries against it), so the GPU box does not spend its minutes on them.
_CACHE_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "floors.json")
_cache = None


def _key(kind, name, log2n, stage, eps, extra):
    return "|".join([kind, name, str(log2n), str(stage), repr(float(eps)), repr(tuple(extra))])


def cached(kind, name, log2n, stage=None, eps=1e-6, extra=()):
    """stage_floor / chain_floor through the committed cache (computed and remembered in memory when the entry is missing)."""
    global _cache
    if _cache is None:
        try:
            with open(_CACHE_PATH) as f:
                _cache = json.load(f)
        except Exception:
            _cache = {}
    k = _key(kind, name, log2n, stage, eps, extra)
    if k not in _cache:
        _cache[k] = stage_floor(name, log2n, stage, eps, extra=tuple(extra)) if kind == "stage" else chain_floor(name, log2n, eps, extra=tuple(extra))
    return _cache[k]


def gate(survey, floor):
    return max(survey, 1.2 * floor)

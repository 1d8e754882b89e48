"""GPU: ccsds_conv_concat_decoder with conv_rate 2/3, 3/4, 5/6, 7/8 (Viterbi_Depunc, common/codings/viterbi/viterbi_punc.cpp + depunc.h)
through the C ABI, bit-exact against the reference on the reference's own soft stream: decoded bits, lock states, CADUs - one shot, in
ragged pushes, and through noise -> signal -> noise (lock search, unlock after viterbi_outsync_after bad calls, lock again with the sliding
buffer's leftover and the depuncturer's held-back symbol carried over)."""
import numpy as np
import pytest

from tests.common import oracle, oracle_demod, oracle_fec, signal

pytestmark = pytest.mark.gpu
RATES = ["qpsk_p23", "qpsk_p34", "qpsk_p56", "qpsk_p78"]


def _soft(name, lg=21):
    O = oracle()
    cfg, raw, clear = signal(name, lg)
    return O, cfg, oracle_demod(O, cfg).run(raw, stages=False)["soft"]


def _gpu(cfg, max_soft):
    from satdump_b200 import capi
    return capi.Fec(capi.fec_cfg_for(cfg, max(max_soft, 65536)))


@pytest.mark.parametrize("name", RATES)
def test_punctured_rates_bit_exact(built, name):
    O, cfg, soft = _soft(name)
    want = oracle_fec(O, cfg).run(soft)
    g = _gpu(cfg, soft.size)
    g.push(soft)
    assert np.array_equal(g.bits(), want["bits"]), (g.bits().size, want["bits"].size)
    got = g.frames()
    ocadu = want["cadu"].reshape(-1, cfg.cadu_bytes)
    assert ocadu.shape[0] >= 100 and got.shape == ocadu.shape and np.array_equal(got, ocadu)
    s = g.stats()
    assert s["viterbi_state"] == int(want["vit_state"][-1])


@pytest.mark.parametrize("name", ["qpsk_p34", "qpsk_p78"])
def test_punctured_ragged_pushes(built, name):
    O, cfg, soft = _soft(name)
    want = oracle_fec(O, cfg).run(soft)["cadu"].reshape(-1, cfg.cadu_bytes)
    g = _gpu(cfg, soft.size)
    parts, prev = [], 0
    for c in [8192 * 3 + 5, 100001, 700000, soft.size]:
        g.push(soft[prev:c])
        parts.append(g.frames())
        prev = c
    got = np.concatenate(parts)
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("name", ["qpsk_p23", "qpsk_p56"])
def test_punctured_noise_signal_noise(built, name):
    """lock search on noise, lock, unlock on noise, lock again: the lock states per module call and every decoded bit follow the reference"""
    O, cfg, soft = _soft(name, 20)
    rng = np.random.default_rng(5)
    noise = lambda n: rng.integers(-60, 61, n).astype(np.int8)
    stream = np.concatenate([noise(8192 * 5 + 16), soft[:400000], noise(8192 * 40 + 2), soft[400000:], noise(8192 * 3)])
    want = oracle_fec(O, cfg).run(stream)
    assert (want["vit_state"] == 0).any() and (want["vit_state"] > 0).any() and want["cadu"].size >= 30 * cfg.cadu_bytes
    for cuts in ([stream.size], [8192 * 7, 300001, 8192 * 70 + 11, stream.size]):
        g = _gpu(cfg, stream.size)
        bits, frames, prev = [], [], 0
        for c in cuts:
            g.push(stream[prev:c])
            bits.append(g.bits())
            frames.append(g.frames())
            prev = c
        assert np.array_equal(np.concatenate(bits), want["bits"]), cuts
        got = np.concatenate(frames)
        ocadu = want["cadu"].reshape(-1, cfg.cadu_bytes)
        assert got.shape == ocadu.shape and np.array_equal(got, ocadu), cuts


@pytest.mark.parametrize("name", ["qpsk_p34", "qpsk_p78"])
def test_punctured_end_to_end_from_iq(built, name):
    """raw IQ -> psk_demod -> ccsds_conv_concat_decoder(conv_rate) through the fused chain (soft symbols stay in HBM), synchronous and
    pipelined: the CADUs are the reference's, and the golden fixture's soft stream decodes to the golden CADUs."""
    import os
    from satdump_b200 import capi
    from tests.common import demod_kwargs, nsamples
    O = oracle()
    cfg, raw, _ = signal(name, 21)
    n = nsamples(raw, cfg)
    want = oracle_fec(O, cfg).run(oracle_demod(O, cfg).run(raw, stages=False)["soft"])["cadu"].reshape(-1, cfg.cadu_bytes)
    assert want.shape[0] >= 100
    for pipelined in (False, True):
        ch = capi.Chain(capi.demod_cfg(max_batch=n, **demod_kwargs(cfg)), capi.fec_cfg_for(cfg, 2 * n))
        if pipelined:
            ch.set_pipelined(True)
        per = 1 if cfg.fmt == "cf32" else 2
        parts, prev = [], 0
        for c in (700001, n):
            ch.push(raw[prev * per:c * per])
            parts.append(ch.frames())
            prev = c
        ch.sync()
        parts.append(ch.frames())
        got = np.concatenate(parts)
        assert got.shape == want.shape and np.array_equal(got, want), pipelined
        ch.close()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"{name}.npz"))
    f = _gpu(cfg, g["soft"].size)
    f.push(g["soft"])
    assert np.array_equal(f.frames().reshape(-1), g["cadu"]) and np.array_equal(np.packbits(f.bits()), g["bits"])

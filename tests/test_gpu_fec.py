"""GPU: decoder parity. All stages are integer: given the oracle's exact int8 soft stream the CUDA path must reproduce the
decoded bits, the lock decisions and the CADUs bit for bit (including RS failures and uncorrected parity bytes)."""
import os

import numpy as np
import pytest

from tests.common import ROOT, gpu_fec_cfg, oracle, oracle_demod, oracle_fec, signal, simple_soft_cases
from satdump_b200 import synth

pytestmark = pytest.mark.gpu
DECODED = ["metop_ahrpt", "bpsk_half", "jpss_hrd"]


def gpu_fec(cfg, max_soft):
    from satdump_b200 import capi
    return capi.Fec(gpu_fec_cfg(cfg, max_soft))


@pytest.mark.parametrize("name", DECODED)
def test_stage_isolated_bit_exact(built, name):
    O = oracle()
    cfg, raw, _ = signal(name, 21)
    soft = oracle_demod(O, cfg).run(raw, stages=False)["soft"]
    f = oracle_fec(O, cfg)
    want = f.run(soft)
    g = gpu_fec(cfg, soft.size).push(soft)
    assert np.array_equal(g.bits(), want["bits"])
    got = g.frames()
    assert got.shape[0] > 0 and np.array_equal(got.reshape(-1), want["cadu"])
    s = g.stats()
    assert s["viterbi_state"] == int(want["vit_state"][-1]) and s["deframer_state"] == int(want["defr_state"][-1])
    assert s["replays"] == 0
    ok = want["rs_err"][want["rs_err"] >= 0]
    assert s["rs_corrected"] == int(ok.sum()) and s["rs_failed"] == int((want["rs_err"] < 0).sum())


@pytest.mark.parametrize("name,esn0", [("metop_ahrpt", 5.5), ("metop_ahrpt", 4.6), ("jpss_hrd", 3.2), ("bpsk_half", 0.6)])
def test_low_snr_stress_bit_exact(built, name, esn0):
    """Viterbi error bursts, RS corrections and failures, parity bytes left uncorrected: still identical to the reference."""
    O = oracle()
    cfg, raw, _ = signal(name, 21, seed=9, esn0=esn0)
    soft = oracle_demod(O, cfg).run(raw, stages=False)["soft"]
    want = oracle_fec(O, cfg).run(soft)
    g = gpu_fec(cfg, soft.size).push(soft)
    assert np.array_equal(g.bits(), want["bits"])
    assert np.array_equal(g.frames().reshape(-1), want["cadu"])
    assert want["rs_err"].size and (want["rs_err"] != 0).any(), "stress case should make RS work"


@pytest.mark.parametrize("name", DECODED)
def test_ragged_streaming_pushes(built, name):
    """Soft bytes arrive in arbitrary pieces: chunking, lock state, NRZ-M carry, deframer shifter and open frames carry over."""
    O = oracle()
    cfg, raw, _ = signal(name, 21)
    soft = oracle_demod(O, cfg).run(raw, stages=False)["soft"]
    want = oracle_fec(O, cfg).run(soft)
    g = gpu_fec(cfg, soft.size)
    cuts = [0, 1, 70001, 70002, 300000, 300000 + g.chunk, soft.size // 2 + 3, soft.size]
    frames, bits = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        g.push(soft[a:b])
        frames.append(g.frames())
        bits.append(g.bits())
    assert np.array_equal(np.concatenate(bits), want["bits"])
    assert np.array_equal(np.concatenate(frames).reshape(-1), want["cadu"])


def test_noise_then_signal_then_noise(built):
    """Lock machine: IDLE over noise (no output), lock when the signal starts, unlock after > outsync bad chunks."""
    O = oracle()
    cfg, raw, _ = signal("metop_ahrpt", 21)
    soft = oracle_demod(O, cfg).run(raw, stages=False)["soft"]
    rng = np.random.default_rng(4)
    noise = lambda n: np.clip(rng.normal(0, 40, n), -127, 127).astype(np.int8)
    # (offsets are multiples of 4 soft bytes: an odd offset would split I/Q pairs and leave the lock test on a knife edge, where the
    # reference itself is not reproducible — its BER scratch buffer has never-written bytes)
    stream = np.concatenate([noise(16384 * 5 + 76), soft[:16384 * 40], noise(16384 * 30), soft[16384 * 40:16384 * 70]])
    want = oracle_fec(O, cfg).run(stream)
    assert 0 in want["vit_state"] and 1 in want["vit_state"]
    g = gpu_fec(cfg, stream.size).push(stream)
    assert np.array_equal(g.bits(), want["bits"])
    assert np.array_equal(g.frames().reshape(-1), want["cadu"])
    assert g.stats()["viterbi_state"] == int(want["vit_state"][-1])
    # and the same stream in two pushes that split the noisy stretch
    g2 = gpu_fec(cfg, stream.size)
    cut = 16384 * 52 + 5
    fr = [g2.push(stream[:cut]).frames(), g2.push(stream[cut:]).frames()]
    assert np.array_equal(np.concatenate(fr).reshape(-1), want["cadu"])


def test_rs_golden_through_the_decoder(built):
    """tests/golden/rs_stress.npz (made by the reference): codewords with 0..20 byte errors, sent noise-free through the conv.
    code so that the Viterbi hands the RS kernel exactly those bytes; corrections and failures must equal the reference's."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "rs_stress.npz"))
    nf = g["noisy"].shape[0]
    frames = np.concatenate([np.tile(np.frombuffer(synth.ASM, np.uint8), (nf, 1)), g["noisy"] ^ synth.ccsds_pn(1020)[None, :]], axis=1)
    # lead-in / lead-out so that the lock search and the chunking see whole chunks
    pad = np.tile(frames[:1], (4, 1))
    bits = np.unpackbits(np.concatenate([pad, frames, pad]).reshape(-1))
    coded = synth.conv_encode(bits)
    # +-60: a clean +-100 stream wraps the reference's 8-bit path metrics (spread > 255) and would not even lock
    soft = np.where(coded > 0, 60, -60).astype(np.int8)
    soft = soft[:soft.size // 8192 * 8192]
    from satdump_b200 import capi
    cfg = capi.ccsds_cfg("qpsk", 8192, 0.3, 20, 4, max_soft=soft.size)
    got = capi.Fec(cfg).push(soft).frames()
    O = oracle()
    want = O.Fec(O.ccsds_cfg("qpsk", 8192, 0.3, 20, 4)).run(soft)["cadu"].reshape(-1, 1024)
    assert want.shape[0] >= nf and np.array_equal(got, want)
    # the frames that carry the golden codewords decode to the golden result
    body = got[:, 4:]
    hits = sum(any(np.array_equal(body[k], g["decoded"][f]) for k in range(body.shape[0])) for f in range(nf))
    assert hits >= nf - 2, hits


def test_metric_wraparound_matches_reference(built):
    """A clean +-100 soft stream makes the reference's uint8 path metrics wrap (spread > 255): it decodes with errors (BER
    ~0.25) — and the CUDA ACS must wrap the same way, bit for bit."""
    rng = np.random.default_rng(8)
    bits = rng.integers(0, 2, 8192 * 24, dtype=np.uint8)
    soft = np.where(synth.conv_encode(bits) > 0, 100, -100).astype(np.int8)
    O = oracle()
    want = O.Fec(O.ccsds_cfg("qpsk", 8192, 0.3, 20, 0)).run(soft)
    from satdump_b200 import capi
    g = capi.Fec(capi.ccsds_cfg("qpsk", 8192, 0.3, 20, 0, max_soft=soft.size)).push(soft)
    assert want["bits"].size > 0 and (want["bits"][:20000] != bits[:20000]).any()
    assert np.array_equal(g.bits(), want["bits"])


def test_soft_fifo_overflow_is_loud(built):
    from satdump_b200 import capi
    f = capi.Fec(capi.metop_cfg(max_soft=65536))
    with pytest.raises(capi.B200Error) as e:
        f.push(np.zeros(1 << 20, np.int8))
    assert e.value.code == -5


SIMPLE = ["bpsk", "bpsk_nrzm", "qpsk_rot0", "qpsk_rot1", "qpsk_rot2", "qpsk_rot3", "qpsk_swapiq_delay", "qpsk_diff_swap1", "qpsk_diff_swap0"]


@pytest.mark.parametrize("mode", SIMPLE)
def test_simple_psk_decoder_bit_exact(built, mode):
    """ccsds_simple_psk_decoder (no convolutional code): hard decisions, NRZ-M / QPSKDiff, the two deframers of plain QPSK, derandomiser
    and RS — bits, frames, deframer state and RS statistics equal the reference's, one shot and in ragged pushes."""
    from satdump_b200 import capi
    O = oracle()
    cases, clear = simple_soft_cases()
    name, kw, soft = next(c for c in cases if c[0] == mode)
    f = O.Fec(O.simple_cfg(cadu_size=8192, rs_i=4, **kw))
    n = soft.size // f.chunk * f.chunk
    want = f.run(soft[:n])
    assert want["cadu"].size >= 20 * 1024 and (want["rs_err"] > 0).any(), "the case should decode frames and give RS work"
    g = capi.Fec(capi.simple_cfg(cadu_size=8192, rs_i=4, max_soft=max(soft.size, 65536), **kw)).push(soft)
    assert np.array_equal(g.bits(), want["bits"])
    assert np.array_equal(g.frames().reshape(-1), want["cadu"])
    s = g.stats()
    assert s["deframer_state"] == int(want["defr_state"][-1]) and s["replays"] == 0
    ok = want["rs_err"][want["rs_err"] >= 0]
    assert s["rs_corrected"] == int(ok.sum()) and s["rs_failed"] == int((want["rs_err"] < 0).sum())
    # ragged pushes: the soft FIFO, the NRZ-M / QPSKDiff / oqpsk_delay registers and both deframers carry over
    g2 = capi.Fec(capi.simple_cfg(cadu_size=8192, rs_i=4, max_soft=max(soft.size, 65536), **kw))
    parts, prev = [], 0
    for c in [5001, 5001 + 8192 * 3 + 1, 70000, 70002, soft.size]:
        parts.append(g2.push(soft[prev:c]).frames().reshape(-1))
        prev = c
    assert np.array_equal(np.concatenate(parts), want["cadu"])


def test_simple_psk_decoder_options(built):
    """rs_i = 0 (no RS), RS(255,239), derandomize off / after RS, rs_usecheck, a non-default ASM, cadu_size 10232 with I=5."""
    from satdump_b200 import capi
    O = oracle()
    rng = np.random.default_rng(9)
    for kw in [dict(rs_i=0, derandomize=False), dict(rs_i=4, rs_type=1), dict(rs_i=4, derand_after_rs=True), dict(rs_i=4, rs_usecheck=True),
               dict(rs_i=4, asm_sync=0xFAF3200D), dict(rs_i=5, cadu_size=10232)]:
        cadu_size = kw.pop("cadu_size", 8192)
        I = max(1, kw["rs_i"])
        pay = rng.integers(0, 256, size=(12, I * (239 if kw.get("rs_type") else 223)), dtype=np.uint8)
        # frames made by the reference's own encoder semantics are not needed here: any bit stream exercises the same code paths;
        # use noise-free random frames with a valid ASM so that the deframer locks, then compare everything with the oracle
        body = rng.integers(0, 256, size=(12, cadu_size // 8 - 4), dtype=np.uint8)
        asm = np.frombuffer(int(kw.get("asm_sync", 0x1ACFFC1D)).to_bytes(4, "big"), np.uint8)
        frames = np.concatenate([np.tile(asm, (12, 1)), body], axis=1)
        bits = np.unpackbits(frames.reshape(-1))
        soft = np.clip(np.round((bits * 2.0 - 1) * 60 + rng.normal(0, 15, bits.size)), -127, 127).astype(np.int8)
        f = O.Fec(O.simple_cfg("bpsk", cadu_size, **kw))
        n = soft.size // f.chunk * f.chunk
        want = f.run(soft[:n])
        g = capi.Fec(capi.simple_cfg("bpsk", cadu_size, max_soft=65536 * 4, **kw)).push(soft)
        assert np.array_equal(g.frames().reshape(-1), want["cadu"]), kw
        assert want["nframes"] >= 10, kw

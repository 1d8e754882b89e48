"""GPU: CADU -> CCSDS space packets (b200_demux_*, satdump_b200/csrc/demux.cuh) against the compiled reference demultiplexer
(ccsds_aos::Demuxer per virtual channel behind parseVCDU, module_metop_instruments.cpp:66-140): the packet bytes, their order, the frame
that returned each, channel, length and APID - bit for bit, on clean streams, on damaged ones (random first header pointers, wrong packet
lengths, frames of another channel, dropped frames), on the crafted corner where an unfinished packet's bytes stay in front of the next
one, with tiny data zones (many packets per frame, headers straddling most boundaries), and whatever the split into pushes."""
import numpy as np
import pytest

from satdump_b200 import capi, synth
from tests.common import gpu_chain, nsamples, oracle, signal

pytestmark = pytest.mark.gpu
ALL = (1 << 63) - 1


def same(got, want):
    gb, gr = got
    wb, wr = want
    assert gr.shape[0] == wr.shape[0], (gr.shape, wr.shape)
    assert np.array_equal(gr[:, :4], wr[:, :4].astype(np.int64))
    assert np.array_equal(gb, wb)
    off = np.concatenate([[0], np.cumsum(6 + wr[:, 2].astype(np.int64))])[:-1]
    assert np.array_equal(gr[:, 4], off)


@pytest.mark.parametrize("mpdu,iz,corrupt,drop,seed", [(884, 0, 0.0, 0.0, 1), (882, 2, 0.0, 0.0, 2), (884, 0, 0.05, 0.01, 3), (882, 2, 0.3, 0.05, 4),
                                                     (100, 0, 0.3, 0.02, 5), (60, 0, 0.5, 0.1, 6), (884, 0, 0.6, 0.1, 7), (40, 2, 0.4, 0.0, 8)])
def test_packets_equal_the_reference(built, mpdu, iz, corrupt, drop, seed):
    O = oracle()
    fr = synth.build_aos_frames(20000, seed=seed, mpdu=mpdu, insert_zone=iz, corrupt=corrupt, drop=drop)
    want = O.Demux(mpdu, iz).run(fr)
    g = capi.Demux(1024, mpdu, iz, max_frames=fr.shape[0], max_packets=fr.shape[0] * 30)
    same(g.run(fr), want)
    assert want[1].shape[0] > 300 and g.stats()["packets_out"] == want[1].shape[0]
    if corrupt == 0.0:
        assert g.stats()["redone_channels"] == 0


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("mpdu", [884, 200])
def test_leftover_bytes_of_an_unfinished_packet(built, mpdu, variant):
    O = oracle()
    fr = synth.craft_leftover_frames(mpdu, variant)
    want = O.Demux(mpdu, 0).run(fr)
    assert want[1][0, 2] > mpdu  # the first packet carries the unfinished one's bytes in front
    g1 = capi.Demux(1024, mpdu, 0, max_frames=64)
    same(g1.run(fr), want)
    assert g1.stats()["redone_channels"] == 1  # the leftover bytes cross a window boundary of the parallel walk: that channel is walked again serially
    g = capi.Demux(1024, mpdu, 0, max_frames=64)  # the same frame by frame: every carried state crosses a push
    parts = [g.run(fr[i:i + 1]) for i in range(fr.shape[0])]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), want[0])
    assert np.array_equal(np.concatenate([p[1][:, :4] for p in parts]), want[1][:, :4].astype(np.int64))


@pytest.mark.parametrize("mpdu,iz,corrupt", [(884, 0, 0.0), (882, 2, 0.2), (60, 0, 0.3)])
def test_any_split_into_pushes_gives_the_same_packets(built, mpdu, iz, corrupt):
    O = oracle()
    fr = synth.build_aos_frames(12000, seed=21, mpdu=mpdu, insert_zone=iz, corrupt=corrupt)
    wb, wr = O.Demux(mpdu, iz).run(fr)
    g = capi.Demux(1024, mpdu, iz, max_frames=12000, max_packets=12000 * 30)
    rng = np.random.default_rng(5)
    cuts = np.sort(np.concatenate([[0, fr.shape[0]], rng.choice(np.arange(1, fr.shape[0]), 9, replace=False), [1, 2, 3]]))
    cuts = np.unique(cuts)
    parts = [g.run(fr[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), wb)
    assert np.array_equal(np.concatenate([p[1][:, :4] for p in parts]), wr[:, :4].astype(np.int64))
    g.reset()  # a new stream after reset
    same(g.run(fr[:5000]), O.Demux(mpdu, iz).run(fr[:5000]))


def test_channel_mask_and_secondary_header_option(built):
    O = oracle()
    fr = synth.build_aos_frames(8000, seed=9, mpdu=884, corrupt=0.1)
    mask = (1 << 9) | (1 << 34)
    want = O.Demux(884, 0, vcid_mask=mask).run(fr)
    assert set(np.unique(want[1][:, 1]).tolist()) == {9, 34}
    same(capi.Demux(1024, 884, 0, vcid_mask=mask, max_frames=8000).run(fr), want)
    want = O.Demux(884, 0, secondary_header_extends=True).run(fr)
    same(capi.Demux(1024, 884, 0, secondary_header_extends=True, max_frames=8000, max_packets=8000 * 30).run(fr), want)


def test_cadus_of_the_chain_go_on_to_packets_on_the_device(built):
    """IQ -> CADUs (fused chain) -> space packets without the frames leaving the device (b200_chain_frames_device): whatever the frames hold
    (here: random payload, i.e. inconsistent M-PDUs), the packets are the reference demultiplexer's for the same CADUs."""
    O = oracle()
    cfg, raw, _ = signal("metop_ahrpt", 22, seed=6)
    n = nsamples(raw, cfg)
    ch = gpu_chain(cfg, n).push(raw)
    ptr, nb = ch.frames_device()
    nfr = nb // 1024
    assert nfr >= 20
    g = capi.Demux(1024, 882, 2, max_frames=nfr, max_packets=nfr * 140)
    got = g.run_device(ptr, nfr)
    cadus = gpu_chain(cfg, n).push(raw).frames()  # (frames_device hands the frames over: a second, identical run brings them to the host)
    assert cadus.shape[0] == nfr
    same(got, O.Demux(882, 2).run(cadus))


def test_errors_are_loud(built):
    with pytest.raises(capi.B200Error, match="does not fit"):
        capi.Demux(1024, 1100, 0)
    g = capi.Demux(1024, 884, 0, max_frames=16)
    with pytest.raises(capi.B200Error, match="max_frames"):
        g.run(np.zeros((17, 1024), np.uint8))
    tiny = capi.Demux(1024, 884, 0, max_frames=64, max_packets=8)  # 64 frames of 884-byte zones hold far more than 8 packets
    with pytest.raises(capi.B200Error, match="packet table"):
        tiny.run(synth.build_aos_frames(64, seed=1, mpdu=884, idle=0.0))

// CADU -> CCSDS space packets on the device (sm_100a): the step behind the decoder.
//
// Reference: one ccsds::ccsds_aos::Demuxer per virtual channel (src-core/common/ccsds/ccsds_aos/demuxer.cpp:64-199) behind parseVCDU's
// channel id (vcdu.cpp:10-17) and parseMPDU's first header pointer (mpdu.cpp:9-13), as the instrument modules use them
// (plugins/noaa_metop_support/metop/module_metop_instruments.cpp:66-140). Demuxer::work is a state machine over the frames of one
// channel; what a frame does splits into
//   (1) a part that depends on the state the previous frames left: finishing the header that straddled the frame boundary, continuing /
//       closing the packet under construction (cut at first_header_pointer + 1, :101) and pushing it;
//   (2) a part that depends on the frame alone: from the first header pointer on, the chain of packets inside the data zone, and what is
//       left open at its end (a packet under construction, or the first bytes of a header).
// (2) is done for all frames at once, one thread per frame (k_dmx_frames -> FrameSum). (1) is a walk over the FrameSums of one channel, O(1)
// per frame, never touching the frame data (k_dmx_walk): DMX_K warps per channel, each starting at a frame whose outgoing state does not
// depend on the past (up to a rare corner that is detected and redone by one warp). The packets themselves are then gathered in parallel:
// per-frame packet counts -> exclusive scan -> records in the reference's order (k_dmx_frames<EMIT> / k_dmx_place) -> scan of the sizes ->
// one warp per packet copies header + payload pieces (k_dmx_copy).
// Every behaviour of the reference on inconsistent frames is kept (bytes of an unfinished packet staying in front of the next one, the
// continuation that takes more than what remains and never completes, frames skipped whole): tests compare against the compiled
// reference on damaged streams.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200
{

struct DmxGeom
{
    int stride;   // bytes per CADU
    int data_off; // offset of the M-PDU data zone in a CADU: 12 (+ insert zone)
    int M;        // MPDU_DATA_SIZE
    int sec_ext;  // SECONDARY_HEADER_EXTENDS_PKT
};

// what a frame does on its own (part 2). 32 bytes.
struct FrameSum
{
    unsigned w0;        // fhp | vcid << 16 | flags << 24
    unsigned d6a, d6b;  // first 6 bytes of the data zone (d6b low 16 bits) | n_local << 16
    unsigned first_cpl; // payload length of the packet at fhp (FS_HDR_FITS)
    unsigned tail;      // tail_pos | tail_taken << 16
    unsigned tail_cpl;
    unsigned hb_a, hb_b; // partial header at the frame end: bytes 0..3, bytes 4..5 | count << 16 (1..6 bytes)
};
enum { FS_VALID = 1, FS_HAS_HDR = 2, FS_HDR_FITS = 4, FS_HAS_SECOND = 8, FS_TAIL_W = 16, FS_TAIL_IH = 32, FS_SELECTED = 64 };

// a payload piece: `len` bytes at data-zone offset `off` of frame `frame`, or of the channel's carry buffer (frame < 0); next = slot of
// the packet's following piece
struct DmxSeg { int frame, off, len, next; };
// a packet the walk produced (part 1), placed later at base[frame] + k
struct WalkPkt { int frame, k, paylen, seg_begin, nseg, hdr_frame, hdr_a, hdr_b; }; // header at (hdr_frame, hdr_a) or inline bytes (hdr_frame < 0)
// a packet in output order
struct PktRec { int frame, paylen, a, b, hdr_frame, hdr_a, hdr_b, vcid; }; // b < 0: single piece at (frame, a + 6); else b chained segments from slot a
// carried state of one channel between pushes (the Demuxer's members)
struct DmxCarry
{
    int working, in_header, ihb, cpl, tpl, rem, npay;
    unsigned hb_a, hb_b, hdr_a, hdr_b; // headerBuffer, currentCCSDSPacket.header.raw
};
struct DmxOut { long offset; int payload_len, frame; short vcid, apid; }; // = b200_packet

#ifdef B200_DEFINE_KERNELS
__device__ __forceinline__ int dmx_cpl(const uint8_t *h, int sec_ext) // readPacket, demuxer.cpp:26-33
{
    const int packet_length = h[4] << 8 | h[5];
    return packet_length + 1 + (sec_ext ? (((h[0] >> 3) & 1) ? 8 : 0) : 0);
}
__device__ __forceinline__ int dmx_cpl_words(unsigned a, unsigned b, int sec_ext) // same from the 6 bytes packed little-endian in (a, b)
{
    const int packet_length = ((b & 0xff) << 8) | ((b >> 8) & 0xff);
    return packet_length + 1 + (sec_ext ? (((a >> 3) & 1) ? 8 : 0) : 0);
}

// ---------------------------------------------------------------- part 2: one thread per frame
// EMIT == false: writes the FrameSum. EMIT == true: writes the records of the frame's own complete packets behind the walk's packets of the
// frame (base[f] + fxa[f] + fxb[f] + j), skipping the first when the walk produced it itself (fxb).
template <bool EMIT>
__global__ void __launch_bounds__(256) k_dmx_frames(const uint8_t *__restrict__ frames, long nframes, long frame0, DmxGeom G, unsigned long long vcid_mask,
                                                    FrameSum *__restrict__ sums, const long *__restrict__ base, const uint8_t *__restrict__ fxa,
                                                    const uint8_t *__restrict__ fxb, PktRec *__restrict__ recs, long cap_recs, int *__restrict__ flags)
{
    const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes)
        return;
    const uint8_t *cadu = frames + f * G.stride;
    const int vcid = cadu[5] & 63; // vcdu.cpp:14
    const uint8_t *mp = cadu + G.data_off - 2;
    const int fhp = (mp[0] & 7) << 8 | mp[1]; // mpdu.cpp:11
    const uint8_t *data = cadu + G.data_off;
    const int M = G.M;
    unsigned fl = 0;
    if ((vcid_mask >> vcid) & 1ull)
        fl |= FS_SELECTED;
    if (!(fhp < 2047 && fhp >= M)) // demuxer.cpp:71-74: such a frame is skipped whole
        fl |= FS_VALID;
    int n_local = 0, first_cpl = 0, tail_pos = 0, tail_taken = 0, tail_cpl = 0, ihb = 0;
    unsigned hb_a = 0, hb_b = 0;
    long slot = 0;
    int skip = 0;
    if (EMIT) {
        if (!(fl & FS_SELECTED) || !(fl & FS_VALID))
            return;
        skip = fxb[f];
        slot = base[f] + fxa[f] + skip;
    }
    auto emit = [&](int pos, int cpl) {
        if (EMIT) {
            if (n_local >= skip) {
                const long idx = slot + (n_local - skip);
                if (idx < cap_recs)
                    recs[idx] = PktRec{(int)(frame0 + f), cpl, pos, -1, (int)f, pos, 0, vcid};
                else
                    atomicOr(flags, 1);
            }
        }
        n_local++;
    };
    if ((fl & FS_VALID) && fhp < 2047) { // demuxer.cpp:121-195 with an empty packet under construction
        fl |= FS_HAS_HDR;
        if (fhp + 6 < M) {
            fl |= FS_HDR_FITS;
            first_cpl = dmx_cpl(data + fhp, G.sec_ext);
            int tpl = first_cpl + 6;
            if (M > fhp + tpl) {
                fl |= FS_HAS_SECOND;
                emit(fhp, first_cpl);
                int next = fhp + tpl;
                while (next < M) {
                    if (next + 6 < M) {
                        const int cpl = dmx_cpl(data + next, G.sec_ext);
                        tpl = cpl + 6;
                        const int room = M - (next + 6);
                        if (cpl <= room)
                            emit(next, cpl);
                        else {
                            fl |= FS_TAIL_W;
                            tail_pos = next;
                            tail_cpl = cpl;
                            tail_taken = room;
                        }
                    } else {
                        fl |= FS_TAIL_IH;
                        tail_pos = next;
                        ihb = M - next;
                        break;
                    }
                    next += tpl;
                }
            } else {
                fl |= FS_TAIL_W;
                tail_pos = fhp;
                tail_cpl = first_cpl;
                const int room = M - (fhp + 6);
                tail_taken = first_cpl > room ? room : first_cpl;
            }
        } else if (fhp < M) {
            fl |= FS_TAIL_IH;
            tail_pos = fhp;
            ihb = M - fhp;
        }
    }
    if (EMIT)
        return;
    if (fl & FS_TAIL_IH) {
        for (int i = 0; i < ihb && i < 4; i++)
            hb_a |= (unsigned)data[tail_pos + i] << (8 * i);
        for (int i = 4; i < ihb && i < 6; i++)
            hb_b |= (unsigned)data[tail_pos + i] << (8 * (i - 4));
        hb_b |= (unsigned)ihb << 16;
    }
    FrameSum s;
    s.w0 = (unsigned)fhp | (unsigned)vcid << 16 | fl << 24;
    s.d6a = (unsigned)data[0] | (unsigned)data[1] << 8 | (unsigned)data[2] << 16 | (unsigned)data[3] << 24;
    s.d6b = (unsigned)data[4] | (unsigned)data[5] << 8 | (unsigned)n_local << 16;
    s.first_cpl = (unsigned)first_cpl;
    s.tail = (unsigned)tail_pos | (unsigned)tail_taken << 16;
    s.tail_cpl = (unsigned)tail_cpl;
    s.hb_a = hb_a;
    s.hb_b = hb_b;
    sums[f] = s;
}

// ---------------------------------------------------------------- part 1: DMX_K warps per virtual channel
// The walk over one channel's summaries is serial, but a frame whose first header fits its data zone (an ANCHOR: FS_HAS_HDR && FS_HDR_FITS)
// leaves a state that does not depend on what came before - except for leftover bytes of a packet that never completed, the rare corner.
// So the batch is cut into DMX_K windows; warp k of a channel starts at the first anchor g0 inside its window, doing only that frame's header
// part with an empty packet under construction (a guess), walks on past the window's end to the next anchor g1 and does only that frame's
// first part (finishing the straddling header, closing and pushing the continued packet) - exactly where the next active warp began.
// If bytes are still under construction there, the guess of the next warp was wrong: the channel is flagged and walked again by one warp
// from the carried state (redo pass; it rewrites everything the guessers wrote for that channel). Warp 0 always starts at frame 0 from the
// carried state. All lanes of a warp run the state machine on the same values (the fields of the frame in turn come by shuffle from the
// lane that loaded it); lane 0 writes.
// Storage needs no counting: a frame owns segment slots 3f .. 3f+2 (slot 3f for its first part, the others for its header part; a packet's
// pieces are chained by `next`), walk-packet slots 2f, 2f+1, and the bytes fxa[f] (packets its first part pushed: 0/1) and fxb[f] (its
// first own packet was emitted by the walk with leftover bytes in front: 0/1). The carry segment of channel v sits in slot 3 * nframes + v.
constexpr int DMX_K = 16;
__global__ void __launch_bounds__(32) k_dmx_walk(const FrameSum *__restrict__ sums, long nframes, long frame0, DmxGeom G, int redo,
                                                 unsigned long long *__restrict__ redo_mask, const uint8_t *__restrict__ frames,
                                                 const DmxCarry *__restrict__ carry_in, DmxCarry *__restrict__ carry_out, uint8_t *__restrict__ fxa,
                                                 uint8_t *__restrict__ fxb, WalkPkt *__restrict__ wp, DmxSeg *__restrict__ segs,
                                                 int *__restrict__ tailseg /* [64][2]: first slot, count of the open packet's segments */)
{
    const int v = blockIdx.y, k = blockIdx.x, lane = threadIdx.x;
    if (redo && (k != 0 || !((*redo_mask >> v) & 1ull)))
        return;
    const int M = G.M;
    const long n = nframes;
    // first anchor of this channel at index >= x (n if none)
    auto first_anchor = [&](long x) -> long {
        for (long c = x & ~31L; c < n; c += 32) {
            const long fi = c + lane;
            unsigned w0 = 0;
            if (fi < n && fi >= x)
                w0 = sums[fi].w0;
            const unsigned fl = w0 >> 24;
            const bool hit = ((w0 >> 16) & 63) == (unsigned)v && (fl & FS_SELECTED) && (fl & FS_VALID) && (fl & FS_HAS_HDR) && (fl & FS_HDR_FITS);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m)
                return c + (__ffs(m) - 1);
        }
        return n;
    };
    const long w_lo = redo ? 0 : (long)k * n / DMX_K, w_hi = redo ? n : (long)(k + 1) * n / DMX_K;
    const bool from_carry = redo || k == 0;
    long g0 = 0;
    if (!from_carry) {
        g0 = first_anchor(w_lo);
        if (g0 >= w_hi)
            return; // no anchor in this window: the warp that is walking through it goes on
    }
    const long g1 = w_hi >= n ? n : first_anchor(w_hi);
    int W = 0, IH = 0, IHB = 0, cpl = 0, tpl = 0, rem = 0, npay = 0;
    unsigned hb_a = 0, hb_b = 0, hdr_a = 0, hdr_b = 0;
    int hdr_frame = -1;
    int head = -1, tail = -1, nseg = 0; // the packet under construction: chain of segment slots
    if (from_carry) {
        const DmxCarry st = carry_in[v];
        W = st.working; IH = st.in_header; IHB = st.ihb; cpl = st.cpl; tpl = st.tpl; rem = st.rem; npay = st.npay;
        hb_a = st.hb_a; hb_b = st.hb_b; hdr_a = st.hdr_a; hdr_b = st.hdr_b;
        if (npay > 0) { // its bytes so far come from the carry buffer
            const int slot = (int)(3 * n) + v;
            if (lane == 0)
                segs[slot] = DmxSeg{-1, 0, npay, -1};
            head = tail = slot;
            nseg = 1;
        }
    }
    for (long c00 = g0 & ~127L; c00 < n && c00 <= g1; c00 += 128) {
      FrameSum s4[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
          const long fq = c00 + 32 * q + lane;
          s4[q].w0 = 0;
          if (fq < n && fq >= g0 && fq <= g1)
              s4[q] = sums[fq];
      }
#pragma unroll 1
      for (int q = 0; q < 4; q++) {
        const long c0 = c00 + 32 * q;
        if (c0 >= n || c0 > g1)
            break;
        const long fi = c0 + lane;
        FrameSum s = s4[0];
        if (q == 1) s = s4[1];
        if (q == 2) s = s4[2];
        if (q == 3) s = s4[3];
        const unsigned myfl = s.w0 >> 24;
        const bool mine = fi < n && fi >= g0 && fi <= g1 && ((s.w0 >> 16) & 63) == (unsigned)v && (myfl & FS_SELECTED) && (myfl & FS_VALID);
        unsigned mask = __ballot_sync(0xffffffffu, mine);
        while (mask) {
            const int l = __ffs(mask) - 1;
            mask &= mask - 1;
            const unsigned w0 = __shfl_sync(0xffffffffu, s.w0, l), d6a = __shfl_sync(0xffffffffu, s.d6a, l), d6b = __shfl_sync(0xffffffffu, s.d6b, l);
            const unsigned first_cpl = __shfl_sync(0xffffffffu, s.first_cpl, l), tl = __shfl_sync(0xffffffffu, s.tail, l);
            const unsigned tail_cpl = __shfl_sync(0xffffffffu, s.tail_cpl, l), t_hb_a = __shfl_sync(0xffffffffu, s.hb_a, l), t_hb_b = __shfl_sync(0xffffffffu, s.hb_b, l);
            const int f = (int)(c0 + l);
            const int fhp = w0 & 0xffff;
            const unsigned fl = w0 >> 24;
            const bool do_first = from_carry || f != g0;  // a guessing warp enters its anchor frame behind the first part
            const bool do_hdr = !(f == g1 && g1 < n);     // and every warp leaves the next anchor frame in front of its header part
            int pre = 0, skip = 0;
            int used = do_first ? 0 : 1; // segment slots of this frame taken so far (slot 0 belongs to the first part)
            auto addseg = [&](int off, int len) {
                if (len > 0) {
                    const int slot = 3 * f + used;
                    used++;
                    if (lane == 0) {
                        segs[slot] = DmxSeg{f, off, len, -1};
                        if (tail >= 0)
                            segs[tail].next = slot;
                    }
                    if (tail < 0)
                        head = slot;
                    tail = slot;
                    nseg++;
                    npay += len;
                }
            };
            auto push_packet = [&]() { // pushPacket, demuxer.cpp:36-44
                if (lane == 0)
                    wp[2 * f + pre] = WalkPkt{(int)(frame0 + f), pre, npay, head, nseg, hdr_frame, (int)hdr_a, (int)hdr_b};
                pre++;
                head = tail = -1;
                nseg = 0;
                npay = 0;
                W = 0;
                cpl = 0;
                rem = 0;
            };
            if (do_first) {
                int offset = 0;
                if (IH) { // :81-92: the header that straddled the frame boundary
                    IH = 0;
                    const unsigned long long have = (unsigned long long)hb_a | (unsigned long long)(hb_b & 0xffff) << 32;
                    const unsigned long long add = (unsigned long long)d6a | (unsigned long long)(d6b & 0xffff) << 32;
                    const unsigned long long keep = IHB >= 6 ? 0xffffffffffffull : ((1ull << (8 * IHB)) - 1);
                    const unsigned long long full = (have & keep) | (add << (8 * IHB));
                    offset = 6 - IHB;
                    IHB = 6;
                    hb_a = (unsigned)full;
                    hb_b = (unsigned)(full >> 32) & 0xffff;
                    hdr_frame = -1;
                    hdr_a = hb_a;
                    hdr_b = hb_b;
                    cpl = dmx_cpl_words(hdr_a, hdr_b, G.sec_ext);
                    tpl = cpl + 6;
                    rem = cpl;
                    W = 1;
                }
                if (rem > 0 && W) { // :95-112
                    if (fl & FS_HAS_HDR) {
                        const int m = (rem + offset) > fhp + 1 ? (fhp + 1) - offset : rem;
                        addseg(offset, m);
                        rem = 0;
                    } else {
                        const int m = (rem + offset) > M - offset ? M - offset : rem;
                        addseg(offset, m);
                        rem -= m;
                    }
                }
                if (rem == 0 && W) // :115-118
                    push_packet();
                if (lane == 0)
                    fxa[f] = (uint8_t)pre;
                if (!do_hdr) {
                    // the next warp started this frame's header part with nothing under construction: true unless bytes are left here
                    if (head >= 0 && lane == 0)
                        atomicOr(redo_mask, 1ull << v);
                    continue;
                }
            }
            used = max(used, 1);
            if (fl & FS_HAS_HDR) {
                if (fl & FS_HDR_FITS) {
                    // readPacket(&mpdu.data[fhp]) keeps whatever the packet under construction already holds (only reachable with rem < 0)
                    hdr_frame = f;
                    hdr_a = (unsigned)fhp;
                    hdr_b = 0;
                    cpl = (int)first_cpl;
                    tpl = cpl + 6;
                    rem = cpl;
                    W = 1;
                    if (fl & FS_HAS_SECOND) {
                        if (head >= 0) { // leftover bytes in front: the walk emits the frame's first own packet itself
                            addseg(fhp + 6, cpl);
                            rem = 0;
                            push_packet();
                            pre--; // it is one of the frame's own packets: fxa counts only the first part's
                            skip = 1;
                        } else {
                            W = 0;
                            cpl = 0;
                            rem = 0;
                        }
                        // the rest of the chain is the frame's own; what it leaves open:
                        if (fl & FS_TAIL_W) {
                            hdr_frame = f;
                            hdr_a = tl & 0xffff;
                            hdr_b = 0;
                            cpl = (int)tail_cpl;
                            tpl = cpl + 6;
                            rem = cpl;
                            W = 1;
                            const int taken = tl >> 16;
                            addseg((int)(tl & 0xffff) + 6, taken);
                            rem -= taken;
                        } else if (fl & FS_TAIL_IH) {
                            IH = 1;
                            IHB = (t_hb_b >> 16) & 0xff;
                            hb_a = t_hb_a;
                            hb_b = t_hb_b & 0xffff;
                        }
                    } else { // :179-186 (workingOnPacket is true: readPacket just set it)
                        const int taken = tl >> 16;
                        addseg(fhp + 6, taken);
                        rem -= taken;
                    }
                } else if (fl & FS_TAIL_IH) { // :188-194
                    IH = 1;
                    IHB = (t_hb_b >> 16) & 0xff;
                    hb_a = t_hb_a;
                    hb_b = t_hb_b & 0xffff;
                }
            }
            if (lane == 0)
                fxb[f] = (uint8_t)skip;
        }
      }
    }
    if (g1 >= n && lane == 0) { // this warp reached the end of the batch: the channel's state for the next push
        // the header of an open packet must outlive the frame buffer: make it inline
        if (hdr_frame >= 0) {
            const uint8_t *h = frames + (long)hdr_frame * G.stride + G.data_off + hdr_a;
            hdr_a = (unsigned)h[0] | (unsigned)h[1] << 8 | (unsigned)h[2] << 16 | (unsigned)h[3] << 24;
            hdr_b = (unsigned)h[4] | (unsigned)h[5] << 8;
        }
        DmxCarry o;
        o.working = W; o.in_header = IH; o.ihb = IHB; o.cpl = cpl; o.tpl = tpl; o.rem = rem; o.npay = npay;
        o.hb_a = hb_a; o.hb_b = hb_b; o.hdr_a = hdr_a; o.hdr_b = hdr_b;
        carry_out[v] = o;
        tailseg[2 * v] = head;
        tailseg[2 * v + 1] = nseg;
    }
}

// packets returned by every frame's work() call: what its first part pushed + its own complete packets
__global__ void k_dmx_cnt(const FrameSum *__restrict__ sums, const uint8_t *__restrict__ fxa, long nframes, int *__restrict__ cnt)
{
    const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes)
        return;
    const FrameSum s = sums[f];
    const unsigned fl = s.w0 >> 24;
    cnt[f] = ((fl & FS_SELECTED) && (fl & FS_VALID)) ? (int)fxa[f] + (int)(s.d6b >> 16) : 0;
}

// ---------------------------------------------------------------- exclusive scan of n ints -> longs (one CTA; n up to a few million)
__global__ void __launch_bounds__(1024) k_dmx_scan(const int *__restrict__ in, long n_host, const long *__restrict__ n_dev, long *__restrict__ out /* n + 1 */)
{
    __shared__ long wsum[32];
    __shared__ long run;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long n = n_dev ? *n_dev : n_host;
    if (t == 0)
        run = 0;
    __syncthreads();
    for (long b = 0; b < n; b += 4096) {
        const long i0 = b + 4L * t;
        long v[4], s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            v[k] = i0 + k < n ? (long)in[i0 + k] : 0;
            s += v[k];
        }
        long incl = s;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const long p = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off)
                incl += p;
        }
        if (lane == 31)
            wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long w = wsum[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const long p = __shfl_up_sync(0xffffffffu, w, off);
                if (lane >= off)
                    w += p;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        long e = run + (warp ? wsum[warp - 1] : 0) + incl - s;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < n)
                out[i0 + k] = e;
            e += v[k];
        }
        __syncthreads();
        if (t == 1023)
            run = e;
        __syncthreads();
    }
    if (t == 0)
        out[n] = run;
}

// places the walk's packets (slots 2f, 2f + 1 of frame f: fxa[f] + fxb[f] of them) at base[f] + slot. One thread per frame.
__global__ void k_dmx_place(const WalkPkt *__restrict__ wp, const FrameSum *__restrict__ sums, const uint8_t *__restrict__ fxa, const uint8_t *__restrict__ fxb,
                            long nframes, const long *__restrict__ base, PktRec *__restrict__ recs, long cap_recs, int *__restrict__ flags)
{
    const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes)
        return;
    const unsigned w0 = sums[f].w0, fl = w0 >> 24;
    if (!(fl & FS_SELECTED) || !(fl & FS_VALID))
        return;
    const int m = (int)fxa[f] + (int)fxb[f];
    for (int j = 0; j < m; j++) {
        const WalkPkt p = wp[2 * f + j];
        const long idx = base[f] + j;
        if (idx < cap_recs)
            recs[idx] = PktRec{p.frame, p.paylen, p.seg_begin, p.nseg, p.hdr_frame, p.hdr_a, p.hdr_b, (int)((w0 >> 16) & 63)};
        else
            atomicOr(flags, 1);
    }
}

__global__ void k_dmx_sizes(const PktRec *__restrict__ recs, const long *__restrict__ npk, int *__restrict__ sizes)
{
    const long n = *npk;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        sizes[i] = 6 + recs[i].paylen;
}

// ---------------------------------------------------------------- one warp per packet: header + payload pieces -> byte stream, and the host record
__global__ void __launch_bounds__(256) k_dmx_copy(const PktRec *__restrict__ recs, const long *__restrict__ npk, const long *__restrict__ offs,
                                                  const DmxSeg *__restrict__ segs, const uint8_t *__restrict__ frames, DmxGeom G,
                                                  const uint8_t *__restrict__ carry_bytes /* [64][cap] */, int carry_cap, uint8_t *__restrict__ out, long cap_bytes,
                                                  DmxOut *__restrict__ orec, int *__restrict__ flags)
{
    const long n = *npk;
    const int lane = threadIdx.x & 31;
    for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += ((long)gridDim.x * blockDim.x) >> 5) {
        const PktRec r = recs[i];
        const long o = offs[i];
        if (o + 6 + r.paylen > cap_bytes) {
            if (lane == 0)
                atomicOr(flags, 2);
            continue;
        }
        unsigned long long hv; // the 6 header bytes, first byte lowest
        if (r.hdr_frame >= 0) {
            const uint8_t *hp = frames + (long)r.hdr_frame * G.stride + G.data_off + r.hdr_a;
            hv = 0;
#pragma unroll
            for (int k = 0; k < 6; k++)
                hv |= (unsigned long long)hp[k] << (8 * k);
        } else
            hv = (unsigned long long)(unsigned)r.hdr_a | (unsigned long long)((unsigned)r.hdr_b & 0xffff) << 32;
        if (lane < 6)
            out[o + lane] = (uint8_t)(hv >> (8 * lane));
        if (lane == 0)
            orec[i] = DmxOut{o, r.paylen, r.frame, (short)r.vcid, (short)(((unsigned)hv & 7) << 8 | ((unsigned)(hv >> 8) & 0xff))};
        uint8_t *dst = out + o + 6;
        if (r.b < 0) {
            const uint8_t *src = frames + (long)r.hdr_frame * G.stride + G.data_off + r.a + 6;
            for (int k = lane; k < r.paylen; k += 32)
                dst[k] = src[k];
        } else {
            long done = 0;
            int si = r.a;
            for (int sidx = 0; sidx < r.b && si >= 0; sidx++) {
                const DmxSeg sg = segs[si];
                const uint8_t *src = sg.frame >= 0 ? frames + (long)sg.frame * G.stride + G.data_off + sg.off : carry_bytes + (long)r.vcid * carry_cap + sg.off;
                for (int k = lane; k < sg.len; k += 32)
                    dst[done + k] = src[k];
                done += sg.len;
                si = sg.next;
            }
        }
    }
}

// the bytes of every channel's open packet -> its next carry buffer (one CTA per channel)
__global__ void __launch_bounds__(256) k_dmx_carry(const int *__restrict__ tailseg, const DmxSeg *__restrict__ segs, const uint8_t *__restrict__ frames, DmxGeom G,
                                                   const uint8_t *__restrict__ carry_in, uint8_t *__restrict__ carry_out, int carry_cap, int *__restrict__ flags)
{
    const int v = blockIdx.x;
    int si = tailseg[2 * v];
    const int count = tailseg[2 * v + 1];
    long done = 0;
    for (int sidx = 0; sidx < count && si >= 0; sidx++) {
        const DmxSeg sg = segs[si];
        si = sg.next;
        if (done + sg.len > carry_cap) {
            if (threadIdx.x == 0)
                atomicOr(flags, 4);
            return;
        }
        const uint8_t *src = sg.frame >= 0 ? frames + (long)sg.frame * G.stride + G.data_off + sg.off : carry_in + (long)v * carry_cap + sg.off;
        for (int k = threadIdx.x; k < sg.len; k += blockDim.x)
            carry_out[(long)v * carry_cap + done + k] = src[k];
        done += sg.len;
    }
}
#endif // B200_DEFINE_KERNELS

} // namespace b200

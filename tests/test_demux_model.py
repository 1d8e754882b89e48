"""CPU: the decomposition behind the device demultiplexer (satdump_b200/csrc/demux.cuh), modelled in Python and checked against the compiled
reference: (2) what a frame does on its own from its first header pointer on (k_dmx_frames -> FrameSum), (1) the per-channel walk over those
summaries that finishes straddling headers, continues / cuts / pushes the packet under construction and handles the leftover-bytes corner
(k_dmx_walk; first as one serial walk, then as the kernel runs it: several warps per channel that start at "anchor" frames with a guess,
and a serial redo when a guess was wrong), and the emission order (walk packets first, then the frame's own, minus the one the walk took
over). The CUDA kernels follow this model statement by statement; tests/test_gpu_demux.py checks them on the device."""
import numpy as np
import pytest

from satdump_b200 import synth


def cplh(h, sec_ext):
    pl = (int(h[4]) << 8) | int(h[5])
    return pl + 1 + ((8 if (int(h[0]) >> 3) & 1 else 0) if sec_ext else 0)

def frame_sum(data, fhp, M, sec_ext):
    s = dict(valid=not (fhp < 2047 and fhp >= M), has_hdr=False, hdr_fits=False, has_second=False, tail_w=False, tail_ih=False,
             n_local=0, first_cpl=0, tail_pos=0, tail_taken=0, tail_cpl=0, ihb=0, hb=b'', d6=bytes(data[:6]), local=[])
    if s['valid'] and fhp < 2047:
        s['has_hdr'] = True
        if fhp + 6 < M:
            s['hdr_fits'] = True
            s['first_cpl'] = cplh(data[fhp:fhp + 6], sec_ext)
            tpl = s['first_cpl'] + 6
            if M > fhp + tpl:
                s['has_second'] = True
                s['local'].append((fhp, s['first_cpl']))
                nxt = fhp + tpl
                while nxt < M:
                    if nxt + 6 < M:
                        c = cplh(data[nxt:nxt + 6], sec_ext)
                        tpl = c + 6
                        room = M - (nxt + 6)
                        if c <= room:
                            s['local'].append((nxt, c))
                        else:
                            s['tail_w'] = True; s['tail_pos'] = nxt; s['tail_cpl'] = c; s['tail_taken'] = room
                    else:
                        s['tail_ih'] = True; s['tail_pos'] = nxt; s['ihb'] = M - nxt
                        break
                    nxt += tpl
            else:
                s['tail_w'] = True; s['tail_pos'] = fhp; s['tail_cpl'] = s['first_cpl']
                room = M - (fhp + 6)
                s['tail_taken'] = min(s['first_cpl'], room)
        elif fhp < M:
            s['tail_ih'] = True; s['tail_pos'] = fhp; s['ihb'] = M - fhp
    if s['tail_ih']:
        s['hb'] = bytes(data[s['tail_pos']:s['tail_pos'] + s['ihb']])
    s['n_local'] = len(s['local'])
    return s

def run(frames, M, insert, sec_ext, mask):
    base = 10 + insert
    out_by_frame = {}
    st = {}
    for f in range(frames.shape[0]):
        cadu = frames[f]
        v = int(cadu[5]) & 63
        if not (mask >> v) & 1: continue
        fhp = ((int(cadu[base]) & 7) << 8) | int(cadu[base + 1])
        data = cadu[base + 2:]
        s = frame_sum(data, fhp, M, sec_ext)
        if not s['valid']: continue
        S = st.setdefault(v, dict(W=0, IH=0, IHB=0, cpl=0, tpl=0, rem=0, hb=bytearray(6), hdr=None, segs=[]))
        pre = []; skip = 0
        def push():
            pay = b''.join(bytes(frames[ff][base + 2 + o: base + 2 + o + l]) for ff, o, l in S['segs'])
            pre.append(bytes(S['hdr']) + pay)
            S['segs'] = []; S['W'] = 0; S['cpl'] = 0; S['rem'] = 0
        def addseg(ff, o, l):
            if l > 0: S['segs'].append((ff, o, l))
        offset = 0
        if S['IH']:
            S['IH'] = 0
            n = 6 - S['IHB']
            S['hb'][S['IHB']:6] = s['d6'][:n]
            offset = n; S['IHB'] = 6
            S['hdr'] = bytes(S['hb']); S['cpl'] = cplh(S['hb'], sec_ext); S['tpl'] = S['cpl'] + 6; S['rem'] = S['cpl']; S['W'] = 1
        if S['rem'] > 0 and S['W']:
            if s['has_hdr']:
                n = (fhp + 1) - offset if (S['rem'] + offset) > fhp + 1 else S['rem']
                addseg(f, offset, n); S['rem'] = 0
            else:
                n = M - offset if (S['rem'] + offset) > M - offset else S['rem']
                addseg(f, offset, n); S['rem'] -= n
        if S['rem'] == 0 and S['W']: push()
        npre = len(pre)
        if s['has_hdr']:
            if s['hdr_fits']:
                S['hdr'] = bytes(data[fhp:fhp + 6]); S['cpl'] = s['first_cpl']; S['tpl'] = S['cpl'] + 6; S['rem'] = S['cpl']; S['W'] = 1
                if s['has_second']:
                    if S['segs']:
                        addseg(f, fhp + 6, S['cpl']); S['rem'] = 0; push(); skip = 1
                    else:
                        S['W'] = 0; S['cpl'] = 0; S['rem'] = 0
                    if s['tail_w']:
                        tp = s['tail_pos']
                        S['hdr'] = bytes(data[tp:tp + 6]); S['cpl'] = s['tail_cpl']; S['tpl'] = S['cpl'] + 6; S['rem'] = S['cpl']; S['W'] = 1
                        addseg(f, tp + 6, s['tail_taken']); S['rem'] -= s['tail_taken']
                    elif s['tail_ih']:
                        S['IH'] = 1; S['IHB'] = s['ihb']; S['hb'][:s['ihb']] = s['hb']
                else:
                    addseg(f, fhp + 6, s['tail_taken']); S['rem'] -= s['tail_taken']
            elif s['tail_ih']:
                S['IH'] = 1; S['IHB'] = s['ihb']; S['hb'][:s['ihb']] = s['hb']
        loc = [bytes(data[p:p + 6 + c]) for p, c in s['local']][skip:]
        out_by_frame[f] = (v, pre + loc)
    outs = []; recs = []
    for f in sorted(out_by_frame):
        v, pk = out_by_frame[f]
        for p in pk:
            outs.append(p); recs.append((f, v, len(p) - 6))
    return b''.join(outs), recs



# ---- the windowed walk (k_dmx_walk): DMX_K warps per channel, each starting at an anchor frame with a guess, redo on a wrong guess
def walk_channel(frames, sums, own, M, base, sec_ext, K):
    """own: sorted list of global frame indices of this channel (valid). Returns dict frame -> (pre list, skip, own_emitted flag) or None on conflict."""
    n = frames.shape[0]
    anchors = [f for f in own if sums[f]['has_hdr'] and sums[f]['hdr_fits']]
    def first_anchor_ge(x):
        for a in anchors:
            if a >= x: return a
        return n
    out_pre = {}; out_skip = {}
    conflict = False
    for k in range(K):
        w0, w1 = k * n // K, (k + 1) * n // K
        g0 = 0 if k == 0 else first_anchor_ge(w0)
        if k > 0 and g0 >= w1: continue  # inactive
        g1 = first_anchor_ge(w1) if w1 < n else n
        S = dict(W=0, IH=0, IHB=0, cpl=0, tpl=0, rem=0, hb=bytearray(6), hdr=None, segs=[])
        for f in [x for x in own if g0 <= x <= min(g1, n - 1)]:
            s = sums[f]; data = frames[f][base + 2:]
            fhp = s['fhp']
            do_first = not (k > 0 and f == g0)
            do_hdr = not (f == g1 and g1 < n)
            if f == g1 and g1 < n and f == g0 and k > 0:
                raise AssertionError("empty range should be inactive")
            pre = []
            def push():
                pay = b''.join(bytes(frames[ff][base + 2 + o: base + 2 + o + l]) for ff, o, l in S['segs'])
                pre.append(bytes(S['hdr']) + pay)
                S['segs'] = []; S['W'] = 0; S['cpl'] = 0; S['rem'] = 0
            def addseg(ff, o, l):
                if l > 0: S['segs'].append((ff, o, l))
            if do_first:
                offset = 0
                if S['IH']:
                    S['IH'] = 0
                    nn = 6 - S['IHB']
                    S['hb'][S['IHB']:6] = s['d6'][:nn]
                    offset = nn; S['IHB'] = 6
                    S['hdr'] = bytes(S['hb']); S['cpl'] = cplh(S['hb'], sec_ext); S['tpl'] = S['cpl'] + 6; S['rem'] = S['cpl']; S['W'] = 1
                if S['rem'] > 0 and S['W']:
                    if s['has_hdr']:
                        m = (fhp + 1) - offset if (S['rem'] + offset) > fhp + 1 else S['rem']
                        addseg(f, offset, m); S['rem'] = 0
                    else:
                        m = M - offset if (S['rem'] + offset) > M - offset else S['rem']
                        addseg(f, offset, m); S['rem'] -= m
                if S['rem'] == 0 and S['W']: push()
                out_pre[f] = list(pre)
                if not do_hdr:
                    if S['segs']:
                        conflict = True
                    continue
            skip = 0
            pre2 = []
            pre = pre2
            if s['has_hdr']:
                if s['hdr_fits']:
                    S['hdr'] = bytes(data[fhp:fhp + 6]); S['cpl'] = s['first_cpl']; S['tpl'] = S['cpl'] + 6; S['rem'] = S['cpl']; S['W'] = 1
                    if s['has_second']:
                        if S['segs']:
                            addseg(f, fhp + 6, S['cpl']); S['rem'] = 0
                            pay = b''.join(bytes(frames[ff][base + 2 + o: base + 2 + o + l]) for ff, o, l in S['segs'])
                            pre2.append(bytes(S['hdr']) + pay); S['segs'] = []; S['W'] = 0; S['cpl'] = 0; S['rem'] = 0
                            skip = 1
                        else:
                            S['W'] = 0; S['cpl'] = 0; S['rem'] = 0
                        if s['tail_w']:
                            tp = s['tail_pos']
                            S['hdr'] = bytes(data[tp:tp + 6]); S['cpl'] = s['tail_cpl']; S['tpl'] = S['cpl'] + 6; S['rem'] = S['cpl']; S['W'] = 1
                            addseg(f, tp + 6, s['tail_taken']); S['rem'] -= s['tail_taken']
                        elif s['tail_ih']:
                            S['IH'] = 1; S['IHB'] = s['ihb']; S['hb'][:s['ihb']] = s['hb']
                    else:
                        addseg(f, fhp + 6, s['tail_taken']); S['rem'] -= s['tail_taken']
                elif s['tail_ih']:
                    S['IH'] = 1; S['IHB'] = s['ihb']; S['hb'][:s['ihb']] = s['hb']
            out_skip[f] = (skip, pre2)
    return None if conflict else (out_pre, out_skip)

def run_windowed(frames, M, insert, sec_ext, K):
    base = 10 + insert
    n = frames.shape[0]
    sums = {}; chans = {}
    for f in range(n):
        cadu = frames[f]
        v = int(cadu[5]) & 63
        fhp = ((int(cadu[base]) & 7) << 8) | int(cadu[base + 1])
        s = frame_sum(cadu[base + 2:], fhp, M, sec_ext); s['fhp'] = fhp
        sums[f] = s
        if s['valid'] and v != 63: chans.setdefault(v, []).append(f)
    per_frame = {}
    nconf = 0
    for v, own in chans.items():
        r = walk_channel(frames, sums, own, M, base, sec_ext, K)
        if r is None:
            nconf += 1
            r = walk_channel(frames, sums, own, M, base, sec_ext, 1)
        out_pre, out_skip = r
        for f in own:
            skip, pre2 = out_skip.get(f, (0, []))
            data = frames[f][base + 2:]
            loc = [bytes(data[p:p + 6 + c]) for p, c in sums[f]['local']][skip:]
            per_frame[f] = (v, out_pre.get(f, []) + pre2 + loc)
    outs = []; recs = []
    for f in sorted(per_frame):
        v, pk = per_frame[f]
        for p in pk:
            outs.append(p); recs.append((f, v, len(p) - 6))
    return b''.join(outs), recs, nconf



def _ref():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return ref


@pytest.mark.parametrize("mpdu,iz,corrupt,drop,seed", [(884, 0, 0.0, 0.0, 1), (882, 2, 0.3, 0.05, 4), (60, 0, 0.5, 0.1, 6)])
def test_decomposition_reproduces_the_reference_demuxer(built, mpdu, iz, corrupt, drop, seed):
    ref = _ref()
    fr = synth.build_aos_frames(2500, seed=seed, mpdu=mpdu, insert_zone=iz, corrupt=corrupt, drop=drop)
    ba, ra = ref.Demux(mpdu, iz).run(fr)
    bb, rb = run(fr, mpdu, iz, False, (1 << 63) - 1)
    assert ra.shape[0] > 50 and bytes(ba) == bb and [tuple(r[:3]) for r in ra.tolist()] == rb


@pytest.mark.parametrize("variant", [0, 1])
def test_decomposition_on_the_leftover_corner(built, variant):
    ref = _ref()
    fr = synth.craft_leftover_frames(200, variant)
    ba, ra = ref.Demux(200, 0).run(fr)
    bb, rb = run(fr, 200, 0, False, (1 << 63) - 1)
    assert bytes(ba) == bb and [tuple(r[:3]) for r in ra.tolist()] == rb


@pytest.mark.parametrize("mpdu,iz,corrupt,drop,seed", [(882, 2, 0.3, 0.05, 4), (60, 0, 0.5, 0.1, 6)])
def test_windowed_walk_with_guesses_reproduces_the_reference(built, mpdu, iz, corrupt, drop, seed):
    ref = _ref()
    fr = synth.build_aos_frames(2500, seed=seed, mpdu=mpdu, insert_zone=iz, corrupt=corrupt, drop=drop)
    ba, ra = ref.Demux(mpdu, iz).run(fr)
    for K in (4, 16):
        bb, rb, _ = run_windowed(fr, mpdu, iz, False, K)
        assert bytes(ba) == bb and [tuple(r[:3]) for r in ra.tolist()] == rb
    for n in (1, 2, 3, 15, 16, 17, 33):  # batches shorter than / around the window count
        ba, ra = ref.Demux(mpdu, iz).run(fr[100:100 + n])
        bb, rb, _ = run_windowed(fr[100:100 + n], mpdu, iz, False, 16)
        assert bytes(ba) == bb and [tuple(r[:3]) for r in ra.tolist()] == rb


@pytest.mark.parametrize("variant", [0, 1])
def test_wrong_guess_at_a_window_boundary_is_redone(built, variant):
    ref = _ref()
    q = synth.craft_leftover_frames(200, variant)
    pad = synth.build_aos_frames(40, seed=3, mpdu=200, vcids=(9,), idle=0.0)
    hit = 0
    for shift in range(8):
        fr = np.concatenate([pad[:shift + 10], q, pad[20:]])
        ba, ra = ref.Demux(200, 0).run(fr)
        bb, rb, nconf = run_windowed(fr, 200, 0, False, 8)
        hit += nconf
        assert bytes(ba) == bb and [tuple(r[:3]) for r in ra.tolist()] == rb
    assert hit > 0  # the leftover bytes did cross a window boundary in some placements

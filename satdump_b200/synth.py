"""Synthetic CCSDS transmitter + channel: the workload generator for tests, smoke() and bench.py.

NOT part of the receive hot path (nothing here is timed or shipped as the product): it manufactures
the baseband the hot path consumes, following SURVEY.md §8(d):

  payload -> CADU (ASM 1ACFFC1D | I x RS(255,223) dual-basis, byte interleaved) -> CCSDS randomiser from byte 4
          -> [NRZ-M] -> conv. code k=7 (polys 79,109 = the reference CCEncoder convention, cc_encoder.cpp:92-104)
          -> [puncture 3/4 in MetOp order: inverse of viterbi_3_4.cpp:84-104] -> BPSK/QPSK/OQPSK mapping
          -> RRC pulse shaping at a fractional samples-per-symbol -> carrier offset/phase, clock offset, AWGN
          -> cf32 / cs16 / cs8.

All coding is written from the CCSDS definitions (GF(256) poly 0x187, generator roots alpha^(11*(112+i)),
dual basis, PN h(x)=x^8+x^7+x^5+x^3+1); tests/test_synth.py cross-checks every encoder against the
reference's own encoders (oracle/_ref) and checks that the reference receiver recovers the payload.
"""
from dataclasses import dataclass, field

import numpy as np

ASM = bytes([0x1A, 0xCF, 0xFC, 0x1D])

# ----------------------------------------------------------------------------- GF(256), RS(255,223)
_GF_POLY = 0x187


def _gf_tables():
    exp = np.zeros(512, np.int32)
    log = np.zeros(256, np.int32)
    x = 1
    for i in range(255):
        exp[i] = x
        log[x] = i
        x <<= 1
        if x & 0x100:
            x ^= _GF_POLY
    exp[255:510] = exp[0:255]
    return exp, log


GF_EXP, GF_LOG = _gf_tables()


def gf_mul(a, b):
    a = np.asarray(a, np.int32)
    b = np.asarray(b, np.int32)
    r = GF_EXP[(GF_LOG[a] + GF_LOG[b]) % 255]
    return np.where((a == 0) | (b == 0), 0, r)


def rs_generator(nroots=32, fcr=112, gap=11):
    """g(x) = prod_i (x - alpha^(gap*(fcr+i))), returned lowest order first, g[nroots] = 1."""
    g = np.array([1], np.int32)
    for i in range(nroots):
        root = GF_EXP[(gap * (fcr + i)) % 255]
        g = np.concatenate([[0], g]) ^ np.concatenate([gf_mul(g, root), [0]])
    return g


# CCSDS dual-basis map is GF(2)-linear: images of the 8 unit vectors of the conventional basis
_TO_DUAL_BASIS_IMAGES = [0x7B, 0xAF, 0x99, 0xFA, 0x86, 0xEC, 0xEF, 0x8D]


def _dual_tables():
    to = np.zeros(256, np.uint8)
    for v in range(256):
        r = 0
        for b in range(8):
            if v >> b & 1:
                r ^= _TO_DUAL_BASIS_IMAGES[b]
        to[v] = r
    frm = np.zeros(256, np.uint8)
    frm[to] = np.arange(256, dtype=np.uint8)
    return to, frm


TO_DUAL, FROM_DUAL = _dual_tables()


def rs_encode(msg, dual=True, nroots=32):
    """msg: (ncw, 255-nroots) uint8 -> (ncw, 255) codewords (message then parity, highest order first)."""
    msg = np.asarray(msg, np.uint8)
    m = FROM_DUAL[msg] if dual else msg
    g = rs_generator(nroots, 112 if nroots == 32 else 120)
    glog = GF_LOG[g[:nroots]][::-1].copy()  # reg[0] is the highest-order remainder coefficient
    gz = (g[:nroots][::-1] == 0)
    reg = np.zeros((m.shape[0], nroots), np.int32)
    for i in range(m.shape[1]):
        fb = m[:, i].astype(np.int32) ^ reg[:, 0]
        reg = np.concatenate([reg[:, 1:], np.zeros((reg.shape[0], 1), np.int32)], axis=1)
        prod = GF_EXP[(GF_LOG[fb][:, None] + glog[None, :]) % 255]
        prod[fb == 0, :] = 0
        prod[:, gz] = 0
        reg ^= prod
    cw = np.concatenate([m, reg.astype(np.uint8)], axis=1)
    return TO_DUAL[cw] if dual else cw


def ccsds_pn(n):
    """CCSDS pseudo-randomiser bytes (h(x)=x^8+x^7+x^5+x^3+1, all-ones start), period 255 bytes... tiled to n."""
    reg = [1] * 8
    bits = []
    for _ in range(255 * 8):
        bits.append(reg[0])
        nb = reg[0] ^ reg[3] ^ reg[5] ^ reg[7]
        reg = reg[1:] + [nb]
    pn = np.packbits(np.array(bits, np.uint8))
    return np.resize(pn, n)


def build_cadus(payload, interleave, dual=True, randomize=True):
    """payload: (nframes, interleave*223) -> (nframes, 4 + interleave*255) CADUs and the clear (pre-randomiser) frames."""
    nf = payload.shape[0]
    msg = payload.reshape(nf, 223, interleave).transpose(0, 2, 1).reshape(nf * interleave, 223)
    cw = rs_encode(msg, dual).reshape(nf, interleave, 255).transpose(0, 2, 1).reshape(nf, 255 * interleave)
    clear = np.concatenate([np.tile(np.frombuffer(ASM, np.uint8), (nf, 1)), cw], axis=1)
    tx = clear.copy()
    if randomize:
        tx[:, 4:] ^= ccsds_pn(255 * interleave)[None, :]
    return tx, clear


# ----------------------------------------------------------------------------- convolutional code
def conv_encode(bits):
    """k=7, polys 79 (taps 0,1,2,3,6) and 109 (taps 0,2,3,5,6), register = (state<<1)|bit, start state 0."""
    b = np.concatenate([np.zeros(6, np.uint8), np.asarray(bits, np.uint8)])
    n = b.size - 6
    s = lambda k: b[6 - k:6 - k + n]
    o0 = s(0) ^ s(1) ^ s(2) ^ s(3) ^ s(6)
    o1 = s(0) ^ s(2) ^ s(3) ^ s(5) ^ s(6)
    return np.stack([o0, o1], axis=1).reshape(-1)


def puncture_34_metop(coded):
    """From each 6 coded symbols e0..e5 transmit e0,e1,e4,e3 (inverse of Viterbi3_4::depuncture, shift=0)."""
    n = coded.size // 6 * 6
    e = coded[:n].reshape(-1, 6)
    return e[:, [0, 1, 4, 3]].reshape(-1)


# Puncturing patterns that Viterbi_Depunc undoes (src-core/common/codings/viterbi/depunc.h): which of the mother code's symbols of one
# period (C1, C2 of consecutive bits) are transmitted = the positions where its depuncturers place data rather than an erasure.
PUNCTURE_KEEP = {"p2/3": (4, [0, 1, 3]), "p3/4": (6, [0, 1, 3, 4]), "p5/6": (10, [0, 1, 3, 4, 7, 8]), "p7/8": (14, [0, 1, 3, 5, 7, 8, 11, 12])}


def puncture_dvb(coded, conv):
    period, keep = PUNCTURE_KEEP[conv]
    n = coded.size // period * period
    return coded[:n].reshape(-1, period)[:, keep].reshape(-1)


def nrzm_encode(bits):
    return (np.cumsum(bits.astype(np.int64)) & 1).astype(np.uint8)


# ----------------------------------------------------------------------------- waveform
def rrc_pulse(t, alpha):
    """Unit-energy root-raised-cosine impulse response at times t (in symbols)."""
    t = np.asarray(t, np.float64)
    out = np.empty_like(t)
    eps = 1e-9
    z = np.abs(t) < eps
    s = np.abs(np.abs(t) - 1.0 / (4 * alpha)) < eps
    r = ~(z | s)
    tr = t[r]
    out[r] = (np.sin(np.pi * tr * (1 - alpha)) + 4 * alpha * tr * np.cos(np.pi * tr * (1 + alpha))) / (np.pi * tr * (1 - (4 * alpha * tr) ** 2))
    out[z] = 1 - alpha + 4 * alpha / np.pi
    out[s] = alpha / np.sqrt(2) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha)) + (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
    return out


@dataclass
class SignalCfg:
    """One of the BASELINE.json configurations."""
    name: str = "metop_ahrpt"
    samplerate: float = 6e6
    symbolrate: float = 2333333
    constellation: str = "qpsk"     # bpsk | qpsk | oqpsk
    conv: str = "3/4"               # "1/2" | "3/4" (MetOp order) | "none" | "p2/3" "p3/4" "p5/6" "p7/8" (Viterbi_Depunc's patterns)
    interleave: int = 4             # RS interleaving depth I
    nrzm: bool = False
    rrc_alpha: float = 0.5
    pll_bw: float = 0.003
    fmt: str = "cs16"
    esn0_db: float = 10.0
    carrier_rad: float = 1e-3       # rad / sample
    phase0: float = 0.7
    clock_ppm: float = 20.0
    rms: float = 0.25
    clock_alpha: float = None       # M&M gains override (DVB-S2 front half uses 1.7e-3)
    # receiver-side decoder parameters (pipeline JSON values)
    decoder: str = "metop"          # metop | ccsds | none
    ber_thresold: float = 0.28
    outsync_after: int = 10
    rs_usecheck: bool = False
    # phase-modulated residual-carrier signal for pm_demod (module_pm_demod.cpp): carrier * exp(j * pm_index * d(t) * sin(2 pi f_sc t)) with
    # d(t) the pulse-shaped BPSK stream and f_sc the subcarrier (0 = the symbol rate, the module's default for "subcarrier_offset")
    pm_index: float = 0.0           # rad; 0 = not a PM signal
    subcarrier: float = 0.0
    pm_pll_bw: float = 0.01         # receiver: carrier PLL bandwidth ("pll_bw" of pm_demod; pll_bw above is its "costas_bw")
    pm_pll_max_offset: float = 3.14
    resample_after_pll: bool = False
    # psk_demod's carrier mode ("has_carrier", module_psk_demod.cpp:93-113): BPSK phase-modulated directly onto a residual carrier
    # (pm_index * d(t), no subcarrier): carrier PLL -> DC blocker -> Costas loop
    has_carrier: bool = False
    carrier_pll_bw: float = 0.001

    @property
    def cadu_bytes(self):
        return 4 + 255 * self.interleave

    @property
    def sps(self):
        return float(np.float32(np.float32(self.samplerate) / np.float32(int(self.symbolrate))))


CONFIGS = {
    # C1/C3: resources/pipelines/MetOp.json:30-47
    "metop_ahrpt": SignalCfg(),
    # C2: BPSK + r=1/2 + RS I=4 (ccsds_conv_concat_decoder), cf32 @3 MS/s, symbolrate chosen so sps=2.5
    "bpsk_half": SignalCfg(name="bpsk_half", samplerate=3e6, symbolrate=1200000, constellation="bpsk", conv="1/2", interleave=4,
                           fmt="cf32", decoder="ccsds", ber_thresold=0.3, outsync_after=20, esn0_db=7.0),
    # C4: JPSS-HRD-type OQPSK r=1/2 NRZ-M RS I=5, cs16 @30 MS/s, 15 Msym/s (sps 2.0) (resources/pipelines/JPSS.json npp_hrd/jpss_hrd)
    "jpss_hrd": SignalCfg(name="jpss_hrd", samplerate=30e6, symbolrate=15000000, constellation="oqpsk", conv="1/2", interleave=5,
                          nrzm=True, pll_bw=0.002, fmt="cs16", decoder="ccsds", ber_thresold=0.3, outsync_after=20, rs_usecheck=True,
                          esn0_db=7.0),
    # C2 at the real NOAA HRPT rate: 665.4 kbaud @ 3 MS/s = 4.51 samples/symbol > MAX_SPS, so BaseDemodModule's front-end resampler
    # runs first (module_demod_base.cpp:66-80,203-204): 3 MS/s -> 2.4 MS/s (4/5), 3.607 samples/symbol
    "hrpt_bpsk": SignalCfg(name="hrpt_bpsk", samplerate=3e6, symbolrate=665400, constellation="bpsk", conv="1/2", interleave=4,
                           fmt="cf32", decoder="ccsds", ber_thresold=0.3, outsync_after=20, esn0_db=7.0),
    # a narrow-roll-off QPSK stream recorded too slowly: 2.4 Msym/s @ 2.6 MS/s = 1.083 samples/symbol < MIN_SPS -> interpolated to
    # 2.64 MS/s (66/65). Exercises the interpolating branch of the resampler; too close to aliasing to be a decoding test
    "qpsk_undersampled": SignalCfg(name="qpsk_undersampled", samplerate=2.6e6, symbolrate=2400000, constellation="qpsk", conv="1/2",
                                   interleave=4, rrc_alpha=0.1, fmt="cs16", decoder="none", esn0_db=14.0),
    # psk_demod -> ccsds_simple_psk_decoder (no convolutional code; 50 shipped pipelines have this shape, e.g. FengYun-3.json:630-639):
    # BPSK + NRZ-M + RS I=4, and QPSK (I rail first on air, so the receiver runs with qpsk_swap_iq) + RS I=4
    "bpsk_simple": SignalCfg(name="bpsk_simple", samplerate=3e6, symbolrate=1200000, constellation="bpsk", conv="none", interleave=4, nrzm=True,
                             fmt="cs16", decoder="simple", esn0_db=9.0),
    "qpsk_simple": SignalCfg(name="qpsk_simple", samplerate=6e6, symbolrate=2400000, constellation="qpsk", conv="none", interleave=4, fmt="cs16",
                             decoder="simple", esn0_db=12.0),
    # 8PSK through psk_demod alone (order-8 Costas loop; no decoder of this path takes 8PSK): demodulator parity only
    "psk8": SignalCfg(name="psk8", samplerate=6e6, symbolrate=2400000, constellation="8psk", conv="none", interleave=4, fmt="cs16",
                      decoder="demod", esn0_db=18.0),
    # MetOp AHRPT recorded at 24 MS/s: 10.3 samples/symbol -> BaseDemodModule resamples to 8 MS/s (round(2333333 / 1e6) * 1e6 * MAX_SPS), a
    # ratio of 3, so SmartResamplerBlock runs its power-of-two decimator (x2: one 69-tap stage) and then the rational resampler 2/3
    "metop_oversampled": SignalCfg(name="metop_oversampled", samplerate=24e6),
    # BPSK r=1/2 at 1 Msym/s recorded at 32 MS/s in cs8: ratio 8 -> power-of-two decimator alone (stages /4 and /2), no rational part
    "bpsk_decim8": SignalCfg(name="bpsk_decim8", samplerate=32e6, symbolrate=1000000, constellation="bpsk", conv="1/2", interleave=4,
                             fmt="cs8", decoder="ccsds", ber_thresold=0.3, outsync_after=20, esn0_db=9.0),
    # ccsds_conv_concat_decoder with conv_rate 2/3 ... 7/8 (Viterbi_Depunc, viterbi_punc.cpp): QPSK + punctured k=7 code + RS I=4
    "qpsk_p23": SignalCfg(name="qpsk_p23", symbolrate=2400000, conv="p2/3", decoder="ccsds", ber_thresold=0.3, outsync_after=20, esn0_db=9.0),
    "qpsk_p34": SignalCfg(name="qpsk_p34", symbolrate=2400000, conv="p3/4", decoder="ccsds", ber_thresold=0.3, outsync_after=20, esn0_db=10.0),
    "qpsk_p56": SignalCfg(name="qpsk_p56", symbolrate=2400000, conv="p5/6", decoder="ccsds", ber_thresold=0.3, outsync_after=20, esn0_db=11.5),
    "qpsk_p78": SignalCfg(name="qpsk_p78", symbolrate=2400000, conv="p7/8", decoder="ccsds", ber_thresold=0.3, outsync_after=20, esn0_db=12.5),
    # pm_demod -> ccsds_conv_concat_decoder (14 shipped pipelines have this shape): residual-carrier PM, BPSK r=1/2 on a subcarrier at the
    # symbol rate, 6 samples per symbol (inside pm_demod's [1.1, 10] window: no resampler), RS I=4
    "pm_bpsk": SignalCfg(name="pm_bpsk", samplerate=3e6, symbolrate=500000, constellation="bpsk", conv="1/2", interleave=4, fmt="cs16",
                         decoder="ccsds", ber_thresold=0.3, outsync_after=20, esn0_db=20.0, pll_bw=0.005, carrier_rad=2e-2, pm_index=1.0),
    # the same at 24 samples per symbol with "resample_after_pll" (27 shipped pipelines set it): carrier PLL and PMToBPSK at the input
    # rate, then SmartResamplerBlock 6 MS/s -> 2 MS/s (decimator /2 + rational 2/3) and the second AGC, 8 samples per symbol
    "pm_bpsk_after": SignalCfg(name="pm_bpsk_after", samplerate=6e6, symbolrate=250000, constellation="bpsk", conv="1/2", interleave=4, fmt="cf32",
                               decoder="ccsds", ber_thresold=0.3, outsync_after=20, esn0_db=26.0, pll_bw=0.005, carrier_rad=1e-2, pm_index=1.0,
                               resample_after_pll=True),
    # the same 24 samples per symbol WITHOUT resample_after_pll: BaseDemodModule resamples 6 MS/s -> 2 MS/s in front (decimator /2 + rational
    # 2/3), the carrier PLL and PMToBPSK run at the working rate
    "pm_bpsk_front": SignalCfg(name="pm_bpsk_front", samplerate=6e6, symbolrate=250000, constellation="bpsk", conv="1/2", interleave=4, fmt="cs16",
                               decoder="ccsds", ber_thresold=0.3, outsync_after=20, esn0_db=26.0, pll_bw=0.005, carrier_rad=1e-2, pm_index=1.0),
    # psk_demod with "has_carrier" (ODIN.json:16-24): BPSK r=1/2 phase-modulated onto a residual carrier, 2.5 samples per symbol
    "bpsk_carrier": SignalCfg(name="bpsk_carrier", samplerate=6e6, symbolrate=2400000, constellation="bpsk", conv="1/2", interleave=4, fmt="cs16",
                              decoder="ccsds", ber_thresold=0.3, outsync_after=20, esn0_db=10.0, pll_bw=0.001, carrier_rad=5e-3, pm_index=1.2,
                              has_carrier=True, carrier_pll_bw=0.001, rrc_alpha=0.35),
    # C5: DVB-S2 front half AGC->RRC->M&M, cs8, 45 Msym/s @ 90 MS/s (sps 2.0), alpha 0.25 (DVB_Test.json:132-137), REC_ALPHA 1.7e-3
    "dvbs2_front": SignalCfg(name="dvbs2_front", samplerate=90e6, symbolrate=45000000, constellation="qpsk", conv="none", interleave=4,
                             rrc_alpha=0.25, fmt="cs8", decoder="none", clock_alpha=1.7e-3, carrier_rad=0.0, phase0=0.0, esn0_db=12.0),
}


def qpsk_diff_encode(dibits, swap=True):
    """Transmit side of the reference's QPSKDiff decoder (src-core/common/codings/differential/qpsk_diff.cpp:5-53): hard symbols
    s = 2*X + Y (X on the Q rail, Y on the I rail, as constellation_t::soft_demod reads them) such that decoding the pair
    (s[i-1], s[i]) gives dibits[i]. Two leading symbols fill the decoder's buffer."""
    def dec(prev, cur):
        xin_1, yin_1, xin, yin = prev & 2, prev & 1, cur & 2, cur & 1
        if ((xin >> 1) ^ yin) == 1:
            ou = ((yin_1 ^ yin) << 1) + ((xin_1 ^ xin) >> 1)
        else:
            ou = (xin_1 ^ xin) + (yin_1 ^ yin)
        return ((ou & 1) << 1 | (ou >> 1)) if swap else ou  # value of the two output bits, first bit = MSB
    nxt = np.zeros((4, 4), np.int64)
    for prev in range(4):
        for cur in range(4):
            nxt[prev, dec(prev, cur)] = cur
    d = np.asarray(dibits, np.int64)
    out = np.zeros(d.size + 2, np.int64)
    cur = 0
    for i, v in enumerate(d.tolist()):
        cur = int(nxt[cur, v])
        out[i + 2] = cur
    return out


def make_bitstream(cfg: SignalCfg, nframes, seed):
    rng = np.random.default_rng(seed)
    payload = rng.integers(0, 256, size=(nframes, cfg.interleave * 223), dtype=np.uint8)
    tx, clear = build_cadus(payload, cfg.interleave)
    bits = np.unpackbits(tx.reshape(-1))
    if cfg.nrzm:
        bits = nrzm_encode(bits)
    if cfg.conv == "none":
        coded = bits
    else:
        coded = conv_encode(bits)
        if cfg.conv == "3/4":
            coded = puncture_34_metop(coded)
        elif cfg.conv in PUNCTURE_KEEP:
            coded = puncture_dvb(coded, cfg.conv)
    return coded, clear


def modulate(cfg: SignalCfg, coded, seed, nsamples=None, device="cpu"):
    """coded bits -> raw IQ in cfg.fmt. Heavy lifting in torch so bench-sized signals are made on the GPU."""
    import torch
    dev = torch.device(device)
    bps = 1 if cfg.constellation == "bpsk" else (3 if cfg.constellation == "8psk" else 2)
    nsym = coded.size // bps
    if bps == 3:  # 8PSK: points at pi/8 + k*pi/4, where the order-8 Costas detector (costas_loop.cpp) has its stable locks
        b = coded[:nsym * 3].reshape(-1, 3).astype(np.int64)
        ang = np.pi / 8 + (b[:, 0] * 4 + b[:, 1] * 2 + b[:, 2]) * (np.pi / 4)
        ai = torch.from_numpy(np.cos(ang).astype(np.float32)).to(dev)
        aq = torch.from_numpy(np.sin(ang).astype(np.float32)).to(dev)
    else:
        a = torch.from_numpy((coded[:nsym * bps].astype(np.float32) * 2 - 1)).to(dev)
        if bps == 2:
            ai, aq = a[0::2].contiguous(), a[1::2].contiguous()
        else:
            ai, aq = a, None
    sps = cfg.samplerate / cfg.symbolrate * (1.0 + cfg.clock_ppm * 1e-6)  # samples per symbol seen by the receiver
    span = 10
    total = int((nsym - 2 * span) * sps)
    if nsamples is None or nsamples > total:
        nsamples = total
    # pulse table, 1/2048-symbol grid, linear interpolation
    res = 2048
    tgrid = np.arange(-span * res, span * res + 2) / res
    tab = torch.from_numpy(rrc_pulse(tgrid, cfg.rrc_alpha).astype(np.float32)).to(dev)
    out_i = torch.empty(nsamples, dtype=torch.float32, device=dev)
    out_q = torch.empty(nsamples, dtype=torch.float32, device=dev)
    step = 1 << 20
    qoff = 0.5 if cfg.constellation == "oqpsk" else 0.0

    def shape(sym, t):  # t: symbol-time of each output sample (float64)
        k0 = torch.floor(t)
        d = (t - k0)  # [0,1)
        k0 = k0.to(torch.int64)
        pos = d * res
        pb = torch.floor(pos)
        fr = (pos - pb).to(torch.float32)
        pb = pb.to(torch.int64)
        acc = torch.zeros(t.shape, dtype=torch.float32, device=t.device)
        for j in range(-span + 1, span + 1):
            p0 = pb + (span - j) * res  # table index of (t - (k0 + j)) on the 1/res grid
            h = tab[p0] * (1 - fr) + tab[p0 + 1] * fr
            acc += sym[k0 + j] * h
        return acc

    for s in range(0, nsamples, step):
        e = min(nsamples, s + step)
        n = torch.arange(s, e, device=dev, dtype=torch.float64)
        t = n / sps + span + 0.37
        out_i[s:e] = shape(ai, t)
        if cfg.pm_index:
            if cfg.has_carrier:
                ph = cfg.pm_index * out_i[s:e]
            else:
                fsc = cfg.subcarrier if cfg.subcarrier else cfg.symbolrate
                sub = torch.sin((2 * np.pi * fsc / cfg.samplerate * n) % (2 * np.pi)).float()
                ph = cfg.pm_index * out_i[s:e] * sub
            out_i[s:e] = torch.cos(ph)
            out_q[s:e] = torch.sin(ph)
        elif aq is not None:
            out_q[s:e] = shape(aq, t - qoff)
        else:
            out_q[s:e] = 0
    x = torch.complex(out_i, out_q)
    p = float((x.real ** 2 + x.imag ** 2).mean())
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) + 7)
    sigma2 = p * (cfg.samplerate / cfg.symbolrate) / (10 ** (cfg.esn0_db / 10))
    for s in range(0, nsamples, step * 8):
        e = min(nsamples, s + step * 8)
        n = torch.arange(s, e, device=dev, dtype=torch.float64)
        ph = (cfg.carrier_rad * n + cfg.phase0) % (2 * np.pi)
        rot = torch.complex(torch.cos(ph).float(), torch.sin(ph).float())
        noise = torch.randn(e - s, 2, generator=gen, device=dev, dtype=torch.float32) * float(np.sqrt(sigma2 / 2))
        x[s:e] = x[s:e] * rot + torch.complex(noise[:, 0], noise[:, 1])
    scale = cfg.rms / float(np.sqrt(p + sigma2))
    x = x * scale
    v = torch.view_as_real(x).reshape(-1)
    if cfg.fmt == "cf32":
        return torch.view_as_complex(v.reshape(-1, 2).contiguous())
    if cfg.fmt == "cs16":
        return torch.clamp(torch.round(v * 32767), -32767, 32767).to(torch.int16)
    if cfg.fmt == "cs8":
        return torch.clamp(torch.round(v * 127), -127, 127).to(torch.int8)
    raise ValueError(cfg.fmt)


def frames_for_samples(cfg: SignalCfg, nsamples):
    bps = 1 if cfg.constellation == "bpsk" else 2
    rate = {"1/2": 0.5, "3/4": 0.75, "none": 1.0, "p2/3": 2 / 3, "p3/4": 0.75, "p5/6": 5 / 6, "p7/8": 7 / 8}[cfg.conv]
    bits_per_sample = bps * rate / (cfg.samplerate / cfg.symbolrate)
    return int(nsamples * bits_per_sample / (cfg.cadu_bytes * 8)) + 4


def make_signal(cfg: SignalCfg, nsamples, seed=0xB2000000, device="cpu"):
    """Returns (raw torch tensor in cfg.fmt with exactly <= nsamples samples, clear CADUs (nframes, cadu_bytes) numpy)."""
    nframes = frames_for_samples(cfg, nsamples)
    coded, clear = make_bitstream(cfg, nframes, seed)
    raw = modulate(cfg, coded, seed, nsamples, device)
    return raw, clear


# ---------------------------------------------------------------------------------------------------------------- AOS frames with space packets
def build_aos_frames(nframes, seed, vcids=(9, 12, 3, 34), mpdu=884, insert_zone=0, cadu_size=1024, corrupt=0.0, drop=0.0, idle=0.05):
    """A CADU stream whose virtual channels carry CCSDS space packets in M-PDUs (the input of ccsds_aos::Demuxer, demuxer.cpp): per channel a
    byte stream of packets (lengths from 1 byte to several frames, so headers straddle frames and first-header-pointers of 2047 occur),
    cut into `mpdu`-byte zones with the first header pointer of each; channels interleaved at random, idle frames (VCID 63, pointer 2046)
    in between. corrupt: fraction of the frames with random damage to the pointer / packet headers / payload (exercises the demuxer's
    behaviour on inconsistent input); drop: fraction of the frames removed after packing (discontinuities). Returns uint8 [frames, cadu_size]."""
    rng = np.random.default_rng(seed)
    base = 10 + (insert_zone if insert_zone else 0)
    assert base + 2 + mpdu <= cadu_size
    streams = {}

    def more(v, need):
        st = streams.setdefault(v, dict(buf=bytearray(), starts=[], pos=0, seq=0))
        while len(st["buf"]) - st["pos"] < need:
            r = rng.random()
            n = int(rng.integers(1, 300)) if r < 0.7 else (int(rng.integers(300, 2000)) if r < 0.9 else int(rng.integers(2000, 14000)))
            apid = int(rng.choice([34, 39, 103, 104, 130, 384, 6]))
            sec = int(rng.integers(0, 2))
            st["seq"] = (st["seq"] + 1) & 0x3FFF
            hdr = bytes([(sec << 3) | (apid >> 8), apid & 0xFF, 0xC0 | (st["seq"] >> 8), st["seq"] & 0xFF, ((n - 1) >> 8) & 0xFF, (n - 1) & 0xFF])
            st["starts"].append(len(st["buf"]))
            st["buf"] += hdr + rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        return st

    out = np.zeros((nframes, cadu_size), np.uint8)
    out[:] = rng.integers(0, 256, (nframes, cadu_size), dtype=np.uint8)
    counters = {}
    for f in range(nframes):
        v = 63 if rng.random() < idle else int(rng.choice(vcids))
        fr = out[f]
        fr[0:4] = (0x1A, 0xCF, 0xFC, 0x1D)
        scid = 0x0B
        fr[4] = (1 << 6) | (scid >> 2)
        fr[5] = ((scid & 3) << 6) | v
        c = counters.get(v, 0)
        counters[v] = (c + 1) & 0xFFFFFF
        fr[6], fr[7], fr[8], fr[9] = (c >> 16) & 0xFF, (c >> 8) & 0xFF, c & 0xFF, 0
        if v == 63:
            fhp = 2046
        else:
            st = more(v, mpdu)
            p0 = st["pos"]
            fr[base + 2:base + 2 + mpdu] = np.frombuffer(bytes(st["buf"][p0:p0 + mpdu]), np.uint8)
            nxt = [s for s in st["starts"] if p0 <= s < p0 + mpdu]
            fhp = nxt[0] - p0 if nxt else 2047
            st["pos"] = p0 + mpdu
            st["starts"] = [s for s in st["starts"] if s >= p0 + mpdu]
            if st["pos"] > 1 << 20:  # keep the buffer short
                st["buf"] = st["buf"][st["pos"]:]
                st["starts"] = [s - st["pos"] for s in st["starts"]]
                st["pos"] = 0
        fr[base] = (int(rng.integers(0, 32)) << 3) | (fhp >> 8)
        fr[base + 1] = fhp & 0xFF
    if corrupt:
        for f in np.nonzero(rng.random(nframes) < corrupt)[0]:
            kind = int(rng.integers(0, 4))
            if kind == 0:  # a random first header pointer
                fhp = int(rng.integers(0, 2048))
                out[f, base] = (out[f, base] & 0xF8) | (fhp >> 8)
                out[f, base + 1] = fhp & 0xFF
            elif kind == 1:  # pointer "no header" although there is one / short pointers
                fhp = int(rng.choice([2047, 0, 1, 2, 3, 4, 5, mpdu - 1, mpdu - 3, mpdu - 6, mpdu - 7, mpdu, mpdu + 1]))
                out[f, base] = (out[f, base] & 0xF8) | (fhp >> 8)
                out[f, base + 1] = fhp & 0xFF
            elif kind == 2:  # a burst of wrong bytes somewhere in the data zone (packet lengths go wrong)
                a = int(rng.integers(0, mpdu - 16))
                out[f, base + 2 + a:base + 2 + a + 16] = rng.integers(0, 256, 16, dtype=np.uint8)
            else:  # the frame lands in another virtual channel
                out[f, 5] = (out[f, 5] & 0xC0) | int(rng.choice(vcids))
    if drop:
        out = out[rng.random(nframes) >= drop]
    return np.ascontiguousarray(out)


def craft_leftover_frames(mpdu, variant=0, vcid=9, cadu_size=1024):
    """Six frames that drive ccsds_aos::Demuxer into its rarest corner: a header straddles a frame boundary, the next frame claims to hold no
    header (pointer 2047) although the packet ends inside it, so the continuation takes more than what remained (demuxer.cpp:109) and the
    packet never completes; its bytes then stay in front of the next packet (readPacket does not clear them, :26-33). variant 0: that next
    packet completes inside its frame; variant 1: it runs on into the following frames."""
    rng = np.random.default_rng(100 + variant)
    fr = rng.integers(0, 256, (6, cadu_size), dtype=np.uint8)
    base = 10

    def head(f, fhp):
        fr[f, 5] = (fr[f, 5] & 0xC0) | vcid
        fr[f, base] = fhp >> 8
        fr[f, base + 1] = fhp & 0xFF

    def d(f):
        return fr[f, base + 2:]

    head(0, mpdu - 3)  # three header bytes at the end of the zone
    fr[0, base + 2 + mpdu - 3] &= 0xF7
    head(1, 2047)
    pl = (mpdu - 3 - 1) - 1  # remaining = mpdu - 4 with offset 3: (rem + 3) > mpdu - 3 but rem < mpdu - 3
    d(1)[1], d(1)[2] = pl >> 8, pl & 0xFF
    head(2, 10)
    n = 20 if variant == 0 else 3000
    d(2)[10 + 4], d(2)[10 + 5] = (n - 1) >> 8, (n - 1) & 0xFF
    if variant == 0:
        d(2)[10 + 6 + n + 4], d(2)[10 + 6 + n + 5] = 0, 49
    head(3, 2047 if variant else 100)
    head(4, 50)
    head(5, 7)
    for f in (3, 4, 5):
        p = ((int(fr[f, base]) & 7) << 8) | int(fr[f, base + 1])
        if p < 2047:
            d(f)[p + 4], d(f)[p + 5] = 0, 30
    return fr

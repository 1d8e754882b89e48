/*
 * TEST INFRASTRUCTURE ONLY — the CPU oracle ("port") for the B200 baseband->CADU hot path.
 *
 * A from-scratch C restatement of the algorithms of the reference's CPU path. Every function cites
 * the reference file:line whose behaviour it follows. It is pinned bit-for-bit against the reference's
 * own code compiled into oracle/_ref (tests/test_oracle_*.py) and against tests/golden fixtures made
 * from it. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product (libb200dsp.so) never does.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
    double samplerate, symbolrate;
    int constellation; /* 0 bpsk, 1 qpsk, 2 oqpsk, 3 8psk, 4 none (AGC->RRC->M&M only) */
    float rrc_alpha;
    int rrc_taps;
    float pll_bw, agc_rate;
    float clock_gain_omega, clock_mu, clock_gain_mu, clock_omega_limit;
    float costas_max_offset;
    int format; /* 0 cf32, 1 cs16, 2 cs8 */
    int buffer_size;
    int iq_swap;             /* re <-> im at the reader (file_source.cpp:31-33) */
    double final_samplerate; /* 0 = samplerate; else the rate BaseDemodModule::initb resamples to (module_demod_base.cpp:59-87) */
    int dc_block;            /* CorrectIQBlock behind the reader (module_demod_base.cpp:113-114) */
    int post_costas_dc;      /* CorrectIQBlock behind the Costas loop (module_psk_demod.cpp:127-134) */
    int clock_recovery;      /* 0 MMClockRecoveryBlock, 1 GardnerClockRecoveryBlock<complex_t> (clock_recovery_gardner.cpp) */
    /* PMDemodModule (module_pm_demod.cpp:61-88): AGC -> PLLCarrierTrackingBlock -> PMToBPSK -> [SmartResampler -> AGC2] -> RRC -> Costas
       (order 2, default frequency limit) -> M&M; pll_bw above is then "costas_bw" */
    int pm;
    float pm_pll_bw, pm_pll_max_offset;
    int pm_resample_after_pll;
    double pm_subcarrier_offset; /* 0 = the symbol rate */
    double freq_shift;           /* FreqShiftBlock behind the reader / DC blocker (module_demod_base.cpp:122-123); 0 = none */
    int has_carrier;             /* psk_demod carrier mode (module_psk_demod.cpp:93-113): RRC -> carrier PLL -> DC blocker -> Costas */
    float carrier_pll_bw, carrier_pll_max_offset;
} orc_demod_cfg;

typedef struct
{
    int kind;          /* 0 metop_ahrpt_decoder, 1 ccsds_conv_concat_decoder r=1/2 */
    int constellation; /* 0 bpsk, 1 qpsk, 2 oqpsk, 5 bpsk_90 */
    int cadu_size, outsync_after;
    float ber_thresold;
    int nrzm, derandomize, derand_after_rs, derand_start;
    int rs_i, rs_dualbasis, rs_fill_bytes, rs_usecheck, rs_type;
    int iq_invert;
    unsigned int asm_sync;
    int qpsk_swap_iq, qpsk_swap_diff, oqpsk_delay; /* kind 2 = ccsds_simple_psk_decoder */
    int conv_rate;                                 /* kind 1: 0 = 1/2, 2 / 3 / 5 / 7 = Viterbi_Depunc rates 2/3, 3/4, 5/6, 7/8 */
} orc_fec_cfg;

int orc_rrc_taps(double gain, double fs, double rs, double alpha, int ntaps, float *out);
void orc_mm_bank(float *out /* 128*8 */);

void *orc_demod_create(const orc_demod_cfg *cfg);
void orc_demod_destroy(void *h);
float orc_demod_sps(void *h);
long orc_demod_run(void *h, const void *raw, long nsamples, float *agc_out, float *fir_out, float *costas_out, float *mm_out,
                   int8_t *soft_out, long sym_cap);
void orc_demod_state(void *h, float *out8);
long orc_demod_last_front(void *h);
/* pm_demod: where the next orc_demod_run calls dump the PLLCarrierTrackingBlock / PMToBPSK outputs; carrier PLL state + AGC2 gain */
void orc_demod_pm_dumps(void *h, float *pll_out, float *pm_out);
void orc_demod_pm_state(void *h, float *out4);
/* front end alone: conversion (+ iq_swap) + SmartResamplerBlock (rational part) on a fresh resampler; returns output samples */
long orc_resample(const orc_demod_cfg *cfg, const void *raw, long nsamples, float *out, long cap);
/* polyphase bank of RationalResamplerBlock(interpolation, decimation): returns taps per arm, *nfilt arms; out[arm*ntaps + k] */
int orc_resampler_taps(unsigned interpolation, unsigned decimation, float *out, int cap, int *nfilt);

/* AGC recurrence in double precision on cf32 input (reference rounding-noise floor, see oracle.c) */
void orc_agc_exact(const float *in, long n, double rate, double ref, double max_gain, float *out);

/* ---- pm_demod / freq_shift blocks (module_pm_demod.cpp:61-88, module_demod_base.cpp:125-126) */
float orc_fast_atan2f(float y, float x); /* common/dsp/utils/fast_trig.cpp:80-154 */
float orc_fast_cos(float x);             /* fast_trig.cpp:158-168 */
float orc_fast_sin(float x);             /* fast_trig.cpp:170-180 */
/* PLLCarrierTrackingBlock (common/dsp/pll/pll_carrier_tracking.cpp:8-71) over n cf32 samples; state[2] = {d_phase, d_freq}, in / out */
void orc_pll_carrier(const float *in, long n, float loop_bw, float max_freq, float min_freq, float *state, float *out);
/* the VOLK rotator behind FreqShiftBlock / PMToBPSK (freq_shift.cpp:16-50, pm_to_bpsk.cpp:10-35), `call` samples per call;
   phase[2] in / out (starts at (1, 0)); imag_only: the input is first reduced to (0, imag) as PMToBPSK does (pm_to_bpsk.cpp:24-25) */
void orc_rotator(const float *in, long n, long call, float inc_re, float inc_im, int imag_only, float *phase, float *out);
/* phase_delta of FreqShiftBlock::set_freq / PMToBPSK's constructor: (cos, sin)(2 pi freq / samplerate) rounded to float */
void orc_rotator_inc(double freq, double samplerate, float *inc2);

void *orc_fec_create(const orc_fec_cfg *cfg);
void orc_fec_destroy(void *h);
int orc_fec_chunk_size(void *h);
int orc_fec_cadu_bytes(void *h);
long orc_fec_run(void *h, const int8_t *soft, long nsoft, uint8_t *cadu_out, long cadu_cap, int *vit_state, float *vit_ber,
                 int *defr_state, uint8_t *bits_out, long *nbits, int *rs_err, long *nframes_seen);

/* primitives for stage-isolated checks */
void orc_cc_decode(const uint8_t *syms, int frame, int ncalls, uint8_t *out_bits);
void orc_cc_encode(const uint8_t *bits, int n, uint8_t *out);
void orc_derand(uint8_t *data, int len);
void orc_rs_decode_interleaved(uint8_t *data, int dual, int interleave, int rs_type, int fill_bytes, int *errors);
int orc_deframe(const uint8_t *bits, int nbits, int cadu_size, int state_synced, uint8_t *out);
void orc_rotate_soft(int8_t *soft, int size, int phase, int iqswap);

/* ---- CADU -> CCSDS space packets: one demultiplexer per virtual channel (common/ccsds/ccsds_aos/demuxer.cpp) behind
   parseVCDU's channel id (vcdu.cpp:10-17). Packets in the order the reference returns them, [6 header bytes][payload] back to back;
   recs[4 i + 0..3] = frame index, vcid, payload length, apid. Returns the packet count, or -1 when a capacity is too small. */
void *orc_demux_create(int mpdu_data_size, int has_insert_zone, int insert_zone_size, int secondary_header_extends);
void orc_demux_destroy(void *h);
long orc_demux_run(void *h, const uint8_t *frames, long nframes, int cadu_size, unsigned long long vcid_mask, long frame0, uint8_t *out, long cap_bytes,
                   long *nbytes, int *recs, long cap_recs);

/* single-thread end-to-end (demod + FEC) used as the CPU "port" baseline; returns CADU bytes */
long orc_pipeline_run(const orc_demod_cfg *dc, const orc_fec_cfg *fc, const void *raw, long nsamples, uint8_t *cadu_out, long cadu_cap);

#ifdef __cplusplus
}
#endif
#endif

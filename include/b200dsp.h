/*
 * b200dsp — C ABI of the B200-native (sm_100a) SatDump baseband -> CADU hot path.
 *
 * This is the boundary a SatDump plugin binds (INTEGRATION.md shows the ProcessingModule shim). It
 * sits one level below SatDump's C++ plugin ABI so that the CUDA library is independent of the host
 * compiler / STL: plain pointers, sizes and POD structs only, caller-owned host buffers, opaque handles.
 *
 * Each entry point names the reference interface it replaces (paths relative to the SatDump tree):
 *
 *   b200_demod_*   the DSP chain PSKDemodModule builds and runs
 *                    src-core/pipeline/modules/demod/module_psk_demod.cpp:86-236   (init / process)
 *                    src-core/pipeline/modules/demod/module_demod_base.cpp:59-208   (initb: source, AGC)
 *                    plugins/dvb_support/dvbs2/module_dvbs2_demod.cpp:98-102        (AGC->RRC->M&M front half; constellation NONE)
 *                    src-core/pipeline/modules/demod/module_pm_demod.cpp:61-160      (PMDemodModule: cfg.pm_demod = 1)
 *   b200_fec_*     the decoder modules
 *                    plugins/noaa_metop_support/metop/module_metop_ahrpt_decoder.cpp:34-90      (kind METOP)
 *                    src-core/pipeline/modules/ccsds/module_ccsds_conv_concat_decoder.cpp:140-200 (kind CCSDS, r=1/2)
 *   b200_chain_*   both modules joined the way Pipeline::run joins them with a byte FIFO
 *                    src-core/pipeline/pipeline_run.cpp:44-117 — here the int8 soft stream never leaves HBM.
 *   b200_demux_*   the step behind the decoder: CADUs -> CCSDS space packets
 *                    src-core/common/ccsds/ccsds_aos/demuxer.cpp:64-199 (one Demuxer per virtual channel), vcdu.cpp:10-17, mpdu.cpp:9-13,
 *                    as plugins/noaa_metop_support/metop/module_metop_instruments.cpp:66-140 calls them
 *
 * All functions return 0 on success or a negative B200_E* code; b200_last_error() gives the text.
 * There is no CPU fallback: without a CUDA device every create() fails with B200_ENODEV.
 */
#ifndef B200DSP_H
#define B200DSP_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_EINVAL -1     /* bad / unsupported parameter (the C++ module turns this into satdump_exception) */
#define B200_ENODEV -2     /* no CUDA device / wrong architecture */
#define B200_ECUDA -3      /* CUDA runtime failure */
#define B200_ENOMEM -4
#define B200_ESTATE -5     /* e.g. pull before push, batch larger than max_batch */
#define B200_EUNSUPPORTED -6 /* option / signal condition this build does not cover (the text says which) */

/* constellation (module param "constellation", module_psk_demod.cpp:17-21; NONE = no Costas loop) */
enum { B200_BPSK = 0, B200_QPSK = 1, B200_OQPSK = 2, B200_8PSK = 3, B200_NONE = 4, B200_BPSK_90 = 5 };
/* baseband_format (common/dsp/io/baseband_type.h:7-21) */
enum { B200_CF32 = 0, B200_CS16 = 1, B200_CS8 = 2 };
/* decoder kind */
enum { B200_FEC_METOP = 0, B200_FEC_CCSDS = 1, B200_FEC_SIMPLE = 2 };
/* debug stage ids for b200_demod_debug_stage */
enum { B200_STAGE_AGC = 0, B200_STAGE_FIR = 1, B200_STAGE_COSTAS = 2, B200_STAGE_RESAMP = 3, B200_STAGE_DC = 4, B200_STAGE_MM = 5,
       B200_STAGE_PLL = 6 /* pm_demod: PLLCarrierTrackingBlock output */, B200_STAGE_PM = 7 /* pm_demod: PMToBPSK output */ };
/* mode bits of b200_demod_debug_run_stage */
#define B200_DEBUG_STRICT 1     /* the reference's operation order: separate multiply and add, left to right (generic VOLK, no FMA) */
#define B200_DEBUG_SEQUENTIAL 2 /* one segment: the feedback loop runs as ONE sequential thread from the initial state */

typedef struct b200_demod_cfg
{
    double samplerate;        /* "samplerate"  (module_demod_base.cpp:17-20)                     */
    double symbolrate;        /* "symbolrate"  (:26-27)                                          */
    int constellation;        /* "constellation"                                                 */
    float rrc_alpha;          /* "rrc_alpha"   (module_psk_demod.cpp:23-26)                      */
    int rrc_taps;             /* "rrc_taps", default 31 (module_psk_demod.h:31); only 31 built   */
    float pll_bw;             /* "pll_bw"                                                        */
    float agc_rate;           /* "agc_rate", default 1e-2 (module_demod_base.h:54)               */
    float clock_gain_omega;   /* default pow(8.7e-3,2)/4 (module_psk_demod.h:36)                 */
    float clock_mu;           /* default 0.5                                                     */
    float clock_gain_mu;      /* default 8.7e-3                                                  */
    float clock_omega_limit;  /* "clock_omega_relative_limit", default 0.005                     */
    float costas_max_offset;  /* rad/sample, default 1.0 (module_psk_demod.cpp:116-118)          */
    int format;               /* "baseband_format": B200_CF32 / CS16 / CS8                       */
    int device;               /* CUDA device ordinal                                             */
    long max_batch;           /* largest nsamples of one push (device buffers are sized for it)  */
    int keep_stages;          /* !=0: keep AGC/FIR/Costas stage outputs for b200_demod_debug_stage */
    int iq_swap;              /* "iq_swap": re <-> im at the reader (dsp/io/file_source.cpp:31-33)    */
    double final_samplerate;  /* 0 = samplerate. Otherwise the rate BaseDemodModule::initb resamples to when samplerate/symbolrate is
                                 outside [min_sps, max_sps] (module_demod_base.cpp:59-87,203-204): use b200_demod_final_samplerate().
                                 Only the rational part of SmartResamplerBlock is built: samplerate / final_samplerate must be < 2 */
    int dc_block;             /* "dc_block": CorrectIQBlock in front (utils/correct_iq.cpp:18-35, module_demod_base.cpp:113-114) */
    int post_costas_dc;       /* "post_costas_dc": CorrectIQBlock between the Costas loop and the clock recovery
                                 (module_psk_demod.cpp:127-134); not with OQPSK                                         */
    int front_resample;       /* 0: the resampler runs iff final_samplerate differs from samplerate. 2: never - BaseDemodModule::initb's
                                 `resample` is false (samplerate / symbolrate inside [min_sps, max_sps]) although "custom_samplerate" set
                                 another final_samplerate: the reference then only designs the RRC and the clock recovery for that rate
                                 (module_demod_base.cpp:66,73-74,203-204); b200_demod_resample_decision() tells                        */
    int clock_recovery;       /* 0: MMClockRecoveryBlock<complex_t> (what psk_demod builds, module_psk_demod.cpp:134-135);
                                 1: dsp::GardnerClockRecoveryBlock<complex_t> with the same arguments
                                 (common/dsp/clock_recovery/clock_recovery_gardner.cpp:33-131)                          */
    /* pm_demod (src-core/pipeline/modules/demod/module_pm_demod.cpp:12-88): AGC -> PLLCarrierTrackingBlock -> PMToBPSK ->
       [SmartResamplerBlock -> AGCBlock(0.001, 1, 1, 1000)] -> RRC -> Costas loop (order 2) -> M&M, soft = clamp(real * 100).
       The working-rate window is then [1.1, 10] samples per symbol (MAX_SPS = 10, :56) */
    int pm_demod;             /* != 0: build PMDemodModule's chain; constellation must be B200_BPSK, pll_bw above is its "costas_bw"  */
    float pm_pll_bw;          /* "pll_bw": carrier PLL loop bandwidth (required by the module, :21-24)                              */
    float pm_pll_max_offset;  /* "pll_max_offset", default 0.5 rad/sample (module_pm_demod.h:29)                                    */
    int pm_resample_after_pll;/* "resample_after_pll": the front-end resampler sits behind PMToBPSK, followed by the second AGC      */
    double pm_subcarrier_offset; /* "subcarrier_offset" Hz (uint64 in the module); 0 = the symbol rate (:67)                           */
    double freq_shift;        /* "freq_shift" Hz (long in the module): FreqShiftBlock behind the reader / DC blocker, in front of the
                                 resampler (module_demod_base.cpp:37,122-123; common/dsp/utils/freq_shift.cpp); 0 = none             */
    /* psk_demod "has_carrier" (module_psk_demod.cpp:39-40,93-116): BPSK on a residual carrier: RRC -> PLLCarrierTrackingBlock ->
       CorrectIQBlock -> Costas loop (whose frequency limit then defaults to 0.2 rad/sample: pass it in costas_max_offset) */
    int has_carrier;
    float carrier_pll_bw;         /* "carrier_pll_bw" (required by the module in this mode)   */
    float carrier_pll_max_offset; /* "carrier_pll_max_offset", default 3.14                    */
} b200_demod_cfg;

typedef struct b200_fec_cfg
{
    int kind;                 /* B200_FEC_METOP (metop_ahrpt_decoder) | B200_FEC_CCSDS (ccsds_conv_concat_decoder) |
                                 B200_FEC_SIMPLE (ccsds_simple_psk_decoder: no convolutional code,
                                 src-core/pipeline/modules/ccsds/module_ccsds_simple_psk_decoder.cpp)                */
    int constellation;        /* ccsds: B200_BPSK / BPSK_90 / QPSK / OQPSK; simple: B200_BPSK / QPSK */
    int cadu_size;            /* bits ("cadu_size"); METOP: 8192                                 */
    int outsync_after;        /* "viterbi_outsync_after"                                         */
    float ber_thresold;       /* "viterbi_ber_thresold"                                          */
    int nrzm, derandomize, derand_after_rs, derand_start;
    int rs_i, rs_dualbasis, rs_fill_bytes, rs_usecheck, rs_type;
    int iq_invert;
    unsigned int asm_sync;    /* "asm", default 0x1ACFFC1D                                       */
    int device;
    long max_soft;            /* largest number of soft bytes of one push                        */
    int qpsk_swap_iq;         /* simple: "qpsk_swap_iq"   (module_ccsds_simple_psk_decoder.cpp:27)    */
    int qpsk_swap_diff;       /* simple: "qpsk_swap_diff", default true (:28); used with nrzm on QPSK  */
    int oqpsk_delay;          /* simple: "oqpsk_delay" (:29)                                          */
    int conv_rate;            /* ccsds: "conv_rate" (module_ccsds_conv_concat_decoder.cpp:33,99-117): 0 = "1/2" (Viterbi1_2); 2, 3, 5, 7 = "2/3",
                                 "3/4", "5/6", "7/8" (Viterbi_Depunc, common/codings/viterbi/viterbi_punc.cpp + depunc.h)           */
} b200_fec_cfg;

typedef struct b200_demod_stats
{
    long samples_in, symbols_out;
    float agc_gain, costas_phase, costas_freq, mm_mu, mm_omega;
    long costas_unconverged;  /* junctions still inconsistent after the repair rounds of the last push (expected 0) */
    long mm_unconverged;      /* same for the M&M loop                                            */
    int agc_clamped;          /* batches so far in which the AGC hit max_gain (silent input): handled by the clamp pass */
    int repairs;              /* segments re-run as exact sequential continuations (junction check failed) */
    long kernel_launches;     /* CUDA kernels launched by this object so far                     */
    long agc_exact_passes;    /* batches whose AGC seeds needed the scanned (exact) pass: weak signal */
    long last_front_samples;  /* samples that entered the AGC in the last push (= nsamples unless the front-end resampler runs) */
    float snr, peak_snr;      /* M2M4SNREstimator over the recovered symbols, dB (module stats keys "snr" / "peak_snr", module_psk_demod.cpp:190-194,242-243):
                                 evaluated at the end of every push; the peak is the maximum of those */
    float pll_freq;           /* pm_demod: carrier PLL frequency, rad/sample (PLLCarrierTrackingBlock::getFreq; module stats key "freq" =
                                 rad_to_hz(pll_freq, final_samplerate), module_pm_demod.cpp:138,186) */
    long pll_unconverged;     /* pm_demod: carrier PLL junctions still inconsistent after the repair rounds of the last push (expected 0) */
} b200_demod_stats;

typedef struct b200_fec_stats
{
    long soft_in, chunks, bits_out, frames_out;
    int viterbi_state;        /* 0 NOSYNC 1 SYNCED  (Viterbi3_4::getState)                       */
    float viterbi_ber;
    int deframer_state;       /* 2 NOSYNC / 6 SYNCING / 12|18 SYNCED (BPSK_CCSDS_Deframer)       */
    long rs_corrected, rs_failed;
    long replays;             /* slow-path re-decodes (lock changes, start-state mis-speculation) */
    long kernel_launches;
    long start_redone;        /* chunks decoded again because their speculated Viterbi start state was wrong (part of replays) */
    long tb_serial;           /* chunks whose parallel chainback blocks disagreed and were chained back serially (part of replays) */
    int spec_steps;           /* current length of the start-state speculation window (adapts to the channel) */
    int tb_overlap;           /* current warm-up length of the parallel chainback blocks (adapts to the channel) */
} b200_fec_stats;

typedef struct b200_demod b200_demod;

/* BaseDemodModule::initb's choice of the working sample rate (module_demod_base.cpp:59-87): returns samplerate when
 * samplerate/symbolrate lies inside [min_sps, max_sps] (psk_demod: 1.1..4.0, OQPSK 1.6..2.4; pass 0 for those defaults), else the
 * rate the reference's front-end resampler converts to. custom_samplerate > 0 overrides ("custom_samplerate"). */
double b200_demod_final_samplerate(double samplerate, double symbolrate, int constellation, float min_sps, float max_sps, double custom_samplerate);
/* BaseDemodModule::initb's `resample` (module_demod_base.cpp:66): 1 when samplerate / symbolrate lies outside [min_sps, max_sps] */
int b200_demod_resample_decision(double samplerate, double symbolrate, int constellation, float min_sps, float max_sps);
/* The polyphase bank the front-end resampler uses for (samplerate -> final_samplerate): RationalResamplerBlock::set_ratio
 * (resamp/rational_resampler.cpp:27-41) = firdes::design_resampler_filter_float + PolyphaseBank::init. Host-only (no device):
 * out[arm * *ntaps + k], *interp arms, reduced ratio *interp / *decim. */
int b200_demod_resampler_bank(double samplerate, double final_samplerate, float *out, long cap, int *ntaps, int *interp, int *decim);
typedef struct b200_fec b200_fec;
typedef struct b200_chain b200_chain;

const char *b200_last_error(void);
int b200_device_count(void);

/* ---- demodulator: raw IQ -> soft symbols ---------------------------------------------------------- */
b200_demod *b200_demod_create(const b200_demod_cfg *cfg);
void b200_demod_destroy(b200_demod *d);
/* One batch of the stream from HOST memory (copied to the device inside the call). Loop state carries over. */
int b200_demod_push_iq(b200_demod *d, const void *host_iq, long nsamples);
/* Same, samples already in device memory of cfg.device (no copy). */
int b200_demod_push_iq_device(b200_demod *d, const void *dev_iq, long nsamples);
/* int8 soft symbols of the LAST push (I,Q interleaved; BPSK: I only) — the bytes PSKDemodModule writes to .soft / the FIFO */
int b200_demod_pull_soft(b200_demod *d, int8_t *host_out, long cap, long *n_out);
/* float symbols of the LAST push (M&M output, interleaved re,im) */
int b200_demod_pull_symbols(b200_demod *d, float *host_out, long cap_symbols, long *n_out);
/* stage outputs of the LAST push (needs keep_stages): nsamples complex values, interleaved re,im */
int b200_demod_debug_stage(b200_demod *d, int stage, float *host_out, long cap_samples);
/* Stage-isolated parity hook: runs ONE stage of a freshly reset demodulator on a caller-supplied stage input
 * (nsamples complex values, interleaved re,im: normally the ORACLE's output of the stage before) and returns the stage output.
 *   B200_STAGE_FIR     in = AGC output.  mode STRICT: fir.cpp:74-83 in the generic VOLK order -> bitwise the oracle's FIR output;
 *                      mode 0: the production kernel (k_agc_fir_w with the AGC switched off), one fma per tap
 *   B200_STAGE_COSTAS  in = FIR output; out = what the clock recovery reads (rotation fix-up, OQPSK delay, post_costas_dc applied)
 *   B200_STAGE_MM      in = the clock recovery's input; out = symbols (*n_out of them). STRICT | SEQUENTIAL -> bitwise the oracle's
 * SEQUENTIAL runs the loop as one segment (no warm-up, no stitching). The demodulator is reset before and after. */
int b200_demod_debug_run_stage(b200_demod *d, int stage, const float *host_in, long nsamples, int mode, float *host_out, long cap_complex, long *n_out);
/* Junction residuals of the LAST push / debug run: per segment s > 0 the Costas phase (mod 2pi/order) and frequency difference
 * between segment s's warmed-up state at its first owned sample and segment s-1's end state (costas_out[2s], [2s+1]), and the M&M
 * sampling-instant difference (mm_out[s], samples). Either pointer may be NULL. */
int b200_demod_debug_junctions(b200_demod *d, double *costas_out, double *mm_out, long cap_segments, long *nseg_out, long *seg_len);
/* test hook: only the sample conversion (BasebandReader::read_samples, baseband_interface.h:170-190) of `nsamples` host samples */
int b200_demod_debug_convert(b200_demod *d, const void *host_iq, long nsamples, float *host_out);
int b200_demod_get_stats(b200_demod *d, b200_demod_stats *out);
/* forget the stream (loop state, resampler counters) but keep every allocation: the next push starts a new stream */
int b200_demod_reset(b200_demod *d);
/* double buffering for host streams, as b200_chain_prefetch_iq: start the H2D copy of a FUTURE batch (pinned memory); the later
 * b200_demod_push_iq() with the same (pointer, nsamples) only waits for it */
int b200_demod_prefetch_iq(b200_demod *d, const void *host_iq, long nsamples);
/* elapsed device milliseconds of the last push (CUDA events on the demodulator's stream): [0] = sum, [1] = (front end +) AGC + FIR,
 * [2] = Costas + rotation (+ OQPSK delay), [3] = M&M + compaction / quantiser */
int b200_demod_last_timing(b200_demod *d, float *ms_out, int n);
/* RRC taps / M&M polyphase bank as designed on the host (for parity tests against firdes / PolyphaseBank) */
int b200_demod_get_taps(b200_demod *d, float *rrc_out, int rrc_cap, float *bank_out /* 128*8 or NULL */);

/* ---- decoder: soft symbols -> CADUs --------------------------------------------------------------- */
b200_fec *b200_fec_create(const b200_fec_cfg *cfg);
void b200_fec_destroy(b200_fec *f);
int b200_fec_push_soft(b200_fec *f, const int8_t *host_soft, long n);
int b200_fec_push_soft_device(b200_fec *f, const int8_t *dev_soft, long n);
/* frames completed by the pushes since the last pull, cadu_bytes each, in stream order */
int b200_fec_pull_frames(b200_fec *f, uint8_t *host_out, long cap, long *nbytes_out);
/* Viterbi output bits (1 bit per byte, after NRZ-M when enabled) of the LAST push, for stage tests */
int b200_fec_debug_bits(b200_fec *f, uint8_t *host_out, long cap, long *n_out);
int b200_fec_get_stats(b200_fec *f, b200_fec_stats *out);
int b200_fec_cadu_bytes(b200_fec *f);
int b200_fec_chunk_size(b200_fec *f);

/* ---- fused chain: raw IQ -> CADUs, soft stream stays in HBM ---------------------------------------- */
b200_chain *b200_chain_create(const b200_demod_cfg *dcfg, const b200_fec_cfg *fcfg);
void b200_chain_destroy(b200_chain *c);
int b200_chain_push_iq(b200_chain *c, const void *host_iq, long nsamples);
int b200_chain_push_iq_device(b200_chain *c, const void *dev_iq, long nsamples);
/* Double buffering for host streams: start the H2D copy of a FUTURE batch (pinned memory) on a copy stream; the later
 * b200_chain_push_iq() with the same (pointer, nsamples) only waits for it. At most two batches may be pending. */
int b200_chain_prefetch_iq(b200_chain *c, const void *host_iq, long nsamples);
int b200_chain_pull_frames(b200_chain *c, uint8_t *host_out, long cap, long *nbytes_out);
/* device-resident result access (no D2H): pointer to the frames produced since the last pull/reset */
int b200_chain_frames_device(b200_chain *c, const uint8_t **dev_ptr, long *nbytes);
int b200_chain_get_stats(b200_chain *c, b200_demod_stats *ds, b200_fec_stats *fs);
/* timing taps used by bench.py: elapsed device milliseconds of the last push per stage (CUDA events on the chain's stream) */
int b200_chain_last_timing(b200_chain *c, float *ms_out, int n); /* n >= 9: [0]=sum of stages [1]=agc+fir [2]=costas(+rotate) [3]=m&m [4]=viterbi stage [5]=deframe+rs [6]=k_vit_main alone [7]=chunks it decoded [8]=whole push, event timed */
/* forget the stream (loop state, lock, FIFOs) but keep every allocation: the next push starts a new stream */
int b200_chain_reset(b200_chain *c);
/* Pipelined mode: the decoder runs on a worker thread and its own CUDA stream one batch behind the demodulator, like the
 * reference's one-thread-per-module pipeline (src-core/pipeline/pipeline_run.cpp:44-117) with a whole batch of soft symbols in HBM
 * as the FIFO element. push(i) returns when batch i is demodulated and batch i-1 decoded; b200_chain_pull_frames() then returns
 * the frames of the batches decoded so far without waiting for the one in flight; b200_chain_sync() waits for the decoder
 * (after it, pull returns everything). Decoder errors surface at the next push / sync. Output is identical to the synchronous mode. */
int b200_chain_set_pipelined(b200_chain *c, int on);
int b200_chain_sync(b200_chain *c);
/* CUDA-event stopwatch over several pushes (both streams): begin syncs the chain and stamps the demodulator's stream, end syncs
 * and stamps the decoder's stream; *ms = elapsed device time */
int b200_chain_span_begin(b200_chain *c);
int b200_chain_span_end(b200_chain *c, float *ms);

/* ---- CADU -> CCSDS space packets (the step behind the decoder) --------------------------------------------
 * Replaces, for all selected virtual channels at once, what the instrument modules do per CADU on the host
 * (plugins/noaa_metop_support/metop/module_metop_instruments.cpp:66-140): ccsds::ccsds_aos::parseVCDU(cadu).vcid
 * (src-core/common/ccsds/ccsds_aos/vcdu.cpp:10-17) selects one ccsds::ccsds_aos::Demuxer per channel, whose work(cadu)
 * (ccsds_aos/demuxer.cpp:64-199; M-PDU header mpdu.cpp:9-13) returns the space packets completed by that frame. The packets come
 * back in exactly that order - frame by frame, inside a frame in the order work() returns them - with exactly the reference's bytes,
 * including its behaviour on inconsistent frames. */
typedef struct b200_demux_cfg
{
    int cadu_size;                /* bytes per CADU as the decoder writes them (1024 for RS interleaving 4)                          */
    int mpdu_data_size;           /* Demuxer(mpdu_data_size, ...), demuxer.h:33 (default 884; MetOp: 882 with a 2-byte insert zone)  */
    int has_insert_zone;          /* Demuxer(.., hasInsertZone, insertZoneSize, ..)                                                  */
    int insert_zone_size;
    int secondary_header_extends; /* Demuxer(.., secondaryHeaderExtendsPkt): + 8 payload bytes when the secondary header flag is set */
    unsigned long long vcid_mask; /* bit v set = virtual channel v is demultiplexed (the modules' `if (vcdu.vcid == N)` chains)      */
    int device;
    long max_frames;              /* largest number of CADUs of one push                                                             */
    long max_packets;             /* room of the packet table of one push; 0 = 8 * max_frames + 1024                                 */
} b200_demux_cfg;
typedef struct b200_packet
{
    long offset;      /* of the packet's 6 header bytes (CCSDSHeader::raw) in the byte stream of the pull; payload behind them      */
    int payload_len;  /* CCSDSPacket::payload.size()                                                                                */
    int frame;        /* stream index (since create / reset) of the CADU whose work() call returned the packet                      */
    short vcid, apid; /* parseVCDU(cadu).vcid; CCSDSHeader::apid                                                                    */
} b200_packet;
typedef struct b200_demux_stats
{
    long frames_in, packets_out, kernel_launches;
    long redone_channels; /* virtual channels walked a second time, serially, because leftover bytes of an unfinished packet crossed a window
                             boundary of the parallel walk (inconsistent frames only; expected 0) */
} b200_demux_stats;
typedef struct b200_demuxer b200_demuxer;
b200_demuxer *b200_demux_create(const b200_demux_cfg *cfg);
void b200_demux_destroy(b200_demuxer *d);
/* nframes CADUs of cadu_size bytes from host memory / already on the device (e.g. b200_chain_frames_device) */
int b200_demux_push_frames(b200_demuxer *d, const uint8_t *host_cadus, long nframes);
int b200_demux_push_frames_device(b200_demuxer *d, const uint8_t *dev_cadus, long nframes);
/* packets of the LAST push: `bytes` = [6 header bytes][payload] back to back, `packets` = one record each */
int b200_demux_pull(b200_demuxer *d, uint8_t *bytes, long cap_bytes, long *nbytes, b200_packet *packets, long cap_packets, long *npackets);
int b200_demux_reset(b200_demuxer *d);
int b200_demux_get_stats(b200_demuxer *d, b200_demux_stats *out);

#ifdef __cplusplus
}
#endif
#endif

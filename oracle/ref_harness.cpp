/*
 * TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
 *
 * oracle/_ref harness: wires the UNMODIFIED reference translation units (compiled from
 * /root/reference by oracle/Makefile, with oracle/shim standing in for VOLK + logger) exactly
 * like the reference modules do, and exposes them through a small C ABI for the Python tests
 * and for bench.py's cpu_baseline / --impl reference arm.
 *
 * Wiring mirrors (reference file:line):
 *   BaseDemodModule::initb            src-core/pipeline/modules/demod/module_demod_base.cpp:59-208
 *   PSKDemodModule::init / process    src-core/pipeline/modules/demod/module_psk_demod.cpp:86-236
 *   PMDemodModule::init / process     src-core/pipeline/modules/demod/module_pm_demod.cpp:61-160
 *   DVBS2DemodModule front half       plugins/dvb_support/dvbs2/module_dvbs2_demod.cpp:98-102
 *   MetOpAHRPTDecoderModule::process  plugins/noaa_metop_support/metop/module_metop_ahrpt_decoder.cpp:34-90
 *   CCSDSConvConcatDecoderModule      src-core/pipeline/modules/ccsds/module_ccsds_conv_concat_decoder.cpp:16-200
 *   Pipeline::run two-module mode     src-core/pipeline/pipeline_run.cpp:44-117
 *   CADU -> CCSDS packets             plugins/noaa_metop_support/metop/module_metop_instruments.cpp:66-140 (parseVCDU + one Demuxer per VCID)
 *
 * No reference source is copied here: this file only *calls* the reference classes.
 */
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <volk/volk.h>
#include "logger.h"
/* work() and the loop state are private/protected in the block classes; the harness drives the
 * blocks synchronously and reads their state, so it opens the access specifiers for ITS view of
 * the headers only (the reference .cpp files are compiled untouched; GCC layout is unaffected). */
#define protected public
#define private public
#include "common/dsp/block.h"
#include "common/dsp/utils/agc.h"
#include "common/dsp/utils/correct_iq.h"
#include "common/dsp/utils/snr_estimator.h"
#include "common/dsp/filter/fir.h"
#include "common/dsp/filter/firdes.h"
#include "common/dsp/pll/costas_loop.h"
#include "common/dsp/demod/delay_one_imag.h"
#include "common/dsp/clock_recovery/clock_recovery_mm.h"
#include "common/dsp/clock_recovery/clock_recovery_gardner.h"
#include "common/dsp/resamp/smart_resampler.h"
#include "common/dsp/resamp/rational_resampler.h"
#include "common/dsp/pll/pll_carrier_tracking.h"
#include "common/dsp/demod/pm_to_bpsk.h"
#include "common/dsp/utils/freq_shift.h"
#include "common/dsp/utils/fast_trig.h"
#undef private
#undef protected
#include "common/dsp/resamp/polyphase_bank.h"
#include "common/dsp/window/window.h"
#include "common/codings/viterbi/viterbi_3_4.h"
#include "common/codings/viterbi/viterbi_1_2.h"
#define private public /* Viterbi_Depunc reads member buffers it never initialises (viterbi_punc.h:52-56): the harness zeroes them */
#include "common/codings/viterbi/viterbi_punc.h"
#undef private
#include "common/codings/viterbi/cc_encoder.h"
#include "common/codings/viterbi/cc_decoder.h"
#include "common/codings/deframing/bpsk_ccsds_deframer.h"
#include "common/codings/randomization.h"
#include "common/codings/reedsolomon/reedsolomon.h"
#include "common/codings/differential/nrzm.h"
#include "common/codings/differential/qpsk_diff.h"
#include "common/dsp/demod/constellation.h"
#include "common/codings/rotation.h"
#include "common/ccsds/ccsds_aos/demuxer.h"
#include "common/ccsds/ccsds_aos/vcdu.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

std::shared_ptr<slog::Logger> logger = std::make_shared<slog::Logger>();

extern "C"
{
    /* ------------------------------------------------------------------ configuration */
    typedef struct
    {
        double samplerate;       /* Hz (module param "samplerate", parsed as long)            */
        double symbolrate;       /* Hz (module param "symbolrate", parsed as long)            */
        int constellation;       /* 0 bpsk, 1 qpsk, 2 oqpsk, 3 8psk, 4 = no Costas (DVB-S2 front half) */
        float rrc_alpha;
        int rrc_taps;            /* default 31 */
        float pll_bw;
        float agc_rate;          /* default 1e-2 */
        float clock_gain_omega;  /* default pow(8.7e-3,2)/4 */
        float clock_mu;          /* default 0.5 */
        float clock_gain_mu;     /* default 8.7e-3 */
        float clock_omega_limit; /* default 0.005 */
        float costas_max_offset; /* rad/sample, default 1.0 */
        int format;              /* 0 cf32, 1 cs16, 2 cs8 */
        int buffer_size;         /* 0 = reference default rule */
        int iq_swap;             /* FileSourceBlock(.., iq_swap): re <-> im at the reader (file_source.cpp:31-33) */
        double final_samplerate; /* 0 = samplerate (sps inside [MIN_SPS, MAX_SPS]); else BaseDemodModule::initb's resampled rate */
        int dc_block;            /* CorrectIQBlock behind the reader (module_demod_base.cpp:113-114) */
        int post_costas_dc;      /* CorrectIQBlock behind the Costas loop (module_psk_demod.cpp:127-134) */
        int clock_recovery;      /* 0: MMClockRecoveryBlock (psk_demod); 1: dsp::GardnerClockRecoveryBlock<complex_t> with the same arguments
                                    (common/dsp/clock_recovery/clock_recovery_gardner.cpp) */
        /* PMDemodModule (module_pm_demod.cpp:61-88): AGC -> PLLCarrierTrackingBlock -> PMToBPSK -> [SmartResampler -> AGC2] -> RRC ->
           Costas(order 2, default frequency limit) -> M&M. `pll_bw` above is then "costas_bw"; constellation must be 0 (bpsk) */
        int pm;
        float pm_pll_bw;          /* "pll_bw" */
        float pm_pll_max_offset;  /* "pll_max_offset" (module_pm_demod.h default) */
        int pm_resample_after_pll;
        double pm_subcarrier_offset; /* "subcarrier_offset": 0 = the symbol rate (module_pm_demod.cpp:67) */
        double freq_shift;        /* "freq_shift": FreqShiftBlock behind the reader / DC blocker (module_demod_base.cpp:125-126); 0 = none */
        /* psk_demod "has_carrier" (module_psk_demod.cpp:93-113): RRC -> PLLCarrierTrackingBlock -> CorrectIQBlock -> Costas (BPSK, limit 0.2) */
        int has_carrier;
        float carrier_pll_bw, carrier_pll_max_offset;
    } ref_demod_cfg;

    typedef struct
    {
        int kind;           /* 0 = metop_ahrpt_decoder, 1 = ccsds_conv_concat_decoder (r=1/2), 2 = ccsds_simple_psk_decoder */
        int constellation;  /* ccsds: 0 bpsk, 1 qpsk, 2 oqpsk, 5 bpsk_90 */
        int cadu_size;      /* bits */
        int outsync_after;
        float ber_thresold;
        int nrzm, derandomize, derand_after_rs, derand_start;
        int rs_i, rs_dualbasis, rs_fill_bytes, rs_usecheck, rs_type; /* rs_type 0 rs223 1 rs239 */
        int iq_invert;
        unsigned int asm_sync;
        /* kind 2 = ccsds_simple_psk_decoder (module_ccsds_simple_psk_decoder.cpp): no convolutional code */
        int qpsk_swap_iq, qpsk_swap_diff, oqpsk_delay;
        int conv_rate; /* kind 1: 0 = "1/2" (Viterbi1_2); 2 / 3 / 5 / 7 = "2/3" / "3/4" / "5/6" / "7/8" (Viterbi_Depunc, module_ccsds_conv_concat_decoder.cpp:99-117) */
    } ref_fec_cfg;
}

namespace
{
    struct RefDemod
    {
        ref_demod_cfg cfg;
        int buffer_size;
        float final_samplerate, final_sps;
        std::shared_ptr<dsp::stream<complex_t>> in;
        std::shared_ptr<dsp::SmartResamplerBlock<complex_t>> resampler; /* module_demod_base.cpp:203-204 */
        std::shared_ptr<dsp::stream<complex_t>> dc_in;
        std::shared_ptr<dsp::CorrectIQBlock<complex_t>> dc; /* module_demod_base.cpp:113-114 */
        std::vector<complex_t> rs_in;
        std::shared_ptr<dsp::AGCBlock<complex_t>> agc;
        std::shared_ptr<dsp::FIRBlock<complex_t>> rrc;
        std::shared_ptr<dsp::CostasLoopBlock> pll;
        std::shared_ptr<dsp::PLLCarrierTrackingBlock> carrier_pll; /* has_carrier */
        std::shared_ptr<dsp::CorrectIQBlock<complex_t>> carrier_dc;
        std::shared_ptr<dsp::CorrectIQBlock<complex_t>> post_pll_dc;
        std::shared_ptr<dsp::DelayOneImagBlock> delay;
        std::shared_ptr<dsp::Block<complex_t, complex_t>> rec; /* the clock recovery in use: one of the two below */
        std::shared_ptr<dsp::MMClockRecoveryBlock<complex_t>> rec_mm;
        std::shared_ptr<dsp::GardnerClockRecoveryBlock<complex_t>> rec_g;
        std::vector<float> rrc_taps;
        long last_front = 0; /* samples that entered the AGC in the last ref_demod_run call */
        /* pm_demod */
        std::shared_ptr<dsp::PLLCarrierTrackingBlock> cpll;
        std::shared_ptr<dsp::PMToBPSK> pm_psk;
        std::shared_ptr<dsp::SmartResamplerBlock<complex_t>> resampler2;
        std::shared_ptr<dsp::stream<complex_t>> mid;
        std::shared_ptr<dsp::AGCBlock<complex_t>> agc2;
        long pm_pos = 0;
        float *pll_dump = nullptr, *pm_dump = nullptr; /* optional dumps of the PLL / PMToBPSK outputs (input-rate sample positions) */
        /* freq_shift */
        std::shared_ptr<dsp::stream<complex_t>> fs_in;
        std::shared_ptr<dsp::FreqShiftBlock> fshift;
    };

    int8_t soft_clamp(float x) /* module_demod_base.h:106-113 semantics */
    {
        if (x < -128.0)
            return -127;
        if (x > 127.0)
            return 127;
        return x;
    }

    void convert_block(const ref_demod_cfg &cfg, const void *raw, long off, int n, complex_t *dst)
    {
        /* baseband_interface.h:170-190 */
        if (cfg.format == 0)
            memcpy(dst, (const complex_t *)raw + off, n * sizeof(complex_t));
        else if (cfg.format == 1)
            volk_16i_s32f_convert_32f_u((float *)dst, (const int16_t *)raw + off * 2, 32767, n * 2);
        else
            volk_8i_s32f_convert_32f_u((float *)dst, (const int8_t *)raw + off * 2, 127, n * 2);
        if (cfg.iq_swap) /* file_source.cpp:31-33 */
            for (int i = 0; i < n; i++)
                dst[i] = complex_t(dst[i].imag, dst[i].real);
    }

    /* reader (+ iq_swap) and the optional DC blocker: n samples into dst */
    void front_block(RefDemod *d, const void *raw, long off, int n, complex_t *dst)
    {
        if (!d->dc)
            convert_block(d->cfg, raw, off, n, dst);
        else
        {
            convert_block(d->cfg, raw, off, n, d->dc_in->writeBuf);
            d->dc_in->swap(n);
            d->dc->work();
            memcpy(dst, d->dc->output_stream->readBuf, n * sizeof(complex_t));
            d->dc->output_stream->flush();
        }
        if (d->fshift) /* module_demod_base.cpp:125-126: behind the DC blocker, in front of the resampler */
        {
            memcpy(d->fs_in->writeBuf, dst, n * sizeof(complex_t));
            d->fs_in->swap(n);
            d->fshift->work();
            memcpy(dst, d->fshift->output_stream->readBuf, n * sizeof(complex_t));
            d->fshift->output_stream->flush();
        }
    }

    struct RefFec
    {
        ref_fec_cfg cfg;
        int buffer_size, cadu_bytes;
        std::shared_ptr<viterbi::Viterbi3_4> v34;
        std::shared_ptr<viterbi::Viterbi1_2> v12;
        std::shared_ptr<viterbi::Viterbi_Depunc> vp;
        std::shared_ptr<deframing::BPSK_CCSDS_Deframer> deframer, deframer_qpsk;
        std::shared_ptr<reedsolomon::ReedSolomon> rs;
        diff::NRZMDiff diff;
        diff::QPSKDiff qpsk_diff;
        int8_t last_q_oqpsk = 0;
        std::vector<uint8_t> bits_buf, qpsk_diff_buffer;
        int errors[16];
        int noSyncsRuns = 0;
        std::vector<uint8_t> viterbi_out, frame_buffer;
        std::vector<int8_t> soft;
    };
}

extern "C"
{
    // M2M4SNREstimator (snr_estimator.cpp) over a symbol stream fed in `chunk`-symbol updates as PSKDemodModule::process does
    // (module_psk_demod.cpp:190-194): out[0] = snr() after the last update, out[1] = the peak over all updates, out[2], out[3] = y1, y2
    void ref_snr_m2m4(const float *syms, long n, long chunk, float *out)
    {
        M2M4SNREstimator est;
        float snr = 0, peak = 0;
        for (long pos = 0; pos < n; pos += chunk)
        {
            long m = std::min(chunk, n - pos);
            est.update((complex_t *)syms + pos, (int)m);
            snr = est.snr();
            if (snr > peak)
                peak = snr;
        }
        out[0] = snr;
        out[1] = peak;
    }

    void *ref_demod_create(const ref_demod_cfg *c)
    {
        RefDemod *d = new RefDemod();
        d->cfg = *c;
        long samplerate = (long)c->samplerate, symbolrate = (long)c->symbolrate;
        d->buffer_size = c->buffer_size > 0 ? c->buffer_size : std::min<int>(dsp::STREAM_BUFFER_SIZE, std::max<int>(8192 + 1, samplerate / 200));
        d->final_samplerate = c->final_samplerate > 0 ? (float)c->final_samplerate : (float)samplerate;
        d->final_sps = d->final_samplerate / (float)symbolrate;
        d->in = std::make_shared<dsp::stream<complex_t>>();
        if (c->dc_block)
        {
            d->dc_in = std::make_shared<dsp::stream<complex_t>>();
            d->dc = std::make_shared<dsp::CorrectIQBlock<complex_t>>(d->dc_in);
        }
        if (c->freq_shift != 0)
        {
            d->fs_in = std::make_shared<dsp::stream<complex_t>>();
            d->fshift = std::make_shared<dsp::FreqShiftBlock>(d->fs_in, (double)samplerate, c->freq_shift);
        }
        const bool resample = c->final_samplerate > 0 && (long)c->final_samplerate != samplerate;
        if (resample)
        {
            /* module_demod_base.cpp:84-87,203-204: buffer scaled by ceil(decimation factor), resampler (final, input) */
            float decimation_factor = samplerate / d->final_samplerate;
            d->buffer_size *= ceil(decimation_factor);
            if (d->buffer_size > 8192 * 20)
                d->buffer_size = 8192 * 20;
            if (!(c->pm && c->pm_resample_after_pll)) /* initb(!d_resample_after_pll), module_pm_demod.cpp:63 */
                d->resampler = std::make_shared<dsp::SmartResamplerBlock<complex_t>>(nullptr, d->final_samplerate, samplerate);
            d->rs_in.resize(d->buffer_size);
        }
        d->agc = std::make_shared<dsp::AGCBlock<complex_t>>(d->in, c->agc_rate, 1.0f, 1.0f, 65536);
        d->rrc_taps = dsp::firdes::root_raised_cosine(1, d->final_samplerate, (int)symbolrate, c->rrc_alpha, c->rrc_taps);
        std::shared_ptr<dsp::stream<complex_t>> rrc_in = d->agc->output_stream;
        if (c->pm)
        {
            /* module_pm_demod.cpp:65-80 */
            d->cpll = std::make_shared<dsp::PLLCarrierTrackingBlock>(d->agc->output_stream, c->pm_pll_bw, c->pm_pll_max_offset, -c->pm_pll_max_offset);
            d->pm_psk = std::make_shared<dsp::PMToBPSK>(d->cpll->output_stream, c->pm_resample_after_pll ? (float)samplerate : d->final_samplerate,
                                                       (unsigned long)c->pm_subcarrier_offset == 0 ? (float)symbolrate : (float)(unsigned long)c->pm_subcarrier_offset);
            rrc_in = d->pm_psk->output_stream;
            if (c->pm_resample_after_pll)
            {
                d->resampler2 = std::make_shared<dsp::SmartResamplerBlock<complex_t>>(nullptr, d->final_samplerate, samplerate);
                d->mid = std::make_shared<dsp::stream<complex_t>>();
                d->agc2 = std::make_shared<dsp::AGCBlock<complex_t>>(d->mid, 0.001, 1.0, 1.0, 1000.0);
                rrc_in = d->agc2->output_stream;
            }
        }
        d->rrc = std::make_shared<dsp::FIRBlock<complex_t>>(rrc_in, d->rrc_taps);
        std::shared_ptr<dsp::stream<complex_t>> last = d->rrc->output_stream;
        if (c->pm) /* module_pm_demod.cpp:84: CostasLoopBlock(rrc->output_stream, d_loop_bw, 2) */
        {
            d->pll = std::make_shared<dsp::CostasLoopBlock>(last, c->pll_bw, 2);
            last = d->pll->output_stream;
        }
        else if (c->constellation != 4)
        {
            int order = c->constellation == 0 ? 2 : (c->constellation == 3 ? 8 : 4);
            if (c->has_carrier) /* module_psk_demod.cpp:93-113 (BPSK only; the Costas limit is then the caller's 0.2 default, :116) */
            {
                d->carrier_pll = std::make_shared<dsp::PLLCarrierTrackingBlock>(last, c->carrier_pll_bw, c->carrier_pll_max_offset, -c->carrier_pll_max_offset);
                d->carrier_dc = std::make_shared<dsp::CorrectIQBlock<complex_t>>(d->carrier_pll->output_stream);
                last = d->carrier_dc->output_stream;
            }
            d->pll = std::make_shared<dsp::CostasLoopBlock>(last, c->pll_bw, order, c->costas_max_offset);
            last = d->pll->output_stream;
            if (c->post_costas_dc)
            {
                d->post_pll_dc = std::make_shared<dsp::CorrectIQBlock<complex_t>>(last);
                last = d->post_pll_dc->output_stream;
            }
            if (c->constellation == 2)
            {
                d->delay = std::make_shared<dsp::DelayOneImagBlock>(last);
                last = d->delay->output_stream;
            }
        }
        if (c->clock_recovery == 1)
        {
            d->rec_g = std::make_shared<dsp::GardnerClockRecoveryBlock<complex_t>>(last, d->final_sps, c->clock_gain_omega, c->clock_mu, c->clock_gain_mu, c->clock_omega_limit);
            d->rec = d->rec_g;
        }
        else
        {
            d->rec_mm = std::make_shared<dsp::MMClockRecoveryBlock<complex_t>>(last, d->final_sps, c->clock_gain_omega, c->clock_mu, c->clock_gain_mu, c->clock_omega_limit);
            d->rec = d->rec_mm;
        }
        return d;
    }

    void ref_demod_destroy(void *h) { delete (RefDemod *)h; }
    int ref_demod_buffer_size(void *h) { return ((RefDemod *)h)->buffer_size; }
    float ref_demod_sps(void *h) { return ((RefDemod *)h)->final_sps; }

    int ref_demod_rrc_taps(void *h, float *out)
    {
        RefDemod *d = (RefDemod *)h;
        memcpy(out, d->rrc_taps.data(), d->rrc_taps.size() * sizeof(float));
        return (int)d->rrc_taps.size();
    }

    /* Polyphase interpolator bank of the M&M block: out[arm*8 + k] (clock_recovery_mm.cpp:18, polyphase_bank.cpp:6-39) */
    int ref_mm_taps(float *out)
    {
        dsp::PolyphaseBank pfb;
        pfb.init(dsp::windowed_sinc(128 * 8, dsp::hz_to_rad(0.5 / 128.0, 1.0), dsp::window::nuttall, 128), 128);
        for (int a = 0; a < pfb.nfilt; a++)
            for (int k = 0; k < pfb.ntaps; k++)
                out[a * pfb.ntaps + k] = pfb.taps[a][k];
        return pfb.nfilt * pfb.ntaps;
    }

    int ref_rrc_design(double gain, double fs, double rs, double alpha, int ntaps, float *out)
    {
        std::vector<float> t = dsp::firdes::root_raised_cosine(gain, fs, rs, alpha, ntaps);
        memcpy(out, t.data(), t.size() * sizeof(float));
        return (int)t.size();
    }

    /*
     * Synchronous run over `nsamples` raw samples, in reference-sized buffers. Any of the dump
     * pointers may be NULL. agc/fir/costas dumps hold nsamples complex values (costas = after the
     * OQPSK delay when present); mm_out / soft_out hold the recovered symbols / int8 soft stream.
     * Returns the number of symbols produced. State persists across calls (streaming).
     */
    long ref_demod_run(void *h, const void *raw, long nsamples, float *agc_out, float *fir_out, float *costas_out, float *mm_out, int8_t *soft_out,
                       long sym_cap)
    {
        RefDemod *d = (RefDemod *)h;
        long nsym = 0, pos = 0; /* pos: samples after the (optional) resampler so far in this call */
        d->pm_pos = 0;          /* pm_demod: samples through the carrier PLL so far in this call (the AGC / PLL / PMToBPSK dumps' position) */
        for (long off = 0; off < nsamples; off += d->buffer_size)
        {
            int n = (int)std::min<long>(d->buffer_size, nsamples - off);
            if (d->resampler)
            {
                front_block(d, raw, off, n, d->rs_in.data());
                n = d->resampler->process(d->rs_in.data(), n, d->in->writeBuf);
                if (n <= 0)
                    continue;
            }
            else
                front_block(d, raw, off, n, d->in->writeBuf);
            d->in->swap(n);
            d->agc->work();
            if (agc_out)
                memcpy(agc_out + (d->cpll ? d->pm_pos : pos) * 2, d->agc->output_stream->readBuf, n * sizeof(complex_t));
            if (d->cpll)
            {
                d->cpll->work();
                if (d->pll_dump)
                    memcpy(d->pll_dump + d->pm_pos * 2, d->cpll->output_stream->readBuf, n * sizeof(complex_t));
                d->pm_psk->work();
                if (d->pm_dump)
                    memcpy(d->pm_dump + d->pm_pos * 2, d->pm_psk->output_stream->readBuf, n * sizeof(complex_t));
                d->pm_pos += n;
                if (d->resampler2)
                {
                    int m = d->resampler2->process(d->pm_psk->output_stream->readBuf, n, d->mid->writeBuf);
                    d->pm_psk->output_stream->flush();
                    n = m;
                    if (n <= 0)
                        continue;
                    d->mid->swap(n);
                    d->agc2->work();
                }
            }
            d->rrc->work();
            if (fir_out)
                memcpy(fir_out + pos * 2, d->rrc->output_stream->readBuf, n * sizeof(complex_t));
            if (d->pll)
            {
                if (d->carrier_pll)
                {
                    d->carrier_pll->work();
                    if (d->pll_dump)
                        memcpy(d->pll_dump + pos * 2, d->carrier_pll->output_stream->readBuf, n * sizeof(complex_t));
                    d->carrier_dc->work();
                    if (d->pm_dump)
                        memcpy(d->pm_dump + pos * 2, d->carrier_dc->output_stream->readBuf, n * sizeof(complex_t));
                }
                d->pll->work();
                if (d->post_pll_dc)
                    d->post_pll_dc->work();
                if (d->delay)
                    d->delay->work();
                if (costas_out) /* what the clock recovery reads */
                    memcpy(costas_out + pos * 2,
                           (d->delay ? d->delay->output_stream : (d->post_pll_dc ? d->post_pll_dc->output_stream : d->pll->output_stream))->readBuf,
                           n * sizeof(complex_t));
            }
            pos += n;
            d->rec->work();
            int m = d->rec->output_stream->getDataSize();
            complex_t *sym = d->rec->output_stream->readBuf;
            if (nsym + m > sym_cap)
                m = (int)(sym_cap - nsym);
            if (mm_out)
                memcpy(mm_out + nsym * 2, sym, m * sizeof(complex_t));
            if (soft_out)
            {
                if (d->cfg.pm) /* module_pm_demod.cpp:141-144 */
                    for (int i = 0; i < m; i++)
                        soft_out[nsym + i] = soft_clamp(sym[i].real * 100);
                else if (d->cfg.constellation == 0) /* module_psk_demod.cpp:199-205 */
                    for (int i = 0; i < m; i++)
                        soft_out[nsym + i] = soft_clamp(sym[i].real * 50);
                else /* module_psk_demod.cpp:207-213 */
                    for (int i = 0; i < m; i++)
                    {
                        soft_out[(nsym + i) * 2] = soft_clamp(sym[i].real * 100);
                        soft_out[(nsym + i) * 2 + 1] = soft_clamp(sym[i].imag * 100);
                    }
            }
            d->rec->output_stream->flush();
            nsym += m;
        }
        d->last_front = pos;
        return nsym;
    }
    long ref_demod_last_front(void *h) { return ((RefDemod *)h)->last_front; }
    /* pm_demod: where the next ref_demod_run calls dump the PLLCarrierTrackingBlock / PMToBPSK outputs (position restarts at 0) */
    void ref_demod_pm_dumps(void *h, float *pll_out, float *pm_out)
    {
        RefDemod *d = (RefDemod *)h;
        d->pll_dump = pll_out;
        d->pm_dump = pm_out;
    }
    /* pm_demod loop state: out[0..1] = carrier PLL phase / frequency (pll_carrier_tracking.h: d_phase, d_freq), out[2] = AGC2 gain */
    void ref_demod_pm_state(void *h, float *out4)
    {
        RefDemod *d = (RefDemod *)h;
        out4[0] = d->cpll ? d->cpll->d_phase : (d->carrier_pll ? d->carrier_pll->d_phase : 0);
        out4[1] = d->cpll ? d->cpll->d_freq : (d->carrier_pll ? d->carrier_pll->d_freq : 0);
        out4[2] = d->agc2 ? d->agc2->gain : 0;
        out4[3] = 0;
    }
    float ref_fast_atan2f(float y, float x) { return dsp::fast_atan2f(y, x); }
    float ref_fast_cos(float x) { return dsp::fast_cos(x); }
    float ref_fast_sin(float x) { return dsp::fast_sin(x); }
    /* the VOLK rotator as the shim restates it (and as FreqShiftBlock / PMToBPSK call it), on `call`-sample calls */
    void ref_rotator(const float *in, long n, long call, double inc_re, double inc_im, float *out)
    {
        lv_32fc_t phase(1, 0), inc((float)inc_re, (float)inc_im);
        for (long pos = 0; pos < n; pos += call)
        {
            long m = std::min(call, n - pos);
            volk_32fc_s32fc_x2_rotator_32fc((lv_32fc_t *)out + pos, (const lv_32fc_t *)in + pos, inc, &phase, (unsigned)m);
        }
    }

    /* The front end alone: conversion (+ iq_swap) and SmartResamplerBlock, in reference-sized buffers on a fresh resampler.
       Returns the number of output samples (<= cap). */
    long ref_resample(const ref_demod_cfg *c, const void *raw, long nsamples, float *out, long cap)
    {
        RefDemod *d = (RefDemod *)ref_demod_create(c);
        std::vector<complex_t> tmp(d->buffer_size * 2 + 16);
        std::vector<complex_t> in(d->buffer_size);
        long pos = 0;
        for (long off = 0; off < nsamples; off += d->buffer_size)
        {
            int n = (int)std::min<long>(d->buffer_size, nsamples - off);
            front_block(d, raw, off, n, in.data());
            int m = n;
            if (d->resampler)
                m = d->resampler->process(in.data(), n, tmp.data());
            else
                memcpy(tmp.data(), in.data(), n * sizeof(complex_t));
            if (pos + m > cap)
                m = (int)(cap - pos);
            memcpy(out + pos * 2, tmp.data(), m * sizeof(complex_t));
            pos += m;
        }
        delete d;
        return pos;
    }

    /* Polyphase bank of the rational resampler for (interpolation, decimation) as RationalResamplerBlock::set_ratio builds it
       (rational_resampler.cpp:27-41): returns ntaps per arm, *nfilt = arms after the gcd reduction; out[arm*ntaps + k] */
    int ref_resampler_taps(unsigned interpolation, unsigned decimation, float *out, int cap, int *nfilt)
    {
        dsp::RationalResamplerBlock<complex_t> r(nullptr, interpolation, decimation);
        *nfilt = r.pfb.nfilt;
        if (r.pfb.nfilt * r.pfb.ntaps > cap)
            return -1;
        for (int a = 0; a < r.pfb.nfilt; a++)
            for (int k = 0; k < r.pfb.ntaps; k++)
                out[a * r.pfb.ntaps + k] = r.pfb.taps[a][k];
        return r.pfb.ntaps;
    }

    /* ONE block of a fresh chain on a caller-supplied cf32 input, in reference-sized buffers: stage 1 = FIRBlock, 2 = CostasLoopBlock
       (+ post-Costas CorrectIQBlock / DelayOneImagBlock: what the clock recovery reads), 5 = MMClockRecoveryBlock. The input is written
       straight into the stream the block reads (stage 0 = AGCBlock on converted samples). Returns the number of complex outputs (<= cap). */
    long ref_demod_run_stage(void *h, int stage, const float *in, long nsamples, float *out, long cap)
    {
        RefDemod *d = (RefDemod *)h;
        long pos = 0;
        for (long off = 0; off < nsamples; off += d->buffer_size)
        {
            int n = (int)std::min<long>(d->buffer_size, nsamples - off);
            std::shared_ptr<dsp::stream<complex_t>> src =
                stage == 0 ? d->in : (stage == 1 ? d->agc->output_stream : (stage == 2 ? d->rrc->output_stream : d->rec->input_stream));
            memcpy(src->writeBuf, in + off * 2, n * sizeof(complex_t));
            src->swap(n);
            std::shared_ptr<dsp::stream<complex_t>> dst;
            if (stage == 0)
            {
                d->agc->work();
                dst = d->agc->output_stream;
            }
            else if (stage == 1)
            {
                d->rrc->work();
                dst = d->rrc->output_stream;
            }
            else if (stage == 2)
            {
                if (!d->pll)
                    return -1;
                d->pll->work();
                dst = d->pll->output_stream;
                if (d->post_pll_dc)
                {
                    d->post_pll_dc->work();
                    dst = d->post_pll_dc->output_stream;
                }
                if (d->delay)
                {
                    d->delay->work();
                    dst = d->delay->output_stream;
                }
            }
            else
            {
                d->rec->work();
                dst = d->rec->output_stream;
            }
            int m = dst->getDataSize();
            if (pos + m > cap)
                m = (int)(cap - pos);
            memcpy(out + pos * 2, dst->readBuf, m * sizeof(complex_t));
            dst->flush();
            pos += m;
        }
        return pos;
    }

    /* Carried loop state, for stage-isolated tests */
    void ref_demod_state(void *h, float *out8)
    {
        RefDemod *d = (RefDemod *)h;
        out8[0] = d->agc->gain;
        out8[1] = d->pll ? d->pll->phase : 0;
        out8[2] = d->pll ? d->pll->freq : 0;
        out8[3] = d->rec_g ? d->rec_g->mu : d->rec_mm->mu;
        out8[4] = d->rec_g ? d->rec_g->omega : d->rec_mm->omega;
        out8[5] = (float)(d->rec_g ? d->rec_g->inc : d->rec_mm->inc);
        out8[6] = d->pll ? d->pll->alpha : 0;
        out8[7] = d->pll ? d->pll->beta : 0;
    }

    /* ------------------------------------------------------------------ FEC side */
    void *ref_fec_create(const ref_fec_cfg *c)
    {
        RefFec *f = new RefFec();
        f->cfg = *c;
        memset(f->errors, 0, sizeof(f->errors));
        if (c->kind == 0)
        {
            f->buffer_size = 8192 * 2;
            f->cadu_bytes = 1024;
            f->v34 = std::make_shared<viterbi::Viterbi3_4>(c->ber_thresold, c->outsync_after, f->buffer_size);
            f->deframer = std::make_shared<deframing::BPSK_CCSDS_Deframer>();
            f->deframer->STATE_SYNCED = 18;
            f->rs = std::make_shared<reedsolomon::ReedSolomon>(reedsolomon::RS223);
        }
        else if (c->kind == 2)
        {
            /* CCSDSSimplePSKDecoderModule ctor, module_ccsds_simple_psk_decoder.cpp:19-98 */
            f->buffer_size = c->cadu_size;
            f->cadu_bytes = (int)ceil(c->cadu_size / 8.0);
            f->deframer = std::make_shared<deframing::BPSK_CCSDS_Deframer>(c->cadu_size, c->asm_sync);
            f->deframer_qpsk = std::make_shared<deframing::BPSK_CCSDS_Deframer>(c->cadu_size, c->asm_sync);
            if (c->rs_i != 0)
                f->rs = std::make_shared<reedsolomon::ReedSolomon>(c->rs_type == 1 ? reedsolomon::RS239 : reedsolomon::RS223, c->rs_fill_bytes);
            if (c->cadu_size % 8 != 0)
            {
                f->deframer->CADU_PADDING = c->cadu_size % 8;
                f->deframer_qpsk->CADU_PADDING = c->cadu_size % 8;
            }
            f->qpsk_diff.swap = c->qpsk_swap_diff;
            f->bits_buf.resize(c->cadu_size * 2);
            f->qpsk_diff_buffer.resize(c->cadu_size * 2);
        }
        else
        {
            f->buffer_size = std::max<int>(c->cadu_size, 8192);
            f->cadu_bytes = (int)ceil(c->cadu_size / 8.0);
            std::vector<phase_t> phases;
            bool oqpsk = c->constellation == 2;
            if (c->constellation == 0)
                phases = {PHASE_0};
            else if (c->constellation == 5)
                phases = {PHASE_90};
            else
                phases = {PHASE_0, PHASE_90};
            if (c->conv_rate == 0)
                f->v12 = std::make_shared<viterbi::Viterbi1_2>(c->ber_thresold, c->outsync_after, f->buffer_size, phases, oqpsk);
            else
            {
                std::shared_ptr<viterbi::puncturing::GenericDepunc> dp;
                if (c->conv_rate == 2)
                    dp = std::make_shared<viterbi::puncturing::Depunc23>();
                else if (c->conv_rate == 3)
                    dp = std::make_shared<viterbi::puncturing::Depunc34>();
                else if (c->conv_rate == 5)
                    dp = std::make_shared<viterbi::puncturing::Depunc56>();
                else
                    dp = std::make_shared<viterbi::puncturing::Depunc78>();
                f->vp = std::make_shared<viterbi::Viterbi_Depunc>(dp, c->ber_thresold, c->outsync_after, f->buffer_size, phases, oqpsk);
                /* members / heap blocks the reference reads before writing them (the tail of the 2054-step test decode beyond the
                   depunctured test symbols, the sliding buffer): pinned to zero here so that the oracle is deterministic */
                memset(f->vp->ber_test_buffer, 0, sizeof(f->vp->ber_test_buffer));
                memset(f->vp->ber_soft_buffer, 0, sizeof(f->vp->ber_soft_buffer));
                memset(f->vp->ber_depunc_buffer, 0, sizeof(f->vp->ber_depunc_buffer));
                memset(f->vp->ber_decoded_buffer, 0, sizeof(f->vp->ber_decoded_buffer));
                memset(f->vp->ber_encoded_buffer, 0, sizeof(f->vp->ber_encoded_buffer));
                memset(f->vp->vit_buffer.buffer_ptr, 0, (size_t)f->buffer_size * 4);
            }
            f->deframer = std::make_shared<deframing::BPSK_CCSDS_Deframer>(c->cadu_size, c->asm_sync);
            if (c->cadu_size % 8 != 0)
                f->deframer->CADU_PADDING = c->cadu_size % 8;
            if (c->rs_i != 0)
                f->rs = std::make_shared<reedsolomon::ReedSolomon>(c->rs_type == 1 ? reedsolomon::RS239 : reedsolomon::RS223, c->rs_fill_bytes);
        }
        f->viterbi_out.resize(f->buffer_size * 8);
        f->frame_buffer.resize(f->buffer_size * 8 + 1024 * 10);
        f->soft.resize(f->buffer_size);
        return f;
    }
    void ref_fec_destroy(void *h) { delete (RefFec *)h; }
    int ref_fec_chunk_size(void *h) { return ((RefFec *)h)->buffer_size; }
    int ref_fec_cadu_bytes(void *h) { return ((RefFec *)h)->cadu_bytes; }

    /*
     * Runs whole chunks of the int8 soft stream (nsoft must be a multiple of the chunk size; the
     * remainder is ignored). Per chunk diagnostics (optional): vit_state[c], vit_ber[c],
     * defr_state[c] (state after the chunk), bits_out (decoded bits, 1 per byte, only for chunks
     * that produced output) and *nbits. rs_err (optional) gets rs_i ints per written frame...
     * Returns number of CADU bytes written to cadu_out.
     */
    long ref_fec_run(void *h, const int8_t *soft, long nsoft, uint8_t *cadu_out, long cadu_cap, int *vit_state, float *vit_ber, int *defr_state,
                     uint8_t *bits_out, long *nbits, int *rs_err, long *nframes_seen)
    {
        RefFec *f = (RefFec *)h;
        long outp = 0, bitp = 0, frames_seen = 0;
        long nchunks = nsoft / f->buffer_size;
        for (long c = 0; c < nchunks; c++)
        {
            memcpy(f->soft.data(), soft + c * f->buffer_size, f->buffer_size);
            int vout = 0;
            if (f->cfg.kind == 0)
            {
                vout = f->v34->work(f->soft.data(), f->buffer_size, f->viterbi_out.data());
                if (vit_state)
                    vit_state[c] = f->v34->getState();
                if (vit_ber)
                    vit_ber[c] = f->v34->ber();
                if (vout > 0)
                {
                    if (bits_out)
                        memcpy(bits_out + bitp, f->viterbi_out.data(), vout);
                    bitp += vout;
                    int frames = f->deframer->work(f->viterbi_out.data(), vout, f->frame_buffer.data());
                    if (f->deframer->getState() == f->deframer->STATE_NOSYNC)
                    {
                        f->noSyncsRuns++;
                        if (f->noSyncsRuns >= 10)
                        {
                            f->v34->reset();
                            f->noSyncsRuns = 0;
                        }
                    }
                    else
                        f->noSyncsRuns = 0;
                    for (int i = 0; i < frames; i++)
                    {
                        uint8_t *cadu = &f->frame_buffer[i * 1024];
                        derand_ccsds(&cadu[4], 1024 - 4);
                        f->rs->decode_interlaved(&cadu[4], true, 4, f->errors);
                        if (rs_err)
                            memcpy(rs_err + frames_seen * 4, f->errors, 4 * sizeof(int));
                        frames_seen++;
                        if (outp + 1024 <= cadu_cap)
                        {
                            memcpy(cadu_out + outp, cadu, 1024);
                            outp += 1024;
                        }
                    }
                }
            }
            else if (f->cfg.kind == 2)
            {
                /* one iteration of CCSDSSimplePSKDecoderModule::process, module_ccsds_simple_psk_decoder.cpp:144-296
                   (oqpsk_method2/3 not wired: the CUDA path rejects them) */
                const ref_fec_cfg &k = f->cfg;
                const int n = f->buffer_size;
                int8_t *sb = f->soft.data();
                uint8_t *bits = f->bits_buf.data();
                dsp::constellation_t qpsk_const(dsp::QPSK);
                int frames = 0;
                if (k.constellation == 0)
                {
                    for (int i = 0; i < n; i++)
                        bits[i] = sb[i] > 0;
                    if (k.nrzm)
                        f->diff.decode_bits(bits, n);
                }
                else
                {
                    if (k.oqpsk_delay)
                        for (int i = 0; i < n / 2; i++)
                        {
                            int8_t back = sb[i * 2 + 0];
                            sb[i * 2 + 0] = f->last_q_oqpsk;
                            f->last_q_oqpsk = back;
                        }
                    if (k.qpsk_swap_iq)
                        rotate_soft(sb, n, PHASE_0, true);
                    if (k.nrzm)
                    {
                        for (int i = 0; i < n / 2; i++)
                            f->qpsk_diff_buffer[i] = qpsk_const.soft_demod(&sb[i * 2]);
                        f->qpsk_diff.work(f->qpsk_diff_buffer.data(), n / 2, bits);
                    }
                    else
                    {
                        for (int i = 0; i < n / 2; i++)
                        {
                            uint8_t sym = qpsk_const.soft_demod(&sb[i * 2]);
                            bits[i * 2 + 0] = sym >> 1;
                            bits[i * 2 + 1] = sym & 1;
                        }
                        frames += f->deframer_qpsk->work(bits, n, &f->frame_buffer[frames * f->cadu_bytes]);
                        rotate_soft(sb, n, PHASE_90, false);
                        for (int i = 0; i < n / 2; i++)
                        {
                            uint8_t sym = qpsk_const.soft_demod(&sb[i * 2]);
                            bits[i * 2 + 0] = sym >> 1;
                            bits[i * 2 + 1] = sym & 1;
                        }
                    }
                }
                if (bits_out)
                    memcpy(bits_out + bitp, bits, n);
                bitp += n;
                frames += f->deframer->work(bits, n, &f->frame_buffer[frames * f->cadu_bytes]);
                for (int i = 0; i < frames; i++)
                {
                    uint8_t *cadu = &f->frame_buffer[i * f->cadu_bytes];
                    if (k.derandomize && !k.derand_after_rs)
                        derand_ccsds(&cadu[k.derand_start], f->cadu_bytes - k.derand_start);
                    if (k.rs_i != 0)
                        f->rs->decode_interlaved(&cadu[4], k.rs_dualbasis, k.rs_i, f->errors);
                    bool valid = true;
                    for (int j = 0; j < k.rs_i; j++)
                        if (f->errors[j] == -1)
                            valid = false;
                    if (k.derandomize && k.derand_after_rs)
                        derand_ccsds(&cadu[k.derand_start], f->cadu_bytes - k.derand_start);
                    if (rs_err)
                        memcpy(rs_err + frames_seen * k.rs_i, f->errors, k.rs_i * sizeof(int));
                    frames_seen++;
                    if (!k.rs_usecheck || valid)
                        if (outp + f->cadu_bytes <= cadu_cap)
                        {
                            memcpy(cadu_out + outp, cadu, f->cadu_bytes);
                            outp += f->cadu_bytes;
                        }
                }
                if (vit_state)
                    vit_state[c] = f->deframer_qpsk->getState(); /* second deframer's state (QPSK without NRZ-M) */
            }
            else
            {
                const ref_fec_cfg &k = f->cfg;
                if (k.constellation == 5 || k.iq_invert)
                    rotate_soft(f->soft.data(), f->buffer_size, PHASE_0, true);
                if (f->vp)
                    vout = f->vp->work(f->soft.data(), f->buffer_size, f->viterbi_out.data());
                else
                    vout = f->v12->work(f->soft.data(), f->buffer_size, f->viterbi_out.data());
                if (vit_state)
                    vit_state[c] = f->vp ? f->vp->getState() : f->v12->getState();
                if (vit_ber)
                    vit_ber[c] = f->vp ? f->vp->d_ber : f->v12->ber();
                if (k.nrzm)
                    f->diff.decode_bits(f->viterbi_out.data(), vout);
                if (bits_out && vout > 0)
                    memcpy(bits_out + bitp, f->viterbi_out.data(), vout);
                bitp += vout;
                int frames = f->deframer->work(f->viterbi_out.data(), vout, f->frame_buffer.data());
                for (int i = 0; i < frames; i++)
                {
                    uint8_t *cadu = &f->frame_buffer[i * f->cadu_bytes];
                    if (k.derandomize && !k.derand_after_rs)
                        derand_ccsds(&cadu[k.derand_start], f->cadu_bytes - k.derand_start);
                    if (k.rs_i != 0)
                        f->rs->decode_interlaved(&cadu[4], k.rs_dualbasis, k.rs_i, f->errors);
                    bool valid = true;
                    for (int j = 0; j < k.rs_i; j++)
                        if (f->errors[j] == -1)
                            valid = false;
                    if (k.derandomize && k.derand_after_rs)
                        derand_ccsds(&cadu[k.derand_start], f->cadu_bytes - k.derand_start);
                    if (rs_err)
                        memcpy(rs_err + frames_seen * k.rs_i, f->errors, k.rs_i * sizeof(int));
                    frames_seen++;
                    if (!k.rs_usecheck || valid)
                        if (outp + f->cadu_bytes <= cadu_cap)
                        {
                            memcpy(cadu_out + outp, cadu, f->cadu_bytes);
                            outp += f->cadu_bytes;
                        }
                }
            }
            if (defr_state)
                defr_state[c] = f->deframer->getState();
        }
        if (nbits)
            *nbits = bitp;
        if (nframes_seen)
            *nframes_seen = frames_seen;
        return outp;
    }

    /* ------------------------------------------------------------------ encoders / primitives (for the synthetic transmitter cross-check) */
    void ref_rs_encode_interleaved(uint8_t *data, int dual, int interleave, int rs_type)
    {
        reedsolomon::ReedSolomon rs(rs_type == 1 ? reedsolomon::RS239 : reedsolomon::RS223);
        rs.encode_interlaved(data, dual, interleave);
    }
    /* returns errors[] like ReedSolomon::decode_interlaved */
    void ref_rs_decode_interleaved(uint8_t *data, int dual, int interleave, int rs_type, int fill_bytes, int *errors)
    {
        reedsolomon::ReedSolomon rs(rs_type == 1 ? reedsolomon::RS239 : reedsolomon::RS223, fill_bytes);
        rs.decode_interlaved(data, dual, interleave, errors);
    }
    void ref_derand(uint8_t *data, int len) { derand_ccsds(data, len); }
    void ref_cc_encode(const uint8_t *bits, int n, uint8_t *out /* 2n */)
    {
        viterbi::CCEncoder enc(n, 7, 2, {79, 109});
        enc.work((uint8_t *)bits, out);
    }
    /* One CCDecoder object decoding `ncalls` consecutive frames of `frame` bits; syms holds
     * ncalls * 2*frame symbols followed by >= 12 bytes the caller controls (tail reads). */
    void ref_cc_decode(const uint8_t *syms, int frame, int ncalls, uint8_t *out_bits)
    {
        viterbi::CCDecoder dec(frame, 7, 2, {79, 109});
        for (int c = 0; c < ncalls; c++)
            dec.work((uint8_t *)syms + (long)c * 2 * frame, out_bits + (long)c * frame);
    }
    void ref_rotate_soft(int8_t *soft, int size, int phase, int iqswap) { rotate_soft(soft, size, (phase_t)phase, iqswap); }
    int ref_deframe(const uint8_t *bits, int nbits, int cadu_size, int state_synced, uint8_t *out)
    {
        deframing::BPSK_CCSDS_Deframer d(cadu_size);
        d.STATE_SYNCED = state_synced;
        std::vector<uint8_t> tmp((size_t)(nbits / cadu_size + 2) * (cadu_size / 8 + 1) + cadu_size);
        int total = 0;
        /* feed in 8192-bit slices like a module would (the deframer keeps state across calls) */
        for (int off = 0; off < nbits; off += 8192)
        {
            int n = std::min(8192, nbits - off);
            int fr = d.work((uint8_t *)bits + off, n, tmp.data());
            memcpy(out + (size_t)total * (cadu_size / 8), tmp.data(), (size_t)fr * (cadu_size / 8));
            total += fr;
        }
        return total;
    }

    /* ------------------------------------------------------------------ threaded pipeline (the reference's own execution model) for CPU timing */
    /*
     * One std::thread per DSP block (Block::start), demod module thread, decoder thread joined by a
     * 1 000 000-byte RingBuffer — pipeline_run.cpp:72-104. Returns wall seconds from first sample
     * to decoder exit; *cadu_bytes gets the number of CADU bytes produced, threads_used the count.
     */
    double ref_pipeline_timed(const ref_demod_cfg *dc, const ref_fec_cfg *fc, const void *raw, long nsamples, uint8_t *cadu_out, long cadu_cap,
                              long *cadu_bytes, int *threads_used)
    {
        RefDemod *d = (RefDemod *)ref_demod_create(dc);
        RefFec *f = fc ? (RefFec *)ref_fec_create(fc) : nullptr; /* fc == NULL: the demodulator module alone (C5: symbols are the product) */
        auto fifo = std::make_shared<dsp::RingBuffer<uint8_t>>(1000000);
        std::atomic<bool> demod_done{false};
        std::atomic<long> out_bytes{0};
        auto t0 = std::chrono::steady_clock::now();

        d->agc->start();
        d->rrc->start();
        if (d->pll)
            d->pll->start();
        if (d->post_pll_dc)
            d->post_pll_dc->start();
        if (d->delay)
            d->delay->start();
        d->rec->start();
        int nthreads = 3 + (d->pll ? 1 : 0) + (d->delay ? 1 : 0);

        std::thread feeder([&]() { /* stands for FileSourceBlock::work */
            for (long off = 0; off < nsamples; off += d->buffer_size)
            {
                int n = (int)std::min<long>(d->buffer_size, nsamples - off);
                convert_block(d->cfg, raw, off, n, d->in->writeBuf);
                if (!d->in->swap(n))
                    break;
            }
        });
        long expected_in = nsamples;
        std::thread demod_mod([&]() { /* PSKDemodModule::process loop */
            std::vector<int8_t> sym_buffer((size_t)d->buffer_size * 4);
            /* we stop once the M&M block has consumed every input sample: count via omega is
               not observable, so the feeder pads and we stop on a sentinel below */
            while (true)
            {
                int m = d->rec->output_stream->read();
                if (m <= 0)
                    break;
                complex_t *sym = d->rec->output_stream->readBuf;
                int nb;
                if (d->cfg.constellation == 0)
                {
                    for (int i = 0; i < m; i++)
                        sym_buffer[i] = soft_clamp(sym[i].real * 50);
                    nb = m;
                }
                else
                {
                    for (int i = 0; i < m; i++)
                    {
                        sym_buffer[i * 2] = soft_clamp(sym[i].real * 100);
                        sym_buffer[i * 2 + 1] = soft_clamp(sym[i].imag * 100);
                    }
                    nb = m * 2;
                }
                d->rec->output_stream->flush();
                if (!f)
                {
                    out_bytes += nb;
                    continue;
                }
                if (fifo->write((uint8_t *)sym_buffer.data(), nb) < 0)
                    break;
            }
            demod_done = true;
        });
        std::thread decoder([&]() {
            if (!f)
                return;
            std::vector<int8_t> chunk(f->buffer_size);
            while (true)
            {
                if (fifo->read((uint8_t *)chunk.data(), f->buffer_size) < 0)
                    break;
                long w = ref_fec_run(f, chunk.data(), f->buffer_size, cadu_out + out_bytes, cadu_cap - out_bytes, 0, 0, 0, 0, 0, 0, 0);
                out_bytes += w;
            }
        });
        nthreads += 3;

        feeder.join();
        /* Wait until every block has drained: poll block idleness through stream readiness */
        auto idle = [&]() {
            return !d->in->getReady() && !d->agc->output_stream->getReady() && !d->rrc->output_stream->getReady() &&
                   (!d->pll || !d->pll->output_stream->getReady()) && (!d->delay || !d->delay->output_stream->getReady()) &&
                   !d->rec->output_stream->getReady();
        };
        int calm = 0;
        while (calm < 20)
        {
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
            calm = idle() ? calm + 1 : 0;
        }
        (void)expected_in;
        /* stop blocks like PSKDemodModule::stop() */
        d->agc->stop();
        d->rrc->stop();
        if (d->pll)
            d->pll->stop();
        if (d->post_pll_dc)
            d->post_pll_dc->stop();
        if (d->delay)
            d->delay->stop();
        d->rec->stop();
        d->rec->output_stream->stopReader();
        demod_mod.join();
        while (f && fifo->getReadable() >= f->buffer_size)
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        fifo->stopReader();
        fifo->stopWriter();
        decoder.join();
        auto t1 = std::chrono::steady_clock::now();
        double secs = std::chrono::duration<double>(t1 - t0).count() - 0.020; /* minus the 20 ms idle-detection window */
        if (cadu_bytes)
            *cadu_bytes = out_bytes;
        if (threads_used)
            *threads_used = nthreads;
        ref_demod_destroy(d);
        if (f)
            ref_fec_destroy(f);
        return secs;
    }
}

/* ------------------------------------------------------------------ CADU -> space packets
 * One ccsds::ccsds_aos::Demuxer per virtual channel, fed the frames of that channel in stream order, exactly like the instrument modules do
 * (module_metop_instruments.cpp:66-140). Output: every packet the demuxers return, in the order they return them, as
 * [6 raw header bytes][payload bytes] back to back; recs[4 * i + 0..3] = frame index, vcid, payload length, apid. */
namespace
{
    struct RefDemux
    {
        int mpdu, insert, insert_size, sec_ext;
        std::map<int, std::shared_ptr<ccsds::ccsds_aos::Demuxer>> by_vcid;
    };
}
extern "C"
{
    void *ref_demux_create(int mpdu_data_size, int has_insert_zone, int insert_zone_size, int secondary_header_extends)
    {
        RefDemux *d = new RefDemux();
        d->mpdu = mpdu_data_size;
        d->insert = has_insert_zone;
        d->insert_size = insert_zone_size;
        d->sec_ext = secondary_header_extends;
        return d;
    }
    void ref_demux_destroy(void *h) { delete (RefDemux *)h; }
    /* frames: nframes * cadu_size bytes; vcid_mask bit v = demultiplex virtual channel v. Returns the number of packets (or -1 - needed
       when a capacity is too small); *nbytes = bytes written. frame0: index of the first frame of this call in the stream. */
    long ref_demux_run(void *h, const uint8_t *frames, long nframes, int cadu_size, unsigned long long vcid_mask, long frame0, uint8_t *out, long cap_bytes,
                       long *nbytes, int *recs, long cap_recs)
    {
        RefDemux *d = (RefDemux *)h;
        long np = 0, nb = 0;
        std::vector<uint8_t> cadu(cadu_size + 16);
        for (long f = 0; f < nframes; f++)
        {
            memcpy(cadu.data(), frames + f * cadu_size, cadu_size);
            ccsds::ccsds_aos::VCDU v = ccsds::ccsds_aos::parseVCDU(cadu.data());
            if (!((vcid_mask >> v.vcid) & 1ull))
                continue;
            auto &dm = d->by_vcid[v.vcid];
            if (!dm)
                dm = std::make_shared<ccsds::ccsds_aos::Demuxer>(d->mpdu, d->insert != 0, d->insert_size, d->sec_ext != 0);
            std::vector<ccsds::CCSDSPacket> pk = dm->work(cadu.data());
            for (ccsds::CCSDSPacket &p : pk)
            {
                if (np >= cap_recs || nb + 6 + (long)p.payload.size() > cap_bytes)
                    return -1;
                memcpy(out + nb, p.header.raw, 6);
                if (!p.payload.empty())
                    memcpy(out + nb + 6, p.payload.data(), p.payload.size());
                recs[4 * np + 0] = (int)(frame0 + f);
                recs[4 * np + 1] = v.vcid;
                recs[4 * np + 2] = (int)p.payload.size();
                recs[4 * np + 3] = p.header.apid;
                nb += 6 + (long)p.payload.size();
                np++;
            }
        }
        *nbytes = nb;
        return np;
    }
}

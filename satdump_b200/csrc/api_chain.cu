// C ABI of the fused chain: demodulator + decoder of one stream on one GPU, the int8 soft stream handed over in HBM
// (the reference joins the two modules with a 1 MB host ring buffer: src-core/pipeline/pipeline_run.cpp:72-104).
#include "demod_host.h"
#include "fec_host.h"

namespace b200
{
struct Chain
{
    Demod *d = nullptr;
    Fec *f = nullptr;
    float t_total = 0;
    cudaEvent_t e0, e1;
    ~Chain()
    {
        delete d;
        delete f;
    }
};
} // namespace b200
using namespace b200;

struct b200_chain
{
    Chain c;
};

static void chain_push(Chain &c, const void *iq, long n, bool on_device)
{
    B200_REQUIRE(n <= c.d->max_batch, B200_ESTATE, "batch of %ld samples exceeds max_batch %ld", n, c.d->max_batch);
    DeviceGuard g(c.d->cfg.device);
    B200_CUDA(cudaEventRecord(c.e0, c.d->stream));
    // upper bound of the soft bytes this batch can produce
    const double omin = c.d->sps * (1.0 - c.d->cfg.clock_omega_limit) - 0.01;
    const long bound = (long)(n / omin + 64) * c.d->bps;
    B200_REQUIRE(bound <= c.f->append_room(), B200_ESTATE, "decoder soft FIFO too small for this batch (max_soft)");
    long syms;
    if (on_device)
        syms = c.d->process(iq, n, c.f->append_ptr());
    else
        syms = c.d->push_host(iq, n, c.f->append_ptr());
    // the decoder works on its own stream: order it after the demodulator (process() returns synchronised)
    c.f->commit(syms * c.d->bps);
    c.f->process();
    B200_CUDA(cudaEventRecord(c.e1, c.f->stream));
    B200_CUDA(cudaEventSynchronize(c.e1));
    cudaEventElapsedTime(&c.t_total, c.e0, c.e1);
}

extern "C" {
b200_chain *b200_chain_create(const b200_demod_cfg *dcfg, const b200_fec_cfg *fcfg)
{
    b200_chain *h = nullptr;
    guarded([&] {
        B200_REQUIRE(dcfg && fcfg, B200_EINVAL, "cfg is NULL");
        B200_REQUIRE(dcfg->device == fcfg->device, B200_EINVAL, "demodulator and decoder of a chain must share a device");
        h = new b200_chain();
        try {
            h->c.d = new Demod(*dcfg);
            h->c.f = new Fec(*fcfg);
            DeviceGuard g(dcfg->device);
            B200_CUDA(cudaEventCreate(&h->c.e0));
            B200_CUDA(cudaEventCreate(&h->c.e1));
        } catch (...) {
            delete h;
            h = nullptr;
            throw;
        }
    });
    return h;
}
void b200_chain_destroy(b200_chain *h) { delete h; }
int b200_chain_push_iq(b200_chain *h, const void *host_iq, long n)
{
    return guarded([&] {
        B200_REQUIRE(h && host_iq, B200_EINVAL, "NULL argument");
        chain_push(h->c, host_iq, n, false);
    });
}
int b200_chain_push_iq_device(b200_chain *h, const void *dev_iq, long n)
{
    return guarded([&] {
        B200_REQUIRE(h && dev_iq, B200_EINVAL, "NULL argument");
        chain_push(h->c, dev_iq, n, true);
    });
}
int b200_chain_pull_frames(b200_chain *h, uint8_t *out, long cap, long *nbytes)
{
    return guarded([&] {
        B200_REQUIRE(h && out && nbytes, B200_EINVAL, "NULL argument");
        *nbytes = h->c.f->pull(out, cap);
    });
}
int b200_chain_frames_device(b200_chain *h, const uint8_t **dev_ptr, long *nbytes)
{
    return guarded([&] {
        B200_REQUIRE(h && dev_ptr && nbytes, B200_EINVAL, "NULL argument");
        *dev_ptr = h->c.f->frames_out.p;
        *nbytes = h->c.f->out_frames * h->c.f->cadu_bytes;
        h->c.f->out_frames = 0;
    });
}
int b200_chain_prefetch_iq(b200_chain *h, const void *host_iq, long n)
{
    return guarded([&] {
        B200_REQUIRE(h && host_iq, B200_EINVAL, "NULL argument");
        h->c.d->prefetch_host(host_iq, n);
    });
}
int b200_chain_reset(b200_chain *h)
{
    return guarded([&] {
        B200_REQUIRE(h, B200_EINVAL, "NULL argument");
        h->c.d->reset();
        h->c.f->reset();
    });
}
int b200_chain_get_stats(b200_chain *h, b200_demod_stats *ds, b200_fec_stats *fs)
{
    return guarded([&] {
        B200_REQUIRE(h, B200_EINVAL, "NULL argument");
        if (ds)
            h->c.d->stats(ds);
        if (fs)
            h->c.f->stats(fs);
    });
}
int b200_chain_last_timing(b200_chain *h, float *ms, int n)
{
    return guarded([&] {
        B200_REQUIRE(h && ms && n >= 9, B200_EINVAL, "need room for 9 floats");
        Chain &c = h->c;
        ms[1] = c.d->t_agcfir;
        ms[2] = c.d->t_costas;
        ms[3] = c.d->t_mm;
        ms[4] = c.f->t_vit;
        ms[5] = c.f->t_frames;
        ms[0] = ms[1] + ms[2] + ms[3] + ms[4] + ms[5];
        ms[6] = c.f->t_vit_main;               // k_vit_main alone (sum over its launches in the last push)
        ms[7] = (float)c.f->last_main_chunks;  // chunks those launches decoded
        ms[8] = c.t_total;                     // CUDA-event time from the start of the push (before any H2D) to the last kernel
    });
}
}

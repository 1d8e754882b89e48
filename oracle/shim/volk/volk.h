/*
 * TEST INFRASTRUCTURE ONLY (oracle). Minimal stand-in for the system VOLK library, which the
 * reference links against but does not vendor (src-core/CMakeLists.txt:221). Written from the
 * VOLK *interface* used at the reference call sites (fir.cpp:82, clock_recovery_mm.cpp:97,
 * baseband_interface.h:178-188, cc_decoder.cpp:56-93). All kernels are the scalar "generic"
 * flavour: strict left-to-right accumulation, division by the scale factor.
 */
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <stdbool.h>
#ifdef __cplusplus
#include <complex>
typedef std::complex<float> lv_32fc_t;
extern "C" {
#else
#include <complex.h>
typedef float _Complex lv_32fc_t;
#endif

static inline size_t volk_get_alignment(void) { return 32; }
static inline void *volk_malloc(size_t size, size_t alignment)
{
    void *p = NULL;
    if (alignment < sizeof(void *)) alignment = sizeof(void *);
    if (posix_memalign(&p, alignment, size ? size : alignment) != 0) return NULL;
    return p;
}
static inline void volk_free(void *p) { free(p); }

/* complex x real dot product; accumulate oldest -> newest */
static inline void volk_32fc_32f_dot_prod_32fc(lv_32fc_t *result, const lv_32fc_t *input, const float *taps, unsigned int num_points)
{
    const float *in = (const float *)input;
    float accr = 0.0f, acci = 0.0f;
    for (unsigned int k = 0; k < num_points; k++) {
        accr += in[2 * k + 0] * taps[k];
        acci += in[2 * k + 1] * taps[k];
    }
    ((float *)result)[0] = accr;
    ((float *)result)[1] = acci;
}
static inline void volk_32fc_32f_dot_prod_32fc_a(lv_32fc_t *result, const lv_32fc_t *input, const float *taps, unsigned int num_points)
{
    volk_32fc_32f_dot_prod_32fc(result, input, taps, num_points);
}
static inline void volk_32f_x2_dot_prod_32f(float *result, const float *input, const float *taps, unsigned int num_points)
{
    float acc = 0.0f;
    for (unsigned int k = 0; k < num_points; k++) acc += input[k] * taps[k];
    *result = acc;
}
static inline void volk_32f_x2_dot_prod_32f_a(float *result, const float *input, const float *taps, unsigned int num_points)
{
    volk_32f_x2_dot_prod_32f(result, input, taps, num_points);
}
static inline void volk_16i_s32f_convert_32f_u(float *out, const int16_t *in, const float scalar, unsigned int num_points)
{
    for (unsigned int k = 0; k < num_points; k++) out[k] = ((float)in[k]) / scalar;
}
static inline void volk_8i_s32f_convert_32f_u(float *out, const int8_t *in, const float scalar, unsigned int num_points)
{
    for (unsigned int k = 0; k < num_points; k++) out[k] = ((float)in[k]) / scalar;
}
static inline void volk_32i_s32f_convert_32f_u(float *out, const int32_t *in, const float scalar, unsigned int num_points)
{
    for (unsigned int k = 0; k < num_points; k++) out[k] = ((float)in[k]) / scalar;
}

/* Rotator (freq_shift.cpp:26-30, pm_to_bpsk.cpp:27-31): out[k] = in[k] * phase; phase *= phase_inc. VOLK (not vendored by the reference;
 * 2.x / 3.x generic flavour as published) brings the phasor back to unit length after every 512 samples of a call and once more at the
 * end of a call whose length is not a multiple of 512. Plain float arithmetic: (a + jb)(c + jd) = (ac - bd) + j(ad + bc), one rounding
 * per operation. The rounding walk of the phasor's angle depends on the VOLK flavour (the SIMD ones advance several samples at once),
 * so downstream comparisons against this restatement carry a tolerance (tests/floors.py measures it). */
#include <math.h>
static inline void b200_shim_rotate(float *out, const float *in, float inc_re, float inc_im, float *phase /* [2] */, unsigned int num_points)
{
    float pr = phase[0], pi = phase[1];
    unsigned int k = 0, since = 0;
    for (k = 0; k < num_points; k++) {
        const float a = in[2 * k], b = in[2 * k + 1];
        out[2 * k] = a * pr - b * pi;
        out[2 * k + 1] = a * pi + b * pr;
        const float nr = pr * inc_re - pi * inc_im, ni = pr * inc_im + pi * inc_re;
        pr = nr;
        pi = ni;
        if (++since == 512) {
            const float m = hypotf(pr, pi);
            pr /= m;
            pi /= m;
            since = 0;
        }
    }
    if (since) {
        const float m = hypotf(pr, pi);
        pr /= m;
        pi /= m;
    }
    phase[0] = pr;
    phase[1] = pi;
}
#ifdef __cplusplus
} /* (the two entry points take a C++ complex by value) */
static inline void volk_32fc_s32fc_x2_rotator_32fc(lv_32fc_t *out, const lv_32fc_t *in, const lv_32fc_t phase_inc, lv_32fc_t *phase, unsigned int num_points)
{
    b200_shim_rotate((float *)out, (const float *)in, phase_inc.real(), phase_inc.imag(), (float *)phase, num_points);
}
static inline void volk_32fc_s32fc_x2_rotator2_32fc(lv_32fc_t *out, const lv_32fc_t *in, const lv_32fc_t *phase_inc, lv_32fc_t *phase, unsigned int num_points)
{
    b200_shim_rotate((float *)out, (const float *)in, phase_inc->real(), phase_inc->imag(), (float *)phase, num_points);
}
extern "C" {
#endif

/* Kernel-description query used by CCDecoder to pick an ACS kernel. We advertise only
 * "generic", so the decoder keeps the reference's own vendored fixed kernel. */
typedef struct volk_func_desc {
    const char **impl_names;
    const int *impl_deps;
    const bool *impl_alignment;
    size_t n_impls;
} volk_func_desc_t;
#ifndef __cplusplus
typedef struct volk_func_desc volk_func_desc;
#endif
static inline struct volk_func_desc volk_8u_x4_conv_k7_r2_8u_get_func_desc(void)
{
    static const char *names[] = {"generic"};
    static const int deps[] = {0};
    static const bool al[] = {false};
    struct volk_func_desc d = {names, deps, al, 1};
    return d;
}
static inline void volk_8u_x4_conv_k7_r2_8u_manual(unsigned char *Y, unsigned char *X, unsigned char *syms, unsigned char *dec,
                                                   unsigned int framebits, unsigned int excess, unsigned char *Branchtab, const char *impl)
{
    (void)Y; (void)X; (void)syms; (void)dec; (void)framebits; (void)excess; (void)Branchtab; (void)impl;
    abort(); /* never selected: only "generic" is advertised */
}
#ifdef __cplusplus
}
#endif

// Diagnostic (not part of the product): warp-instruction throughput per SM of the instruction kinds the AGC + FIR kernel is made of
// (FFMA, FFMA2 with a register / uniform scalar operand, FMUL, FADD, I2F, MUFU, PRMT-based conversion, LDS.128), alone and mixed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/tools/ubench_fp tests/tools/ubench_fp.cu
#include <cuda_runtime.h>
#include <cstdio>
#define ITERS 2048
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int OP> __global__ void k(float *out, float seed, int one)
{
    float x[8];
    unsigned long long y[8];
    __shared__ float4 sm[256 * 2];
    sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    sm[threadIdx.x + 256] = make_float4(seed, seed, seed, seed);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = seed * (i + 1) + threadIdx.x; y[i] = (unsigned long long)__float_as_uint(x[i]) * 0x100000001ull; }
    const float c = seed * 0.5f;
    const unsigned long long cc = (unsigned long long)__float_as_uint(c) * 0x100000001ull;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) x[i] = fmaf(x[i], c, 1.0f);
            if (OP == 1) y[i] = ffma2(y[i], cc, y[i]);
            if (OP == 2) x[i] = x[i] * c;
            if (OP == 3) x[i] = x[i] + c;
            if (OP == 4) x[i] = (float)(__float_as_int(x[i]) >> 9);                       // I2F
            if (OP == 5) x[i] = rsqrtf(x[i]);                                             // MUFU.RSQ
            if (OP == 6) x[i] = __uint_as_float(__byte_perm(__float_as_uint(x[i]), 0x43C00000u, 0x7610)) - 385.0f; // PRMT + FADD conversion
            if (OP == 7) { float4 v = sm[(threadIdx.x + (it & one)) & 511]; x[i] += v.x; x[(i + 1) & 7] += v.w; } // LDS.128 + 2 FADD
            if (OP == 8) { y[i] = ffma2(y[i], cc, y[i]); x[i] = fmaf(x[i], c, 1.0f); }   // FFMA2 + FFMA interleaved
            if (OP == 9) { y[i] = ffma2(y[i], cc, y[i]); x[i] = __uint_as_float(__float_as_uint(x[i]) ^ (unsigned)it); } // FFMA2 + LOP3
            if (OP == 10) { y[i] = ffma2(y[i], cc, y[i]); y[(i + 4) & 7] = ffma2(y[(i + 4) & 7], cc, cc); x[i] = __uint_as_float(__float_as_uint(x[i]) ^ (unsigned)it); } // 2 FFMA2 + LOP3
            if (OP == 11) x[i] = fmaxf(x[i], c);                                          // FMNMX
            if (OP == 12) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i])); x[i] = r; } // MUFU.SQRT
            if (OP == 13) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i])); x[i] = r; } // MUFU.RSQ
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + __uint_as_float((unsigned)(y[i] >> 32)) + __uint_as_float((unsigned)y[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char *name, float *d, int sms, double per_iter)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 4, threads = 256; // 32 warps per SM
    k<OP><<<blocks, threads>>>(d, 1.0001f, 1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(d, 1.0001f, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double inst_per_sm = (double)ITERS * 8 * 32 * per_iter;
    printf("%-28s %.3f ms  -> %.2f warp-inst / clk / SM (nominal %.0f MHz; %.0f listed inst per iteration)\n", name, ms, inst_per_sm / (ms * 1e-3 * clk * 1e3), clk / 1e3, per_iter);
}
int main()
{
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *d; cudaMalloc(&d, sms * 4 * 256 * 4);
    run<0>("FFMA", d, sms, 1); run<1>("FFMA2", d, sms, 1); run<2>("FMUL", d, sms, 1); run<3>("FADD", d, sms, 1); run<4>("SHF+I2F", d, sms, 2);
    run<5>("MUFU.RSQ(+fixup)", d, sms, 1); run<6>("PRMT+FADD", d, sms, 2); run<7>("LDS.128+2FADD", d, sms, 3); run<8>("FFMA2+FFMA", d, sms, 2);
    run<12>("MUFU.SQRT", d, sms, 1); run<13>("MUFU.RSQ", d, sms, 1); run<9>("FFMA2+LOP3", d, sms, 2); run<10>("2xFFMA2+LOP3", d, sms, 3); run<11>("FMNMX", d, sms, 1);
    return 0;
}

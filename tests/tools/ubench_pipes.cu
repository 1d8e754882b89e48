// Diagnostic (not part of the product): warp-instruction throughput per SM of the instruction kinds the Viterbi ACS kernel is made of.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/tools/ubench_pipes tests/tools/ubench_pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
#define ITERS 2048
template <int OP> __device__ __forceinline__ unsigned op(unsigned a, unsigned b, int lane)
{
    if (OP == 0) return __vadd2(a, b);                                   // VIADD.16x2
    if (OP == 1) return __vminu2(a, b);                                  // VIMNMX.U16x2
    if (OP == 2) { bool p, q; unsigned r = __vibmin_u16x2(a, b, &p, &q); return r + (p ? 1u : 0u) + (q ? 2u : 0u); } // + predicate use
    if (OP == 3) return __byte_perm(a, b, 0x3210u ^ (b & 0x3333u));      // PRMT
    if (OP == 4) return (a ^ b) & 0x00FF00FFu;                           // LOP3
    if (OP == 5) return a + b;                                           // IADD3 / IMAD.IADD
    if (OP == 6) return a * 0xFFFF0001u + b;                             // IMAD
    if (OP == 7) return __shfl_sync(0xffffffffu, a, (lane + 1) & 31) ^ b; // SHFL.IDX
    if (OP == 8) return __ballot_sync(0xffffffffu, (a & 1) != 0) + b;    // VOTE
    if (OP == 9) return __reduce_min_sync(0xffffffffu, a) + b;           // CREDUX.MIN
    if (OP == 10) return min(a, b);                                      // VIMNMX.U32
    if (OP == 11) return __viaddmin_u16x2(a, b, a ^ 0x10001u);           // VIADDMNMX.U16x2
    return a;
}
template <int OP> __global__ void k(unsigned *out, unsigned seed)
{
    const int lane = threadIdx.x & 31;
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = seed * (i + 1) + threadIdx.x;
    unsigned y = seed | 1;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = op<OP>(x[i], y, lane);
        y += 0x10001u;
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char *name, unsigned *d, int sms)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 4, threads = 256; // 32 warps per SM
    k<OP><<<blocks, threads>>>(d, 12345u);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(d, 12345u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double inst_per_sm = (double)ITERS * 8 * 32; // warp-instructions of the measured kind per SM
    printf("%-18s %.3f ms  -> %.2f warp-inst / clk / SM at the nominal %.0f MHz (plus ~1/8 loop overhead)\n", name, ms, inst_per_sm / (ms * 1e-3 * clk * 1e3), clk / 1e3);
}
int main()
{
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned *d; cudaMalloc(&d, sms * 4 * 256 * 4);
    run<0>("VIADD.16x2", d, sms); run<1>("VIMNMX.U16x2", d, sms); run<2>("VIMNMX.U16x2+P", d, sms); run<3>("PRMT", d, sms); run<4>("LOP3", d, sms);
    run<5>("IADD", d, sms); run<6>("IMAD", d, sms); run<7>("SHFL.IDX", d, sms); run<8>("VOTE", d, sms); run<9>("CREDUX.MIN", d, sms);
    run<10>("VIMNMX.U32", d, sms); run<11>("VIADDMNMX.U16x2", d, sms);
    return 0;
}

// Diagnostic (not part of the product): times the Viterbi ACS kernels of satdump_b200/csrc/fec.cuh against each other on random soft
// symbols and checks that every variant produces the survivor decisions and end states of the round-1 kernel bit for bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/tools/bench_acs tests/tools/bench_acs.cu
#define B200_DEFINE_KERNELS
#include "../../satdump_b200/csrc/fec.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace b200;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char **argv)
{
    const int nchunks = argc > 1 ? atoi(argv[1]) : 6372, rate34 = argc > 2 ? atoi(argv[2]) : 1;
    const bool check = argc <= 3 || atoi(argv[3]) != 0; // 0: timing only (large runs)
    VitGeom g{};
    g.rate34 = rate34; g.chunk = rate34 ? 16384 : 10232; g.F = rate34 ? g.chunk * 3 / 4 : g.chunk / 2;
    g.dec_stride = (g.F + 6 + 7) & ~7; g.bit_words = (g.F + 31) / 32 + 1;
    VitHyp h{0, 1, 1};
    std::vector<int8_t> soft((size_t)nchunks * g.chunk + 64);
    srand(7);
    for (auto &v : soft) { int s = (rand() & 1) ? 90 : -90; s += (rand() % 121) - 60; v = (int8_t)(s > 127 ? 127 : (s < -128 ? -128 : s)); }
    std::vector<int> ss(nchunks);
    for (int i = 0; i < nchunks; i++) ss[i] = i == 0 ? -1 : rand() & 63;
    int8_t *d_soft; int *d_ss; uint2 *d_dec[2]; VitRec *d_rec[2];
    const size_t dec_bytes = (size_t)nchunks * g.dec_stride * 8;
    CK(cudaMalloc(&d_soft, soft.size())); CK(cudaMalloc(&d_ss, nchunks * 4));
    for (int k = 0; k < 2; k++) { CK(cudaMalloc(&d_dec[k], dec_bytes)); CK(cudaMalloc(&d_rec[k], nchunks * sizeof(VitRec))); CK(cudaMemset(d_rec[k], 0, nchunks * sizeof(VitRec))); CK(cudaMemset(d_dec[k], 0, dec_bytes)); }
    CK(cudaMemcpy(d_soft, soft.data(), soft.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ss, ss.data(), nchunks * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int nb = (nchunks + 3) / 4;
    const char *names[6] = {"k_vit_acs (round 1)", "acs3 shfl/sts", "acs3 smem-metrics/sts", "acs3 shfl/transpose", "acs3 smem-metrics/transpose", "acs3 smem-metrics/sel"};
    std::vector<uint2> a(check ? (size_t)nchunks * g.dec_stride : 1), b(a.size());
    std::vector<VitRec> ra(nchunks), rb(nchunks);
    long total_bad = 0;
    for (int variant = 0; variant < 6; variant++) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; rep++) {
            cudaEventRecord(e0);
            if (variant == 0) k_vit_acs<<<nb, 128>>>(d_soft, 0, nchunks, g, h, d_ss, d_dec[0], d_rec[0]);
            else if (variant == 1) k_vit_acs3<false, 0><<<nb, 128>>>(d_soft, 0, nchunks, g, h, d_ss, d_dec[1], d_rec[1], nullptr);
            else if (variant == 2) k_vit_acs3<true, 0><<<nb, 128>>>(d_soft, 0, nchunks, g, h, d_ss, d_dec[1], d_rec[1], nullptr);
            else if (variant == 3) k_vit_acs3<false, 2><<<nb, 128>>>(d_soft, 0, nchunks, g, h, d_ss, d_dec[1], d_rec[1], nullptr);
            else if (variant == 4) k_vit_acs3<true, 2><<<nb, 128>>>(d_soft, 0, nchunks, g, h, d_ss, d_dec[1], d_rec[1], nullptr);
            else k_vit_acs3<true, 1><<<nb, 128>>>(d_soft, 0, nchunks, g, h, d_ss, d_dec[1], d_rec[1], nullptr);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        const double steps = (double)nchunks * (g.F + 6);
        printf("variant %d (%s): %.3f ms for %d chunks, %.2f ns per 1000 trellis steps, %.1f Gstep/s\n", variant, names[variant], best, nchunks,
               best * 1e6 / steps * 1e3, steps / best / 1e6);
        if (!check) continue;
        if (variant == 0) {
            CK(cudaMemcpy(a.data(), d_dec[0], dec_bytes, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(ra.data(), d_rec[0], nchunks * sizeof(VitRec), cudaMemcpyDeviceToHost));
            continue;
        }
        CK(cudaMemcpy(b.data(), d_dec[1], dec_bytes, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rb.data(), d_rec[1], nchunks * sizeof(VitRec), cudaMemcpyDeviceToHost));
        long bad = 0, first = -1;
        for (int q = 0; q < nchunks; q++) {
            for (int t = 0; t < g.F + 6; t++) { const size_t i = (size_t)q * g.dec_stride + t; if (a[i].x != b[i].x || a[i].y != b[i].y) { if (first < 0) first = (long)i; bad++; } }
            if (ra[q].end_state != rb[q].end_state) bad++;
        }
        printf("   decision rows / end states differing from round 1: %ld (first at %ld)\n", bad, first);
        total_bad += bad;
        CK(cudaMemset(d_dec[1], 0, dec_bytes));
    }
    return total_bad != 0;
}

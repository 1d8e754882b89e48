// Diagnostic (not a test): bias / spread of the float magnitude variants against sqrt in double.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ float rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__global__ void k(int n, double *sum, double *sumabs, double *mx)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // s2 sweeps [0.25, 4)
    float s2 = 0.25f * exp2f(4.0f * (float)i / (float)n);
    double ref = sqrt((double)s2);
    float r = rsq(s2);
    float m0 = s2 * r;
    float m1 = fmaf(fmaf(-m0, m0, s2), 0.5f * r, m0);
    float m2 = sqrtf(s2);
    float v[3] = {m0, m1, m2};
    for (int j = 0; j < 3; j++) {
        double e = ((double)v[j] - ref) / ref;
        atomicAdd(&sum[j], e);
        atomicAdd(&sumabs[j], fabs(e));
        // max via compare-and-swap on the bit pattern (positive doubles order like integers)
        unsigned long long *p = (unsigned long long *)&mx[j];
        atomicMax(p, (unsigned long long)__double_as_longlong(fabs(e)));
    }
}
int main()
{
    const int n = 1 << 24;
    double *d;
    cudaMalloc(&d, 9 * sizeof(double));
    cudaMemset(d, 0, 9 * sizeof(double));
    k<<<n / 256, 256>>>(n, d, d + 3, d + 6);
    double h[9];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char *nm[3] = {"s2*rsqrt", "newton", "sqrtf"};
    for (int j = 0; j < 3; j++)
        printf("%-9s mean rel err %+.3e  mean |err| %.3e  max %.3e  (float ulp ~ 6e-8..1.2e-7)\n", nm[j], h[j] / n, h[3 + j] / n, h[6 + j]);
    return 0;
}

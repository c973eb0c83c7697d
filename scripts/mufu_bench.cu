// MUFU.EX2 throughput per SM (development tool)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__global__ void k(float* out, int iters) {
    float v[16];
    for (int j = 0; j < 16; ++j) v[j] = -0.001f * (threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = ex2(v[j]) - 1.5f;
    }
    float s = 0; for (int j = 0; j < 16; ++j) s += v[j];
    if (s == 123.456f) out[0] = s;
}
__global__ void k2(float* out, int iters) {   // pure ex2 chains, no FADD
    float v[16];
    for (int j = 0; j < 16; ++j) v[j] = -0.001f * (threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = ex2(-v[j]);
    }
    float s = 0; for (int j = 0; j < 16; ++j) s += v[j];
    if (s == 123.456f) out[0] = s;
}
int main() {
    float* out; cudaMalloc(&out, 4);
    int dev_clk; cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
    for (int warps = 1; warps <= 16; warps *= 2) {
        for (int which = 0; which < 2; ++which) {
            const int iters = 4096;
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                if (which == 0) k<<<148, warps * 32>>>(out, iters); else k2<<<148, warps * 32>>>(out, iters);
                cudaEventRecord(b); cudaEventSynchronize(b);
            }
            float ms; cudaEventElapsedTime(&ms, a, b);
            double exps = 16.0 * iters * warps * 32;   // per SM
            printf("warps/SM %2d kernel %d: %.3f ms -> %.2f ex2/clk/SM at %.0f MHz nominal\n", warps, which, ms,
                   exps / (ms * 1e-3 * dev_clk * 1e3), dev_clk / 1e3);
        }
    }
    return 0;
}

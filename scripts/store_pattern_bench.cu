// Write-only streaming: (A) every thread writes its own contiguous 128 bytes with 4 x st.v4.f64 (the RAMBO row
// layout written straight from registers), (B) the warp writes 512 contiguous bytes per st.v2.f64 instruction
// (rows transposed through shared memory first).  Same bytes, same grid-stride structure.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__global__ void __launch_bounds__(128) pat_a(double* out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)i;
#pragma unroll
        for (int p = 0; p < 4; ++p) st256(out + i * 16 + p * 4, v, v + 1, v + 2, v + 3);
    }
}
__global__ void __launch_bounds__(128) pat_b(double* out, int64_t n) {
    __shared__ double2 tile[4][32 * 8 + 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + warp * 32); i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)(i0 + lane);
#pragma unroll
        for (int j = 0; j < 8; ++j) tile[warp][lane * 8 + ((j + lane) & 7)] = make_double2(v + j, v - j);   // skewed rows
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int idx = j * 32 + lane;            // element of the warp's 256 x 16 B block
            const int row = idx >> 3, col = idx & 7;
            const double2 w = tile[warp][row * 8 + ((col + row) & 7)];
            asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(out + i0 * 16 + idx * 2), "d"(w.x), "d"(w.y) : "memory");
        }
        __syncwarp();
    }
}
int main() {
    const int64_t n = 1LL << 27;
    double* out; cudaMalloc(&out, (size_t)n * 128);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int pat = 0; pat < 2; ++pat)
        for (int bps : {8, 16, 32}) {
            float sum = 0;
            for (int i = 0; i < 5; ++i) {
                cudaEventRecord(e0);
                if (pat == 0) pat_a<<<sms * bps, 128>>>(out, n); else pat_b<<<sms * bps, 128>>>(out, n);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (i >= 1) sum += ms;
            }
            printf("pattern %c  %2d CTAs/SM  %.3f ms  %.0f GB/s\n", pat ? 'B' : 'A', bps, sum / 4, n * 128.0 / (sum / 4 * 1e-3) / 1e9);
        }
    return 0;
}

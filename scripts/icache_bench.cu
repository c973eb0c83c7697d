// Does a long straight-line loop body run at a lower IPC than a short one (instruction fetch)?
// Same instruction mix (4 DFMA + 8 LOP3 + 8 IMAD per step, independent chains), body = R steps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int R>
__global__ void __launch_bounds__(256) body(int steps, double* sinkd, uint32_t* sinki, double m, double c, uint32_t k) {
    double d[4];
    uint32_t l[8], q[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { l[i] = threadIdx.x * 7 + i; q[i] = threadIdx.x * 13 + i; }
    for (int it = 0; it < steps / R; ++it) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i]) : "r"(k), "r"(l[(i + 1) & 7]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(q[i]) : "r"(k), "r"(q[(i + 3) & 7]));
            }
        }
    }
    double sd = 0; uint32_t si = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) sd += d[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) si += l[i] + q[i];
    if (sd == 1.2345) sinkd[0] = sd;
    if (si == 12345u) sinki[0] = si;
}

template <int R>
static void run(int bps) {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* sd; uint32_t* si; cudaMalloc(&sd, 8); cudaMalloc(&si, 4);
    const int steps = 1 << 16;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    body<R><<<sms * bps, 256>>>(1 << 10, sd, si, 0.999999, 1e-9, 0x9E3779B9u);
    cudaEventRecord(e0);
    body<R><<<sms * bps, 256>>>(steps, sd, si, 0.999999, 1e-9, 0x9E3779B9u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * khz * 1e3 / ((double)steps * bps * 2.0);
    printf("body %5d instr (%6.1f KB)  %d warps/SMSP  %8.3f ms  %.3f cycles per instr\n", R * 20, R * 20 * 16 / 1024.0, bps * 2, ms, cyc / 20);
    cudaFree(sd); cudaFree(si);
}

int main() {
    for (int bps = 3; bps <= 3; ++bps) {
        run<1>(bps); run<4>(bps); run<8>(bps); run<16>(bps); run<24>(bps); run<32>(bps); run<48>(bps); run<64>(bps); run<128>(bps);
    }
    return 0;
}

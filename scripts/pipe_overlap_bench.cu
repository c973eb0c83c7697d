// Do the half-rate pipes of sm_100 overlap?  Each kernel runs N iterations of a fixed mix of
// independent dependency chains per thread: D = DFMA (FP64 pipe), L = LOP3 (ALU pipe), M = IMAD
// (FMA pipe).  If pipes overlap, time(D+L) ~ max(time(D), time(L)); if the dispatch port is held for
// both cycles of a half-rate instruction, time(D+L) ~ time(D) + time(L).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/pipe_overlap_bench.cu -o scripts/_build/pipe_overlap
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int ND, int NL, int NM>
__global__ void __launch_bounds__(256) mix(int iters, double* sinkd, uint32_t* sinki, double m, double c, uint32_t k) {
    double d[8];
    uint32_t l[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i] = threadIdx.x + i; l[i] = threadIdx.x * 7 + i; q[i] = threadIdx.x * 13 + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < ND) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
                if (i < NL) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i]) : "r"(k), "r"(l[(i + 1) & 7]));
                if (i < NM) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(q[i]) : "r"(k), "r"(q[(i + 3) & 7]));
            }
        }
    }
    double sd = 0; uint32_t si = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sd += d[i]; si += l[i] + q[i]; }
    if (sd == 1.2345) sinkd[0] = sd;
    if (si == 12345u) sinki[0] = si;
}

template <int ND, int NL, int NM>
static void run(const char* tag, int blocks_per_sm) {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* sd; uint32_t* si; cudaMalloc(&sd, 8); cudaMalloc(&si, 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<ND, NL, NM><<<sms * blocks_per_sm, 256>>>(100, sd, si, 0.999999, 1e-9, 0x9E3779B9u);
    cudaEventRecord(e0);
    mix<ND, NL, NM><<<sms * blocks_per_sm, 256>>>(iters, sd, si, 0.999999, 1e-9, 0x9E3779B9u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    // cycles per (warp, inner step) on one SMSP, assuming max clock: warps per SMSP = blocks_per_sm * 8 / 4
    const double warps_smsp = blocks_per_sm * 2.0;
    const double steps = (double)iters * 4;
    const double cyc = ms * 1e-3 * khz * 1e3 / (steps * warps_smsp);
    printf("%-14s D=%d L=%d M=%d  %8.3f ms  %6.2f cycles per warp-step (%.2f per instr)\n", tag, ND, NL, NM, ms, cyc, cyc / (ND + NL + NM));
    cudaFree(sd); cudaFree(si);
}

int main() {
    for (int bps = 2; bps <= 4; bps += 2) {
        printf("-- %d CTAs/SM of 256 threads (%d warps per SMSP)\n", bps, bps * 2);
        run<8, 0, 0>("fp64", bps);
        run<0, 8, 0>("lop3", bps);
        run<0, 0, 8>("imad", bps);
        run<8, 8, 0>("fp64+lop3", bps);
        run<8, 0, 8>("fp64+imad", bps);
        run<0, 8, 8>("lop3+imad", bps);
        run<8, 8, 8>("all three", bps);
        run<4, 4, 0>("fp64+lop3 4+4", bps);
        run<4, 8, 8>("4D+8L+8M", bps);
    }
    return 0;
}

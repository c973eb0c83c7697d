// NOTE: every variant here issues from ONE thread inside a divergent branch -- the form that makes
// ptxas wrap each MMA in an R2UR waterfall loop (55+ cycles per MMA).  The product issues from a
// warp-uniform loop with elect.sync inside the asm (elm_tc.cuh), which removes that cost.
// Microbenchmark of tcgen05.mma issue/throughput for the two GEMM shapes of the ELM pipeline
// (development tool, not product code).  One CTA per SM; one thread issues `reps` groups of MMAs
// back to back, commits, waits; cycles measured with clock64 around issue+completion.
//   mode 0: G1 shape only   (M128 N128 K8 tf32, A from TMEM), g1 MMAs per group
//   mode 1: G2 shape only   (M128 N16  K8 tf32, A from TMEM), g2 MMAs per group, D fixed
//   mode 2: alternate G1 group (D = S[b]) and G2 group (A = S[b'], D = D2)
//   mode 3: G2 shape with A from shared memory (SS)
//   mode 4: like 2 but the G2 group reads A from the fixed A1 columns (never written by MMAs)
//   mode 5: G2-shape with N = nvar
#include <cstdio>
#include <cstdlib>
#include "../hepmc_b200/csrc/elm_tc.cuh"

using namespace hepmc::tc;

// single-thread issue form (what the pipeline used before the warp-uniform wrappers of elm_tc.cuh)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
template <int ACC>
__device__ __forceinline__ void umma_1t(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 bd;\nsetp.ne.b32 p, %5, 0;\nmov.b64 bd, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n}" ::"r"(d_tmem), "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "n"(ACC) : "memory");
}
__device__ __forceinline__ void commit_1t(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(160) bench(int mode, int reps, int g1, int g2, int nvar, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_ptr;
    float* f = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) f[i] = 0.001f * (i % 97);
    const uint32_t bar_a = smem_u32(&bar);
    if (threadIdx.x == 0) { mbar_init(bar_a, 1); mbar_init(smem_u32(&bar2), 1); }
    if (threadIdx.x / 32 == 4) tmem_alloc(smem_u32(&tmem_ptr), 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tmem_ptr;
    if (threadIdx.x < 128) {   // fill TMEM with finite values
        uint32_t v[32];
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(0.01f * j);
        const uint32_t lane_base = tb + ((threadIdx.x & ~31u) << 16);
        for (int c = 0; c < 512; c += 32) tmem_st32(lane_base + c, v);
        tmem_wait_st();
    }
    fence_before();
    fence_async_smem();
    __syncthreads();
    fence_after();
    if (mode >= 11 && mode <= 13) {
        // hand-off ping-pong: issuer (thread 128): n MMAs + commit(bar) ; owner warp 0: wait bar, fences,
        // arrive bar2 ; issuer waits bar2.  mode 11: 1 MMA, mode 12: no MMA (plain arrive), mode 13: 4 MMAs N=64
        const uint32_t b1 = bar_a, b2 = smem_u32(&bar2);
        if (threadIdx.x < 32) {
            uint32_t ph = 0;
            for (int r = 0; r < reps; ++r) {
                mbar_wait(b1, ph); ph ^= 1;
                fence_after();
                fence_before();
                __syncwarp();
                if (threadIdx.x == 0) mbar_arrive(b2);
            }
        } else if (threadIdx.x == 128) {
            const uint32_t sb = smem_u32(smem);
            const uint64_t b1d = make_desc(sb, TILE_N * 16, 128);
            const uint32_t id1 = make_idesc(128, 64);
            uint32_t ph = 0;
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                if (mode == 12) mbar_arrive(b1);
                else {
                    const int n = mode == 13 ? 4 : 1;
                    for (int k = 0; k < n; ++k) umma_tf32_ts(tb + COL_S0 + (r % 3) * 64, tb + COL_A1, b1d, id1, k > 0);
                    commit_1t(b1);
                }
                mbar_wait(b2, ph); ph ^= 1;
                fence_after();
            }
            const long long t1 = clock64();
            if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t1 - t0; }
        }
        __syncthreads();
        if (threadIdx.x / 32 == 4) tmem_dealloc(tb, 512);
        return;
    }
    if (threadIdx.x == 0 && mode == 10) {
        const uint32_t sb = smem_u32(smem);
        const uint64_t b2d = make_desc(sb + 16384, N2 * 16, 128);
        const uint32_t id2 = make_idesc(128, 16);
        const uint32_t lo = (uint32_t)b2d, hi = (uint32_t)(b2d >> 32);
        for (int r = 0; r < reps; ++r) {
            const uint32_t a = tb + COL_S0 + ((r + 2) % 3) * 128;
#pragma unroll
            for (int k = 0; k < 16; ++k) umma_1t<1>(tb + COL_D2 + 16, a + k * 8, lo + k * 32, hi, id2);
        }
        commit_1t(smem_u32(&bar2));
        mbar_wait(smem_u32(&bar2), 0);
    }
    if (threadIdx.x == 128) {
        const uint32_t sb = smem_u32(smem);
        const uint64_t b1d = make_desc(sb, TILE_N * 16, 128);
        const uint64_t b2d = make_desc(sb + 16384, N2 * 16, 128);
        const uint64_t asd = make_desc(sb + 32768, TILE_M * 16, 128);
        const uint32_t id1 = make_idesc(128, 128), id2 = make_idesc(128, mode == 5 ? nvar : 16);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const int b = r % 3;
            if (mode == 0 || mode == 2 || mode == 4) {
                for (int k = 0; k < g1; ++k)
                    umma_tf32_ts(tb + COL_S0 + b * 128, tb + COL_A1 + (k % 4) * 8, b1d + (uint64_t)((k % 4) * 256), id1, k > 0);
            }
            if (mode == 1 || mode == 2 || mode == 5) {
                const int bb = (r + 1) % 3;
                for (int k = 0; k < g2; ++k)
                    umma_tf32_ts(tb + COL_D2, tb + COL_S0 + bb * 128 + (k % 16) * 8, b2d + (uint64_t)((k % 16) * 32), id2, 1);
            }
            if (mode == 4) {
                for (int k = 0; k < g2; ++k)
                    umma_tf32_ts(tb + COL_D2, tb + COL_A1 + (k % 4) * 8, b2d + (uint64_t)((k % 16) * 32), id2, 1);
            }
            if (mode == 6 || mode == 8 || mode == 10) {   // unrolled, precomputed operands (the kernel's issue path)
                const uint32_t lo = (uint32_t)b2d, hi = (uint32_t)(b2d >> 32);
                const uint32_t a = tb + COL_S0 + ((r + 1) % 3) * 128;
#pragma unroll
                for (int k = 0; k < 16; ++k) umma_1t<1>(tb + COL_D2, a + k * 8, lo + k * 32, hi, id2);
            }
            if (mode == 7 || mode == 8) {
                const uint32_t lo = (uint32_t)b1d, hi = (uint32_t)(b1d >> 32);
                const uint32_t d = tb + COL_S0 + b * 128;
                umma_1t<0>(d, tb + COL_A1, lo, hi, id1);
#pragma unroll
                for (int k = 1; k < 4; ++k) umma_1t<1>(d, tb + COL_A1 + k * 8, lo + k * 256, hi, id1);
            }
            if (mode == 9) {   // unrolled N=16 MMAs, same A columns every time
                const uint32_t lo = (uint32_t)b2d, hi = (uint32_t)(b2d >> 32);
#pragma unroll
                for (int k = 0; k < 16; ++k) umma_1t<1>(tb + COL_D2, tb + COL_A1, lo, hi, id2);
            }
            if (mode == 3) {
                for (int k = 0; k < g2; ++k)
                    umma_tf32_ss(tb + COL_D2, asd + (uint64_t)((k % 4) * 256), b2d + (uint64_t)((k % 16) * 32), id2, 1);
            }
        }
        commit_1t(bar_a);
        const long long t1 = clock64();
        mbar_wait(bar_a, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncthreads();
    if (threadIdx.x / 32 == 4) tmem_dealloc(tb, 512);
}

int main(int argc, char** argv) {
    long long* out;
    cudaMallocManaged(&out, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    struct Case { int mode, g1, g2, nvar; const char* what; };
    const Case cases[] = {
        {0, 4, 0, 0, "G1 only, 4 MMA/group (N=128)"},
        {0, 1, 0, 0, "G1 only, 1 MMA/group"},
        {1, 0, 16, 0, "G2 only, 16 MMA/group (N=16, TS)"},
        {1, 0, 2, 0, "G2 only, 2 MMA/group"},
        {2, 4, 16, 0, "alternate G1(4) / G2(16)"},
        {2, 4, 2, 0, "alternate G1(4) / G2(2)"},
        {2, 1, 2, 0, "alternate G1(1) / G2(2)"},
        {4, 4, 16, 0, "alternate G1(4) / G2(16), G2 reads A1 columns"},
        {3, 0, 16, 0, "G2 only, SS (A from smem), 16 MMA/group"},
        {6, 0, 16, 0, "unrolled G2 x16 (N=16)"},
        {7, 4, 0, 0, "unrolled G1 x4 (N=128)"},
        {8, 4, 16, 0, "unrolled G2 x16 then G1 x4"},
        {9, 0, 16, 0, "unrolled N=16 x16, identical operands"},
        {10, 0, 16, 0, "two issuing threads, each unrolled G2 x16"},
        {11, 0, 0, 0, "ping-pong: 1 MMA + commit -> owner -> arrive -> issuer"},
        {12, 0, 0, 0, "ping-pong: plain arrive both ways"},
        {13, 0, 0, 0, "ping-pong: 4 MMA (N=64) + commit"},
        {5, 0, 16, 32, "G2 shape N=32"},
        {5, 0, 16, 64, "G2 shape N=64"},
        {5, 0, 16, 128, "G2 shape N=128"},
        {5, 0, 16, 8, "G2 shape N=8"},
    };
    const int reps = 200;
    for (const Case& c : cases) {
        for (int it = 0; it < 2; ++it) {
            bench<<<1, 160, 64 * 1024>>>(c.mode, reps, c.g1, c.g2, c.nvar, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", c.what, cudaGetErrorString(e)); return 1; }
        }
        printf("%-50s issue %8.1f  total %8.1f cycles/group\n", c.what, (double)out[0] / reps, (double)out[1] / reps);
    }
    return 0;
}

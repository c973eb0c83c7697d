// Shared device/host helpers: error plumbing, Philox4x32-10 in registers, bits->double,
// warp reductions, Chan/Welford merge.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/hepmc_b200.h"
#include "fastmath.cuh"

namespace hepmc {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define HEPMC_CUDA(call)                                        \
    do {                                                        \
        int _e = hepmc::check_cuda((call), #call);              \
        if (_e) return _e;                                      \
    } while (0)

#define HEPMC_REQUIRE(cond, code, ...)                          \
    do {                                                        \
        if (!(cond)) {                                          \
            hepmc::set_error(__VA_ARGS__);                      \
            return (code);                                      \
        }                                                       \
    } while (0)

#define HEPMC_LAUNCH_CHECK() HEPMC_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
int sm_count();

// ---------------------------------------------------------------- Philox4x32-10
// Salmon et al., SC'11 (Random123).  Round keys are warp-uniform (functions of the
// seed only), so they are expanded once per thread block into registers by the caller.
struct PhiloxKey {
    uint32_t k0[10];
    uint32_t k1[10];
};

__host__ __device__ __forceinline__ PhiloxKey philox_expand(uint64_t seed) {
    PhiloxKey k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = a;
        k.k1[r] = b;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    return k;
}

struct U4 {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                             const PhiloxKey& key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.k0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.k1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return U4{c0, c1, c2, c3};
}

// 52 random bits k = (20 high bits of a : all 32 bits of b) -> (2k+1) * 2^-53, strictly
// inside (0,1), exact (oracle: hepmc_oracle.u52).
__device__ __forceinline__ double u52(uint32_t a, uint32_t b) {
    const uint32_t hi = 0x3FF00000u | (a >> 12);
    return __hiloint2double((int)hi, (int)b) - 0.99999999999999988897769753748;  // 1 - 2^-53
}

// u52 leaves the low 12 bits of its first word unused.  The Metropolis test of the random-walk
// kernel takes its uniform from those spare bits of the step's proposal draw, (a, c) = the first
// words of the two u52 pairs of one Philox call: 24 bits -> (k + 1/2) 2^-24, strictly inside (0,1).
// This saves one Philox call per step (52 of 412 instructions); the realised acceptance probability
// is the ratio rounded to a multiple of 2^-24 (absolute error <= 3e-8, invisible next to the Monte
// Carlo error of any feasible run).  Oracle: hepmc_oracle.u24_spare.
__device__ __forceinline__ double u24_spare(uint32_t a, uint32_t c) {
    const uint32_t k = ((a & 0xFFFu) << 12) | (c & 0xFFFu);
    return ((double)k + 0.5) * 5.9604644775390625e-08;   // 2^-24
}

// VEGAS: one 64-bit random word (a:b) -> (bin, frac) with t = K * (a:b) / 2^64 computed
// exactly in integers: bin = floor(t) in [0, K), frac = top 52 bits of the remainder.
__device__ __forceinline__ void bin_frac(uint32_t a, uint32_t b, uint32_t K, uint32_t& bin, double& frac) {
    const uint64_t plo = (uint64_t)K * b;
    const uint64_t phi = (uint64_t)K * a + (plo >> 32);
    bin = (uint32_t)(phi >> 32);
    const uint32_t f_hi = (uint32_t)phi;   // top 32 bits of the 64-bit fraction
    const uint32_t f_lo = (uint32_t)plo;   // low 32 bits
    const uint32_t hi = 0x3FF00000u | (f_hi >> 12);
    const uint32_t lo = __funnelshift_r(f_lo, f_hi, 12);
    frac = __hiloint2double((int)hi, (int)lo) - 0.99999999999999988897769753748;
}

// ---------------------------------------------------------------- reductions
struct Stats {  // count, mean, M2 = sum (x - mean)^2
    double n, mean, m2;
};

// Chan, Golub, LeVeque pairwise merge; a then b (order matters only in the last bits,
// and every caller uses a fixed order so results are reproducible).
__host__ __device__ __forceinline__ Stats merge(const Stats& a, const Stats& b) {
    if (b.n == 0.0) return a;
    if (a.n == 0.0) return b;
    Stats r;
    r.n = a.n + b.n;
    const double delta = b.mean - a.mean;
    const double fb = b.n / r.n;
    r.mean = a.mean + delta * fb;
    r.m2 = a.m2 + b.m2 + delta * delta * a.n * fb;
    return r;
}

__device__ __forceinline__ Stats warp_merge(Stats s) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        Stats o;
        o.n = __shfl_xor_sync(0xffffffffu, s.n, off);
        o.mean = __shfl_xor_sync(0xffffffffu, s.mean, off);
        o.m2 = __shfl_xor_sync(0xffffffffu, s.m2, off);
        // fixed orientation: lower lane's value is always the left operand
        const bool low = ((threadIdx.x & off) == 0);
        s = low ? merge(s, o) : merge(o, s);
    }
    return s;
}

// warp_merge for lanes that all hold the SAME count (a full chunk): b.n / (a.n + b.n) = 1/2 exactly at every level of
// the butterfly, so the two divisions of merge() become multiplications by 1/2 -- bit-identical to warp_merge (the
// products are formed in merge()'s order), ~25 instead of ~150 instructions per level.
__device__ __forceinline__ Stats warp_merge_equal(Stats s) {
    double n = s.n;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double om = __shfl_xor_sync(0xffffffffu, s.mean, off);
        const double o2 = __shfl_xor_sync(0xffffffffu, s.m2, off);
        const bool low = ((threadIdx.x & off) == 0);          // the lower lane's value is the left operand
        const double am = low ? s.mean : om, bm = low ? om : s.mean;
        const double a2 = low ? s.m2 : o2, b2 = low ? o2 : s.m2;
        const double delta = bm - am;
        s.mean = am + delta * 0.5;
        s.m2 = a2 + b2 + delta * delta * n * 0.5;
        n += n;
    }
    s.n = n;
    return s;
}

// Per-thread shifted accumulator: shift = first value seen, so (mean - shift)^2 <~ n * var
// and the single pass loses at most log10(n) digits (n <= chunk/threads).
struct ShiftedAcc {
    double shift, a, b, n;
    __device__ __forceinline__ void init() { shift = 0.0; a = 0.0; b = 0.0; n = 0.0; }
    __device__ __forceinline__ void add(double v) {
        if (n == 0.0) shift = v;
        const double d = v - shift;
        a += d;
        b = fma(d, d, b);
        n += 1.0;
    }
    __device__ __forceinline__ Stats stats() const {
        Stats s;
        s.n = n;
        if (n == 0.0) { s.mean = 0.0; s.m2 = 0.0; return s; }
        const double am = a / n;
        s.mean = shift + am;
        s.m2 = b - a * am;
        if (s.m2 < 0.0) s.m2 = 0.0;
        return s;
    }
};

// block-level merge of per-thread Stats in thread order; result valid in thread 0.
template <int THREADS>
__device__ __forceinline__ Stats block_merge(Stats s, Stats* smem /* THREADS/32 */) {
    s = warp_merge(s);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) smem[warp] = s;
    __syncthreads();
    Stats r{0.0, 0.0, 0.0};
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) r = merge(r, smem[w]);
    }
    return r;
}

__device__ __forceinline__ int64_t warp_sum_i64(int64_t v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
}

}  // namespace hepmc

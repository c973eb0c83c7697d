// ELM Gaussian-basis gradient on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   grad_jk = sum_i g_ji (x_jk - c_ik),   g_ji = -exp(-h_i |x_j - c_i|^2) * coef_i
//
// is two GEMMs around an exponential, structurally flash-attention with head dim 8:
//   GEMM-1  S = A1 . B1^T     A1 row j = [x', |x'|^2, 1]          (x' = x - 1/2, K = D+2)
//                             B1 row i = log2e * [2 h c', -h, -h |c'|^2]
//           so S_ji = log2e * (-h_i |x_j - c_i|^2) directly.  Operands are TF32; to keep FP32-level
//           accuracy through the cancellation both sides are split hi/lo and concatenated along K
//           ("3xTF32": hi*hi + hi*lo + lo*hi), K1 = 3 (D+2) <= 32 for D <= 8.
//   P = exp2(S)  (one MUFU.EX2 per pair; the MMA truncates FP32 -> TF32, a uniform -2^-12 relative
//           scale of the whole gradient; an explicit cvt.rna would double the XU-pipe load),
//           written back IN PLACE into the S columns of TMEM.  P and B2 stay TF32 (8-bit exponent):
//           trained ELM weights reach 1e6..1e10 with heavy cancellation, so tiny P still matter;
//           an F16 GEMM-2 (half the MMA count) was measured 11% faster but loses those terms.
//   GEMM-2  D2 += P . B2^T    B2 row n = coef * c'_n (n < D), coef (n = D)      -> D2 (128 x 16)
//           grad_jk = D2_jk - x'_jk * D2_jD
// Both GEMMs take A from TMEM (tcgen05.mma [d], [a], b_desc) and B from shared memory in the
// canonical K-major no-swizzle layout; node tables for all tiles stay resident in smem
// (B1: 8 KB, B2: 4 KB per 64 nodes).
//
// One CTA runs TWO independent pipelines ("groups"), each = 128 chains (TMEM lanes 0..127, its own
// 256 TMEM columns and mbarriers) = 4 owner warps + 1 MMA-issuing warp, sharing only the node
// tables, the MUFU and the tensor pipe.  HMC evaluates its gradients strictly one after the other,
// and one evaluation is a chain of ~10 hand-offs (A1 -> G1 -> S_full -> exp -> P_ready -> G2 ->
// D2_full -> epilogue -> leapfrog) of ~500 cycles each: a single pipeline spends 5-7 k cycles per
// evaluation waiting (measured: 7.2 k cycles for a ONE-tile evaluation).  Two groups fill each
// other's bubbles.  A tcgen05.mma costs its issuing thread 55-80 cycles whatever its N
// (scripts/tc_microbench.cu: 16 x (M128 N16 K8) = 900 cycles from one thread; a second thread
// issues another 16 in the same time), so the 12 MMAs of a 64-node tile (4 GEMM-1 + 8 GEMM-2)
// are spread over N_ISS issuing threads per group: issuer 0 does GEMM-1 and its share of the
// GEMM-2 K-steps, the others only GEMM-2, each into its own D2 accumulator (summed in the epilogue).
// Per group S is triple buffered in TMEM (64 columns each) so GEMM-1 of node tile t+2 / GEMM-2 of
// tile t overlap the exponentials of the tiles in between.  mbarrier pipeline per group:
//   A_ready (4 warps) -> [G1(0..2)] -> S_full[b] (commit) -> exp -> P_ready[b] (4 warps)
//   -> every issuer: its K-steps of G2(t); issuers > 0 commit G2_done[b]
//   -> issuer 0: G1(t+3) once G2_done[b] says S[b] is no longer read ... -> D2_full (N_ISS commits) -> epilogue.
#pragma once
#include "common.cuh"

namespace hepmc {
namespace tc {

constexpr int TILE_M = 128;      // chains per group tile (TMEM lanes)
constexpr int TILE_N = 64;       // nodes per tile
constexpr int K1 = 32;           // 3 * (D + 2) padded to a multiple of 8, D <= 8
constexpr int N2 = 16;           // D + 1 padded to the MMA N granularity at M = 128
constexpr int MAX_D = 8;
constexpr int N_GROUPS = 2;      // independent pipelines per CTA
#ifndef HEPMC_TC_ISSUERS
#define HEPMC_TC_ISSUERS 1
#endif
constexpr int N_ISS = HEPMC_TC_ISSUERS;   // MMA-issuing threads per group; issuer q owns K-steps q, q + N_ISS, ... of GEMM-2
constexpr int N_SBUF = 3;        // S/P tiles in flight per group
constexpr int GROUP_COLS = 256;  // TMEM columns per group
constexpr int COL_A1 = 0, COL_D2 = 32, COL_S0 = 64, TMEM_COLS = 512;   // per group; S buffer b at COL_S0 + TILE_N b
constexpr int B1_TILE_BYTES = (K1 / 4) * TILE_N * 16;   // [K chunk][node row][16 B]
constexpr int B2_TILE_BYTES = (TILE_N / 4) * N2 * 16;   // [node chunk][n row][16 B]
constexpr int N_OWNER = TILE_M;                         // owner threads per group: one chain = one TMEM lane each
constexpr int N_COMPUTE = N_GROUPS * TILE_M;            // owner threads of all groups
constexpr int N_THREADS = N_COMPUTE + 32 * N_GROUPS * N_ISS;   // + MMA-issuing warps (the first allocates TMEM)
constexpr int MMA_WARP = N_COMPUTE / 32;
constexpr int N_BARS = 2 + 3 * N_SBUF;                  // mbarriers per group
static_assert(N_ISS == 1 || N_ISS == 2, "issuers must divide the 2 K-steps of the narrowest tile");
static_assert(COL_D2 + N_ISS * N2 <= COL_S0, "TMEM column plan");
static_assert(COL_S0 + N_SBUF * TILE_N <= GROUP_COLS && N_GROUPS * GROUP_COLS <= TMEM_COLS, "TMEM column plan");

__host__ __device__ inline size_t smem_bytes(int nodes) {
    const int nt = (nodes + TILE_N - 1) / TILE_N;
    return (size_t)nt * (B1_TILE_BYTES + B2_TILE_BYTES) + 256;
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#ifndef HEPMC_TC_SPIN
#define HEPMC_TC_SPIN 0   // 1: non-blocking test_wait polling -- same hand-off latency, but the polling warps take issue slots from the exp stage
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if HEPMC_TC_SPIN
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
#else
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// The MMA-issuing warps run their loops with ALL 32 lanes (warp-uniform control flow and operands)
// and elect one lane per instruction inside the asm block.  Issuing from a single-lane branch
// (if (threadIdx.x == ...)) makes ptxas wrap every tcgen05.mma in an ELECT / 6 x R2UR.BROADCAST /
// BRA.U.ANY waterfall loop, because it cannot prove the operands warp-uniform: measured 55-130
// cycles per MMA for the issuing thread instead of the ~10 the uniform datapath needs.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile(
        "{\n"
        ".reg .pred e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]^T, TF32 operands, FP32 accumulate; descriptor passed as
// (lo, hi) words (advanced with ONE 32-bit add per MMA), accumulate flag a compile-time constant
template <int ACC>
__device__ __forceinline__ void umma_tf32_ts_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi,
                                                uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred p, e;\n"
        ".reg .b64 bd;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 bd, {%2, %3};\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n"
        "}" ::"r"(d_tmem), "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "n"(ACC) : "memory");
}

#define HEPMC_R8(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define HEPMC_W8(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : HEPMC_R8(v, 0), HEPMC_R8(v, 8), HEPMC_R8(v, 16), HEPMC_R8(v, 24)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : HEPMC_R8(v, 0), HEPMC_R8(v, 8)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        HEPMC_W8(v, 0), HEPMC_W8(v, 8), HEPMC_W8(v, 16), HEPMC_W8(v, 24)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        HEPMC_W8(v, 0), HEPMC_W8(v, 8)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// exp2 on the FMA pipe: round-to-nearest split x = xi + f, f in [-0.5, 0.5], degree-4 minimax
// polynomial for 2^f (max relative error 2.7e-6 in float arithmetic, 200x below the TF32
// truncation P gets in GEMM-2), exponent inserted with one integer shift-add.  9 instructions
// per value (~11 issue slots with the surrounding moves) against one MUFU.EX2 that occupies the
// 4-lane MUFU pipe for 8 cycles per warp: the exp stage gives POLY_MASK's share of every 16-value
// chunk to this routine.  In isolation (scripts/exp_microbench.cu) 4-6 of 16 is the balance point
// (300 -> 235 cycles per chunk with two warps per sub-partition); inside the pipeline issue slots
// are scarcer and 3 of 16 is best.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.0f);
    const float t = x + 12582912.0f;             // 1.5 * 2^23: xi lands in the low mantissa bits
    const float f = x - (t - 12582912.0f);
    float p = fmaf(0.009570102207362652f, f, 0.05591785907745361f);
    p = fmaf(p, f, 0.240247443318367f);
    p = fmaf(p, f, 0.6931217908859253f);
    p = fmaf(p, f, 0.9999992847442627f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
#ifndef HEPMC_TC_POLY_MASK
#define HEPMC_TC_POLY_MASK 0x0421   // values 0, 5, 10 of each 16-value chunk (measured optimum: 2/16 -7 %, 3/16 -8 %, 4/16 -5 %, 5/16 +2 %, 6/16 +10 % cycles)
#endif
template <int J>
__device__ __forceinline__ float ex2_mixed(float x) {
    if constexpr ((HEPMC_TC_POLY_MASK >> J) & 1) return ex2_poly(x);
    else return ex2_approx(x);
}
__device__ __forceinline__ void ex2_chunk(uint32_t (&v)[16]) {
#define HEPMC_EX2(J) v[J] = __float_as_uint(ex2_mixed<J>(__uint_as_float(v[J])));
    HEPMC_EX2(0) HEPMC_EX2(1) HEPMC_EX2(2) HEPMC_EX2(3) HEPMC_EX2(4) HEPMC_EX2(5) HEPMC_EX2(6) HEPMC_EX2(7)
    HEPMC_EX2(8) HEPMC_EX2(9) HEPMC_EX2(10) HEPMC_EX2(11) HEPMC_EX2(12) HEPMC_EX2(13) HEPMC_EX2(14) HEPMC_EX2(15)
#undef HEPMC_EX2
}
__device__ __forceinline__ uint32_t to_tf32_rna(float x) {
    uint32_t y;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// shared-memory matrix descriptor, K-major, no swizzle (canonical ((8,n),2):((1,SBO),LBO) in 16-B units)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// development knob (scripts only): 1 = skip GEMM-2, 2 = skip exp, 4 = GEMM-2 with 2 MMAs per tile,
// 8 = GEMM-1 with 1 MMA per tile, 16 = exp stage without TMEM access, 32 = group 1 idle.
// Read ONCE per kernel into a register: every tcgen05 asm statement is a compiler memory barrier,
// so reading the global inside the issue loops cost a ~200-cycle load per use (measured with
// the trace below: 110 instead of 55 cycles per issued MMA).
// The knob exists only in development builds (make EXTRA=-DHEPMC_TC_DEBUG); in the product library it is the
// constant 0 and every branch on it is compiled out.
#ifdef HEPMC_TC_DEBUG
__device__ int g_tc_debug = 0;
#define HEPMC_TC_DBG() g_tc_debug
#else
#define HEPMC_TC_DBG() 0
#endif

// development trace (make EXTRA=-DHEPMC_TC_TRACE): clock64 stamps of the hand-offs of group 0 of
// CTA 0; owner thread 0 writes records [0, 4096), the issuer [4096, 8192): (code, tile, clock)
#ifdef HEPMC_TC_TRACE
__device__ long long* g_tc_trace = nullptr;
#define HEPMC_TRACE(st, base, code, tile)                                                      \
    do {                                                                                       \
        if ((st).trp && (threadIdx.x & 31) == 0 && (st).tr < 1360) {                                                      \
            long long* r_ = (st).trp + 3 * ((base) + (st).tr++);                               \
            r_[0] = (code); r_[1] = (tile); r_[2] = clock64();                                 \
        }                                                                                      \
    } while (0)
#else
#define HEPMC_TRACE(st, base, code, tile) do { } while (0)
#endif

struct Shared {
    float* b1;           // [n_tiles][K1/4][TILE_N][4]
    float* b2;           // [n_tiles][TILE_N/4][N2][4]
    uint64_t* bars;      // per group: A_ready, D2_full, S_full[N_SBUF], P_ready[N_SBUF]
    uint32_t* tmem_ptr;
};

__device__ __forceinline__ Shared carve_tc(unsigned char* base, int n_tiles) {
    Shared s;
    s.b1 = reinterpret_cast<float*>(base);
    s.b2 = reinterpret_cast<float*>(base + (size_t)n_tiles * B1_TILE_BYTES);
    s.bars = reinterpret_cast<uint64_t*>(base + (size_t)n_tiles * (B1_TILE_BYTES + B2_TILE_BYTES));
    s.tmem_ptr = reinterpret_cast<uint32_t*>(s.bars + N_GROUPS * N_BARS);
    return s;
}

// S_full[b] = BAR_S0 + b, P_ready[b] = BAR_P0 + b, G2_done[b] = BAR_G0 + b (indices within a group's N_BARS barriers)
enum { BAR_A = 0, BAR_D2 = 1, BAR_S0 = 2, BAR_P0 = 2 + N_SBUF, BAR_G0 = 2 + 2 * N_SBUF };

// Build the resident node tables (all threads).  centers (I, D), widths (I,), out_w (I,).
// FMA_G2 = true ("tc32" mode): the contraction with B2 runs in FP32 on the FMA pipe, so the B2
// region holds plain FP32 values w[n] = [coef c'_0 .. coef c'_{D-1}, coef, 0] per node
// (read with warp-uniform, i.e. broadcast, LDS) instead of the TF32 MMA tile.  The contraction is
// bound by the LDS write-back into registers (D + 1 floats per lane and node), not by the FMA
// issue: packed fma.rn.f32x2 with node-pair interleaved rows was measured no faster / slower
// (26.9 k / 30.0 k vs 27.1 k cycles per evaluation at I = 1024), so the plain form stays and
// only the D + 1 used floats of a row are loaded.
constexpr int B2F_ROW = 12;
static_assert(TILE_N * B2F_ROW * 4 <= B2_TILE_BYTES && MAX_D + 1 <= B2F_ROW, "FP32 B2 rows fit the B2 tile");

template <int D, bool FMA_G2 = false>
__device__ __forceinline__ void build_tables(const Shared& s, const double* __restrict__ centers,
                                             const double* __restrict__ widths, const double* __restrict__ out_w,
                                             int nodes, int n_tiles) {
    const double L2E = 1.4426950408889634;
    for (int i = threadIdx.x; i < n_tiles * TILE_N; i += blockDim.x) {
        const int t = i / TILE_N, r = i % TILE_N;
        float v[K1];
#pragma unroll
        for (int k = 0; k < K1; ++k) v[k] = 0.f;
        float b2v[N2];
#pragma unroll
        for (int n = 0; n < N2; ++n) b2v[n] = 0.f;
        double vals[D + 2];
        if (i < nodes) {
            const double w = widths[i];
            const double h = 1.0 / (2.0 * (w * w));
            const double coef = out_w[i] / (w * w);
            double c2 = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const double c = centers[(size_t)i * D + k] - 0.5;
                c2 += c * c;
                vals[k] = 2.0 * h * L2E * c;
                b2v[k] = FMA_G2 ? (float)(coef * c) : __uint_as_float(to_tf32_rna((float)(coef * c)));
            }
            vals[D] = -h * L2E;
            vals[D + 1] = -h * L2E * c2;
            b2v[D] = FMA_G2 ? (float)coef : __uint_as_float(to_tf32_rna((float)coef));
        } else {  // padding node: S = -1e4 -> P = 0
#pragma unroll
            for (int k = 0; k < D + 1; ++k) vals[k] = 0.0;
            vals[D + 1] = -1.0e4;
        }
#pragma unroll
        for (int k = 0; k < D + 2; ++k) {
            const float hi = tf32_trunc((float)vals[k]);
            const float lo = (float)(vals[k] - (double)hi);
            v[k] = hi;                  // pairs with A hi
            v[(D + 2) + k] = lo;        // pairs with A hi
            v[2 * (D + 2) + k] = hi;    // pairs with A lo
        }
        float* b1t = s.b1 + (size_t)t * (B1_TILE_BYTES / 4);
#pragma unroll
        for (int c = 0; c < K1 / 4; ++c)
            *reinterpret_cast<float4*>(b1t + ((size_t)c * TILE_N + r) * 4) =
                make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        float* b2t = s.b2 + (size_t)t * (B2_TILE_BYTES / 4);
        if constexpr (FMA_G2) {
#pragma unroll
            for (int n = 0; n < B2F_ROW; ++n) b2t[r * B2F_ROW + n] = b2v[n];
        } else {
            const int chunk = r >> 2, within = r & 3;
#pragma unroll
            for (int n = 0; n < N2; ++n) b2t[((size_t)chunk * N2 + n) * 4 + within] = b2v[n];
        }
    }
}

// ---- MMA-issuer side: one gradient evaluation of one group over all node tiles (one elected thread).
// `tmem_base` and `bars` are the group's.
struct IssuerState {
    uint32_t ph_a = 0;
    uint32_t ph_p[N_SBUF] = {};
    uint32_t ph_g[N_SBUF] = {};
    int tr = 0;                 // trace cursor
    long long* trp = nullptr;   // trace buffer (NULL: not traced)
};

__device__ __forceinline__ void issue_g1(uint32_t tmem_base, uint32_t b1_saddr, int t, int n_cols, int dbg) {
    const uint32_t d = tmem_base + COL_S0 + (t % N_SBUF) * TILE_N;
    const uint32_t idesc = make_idesc(TILE_M, n_cols);
    const uint64_t bd0 = make_desc(b1_saddr + t * B1_TILE_BYTES, TILE_N * 16, 128);
    const uint32_t lo = (uint32_t)bd0, hi = (uint32_t)(bd0 >> 32);
    const uint32_t a = tmem_base + COL_A1;
    constexpr uint32_t STEP = (2 * TILE_N * 16) >> 4;   // two 16-byte K chunks per MMA, in descriptor units
    umma_tf32_ts_lh<0>(d, a, lo, hi, idesc);
    if (dbg & 8) return;
#pragma unroll
    for (int ks = 1; ks < K1 / 8; ++ks) umma_tf32_ts_lh<1>(d, a + ks * 8, lo + ks * STEP, hi, idesc);
}
// K-steps q, q + N_ISS, ... of GEMM-2 of tile t into issuer q's accumulator
__device__ __forceinline__ void issue_g2(uint32_t tmem_base, uint32_t b2_saddr, int t, int n_cols, bool first, int q, int dbg) {
    const uint32_t a = tmem_base + COL_S0 + (t % N_SBUF) * TILE_N;
    const uint32_t d = tmem_base + COL_D2 + q * N2;
    const uint32_t idesc = make_idesc(TILE_M, N2);
    const uint64_t bd0 = make_desc(b2_saddr + t * B2_TILE_BYTES, N2 * 16, 128);
    const uint32_t lo = (uint32_t)bd0, hi = (uint32_t)(bd0 >> 32);
    constexpr uint32_t STEP = (2 * N2 * 16) >> 4;
    const int nk = (dbg & 4) ? N_ISS : (n_cols >> 3);
#pragma unroll
    for (int j = 0; j < TILE_N / 8 / N_ISS; ++j) {
        const int ks = j * N_ISS + q;
        if (ks < nk) {
            if (first && j == 0) umma_tf32_ts_lh<0>(d, a + ks * 8, lo + ks * STEP, hi, idesc);
            else umma_tf32_ts_lh<1>(d, a + ks * 8, lo + ks * STEP, hi, idesc);
        }
    }
}

// node-tile widths: full tiles are TILE_N; the last one is padded up to a multiple of 16
// (padding nodes have P = 0 and B2 = 0)
__device__ __forceinline__ int tile_cols(int nodes, int t) {
    const int rem = nodes - t * TILE_N;
    return rem >= TILE_N ? TILE_N : ((rem + 15) & ~15);
}

__device__ __forceinline__ void issuer_eval(IssuerState& st, uint32_t bars, uint32_t tmem_base, uint32_t b1_saddr,
                                            uint32_t b2_saddr, int nodes, int n_tiles, int dbg) {
    HEPMC_TRACE(st, 4096, 10, -1);
    mbar_wait(bars + 8 * BAR_A, st.ph_a);
    st.ph_a ^= 1;
    fence_after();
    HEPMC_TRACE(st, 4096, 11, -1);
#pragma unroll
    for (int b = 0; b < N_SBUF; ++b) {
        if (b < n_tiles) {
            issue_g1(tmem_base, b1_saddr, b, tile_cols(nodes, b), dbg);
            umma_commit(bars + 8 * (BAR_S0 + b));
            HEPMC_TRACE(st, 4096, 12, b);
        }
    }
    for (int t = 0; t < n_tiles; ++t) {
        const int b = t % N_SBUF;
        mbar_wait(bars + 8 * (BAR_P0 + b), st.ph_p[b]);
        st.ph_p[b] ^= 1;
        fence_after();
        HEPMC_TRACE(st, 4096, 13, t);
        if (!(dbg & 1)) issue_g2(tmem_base, b2_saddr, t, tile_cols(nodes, t), t == 0, 0, dbg);
        HEPMC_TRACE(st, 4096, 14, t);
        if (N_ISS == 1) {
            if (t + N_SBUF < n_tiles) {   // same thread: G1(t+3) is ordered behind G2(t), which reads S[b]
                issue_g1(tmem_base, b1_saddr, t + N_SBUF, tile_cols(nodes, t + N_SBUF), dbg);
                umma_commit(bars + 8 * (BAR_S0 + b));
                HEPMC_TRACE(st, 4096, 12, t + N_SBUF);
            }
        } else if (t >= 1 && t - 1 + N_SBUF < n_tiles) {
            // refill the buffer of the PREVIOUS tile: the other issuer's reads of it were committed a
            // tile ago, so this wait does not stall the issue stream
            const int pb = (t - 1) % N_SBUF;
            mbar_wait(bars + 8 * (BAR_G0 + pb), st.ph_g[pb]);
            st.ph_g[pb] ^= 1;
            fence_after();
            issue_g1(tmem_base, b1_saddr, t - 1 + N_SBUF, tile_cols(nodes, t - 1 + N_SBUF), dbg);
            umma_commit(bars + 8 * (BAR_S0 + pb));
            HEPMC_TRACE(st, 4096, 12, t - 1 + N_SBUF);
        }
    }
    umma_commit(bars + 8 * BAR_D2);
    HEPMC_TRACE(st, 4096, 15, -1);
}

// "tc32" mode: only GEMM-1 is issued; BAR_P0 + b means "S[b] has been read into registers"
__device__ __forceinline__ void issuer_eval_g1(IssuerState& st, uint32_t bars, uint32_t tmem_base, uint32_t b1_saddr,
                                               int nodes, int n_tiles, int dbg) {
    mbar_wait(bars + 8 * BAR_A, st.ph_a);
    st.ph_a ^= 1;
    fence_after();
#pragma unroll
    for (int b = 0; b < N_SBUF; ++b) {
        if (b < n_tiles) {
            issue_g1(tmem_base, b1_saddr, b, tile_cols(nodes, b), dbg);
            umma_commit(bars + 8 * (BAR_S0 + b));
        }
    }
    for (int t = 0; t < n_tiles; ++t) {
        const int b = t % N_SBUF;
        mbar_wait(bars + 8 * (BAR_P0 + b), st.ph_p[b]);
        st.ph_p[b] ^= 1;
        if (t + N_SBUF < n_tiles) {
            fence_after();
            issue_g1(tmem_base, b1_saddr, t + N_SBUF, tile_cols(nodes, t + N_SBUF), dbg);
            umma_commit(bars + 8 * (BAR_S0 + b));
        }
    }
}

// issuers 1 .. N_ISS-1 of a group: their share of GEMM-2
__device__ __forceinline__ void aux_issuer_eval(IssuerState& st, uint32_t bars, uint32_t tmem_base, uint32_t b2_saddr,
                                                int nodes, int n_tiles, int q, int dbg) {
    for (int t = 0; t < n_tiles; ++t) {
        const int b = t % N_SBUF;
        mbar_wait(bars + 8 * (BAR_P0 + b), st.ph_p[b]);
        st.ph_p[b] ^= 1;
        fence_after();
        if (!(dbg & 1)) issue_g2(tmem_base, b2_saddr, t, tile_cols(nodes, t), t == 0, q, dbg);
        if (t + N_SBUF < n_tiles) umma_commit(bars + 8 * (BAR_G0 + b));   // S[b] will be refilled
    }
    umma_commit(bars + 8 * BAR_D2);
}

// ---- compute side: thread = chain = TMEM lane; `bars` and the column part of `tmem_lane_base` are the group's ----
struct ComputeState {
    uint32_t ph_s[N_SBUF] = {};
    uint32_t ph_d2 = 0;
    int tr = 0;                 // trace cursor
    long long* trp = nullptr;   // trace buffer (NULL: not traced)
};

// Plain form of exp_tiles (one tile after the other, every wait exposed): kept for the development
// knobs (skip exp / MUFU without TMEM) and as the readable statement of the protocol.
__device__ __forceinline__ void exp_tiles_simple(ComputeState& st, uint32_t bars, uint32_t tmem_lane_base, int nodes,
                                          int n_tiles, int dbg) {
    for (int t = 0; t < n_tiles; ++t) {
        const int b = t % N_SBUF;
        const uint32_t col = COL_S0 + b * TILE_N;
        HEPMC_TRACE(st, 0, 1, t);
        mbar_wait(bars + 8 * (BAR_S0 + b), st.ph_s[b]);
        st.ph_s[b] ^= 1;
        fence_after();
        HEPMC_TRACE(st, 0, 2, t);
        if (!(dbg & 2)) {
            // software pipeline over 16-column chunks: the TMEM read of chunk c+1 is in flight
            // while the MUFU works on chunk c
            const int nch = tile_cols(nodes, t) >> 4;
            uint32_t va[16], vb[16];
            if (dbg & 16) {  // experiment: MUFU work without touching TMEM
#pragma unroll
                for (int j = 0; j < 16; ++j) { va[j] = threadIdx.x + j; vb[j] = threadIdx.x * 3 + j; }
                for (int c = 0; c < nch; ++c) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) va[j] = __float_as_uint(ex2_approx(__uint_as_float(va[j] ^ vb[j])));
                }
                if (va[3] == 0x12345u) tmem_st16(tmem_lane_base + col, va);
            } else {
                tmem_ld16(tmem_lane_base + col, va);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < TILE_N / 16; ++c) {
                    if (c < nch) {
                        if (c & 1) {
                            if (c + 1 < nch) tmem_ld16(tmem_lane_base + col + (c + 1) * 16, va);
#pragma unroll
                            for (int j = 0; j < 16; ++j) vb[j] = __float_as_uint(ex2_approx(__uint_as_float(vb[j])));
                            tmem_st16(tmem_lane_base + col + c * 16, vb);
                        } else {
                            if (c + 1 < nch) tmem_ld16(tmem_lane_base + col + (c + 1) * 16, vb);
#pragma unroll
                            for (int j = 0; j < 16; ++j) va[j] = __float_as_uint(ex2_approx(__uint_as_float(va[j])));
                            tmem_st16(tmem_lane_base + col + c * 16, va);
                        }
                        if (c + 1 < nch) tmem_wait_ld();
                    }
                }
            }
            tmem_wait_st();
        }
        fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(bars + 8 * (BAR_P0 + b));  // one arrival per warp
        HEPMC_TRACE(st, 0, 3, t);
    }
}

// S -> P = exp2(S) in place for every node tile of one gradient evaluation (64 columns per tile
// and TMEM lane), as ONE stream of 16-column chunks across tile boundaries: the TMEM read of the
// next chunk, the poll of the next tile's S_full and the wait::st / P_ready arrival of the
// previous tile are all issued right after a batch of 16 MUFU.EX2 (128 cycles of MUFU time), so
// none of those latencies is exposed (measured with the trace: ~400 cycles per tile before).
__device__ __forceinline__ void exp_tiles(ComputeState& st, uint32_t bars, uint32_t tmem_lane_base, int nodes,
                                          int n_tiles, int dbg) {
    if (dbg & 18) {
        exp_tiles_simple(st, bars, tmem_lane_base, nodes, n_tiles, dbg);
        return;
    }
    uint32_t va[16], vb[16];
    HEPMC_TRACE(st, 0, 1, 0);
    mbar_wait(bars + 8 * BAR_S0, st.ph_s[0]);
    st.ph_s[0] ^= 1;
    fence_after();
    HEPMC_TRACE(st, 0, 2, 0);
    tmem_ld16(tmem_lane_base + COL_S0, va);
    int prev_b = -1;   // tile whose P_ready arrival is still owed
    for (int t = 0; t < n_tiles; ++t) {
        const int b = t % N_SBUF;
        const uint32_t col = COL_S0 + b * TILE_N;
        const int nch = tile_cols(nodes, t) >> 4;
        const bool more = t + 1 < n_tiles;        // full tile (nch == 4) whenever another one follows
        const int nb = (t + 1) % N_SBUF;
#pragma unroll
        for (int c = 0; c < TILE_N / 16; ++c) {
            if (c < nch) {
                uint32_t(&cur)[16] = (c & 1) ? vb : va;
                uint32_t(&oth)[16] = (c & 1) ? va : vb;
                tmem_wait_ld();
                if (c + 1 < nch) tmem_ld16(tmem_lane_base + col + (c + 1) * 16, oth);
                ex2_chunk(cur);
                // --- in the shadow of that batch ---
                if (c == 0 && prev_b >= 0) {     // previous tile: stores done -> P_ready
                    tmem_wait_st();
                    fence_before();
                    __syncwarp();
                    if ((threadIdx.x & 31) == 0) mbar_arrive(bars + 8 * (BAR_P0 + prev_b));
                    HEPMC_TRACE(st, 0, 3, t - 1);
                }
                if (c + 1 == nch && more) {      // next tile: S_full -> first chunk on its way
                    mbar_wait(bars + 8 * (BAR_S0 + nb), st.ph_s[nb]);
                    st.ph_s[nb] ^= 1;
                    fence_after();
                    HEPMC_TRACE(st, 0, 2, t + 1);
                    tmem_ld16(tmem_lane_base + COL_S0 + nb * TILE_N, oth);
                }
                tmem_st16(tmem_lane_base + col + c * 16, cur);
            }
        }
        prev_b = b;
    }
    tmem_wait_st();
    fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bars + 8 * (BAR_P0 + prev_b));
    HEPMC_TRACE(st, 0, 3, n_tiles - 1);
}

// "tc32" mode: P = exp2(S) stays in registers and is contracted with the FP32 B2 rows on the FMA
// pipe (D + 1 FFMA per pair): acc[n] += P_ji * b2f[i][n].  Same chunk stream as exp_tiles; a tile's
// S buffer is released (BAR_P0 + b) as soon as its last chunk is in registers.
template <int D>
__device__ __forceinline__ void exp_fma_tiles(ComputeState& st, uint32_t bars, uint32_t tmem_lane_base,
                                              const float* __restrict__ b2f, int nodes, int n_tiles, float (&acc)[D + 1]) {
#pragma unroll
    for (int n = 0; n <= D; ++n) acc[n] = 0.f;
    uint32_t va[16], vb[16];
    mbar_wait(bars + 8 * BAR_S0, st.ph_s[0]);
    st.ph_s[0] ^= 1;
    fence_after();
    tmem_ld16(tmem_lane_base + COL_S0, va);
    for (int t = 0; t < n_tiles; ++t) {
        const int b = t % N_SBUF;
        const uint32_t col = COL_S0 + b * TILE_N;
        const int nch = tile_cols(nodes, t) >> 4;
        const bool more = t + 1 < n_tiles;
        const int nb = (t + 1) % N_SBUF;
        const float* rows = b2f + (size_t)t * (B2_TILE_BYTES / 4);
#pragma unroll
        for (int c = 0; c < TILE_N / 16; ++c) {
            if (c < nch) {
                uint32_t(&cur)[16] = (c & 1) ? vb : va;
                uint32_t(&oth)[16] = (c & 1) ? va : vb;
                tmem_wait_ld();
                if (c + 1 < nch) {
                    tmem_ld16(tmem_lane_base + col + (c + 1) * 16, oth);
                } else {
                    // the whole tile is in registers: release its buffer, then start on the next tile
                    fence_before();
                    __syncwarp();
                    if ((threadIdx.x & 31) == 0) mbar_arrive(bars + 8 * (BAR_P0 + b));
                    if (more) {
                        mbar_wait(bars + 8 * (BAR_S0 + nb), st.ph_s[nb]);
                        st.ph_s[nb] ^= 1;
                        fence_after();
                        tmem_ld16(tmem_lane_base + COL_S0 + nb * TILE_N, oth);
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float p = ex2_approx(__uint_as_float(cur[j]));
                    const float* row = rows + (c * 16 + j) * B2F_ROW;
                    float w[B2F_ROW];
#pragma unroll
                    for (int q = 0; q < (D + 1) / 4; ++q) {
                        const float4 v = reinterpret_cast<const float4*>(row)[q];
                        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
                    }
#pragma unroll
                    for (int n = ((D + 1) / 4) * 4; n <= D; ++n) w[n] = row[n];
#pragma unroll
                    for (int n = 0; n <= D; ++n) acc[n] = fmaf(p, w[n], acc[n]);
                }
            }
        }
    }
}

template <int D, bool FMA_G2 = false>
__device__ __forceinline__ void compute_eval(ComputeState& st, uint32_t bars, uint32_t tmem_lane_base, int nodes,
                                             int n_tiles, int dbg, const double (&x)[D], double (&g)[D],
                                             const float* __restrict__ b2f = nullptr) {
    // A1 row: [hi(a), hi(a), lo(a)] with a = [x', |x'|^2, 1]; single precision is enough for the
    // split (hi + lo carries 22 bits), the double -> float conversion happens once per coordinate
    uint32_t a1[32];
    float xc[D];
    {
        float av[D + 2];
        float r2 = 0.f;
        bool finite = true;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            av[k] = (float)(x[k] - 0.5);
            finite = finite && (fabsf(av[k]) < 1.0e6f);
            r2 = fmaf(av[k], av[k], r2);
        }
        av[D] = r2;
        av[D + 1] = 1.f;
        // non-finite / runaway positions (diverged trajectories) are mapped to a far-away finite
        // point: every exponent underflows, the gradient is 0 and the MH test rejects on pdf = 0
        if (!finite) {
#pragma unroll
            for (int k = 0; k < D; ++k) av[k] = 0.f;
            av[D] = 1.0e12f;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) xc[k] = av[k];
#pragma unroll
        for (int k = 0; k < 32; ++k) a1[k] = 0u;
#pragma unroll
        for (int k = 0; k < D + 2; ++k) {
            const float hi = tf32_trunc(av[k]);
            const float lo = av[k] - hi;
            a1[k] = __float_as_uint(hi);
            a1[(D + 2) + k] = __float_as_uint(hi);
            a1[2 * (D + 2) + k] = __float_as_uint(lo);
        }
    }
    HEPMC_TRACE(st, 0, 4, -1);
    tmem_st32(tmem_lane_base + COL_A1, a1);
    tmem_wait_st();
    fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bars + 8 * BAR_A);  // one arrival per warp
    HEPMC_TRACE(st, 0, 5, -1);
    if constexpr (FMA_G2) {
        float accf[D + 1];
        exp_fma_tiles<D>(st, bars, tmem_lane_base, b2f, nodes, n_tiles, accf);
#pragma unroll
        for (int k = 0; k < D; ++k) g[k] = (double)(accf[k] - xc[k] * accf[D]);
        return;
    }
    exp_tiles(st, bars, tmem_lane_base, nodes, n_tiles, dbg);
    mbar_wait(bars + 8 * BAR_D2, st.ph_d2);
    st.ph_d2 ^= 1;
    fence_after();
    HEPMC_TRACE(st, 0, 6, -1);
    uint32_t d2[N_ISS][16];
#pragma unroll
    for (int q = 0; q < N_ISS; ++q) tmem_ld16(tmem_lane_base + COL_D2 + q * N2, d2[q]);
    tmem_wait_ld();
    float acc[D + 1];
#pragma unroll
    for (int k = 0; k <= D; ++k) {
        acc[k] = __uint_as_float(d2[0][k]);
#pragma unroll
        for (int q = 1; q < N_ISS; ++q) acc[k] += __uint_as_float(d2[q][k]);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) g[k] = (double)(acc[k] - xc[k] * acc[D]);
}

}  // namespace tc
}  // namespace hepmc

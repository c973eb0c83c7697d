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
// (B1: 16 KB, B2: 8 KB per 128 nodes).  One CTA = 128 chains (TMEM lanes) = 4 compute warps +
// 4 helper warps (upper half of the S columns) + 1 MMA-issuing warp; S is double buffered in TMEM so GEMM-1 of node tile t+1 / GEMM-2 of tile t
// overlap the exponentials of the tiles in between.  mbarrier pipeline:
//   A_ready (128 arrivals) -> [G1(0), G1(1)] -> S_full[b] (commit) -> exp -> P_ready[b] (256)
//   -> G2(t), G1(t+2) ... -> D2_full (commit) -> epilogue.
#pragma once
#include "common.cuh"

namespace hepmc {
namespace tc {

constexpr int TILE_M = 128;      // chains per CTA tile (TMEM lanes)
constexpr int TILE_N = 128;      // nodes per tile
constexpr int K1 = 32;           // 3 * (D + 2) padded to a multiple of 8, D <= 8
constexpr int N2 = 16;           // D + 1 padded to the MMA N granularity at M = 128
constexpr int MAX_D = 8;
constexpr int N_ACC = 1;          // GEMM-2 accumulators (round-robin over 2 or 4 measured: no gain, a tcgen05.mma costs ~100 cycles here)
constexpr int N_SBUF = 3;         // S/P tiles in flight: the commit -> mbarrier -> waiter round trip is ~1000 cycles
constexpr int COL_A1 = 0, COL_D2 = 32, COL_S0 = 64, TMEM_COLS = 512;   // S buffer b at COL_S0 + 128 b
constexpr int B1_TILE_BYTES = (K1 / 4) * TILE_N * 16;   // [K chunk][node row][16 B]
constexpr int B2_TILE_BYTES = (TILE_N / 4) * N2 * 16;   // [node chunk][n row][16 B]
constexpr int N_OWNER = TILE_M;                         // owner threads: one chain = one TMEM lane each
constexpr int N_COMPUTE = 2 * TILE_M;                   // + helper warps: same lanes, upper half of the S columns
constexpr int N_THREADS = N_COMPUTE + 32;               // + MMA / allocator warp
constexpr int MMA_WARP = N_COMPUTE / 32;

__host__ __device__ inline size_t smem_bytes(int nodes) {
    const int nt = (nodes + TILE_N - 1) / TILE_N;
    return (size_t)nt * (B1_TILE_BYTES + B2_TILE_BYTES) + 128;
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]^T, TF32 operands, FP32 accumulate
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// Same, descriptor passed as (lo, hi) words and the accumulate flag as a compile-time constant:
// the issuing thread is a single lane, so every scalar instruction per MMA is ~5 cycles of
// issue latency -- the descriptor is advanced with ONE 32-bit add per MMA.
template <int ACC>
__device__ __forceinline__ void umma_tf32_ts_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi,
                                                uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 bd;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 bd, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n"
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
// 8 = GEMM-1 with 1 MMA per tile, 16 = exp stage without TMEM access
__device__ int g_tc_debug = 0;

struct Shared {
    float* b1;           // [n_tiles][K1/4][TILE_N][4]
    float* b2;           // [n_tiles][TILE_N/4][N2][4]
    uint64_t* bars;      // A_ready, S_full[2], P_ready[2], D2_full
    uint32_t* tmem_ptr;
};

__device__ __forceinline__ Shared carve_tc(unsigned char* base, int n_tiles) {
    Shared s;
    s.b1 = reinterpret_cast<float*>(base);
    s.b2 = reinterpret_cast<float*>(base + (size_t)n_tiles * B1_TILE_BYTES);
    s.bars = reinterpret_cast<uint64_t*>(base + (size_t)n_tiles * (B1_TILE_BYTES + B2_TILE_BYTES));
    s.tmem_ptr = reinterpret_cast<uint32_t*>(s.bars + 2 + 2 * N_SBUF);
    return s;
}

enum { BAR_A = 0, BAR_D2 = 1, BAR_S0 = 2, BAR_P0 = 2 + N_SBUF };   // S_full[b] = BAR_S0 + b, P_ready[b] = BAR_P0 + b

// Build the resident node tables (all threads).  centers (I, D), widths (I,), out_w (I,).
template <int D>
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
                b2v[k] = __uint_as_float(to_tf32_rna((float)(coef * c)));
            }
            vals[D] = -h * L2E;
            vals[D + 1] = -h * L2E * c2;
            b2v[D] = __uint_as_float(to_tf32_rna((float)coef));
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
        const int chunk = r >> 2, within = r & 3;
#pragma unroll
        for (int n = 0; n < N2; ++n) b2t[((size_t)chunk * N2 + n) * 4 + within] = b2v[n];
    }
}

// ---- MMA-issuer side: one gradient evaluation over all node tiles (one elected thread) ----
struct IssuerState {
    uint32_t ph_a = 0;
    uint32_t ph_p[N_SBUF] = {};
};

__device__ __forceinline__ void issue_g1(uint32_t tmem_base, uint32_t b1_saddr, int t, int n_cols) {
    const uint32_t d = tmem_base + COL_S0 + (t % N_SBUF) * TILE_N;
    const uint32_t idesc = make_idesc(TILE_M, n_cols);
    const uint64_t bd0 = make_desc(b1_saddr + t * B1_TILE_BYTES, TILE_N * 16, 128);
    const uint32_t lo = (uint32_t)bd0, hi = (uint32_t)(bd0 >> 32);
    const uint32_t a = tmem_base + COL_A1;
    constexpr uint32_t STEP = (2 * TILE_N * 16) >> 4;   // two 16-byte K chunks per MMA, in descriptor units
    umma_tf32_ts_lh<0>(d, a, lo, hi, idesc);
    if (g_tc_debug & 8) return;
#pragma unroll
    for (int ks = 1; ks < K1 / 8; ++ks) umma_tf32_ts_lh<1>(d, a + ks * 8, lo + ks * STEP, hi, idesc);
}
__device__ __forceinline__ void issue_g2(uint32_t tmem_base, uint32_t b2_saddr, int t, int n_cols, bool first) {
    const uint32_t a = tmem_base + COL_S0 + (t % N_SBUF) * TILE_N;
    const uint32_t d = tmem_base + COL_D2;
    const uint32_t idesc = make_idesc(TILE_M, N2);
    const uint64_t bd0 = make_desc(b2_saddr + t * B2_TILE_BYTES, N2 * 16, 128);
    const uint32_t lo = (uint32_t)bd0, hi = (uint32_t)(bd0 >> 32);
    constexpr uint32_t STEP = (2 * N2 * 16) >> 4;
    const int nk = (g_tc_debug & 4) ? 2 : (n_cols >> 3);
#pragma unroll
    for (int ks = 0; ks < TILE_N / 8; ++ks) {
        if (ks < nk) {
            const uint32_t dq = d + (ks % N_ACC) * N2;
            if (first && ks < N_ACC) umma_tf32_ts_lh<0>(dq, a + ks * 8, lo + ks * STEP, hi, idesc);
            else umma_tf32_ts_lh<1>(dq, a + ks * 8, lo + ks * STEP, hi, idesc);
        }
    }
}

// node-tile widths: full tiles are TILE_N; the last one is padded up to a multiple of 16
__device__ __forceinline__ int tile_cols(int nodes, int t) {
    const int rem = nodes - t * TILE_N;
    return rem >= TILE_N ? TILE_N : ((rem + 15) & ~15);
}

__device__ __forceinline__ void issuer_eval(IssuerState& st, uint32_t bars, uint32_t tmem_base, uint32_t b1_saddr,
                                            uint32_t b2_saddr, int nodes, int n_tiles) {
    mbar_wait(bars + 8 * BAR_A, st.ph_a);
    st.ph_a ^= 1;
    fence_after();
#pragma unroll
    for (int b = 0; b < N_SBUF; ++b) {
        if (b < n_tiles) {
            issue_g1(tmem_base, b1_saddr, b, tile_cols(nodes, b));
            umma_commit(bars + 8 * (BAR_S0 + b));
        }
    }
    for (int t = 0; t < n_tiles; ++t) {
        const int b = t % N_SBUF;
        mbar_wait(bars + 8 * (BAR_P0 + b), st.ph_p[b]);
        st.ph_p[b] ^= 1;
        fence_after();
        if (!(g_tc_debug & 1)) issue_g2(tmem_base, b2_saddr, t, tile_cols(nodes, t), t == 0);
        if (t + N_SBUF < n_tiles) {
            issue_g1(tmem_base, b1_saddr, t + N_SBUF, tile_cols(nodes, t + N_SBUF));
            umma_commit(bars + 8 * (BAR_S0 + b));
        }
    }
    umma_commit(bars + 8 * BAR_D2);
}

// ---- compute side: thread = chain = TMEM lane ----
struct ComputeState {
    uint32_t ph_s[N_SBUF] = {};
    uint32_t ph_d2 = 0;
};

// S -> P = exp2(S) in place for one half (64 columns) of every node tile of one gradient
// evaluation.  half 0: owner warps, half 1: helper warps (same TMEM lanes, warp % 4 equal).
__device__ __forceinline__ void exp_half_tiles(ComputeState& st, uint32_t bars, uint32_t tmem_lane_base, int nodes,
                                               int n_tiles, int half) {
    for (int t = 0; t < n_tiles; ++t) {
        const int b = t % N_SBUF;
        const uint32_t col = COL_S0 + b * TILE_N + half * 64;
        mbar_wait(bars + 8 * (BAR_S0 + b), st.ph_s[b]);
        st.ph_s[b] ^= 1;
        fence_after();
        const int my_cols = tile_cols(nodes, t) - half * 64;   // columns of this half that GEMM-2 reads
        if (my_cols > 0 && !(g_tc_debug & 2)) {
            // software pipeline over 16-column chunks: the TMEM read of chunk c+1 is in flight
            // while the MUFU works on chunk c (TMEM read and MUFU both run at 4 values/clk/SMSP)
            const int nch = my_cols >= 64 ? 4 : (my_cols >> 4);
            uint32_t va[16], vb[16];
            if (g_tc_debug & 16) {  // experiment: MUFU work without touching TMEM
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
            for (int c = 0; c < 4; ++c) {
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
    }
}

template <int D>
__device__ __forceinline__ void compute_eval(ComputeState& st, uint32_t bars, uint32_t tmem_lane_base, int nodes,
                                             int n_tiles, const double (&x)[D], double (&g)[D]) {
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
    tmem_st32(tmem_lane_base + COL_A1, a1);
    tmem_wait_st();
    fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bars + 8 * BAR_A);  // one arrival per warp
    exp_half_tiles(st, bars, tmem_lane_base, nodes, n_tiles, 0);
    mbar_wait(bars + 8 * BAR_D2, st.ph_d2);
    st.ph_d2 ^= 1;
    fence_after();
    uint32_t d2[N_ACC][16];
#pragma unroll
    for (int q = 0; q < N_ACC; ++q) tmem_ld16(tmem_lane_base + COL_D2 + q * N2, d2[q]);
    tmem_wait_ld();
    float acc[D + 1];
#pragma unroll
    for (int k = 0; k <= D; ++k) {
        acc[k] = __uint_as_float(d2[0][k]);
#pragma unroll
        for (int q = 1; q < N_ACC; ++q) acc[k] += __uint_as_float(d2[q][k]);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) g[k] = (double)(acc[k] - xc[k] * acc[D]);
}

}  // namespace tc
}  // namespace hepmc

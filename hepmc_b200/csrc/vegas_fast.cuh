// Throughput kernel of the VEGAS iteration (vegas.py:100-111 fused with
// stratified_volume.py:104-122 and vegas.py:68-72): native Philox draws, compile-time
// dimensionality and integrand, one CTA per chunk, deterministic partial rows.
//
// Round 2 rewrite of the hot loop, driven by ncu (profiles/r02_vegas_*):
//   * RNG mode is a template parameter (hepmc_set_rng_mode): 64 or 32 random bits per axis,
//     Philox4x32-10 or -7.  32-bit law: t = K*w (one IMAD.WIDE): bin = t >> 32, offset =
//     ((t mod 2^32) + 1/2) 2^-32, built with one DADD from the raw word (exponent trick).
//     Round keys come expanded in the kernel parameters (constant-bank operands).
//   * shared memory is the binding resource once the integer work is cut (one 128-byte wavefront
//     per cycle and SM): for K <= 16 the grid is staged as separate rows of bounds and sizes,
//     16 doubles = 128 B each, so a random-bin look-up never has a bank conflict (the
//     (bound, size) pair layout cost 0.8 extra wavefronts per look-up).
//   * histogram: bins of a warp tile are staged word-major ([4 axes][32 points], rows padded
//     to 36 words), the lane that owns axis d fetches the bins of 4 points with one LDS.128
//     and the f*w values of 2 points with one LDS.128; for K <= 16 every warp owns 16 rows of
//     the CTA's histogram and the staged byte is warp*16 + bin, so the column address
//     (byte << 8 | lane << 3) is ONE byte-permute and the load/store use [R + UR] addressing:
//     4 instructions per (point, axis): PRMT, LDS, DADD, STS (round 1: 10).
//   * camel in centred form (3 FP64 instructions per axis instead of 4, see below); the open-cube
//     test is an integer max over the high words of x (ALU pipe instead of 16 DSETP).
//   * 32-bit modes, K <= 16: the offset word becomes a double by ONE conversion on the idle XU pipe
//     (I2F.F64.U32), the 2^-32 scale and the half-quantum shift live in the staged table; the
//     integrand's exponentials use a 64-entry table of 2^(j/64) + a degree-5 polynomial (Tang's method).
//   * full chunks run a loop without bounds checks; the shift of the single-pass variance
//     accumulator is the thread's first value, counts are implied.
// Determinism: the point -> thread -> lane order is fixed by (chunk, NDIM) only, every sum has
// a fixed order, there are no atomics: any partition of the chunk range over GPUs gives
// bit-identical partial rows.
#pragma once
// included by integrate_kernels.cuh (SampleArgs, Philox modes, bin_frac32)

namespace hepmc {

#ifndef HEPMC_VFAST_MINBLOCKS
#define HEPMC_VFAST_MINBLOCKS 3
#endif
#ifndef HEPMC_VFAST_OUT_MINBLOCKS
#define HEPMC_VFAST_OUT_MINBLOCKS 3
#endif
#ifndef HEPMC_VFAST_STAGGER_NS
#define HEPMC_VFAST_STAGGER_NS 400   // start offset between the CTAs sharing an SM (0 = none); 200, 400, 1000 measured alike
                                     // (22.15 ms per 1e9 points; powers of two -- a shift instead of a multiply -- 22.23; none 22.45)
#endif
#ifndef HEPMC_VFAST_TABEXP
#define HEPMC_VFAST_TABEXP 1   // the integrand's exponentials by a 64-entry table of 2^(j/64) + degree-5 polynomial
#endif
#ifndef HEPMC_VFAST_I2F
#define HEPMC_VFAST_I2F 1      // 32-bit modes, K <= 16: offset word -> double by I2F (XU pipe), scaled table (b + s 2^-33, s 2^-32)
#endif
template <int NDIM>
struct HistGeom {
    static constexpr int ngrp = 32 / NDIM;                    // lane groups that could work in parallel
    static constexpr int need = (32 + ngrp - 1) / ngrp;       // points per group
    static constexpr int NUPD = need <= 4 ? 4 : need <= 8 ? 8 : need <= 16 ? 16 : 32;  // rounded: LDS.128 of 4 points
    static constexpr int NG = 32 / NUPD;                      // active groups (NG * NDIM <= 32)
    static constexpr int NQ = (NDIM + 3) / 4;                 // words of packed bins per point
    static constexpr int ROW = 36;                            // padded row of the word-major tile (bank spread)
};

__host__ __device__ inline int vegas_fast_ks(int K) { return K <= 16 ? 16 : K; }
// histogram rows per warp: fixed 16 for K <= 16 (the staged byte carries warp*16 + bin), else K
__host__ __device__ inline int vegas_fast_rows(int K) { return K <= 16 ? 16 : K; }

__host__ __device__ inline size_t vegas_fast_smem_bytes(int ndim, int K, int threads) {
    const int nw = threads / 32;
    const int nq = (ndim + 3) / 4;
    size_t b = 64 * 8;                                 // 2^(j/T) table of the exponentials (up to 64 entries), first
    b += (size_t)ndim * vegas_fast_ks(K) * 16;         // grid (bound, size) pairs / rows
    b += 8 * sizeof(Stats);                            // block merge scratch
    b += (size_t)nw * vegas_fast_rows(K) * 32 * 8;   // private histogram columns
    b += (size_t)nw * 32 * 8;                          // f*w tile
    b += (size_t)nw * nq * 36 * 4;                     // packed-bins tile
    return (b + 15) & ~(size_t)15;
}

// Camel in "centred" form (means equal on every axis, the reference's only use: camel.py:13-18):
// with c = (a + b)/2 and h = (b - a)/2,
//   sum (x-a)^2 = S2 + 2 h S1 + n h^2,  sum (x-b)^2 = S2 - 2 h S1 + n h^2,  S1 = sum (x-c), S2 = sum (x-c)^2:
// 3 FP64 instructions per axis instead of 4.  The cancellation costs a few ulp of max(S2, n h^2) ~ 1:
// relative error of the pdf <~ 1e-14 (tests gate the native path at 1e-12 against the oracle).
// Camels with different means per axis take the generic kernel.
__host__ inline bool camel_uniform_means(const Dens& p) {
    for (int d = 1; d < p.ndim; ++d)
        if (p.a[d] != p.a[0] || p.b[d] != p.b[0]) return false;
    return true;
}

#ifndef HEPMC_VFAST_EXPTAB_BITS
#define HEPMC_VFAST_EXPTAB_BITS 6      // 64 entries; 16 (one conflict-free 128-byte row, degree-7 polynomial) measured the same
#endif
constexpr int kExpBits = HEPMC_VFAST_EXPTAB_BITS;
constexpr int kExpTab = 1 << kExpBits;

// exp(x) by Tang's table method: x = (T k + j) ln2/T + r, |r| <= ln2/(2T); exp(x) = 2^k * tab[j] * e^r,
// tab[j] = 2^(j/T) in shared memory.  T = 16: degree-7 polynomial (r^8/8! < 2e-18); T = 64: degree 5.
// Valid for |x| < 708 (normal results); callers test exp_tab_ok on the arguments and take the library otherwise.
__device__ __forceinline__ double exp_tab(double x, const double* tab) {
    const double MAGIC = 6755399441055744.0;                       // 1.5 * 2^52
    constexpr double INV = kExpBits == 4 ? 23.083120654223414 : 92.332482616893657;          // T / ln 2
    constexpr double HI = kExpBits == 4 ? 0.04332169877307024 : 0.01083042469326756;         // ln2/T, top 32 mantissa bits
    constexpr double LO = kExpBits == 4 ? 1.1926343307941173e-11 : 2.9815858269852933e-12;   // ln2/T - HI
    const double t = fma(x, INV, MAGIC);
    const int n = __double2loint(t);
    const double nd = t - MAGIC;
    double r = fma(nd, -HI, x);                                    // nd * HI is exact
    r = fma(nd, -LO, r);
    double p;
    if constexpr (kExpBits == 4) {
        p = fma(r, 1.0 / 5040.0, 1.0 / 720.0);
        p = fma(r, p, 1.0 / 120.0);
        p = fma(r, p, 1.0 / 24.0);
    } else {
        p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    }
    p = fma(r, p, 1.0 / 6.0);
    p = fma(r, p, 0.5);
    p = fma(r * r, p, r);                                          // e^r - 1
    const double tj = tab[n & (kExpTab - 1)];
    const double v = fma(tj, p, tj);
    const int k = n >> kExpBits;
    return __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
}

// |x| < 708 (false for NaN and inf): one mask and one integer compare on the high word
__device__ __forceinline__ bool exp_tab_ok(double x) { return (__double2hiint(x) & 0x7fffffff) < 0x40862000; }

// KP: compile-time grid row stride (16) or 0 = runtime stride K
template <int KIND, int NDIM, int RNG, bool OUT, int KP>
__global__ void __launch_bounds__(256, OUT ? HEPMC_VFAST_OUT_MINBLOCKS : HEPMC_VFAST_MINBLOCKS) vegas_fast_kernel(const __grid_constant__ Dens dens,
                                                                               const __grid_constant__ SampleArgs a) {
    extern __shared__ __align__(128) double smem_raw[];
    using G = HistGeom<NDIM>;
    constexpr bool fused = KIND != HEPMC_NONE;
    constexpr bool camel = KIND == HEPMC_CAMEL || KIND == HEPMC_UCAMEL;
    constexpr int NQ = G::NQ;
    constexpr bool ROWS16 = KP == 16;
    const int K = a.K;
    const int KS = KP ? KP : K;
    const int RW = ROWS16 ? 16 : K;                      // histogram rows per warp
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int nw = 8;

    // the exponential's table first (fixed offset: the look-ups stay LDS -- an address rounded up through an integer
    // cast made the compiler emit generic LD.E at 4.9 wavefronts per look-up), then the grid rows (128-byte aligned)
    double* s_exp = smem_raw;                                                      // [kExpTab]
    double2* s_grid2 = reinterpret_cast<double2*>(smem_raw + kExpTab);             // [NDIM][KS]
    Stats* s_stats = reinterpret_cast<Stats*>(s_grid2 + (size_t)NDIM * KS);
    double* s_hist = reinterpret_cast<double*>(s_stats + 8);                       // [nw][RW][32]
    double* s_wf = s_hist + (size_t)nw * RW * 32;                           // [nw][32]
    uint32_t* s_bins = reinterpret_cast<uint32_t*>(s_wf + nw * 32);                // [nw][NQ][36]
    constexpr bool TABEXP = HEPMC_VFAST_TABEXP && (camel || KIND == HEPMC_BANANA || KIND == HEPMC_GAUSSIAN);
    if constexpr (TABEXP) {
        if (threadIdx.x < kExpTab) s_exp[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / kExpTab));
    }

    constexpr bool W32 = RNG != HEPMC_RNG_U52_R10;
    // K <= 16: bounds and sizes as separate rows of 16 doubles (128 B: every entry in its own bank pair, so
    // the random-bin look-ups of a half-warp never conflict); else (bound, size) pairs, one LDS.128
    constexpr bool SOA = KP == 16;
    // 32-bit modes on the SoA layout: x = b + s (lo + 1/2) 2^-32 = (b + s 2^-33) + (s 2^-32) lo.  The table holds
    // b2 = b + s 2^-33 (one rounding, <= ulp(b)/2) and s32 = s 2^-32 (exact); the offset word becomes a double with ONE
    // conversion instruction on the otherwise idle XU pipe (I2F.F64.U32) instead of building the (lo, 0x41300000)
    // register pair (2 moves) and subtracting: 2 issue slots and 1 FP64 instruction less per axis.  The weight
    // prod(s) is taken from s32 as well and rescaled by exact powers of two (2^256 after 8 axes, the rest with K^n).
    constexpr bool I2F = HEPMC_VFAST_I2F && SOA && W32;
    double* s_gb = reinterpret_cast<double*>(s_grid2);
    double* s_gs = s_gb + (size_t)NDIM * KS;
    for (int i = threadIdx.x; i < NDIM * K; i += 256) {
        const int d = i / K, k = i - d * K;
        const double* gd = a.grid + d * (2 * K + 1);
        const double gbv = gd[k], gsv = gd[K + 1 + k];
        if constexpr (I2F) { s_gb[d * KS + k] = fma(gsv, 0x1p-33, gbv); s_gs[d * KS + k] = gsv * 0x1p-32; }
        else if constexpr (SOA) { s_gb[d * KS + k] = gbv; s_gs[d * KS + k] = gsv; }
        else s_grid2[d * KS + k] = make_double2(gbv, gsv);
    }
    if constexpr (fused)
        for (int i = threadIdx.x; i < nw * RW * 32; i += 256) s_hist[i] = 0.0;
    __syncthreads();

    double* wf_w = s_wf + warp * 32;
    uint32_t* bins_w = s_bins + warp * NQ * G::ROW;

    // consumer role of this lane in the histogram phase: axis hd, points hg*NUPD .. hg*NUPD + NUPD - 1
    const int hd = lane % NDIM, hg = lane / NDIM;
    const bool h_active = hg < G::NG;
    const uint32_t lane8 = (uint32_t)lane << 3;
    // column address = base + (staged byte << 8 | lane << 3): ONE byte-permute of the staged word
    const uint32_t h_sel = 0x6504u | ((uint32_t)(hd & 3) << 4);
    const uint32_t* h_bins = bins_w + (hd >> 2) * G::ROW + (h_active ? hg : 0) * G::NUPD;
    const double* h_wf = wf_w + (h_active ? hg : 0) * G::NUPD;
    // ROWS16: the staged byte is warp*16 + bin, the base is the CTA's histogram (uniform register);
    // otherwise the staged byte is the bin and the base is this warp's block of K rows
    const uint32_t pack0 = ROWS16 ? (uint32_t)(warp * 16) * 0x01010101u : 0u;
    char* const hist_base = reinterpret_cast<char*>(s_hist) + (ROWS16 ? 0 : (size_t)warp * K * 256);

    const PhiloxKey& key = a.keys;
    const int64_t chunk0 = (int64_t)blockIdx.x * a.chunk;
    const int ppt = (int)(a.chunk >> 8);
    constexpr int ROUNDS = RNG == HEPMC_RNG_U32_R7 ? 7 : 10;
    constexpr int PER = W32 ? 4 : 2;
    constexpr int NCALL = (NDIM + PER - 1) / PER;

    double shift = 0.0, sa = 0.0, sb = 0.0;
    int cnt = 0;
    // The three CTAs that share an SM start together and, having equal work, stay in step wave after wave (all in the
    // Philox stage, then all in the histogram stage).  A start offset of a fraction of a microsecond between them
    // breaks the lock-step for good: -1.4 % kernel time at 1e9 points, any offset from 200 ns to 3 us measured the
    // same; per-warp offsets alone did nothing (scripts/vegas_kernel_bench.cu).  Consecutive block indices land on
    // different SMs, blockIdx % 3 separates the co-resident ones (148 % 3 = 1).  A pure delay: results unchanged.
    if (HEPMC_VFAST_STAGGER_NS) __nanosleep((blockIdx.x % 3) * HEPMC_VFAST_STAGGER_NS);

    auto draw = [&](int64_t li, U4 (&r)[NCALL]) {
        const uint64_t g = (uint64_t)(a.first_index + li);
#pragma unroll
        for (int call = 0; call < NCALL; ++call)
            r[call] = philox4x32_r<ROUNDS>((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)call, a.stream_id, key);
    };

    auto hist_phase = [&]() {
        if (G::NG * NDIM == 32 || h_active) {
            uint4 bw[G::NUPD / 4];
#pragma unroll
            for (int j = 0; j < G::NUPD / 4; ++j) bw[j] = *reinterpret_cast<const uint4*>(h_bins + 4 * j);
#pragma unroll
            for (int i = 0; i < G::NUPD; i += 2) {
                const double2 v2 = *reinterpret_cast<const double2*>(h_wf + i);
                const uint4 b4 = bw[i >> 2];
                const uint32_t w0 = (i & 3) == 0 ? b4.x : b4.z, w1 = (i & 3) == 0 ? b4.y : b4.w;
                *reinterpret_cast<double*>(hist_base + __byte_perm(w0, lane8, h_sel)) += v2.x;
                *reinterpret_cast<double*>(hist_base + __byte_perm(w1, lane8, h_sel)) += v2.y;
            }
        }
    };

    auto run = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        for (int it = 0; it < ppt; ++it) {
            const int64_t li = chunk0 + ((int64_t)it << 8) + threadIdx.x;
            const bool ok = FULL || li < a.n;
            // bin indices are an output of the sample-only instances (two-kernel path); fused instances asked for
            // them take the generic kernel (launch_sample)
            constexpr bool IDX = OUT && KIND == HEPMC_NONE;
            double* xp = (OUT && a.x_out && ok) ? a.x_out + li : nullptr;
            int64_t* ip = (IDX && a.idx_out && ok) ? a.idx_out + li : nullptr;
            (void)xp; (void)ip;
            U4 rc[NCALL];
            draw(li, rc);
            PdfAcc pa;
            pa.init();
            double w = 1.0;
            uint32_t packed[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) packed[q] = pack0;
            int xmax = 0;
#pragma unroll
            for (int call = 0; call < NCALL; ++call) {
                const uint32_t rw[4] = {rc[call].x, rc[call].y, rc[call].z, rc[call].w};
#pragma unroll
                for (int h = 0; h < PER; ++h) {
                    const int d = PER * call + h;
                    if (d < NDIM) {
                        uint32_t bin;
                        double frac;
                        if constexpr (I2F) {
                            const uint64_t t = (uint64_t)(uint32_t)K * rw[h];
                            bin = (uint32_t)(t >> 32);
                            frac = __uint2double_rn((uint32_t)t);            // (t mod 2^32) as a double; scaling lives in the table
                        } else if constexpr (W32) bin_frac32(rw[h], (uint32_t)K, bin, frac);
                        else bin_frac(rw[2 * h], rw[2 * h + 1], (uint32_t)K, bin, frac);
                        double gb, gs;
                        if constexpr (SOA) { gb = s_gb[d * KS + bin]; gs = s_gs[d * KS + bin]; }
                        else { const double2 bs = s_grid2[d * KS + bin]; gb = bs.x; gs = bs.y; }
                        const double x = fma(gs, frac, gb);      // stratified_volume.py:119-121
                        w *= gs;                                 // vegas.py:69-70
                        if constexpr (I2F && NDIM > 8) {
                            if (d == 7) w *= 0x1p+256;           // exact: keeps the partial product of scaled sizes in range
                        }
                        if constexpr (camel) {
                            const double tc = x - a.camel_c;
                            pa.m0 += tc;                         // S1
                            pa.m1 = fma(tc, tc, pa.m1);          // S2
                            // x > 0 by construction; x < 1 can fail by rounding in the last bin: positive
                            // doubles order like their high words, so one integer max per axis
                            if constexpr (KIND == HEPMC_CAMEL) xmax = max(xmax, __double2hiint(x));
                        } else if constexpr (fused) {
                            pdf_acc_dim<KIND, 1>(dens, pa, d, x);
                        }
                        packed[d >> 2] = bin * (1u << (8 * (d & 3))) + packed[d >> 2];
                        if constexpr (OUT) {
                            if (xp) { *xp = x; xp += a.n; }
                            if constexpr (IDX) {
                                if (ip) { *ip = (int64_t)bin; ip += a.n; }
                            }
                        }
                    }
                }
            }
            // vegas.py:72 (partition_count in floating point, App. B1); I2F: times the remaining power of two
            if constexpr (I2F) w *= a.kpow_scaled;
            else w *= a.kpow;
            if constexpr (OUT) {
                if (a.w_out && ok) a.w_out[li] = w;
            }
            if constexpr (fused) {
                if constexpr (camel) {
                    if constexpr (KIND == HEPMC_CAMEL) pa.inside = xmax < 0x3ff00000;   // all x < 1.0
                    const double base = pa.m1 + a.camel_h2, cross = a.camel_2h * pa.m0;
                    pa.m0 = base + cross;
                    pa.m1 = base - cross;
                }
                double f;
                if constexpr (TABEXP && camel) {
                    const double s2 = dens.s0 * dens.s0;             // camel.py:31 with the table exponential
                    const double ea = -0.5 * (dens.s1 + s2 * pa.m0), eb = -0.5 * (dens.s1 + s2 * pa.m1);
                    if (__builtin_expect(exp_tab_ok(ea) && exp_tab_ok(eb), 1)) f = (exp_tab(ea, s_exp) + exp_tab(eb, s_exp)) / 2.0;
                    else f = (exp(ea) + exp(eb)) / 2.0;              // subnormal / zero / overflowing terms: the library
                    if (KIND == HEPMC_CAMEL && !pa.inside) f = 0.0;
                } else if constexpr (TABEXP) {
                    const double ea = -0.5 * (dens.s1 + pa.m0);      // banana.py / gaussian.py: exp(logpdf)
                    f = __builtin_expect(exp_tab_ok(ea), 1) ? exp_tab(ea, s_exp) : exp(ea);
                } else {
                    f = pdf_acc_finish<KIND>(dens, pa);
                }
                double wf = f * w;  // vegas.py:104
                if constexpr (OUT) {
                    if (a.f_out && ok) a.f_out[li] = f;
                }
                if constexpr (FULL) {
                    if (it == 0) shift = wf;
                    const double dv = wf - shift;
                    sa += dv;
                    sb = fma(dv, dv, sb);
                } else {
                    if (ok) {
                        if (cnt == 0) shift = wf;
                        const double dv = wf - shift;
                        sa += dv;
                        sb = fma(dv, dv, sb);
                        ++cnt;
                    } else {
                        wf = 0.0;
                    }
                }
                wf_w[lane] = wf;
#pragma unroll
                for (int q = 0; q < NQ; ++q) bins_w[q * G::ROW + lane] = packed[q];
                __syncwarp();
                hist_phase();
                __syncwarp();
            }
        }
        if constexpr (FULL) cnt = ppt;
    };
    if (chunk0 + a.chunk <= a.n) run(std::true_type{});
    else run(std::false_type{});

    if constexpr (!fused) return;
    // ---- chunk epilogue.  Statistics first (they need no other warp's data), so ONE barrier serves both the
    // histogram flush and the block merge; the last thread (idle in the flush when NDIM*K < 256) merges the warps
    // while the others flush.
    Stats mine;
    mine.n = (double)cnt;
    // full chunk with a power-of-two count per thread (every default chunk size): a / n is a multiplication by an
    // exact reciprocal; full chunk: every lane holds the same count -> warp_merge_equal (no divisions, same bits)
    const bool pow2_full = (chunk0 + a.chunk <= a.n) && (ppt & (ppt - 1)) == 0;
    if (cnt == 0) { mine.mean = 0.0; mine.m2 = 0.0; }
    else {
        const double am = pow2_full ? sa * __hiloint2double((992 + __clz(ppt)) << 20, 0) : sa / mine.n;   // 2^-log2(ppt)
        mine.mean = shift + am;
        mine.m2 = sb - sa * am;
        if (mine.m2 < 0.0) mine.m2 = 0.0;
    }
    const Stats st = (chunk0 + a.chunk <= a.n) ? warp_merge_equal(mine) : warp_merge(mine);
    if (lane == 0) s_stats[warp] = st;
    __syncthreads();
    if (threadIdx.x == 255) {
        Stats r{0.0, 0.0, 0.0};
        for (int w = 0; w < nw; ++w) r = merge(r, s_stats[w]);
        double* o = a.stats_partials + (size_t)blockIdx.x * 3;
        o[0] = r.n; o[1] = r.mean; o[2] = r.m2;
    }
    {   // flush: out[d*K + k] = sum over warps, groups (fixed order).  Thread j -> (k = j / NDIM, d = j % NDIM): a warp
        // reads whole histogram rows (no bank conflict beyond the two wavefronts of a 64-bit access); the loads are
        // independent, the additions keep their order
        double* out_row = a.hist_partials + (size_t)blockIdx.x * NDIM * K;
        for (int j = threadIdx.x; j < NDIM * K; j += 256) {
            const int k = j / NDIM, d = j - k * NDIM;
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < nw; ++w) {
                double v[G::NG];
#pragma unroll
                for (int g = 0; g < G::NG; ++g) v[g] = s_hist[((size_t)w * RW + k) * 32 + g * NDIM + d];
#pragma unroll
                for (int g = 0; g < G::NG; ++g) s += v[g];
            }
            out_row[d * K + k] = s;
        }
    }
}

}  // namespace hepmc

// Kernel templates of the integration path, shared by integrate.cu (generic instances:
// runtime ndim, replay tapes, per-point outputs) and integrate_fast_*.cu (throughput
// instances: compile-time ndim and density kind, native RNG; one instance without and one with
// per-point outputs -- keep_samples=True, the reference's default -- per (kind, ndim), and
// sample-only instances (kind NONE) for user integrands).
#pragma once
#include <type_traits>
#include "density.cuh"

namespace hepmc {

constexpr int kBinWords = HEPMC_MAX_DIM / 4 + 1;  // packed bin bytes per point, padded (bank-conflict-free)
constexpr int kMaxPPI = 2;        // warp tiles staged per loop trip (shared-memory sizing)
#ifndef HEPMC_FAST_PPI
#define HEPMC_FAST_PPI 1
#endif
#ifndef HEPMC_FAST_MINBLOCKS
#define HEPMC_FAST_MINBLOCKS 3
#endif
constexpr int kFastPPI = HEPMC_FAST_PPI;
constexpr int kFastMinBlocks = HEPMC_FAST_MINBLOCKS;

struct SampleArgs {
    const double* grid;  // VEGAS: (ndim, 2K+1); plain: unused
    int ndim, K;
    int64_t n, first_index, chunk;
    uint64_t seed;
    uint32_t stream_id;
    const int64_t* idx_replay;
    const double* u_replay;
    double* x_out;
    double* f_out;
    double* w_out;
    int64_t* idx_out;
    double* stats_partials;
    double* hist_partials;
    double kpow;
    double kpow_scaled;  // kpow * 2^(32 * (ndim > 8 ? ndim - 8 : ndim)): the scaled-size products of vegas_fast_kernel
    PhiloxKey keys;  // round keys of `seed`, expanded on the host: constant-bank operands of the Philox rounds
    double camel_c, camel_2h, camel_h2;  // centred camel of the throughput kernel (vegas_fast.cuh): (a+b)/2, b-a, n((b-a)/2)^2
    int rng_mode;  // HEPMC_RNG_*: bits per draw and Philox rounds of the native generator (hepmc_set_rng_mode)
    // two-kernel path inputs (accumulate)
    const double* f_in;
    const double* w_in;
    const int64_t* idx_in;
};

// Histogram phase of a warp tile of 32 points, runtime geometry (generic kernels).  Lane L: axis
// d = L % ndim, group g = L / ndim; group g walks the points g*nupd .. g*nupd + nupd - 1 in order,
// nupd = 32 / ngrp rounded up to 4, 8, 16 or 32.  The SAME assignment as the throughput kernel
// (HistGeom in vegas_fast.cuh), so fused, two-kernel and replay paths add in the same order.
__device__ __forceinline__ void warp_hist_update(double* hist_w, const double* wf_w, const uint32_t* bins_w,
                                                 int ndim, int lane) {
    const int ngrp = 32 / ndim;
    const int need = (32 + ngrp - 1) / ngrp;
    const int nupd = need <= 4 ? 4 : need <= 8 ? 8 : need <= 16 ? 16 : 32;
    const int d = lane % ndim, grp = lane / ndim;
    if (grp < 32 / nupd) {
        const int word = d >> 2, shift = 8 * (d & 3);
        for (int p = grp * nupd; p < (grp + 1) * nupd; ++p) {
            const double v = wf_w[p];
            const uint32_t bin = (bins_w[p * kBinWords + word] >> shift) & 0xffu;
            hist_w[bin * 32 + lane] += v;
        }
    }
}

__device__ __forceinline__ void block_hist_flush(const double* s_hist, int ndim, int K, int nwarps,
                                                 double* out_row) {
    const int ngrp = 32 / ndim;
    for (int j = threadIdx.x; j < ndim * K; j += blockDim.x) {
        const int d = j / K, k = j % K;
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w)
            for (int g = 0; g < ngrp; ++g) s += s_hist[(w * K + k) * 32 + g * ndim + d];
        out_row[j] = s;
    }
}

inline void set_camel_centre(SampleArgs& a, const Dens& p) {
    const double h = 0.5 * (p.b[0] - p.a[0]);
    a.camel_c = 0.5 * (p.a[0] + p.b[0]);
    a.camel_2h = 2.0 * h;
    a.camel_h2 = (double)p.ndim * h * h;
}

struct SmemLayout {
    double* grid;
    Stats* stats;
    double* hist;
    double* wf;
    uint32_t* bins;
};

__host__ __device__ inline size_t vegas_smem_bytes(int ndim, int K, int threads) {
    const int nw = threads / 32;
    size_t b = (size_t)ndim * (2 * K + 1) * 8;  // grid
    b += 32 * sizeof(Stats);                    // block merge scratch
    b += (size_t)nw * K * 32 * 8;               // private histogram columns
    b += (size_t)nw * 32 * 8 * kMaxPPI;                   // wf tiles
    b += (size_t)nw * 32 * kBinWords * 4 * kMaxPPI;       // packed bins tiles
    return b;
}

__device__ __forceinline__ SmemLayout carve(double* base, int ndim, int K, int nw) {
    SmemLayout s;
    s.grid = base;
    s.stats = reinterpret_cast<Stats*>(s.grid + (size_t)ndim * (2 * K + 1));
    s.hist = reinterpret_cast<double*>(s.stats + 32);
    s.wf = s.hist + (size_t)nw * K * 32;
    s.bins = reinterpret_cast<uint32_t*>(s.wf + nw * 32 * kMaxPPI);
    return s;
}

// ------------------------------------------------------------------ draws of the native generator
// modes: enum hepmc_rng_mode (include/hepmc_b200.h)

template <int R>
__device__ __forceinline__ U4 philox4x32_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKey& key) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
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

__device__ __forceinline__ U4 philox4x32_mode(int mode, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKey& key) {
    return mode == HEPMC_RNG_U32_R7 ? philox4x32_r<7>(c0, c1, c2, c3, key) : philox4x32_r<10>(c0, c1, c2, c3, key);
}

// 32 random bits -> uniform (w + 1/2) 2^-32 in (0,1), exact: as_double(0x41300000 : w) = 2^20 + w 2^-32,
// minus (2^20 - 2^-33).  Oracle: hepmc_oracle.u32.
__device__ __forceinline__ double u32(uint32_t w) {
    return __hiloint2double(0x41300000, (int)w) - 1048575.99999999988358467817306518554687500;
}

// 32 random bits -> (bin, offset): t = K*w; bin = floor(t / 2^32); offset = (t mod 2^32 + 1/2) 2^-32.
// Same joint law as randint + rand (stratified_volume.py:111,121) at 32-bit resolution: bin
// probabilities differ from 1/K by at most 2^-32.  Oracle: hepmc_oracle.vegas_bin_frac32.
__device__ __forceinline__ void bin_frac32(uint32_t w, uint32_t K, uint32_t& bin, double& frac) {
    const uint64_t t = (uint64_t)K * w;
    bin = (uint32_t)(t >> 32);
    frac = u32((uint32_t)t);
}

// ------------------------------------------------------------------ VEGAS sample, generic instance
// Runtime dimensionality, replay tapes, all per-point outputs: parity runs and shapes without a
// throughput instance (ndim > 16).  The throughput kernel is vegas_fast_kernel (vegas_fast.cuh).
// KIND: integrand fused into the kernel (HEPMC_NONE: sample/weight only).
// Shared grid layout: per axis K entries of (bound_k, size_k) as double2 -> one LDS.128.
template <int KIND>
__global__ void __launch_bounds__(256, 1) vegas_sample_kernel(const __grid_constant__ Dens dens,
                                                              const __grid_constant__ SampleArgs a) {
    extern __shared__ double smem_raw[];
    const int ndim = a.ndim;
    const int K = a.K;
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SmemLayout s = carve(smem_raw, ndim, K, nw);
    constexpr bool fused = KIND != HEPMC_NONE;
    double2* s_grid2 = reinterpret_cast<double2*>(s.grid);  // [ndim][K] (bound, size); 2K <= 2K+1 doubles per axis

    for (int i = threadIdx.x; i < ndim * K; i += blockDim.x) {
        const int d = i / K, k = i - d * K;
        const double* gd = a.grid + d * (2 * K + 1);
        s_grid2[i] = make_double2(gd[k], gd[K + 1 + k]);
    }
    if (fused)
        for (int i = threadIdx.x; i < nw * K * 32; i += blockDim.x) s.hist[i] = 0.0;
    __syncthreads();

    double* hist_w = s.hist + (size_t)warp * K * 32;
    double* wf_w = s.wf + warp * 32 * kMaxPPI;
    uint32_t* bins_w = s.bins + warp * 32 * kBinWords * kMaxPPI;

    const PhiloxKey& key = a.keys;   // expanded on the host: constant-bank operands of the rounds
    const bool replay = a.u_replay != nullptr;
    const bool w32 = a.rng_mode != HEPMC_RNG_U52_R10;
    const int per = w32 ? 4 : 2;                       // axes served by one Philox call
    const int ncalls = (ndim + per - 1) / per;
    const int64_t chunk0 = (int64_t)blockIdx.x * a.chunk;
    const int ppt = (int)(a.chunk / blockDim.x);
    constexpr int NPACK = kBinWords - 1;
    ShiftedAcc acc;
    acc.init();

    for (int it = 0; it < ppt; ++it) {
        const int64_t li = chunk0 + (int64_t)it * blockDim.x + threadIdx.x;
        const bool valid = li < a.n;
        double wf = 0.0;
        uint32_t packed[NPACK];
#pragma unroll
        for (int q = 0; q < NPACK; ++q) packed[q] = 0u;
        if (valid) {
            const uint64_t g = (uint64_t)(a.first_index + li);
            double* xp = a.x_out ? a.x_out + li : nullptr;
            int64_t* ip = a.idx_out ? a.idx_out + li : nullptr;
            PdfAcc pa;
            pa.init();
            double w = 1.0;
            for (int call = 0; call < ncalls; ++call) {
                U4 r{0u, 0u, 0u, 0u};
                if (!replay) r = philox4x32_mode(a.rng_mode, (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)call, a.stream_id, key);
                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int d = per * call + h;
                    if (h < per && d < ndim) {
                        uint32_t bin;
                        double frac;
                        if (replay) {
                            bin = (uint32_t)a.idx_replay[(int64_t)d * a.n + li];
                            frac = a.u_replay[(int64_t)d * a.n + li];
                        } else if (w32) {
                            bin_frac32(rw[h], (uint32_t)K, bin, frac);
                        } else {
                            bin_frac(rw[(2 * h) & 3], rw[(2 * h + 1) & 3], (uint32_t)K, bin, frac);
                        }
                        const double2 bs = s_grid2[d * K + bin];
                        const double x = fma(bs.y, frac, bs.x);  // stratified_volume.py:119-121
                        w *= bs.y;                               // vegas.py:69-70
                        if constexpr (fused) pdf_acc_dim<KIND, 1>(dens, pa, d, x);
#pragma unroll
                        for (int q = 0; q < NPACK; ++q)
                            if (q == (d >> 2)) packed[q] |= bin << (8 * (d & 3));
                        // dim-major outputs: running pointers
                        if (xp) { *xp = x; xp += a.n; }
                        if (ip) { *ip = (int64_t)bin; ip += a.n; }
                    }
                }
            }
            w *= a.kpow;  // vegas.py:72 (partition_count evaluated in floating point, App. B1)
            if (a.w_out) a.w_out[li] = w;
            if constexpr (fused) {
                const double f = pdf_acc_finish<KIND>(dens, pa);
                wf = f * w;  // vegas.py:104
                if (a.f_out) a.f_out[li] = f;
                acc.add(wf);
            }
        }
        if constexpr (fused) {
            wf_w[lane] = wf;
#pragma unroll
            for (int q = 0; q < NPACK; ++q)
                if (q * 4 < ndim) bins_w[lane * kBinWords + q] = packed[q];
            __syncwarp();
            warp_hist_update(hist_w, wf_w, bins_w, ndim, lane);
            __syncwarp();
        }
    }
    if constexpr (!fused) return;
    __syncthreads();
    block_hist_flush(s.hist, ndim, K, nw, a.hist_partials + (size_t)blockIdx.x * ndim * K);
    // per-thread -> warp -> block (fixed order)
    // full chunk: equal counts in every lane -> the division-free butterfly (same bits, common.cuh)
    Stats st = (chunk0 + a.chunk <= a.n) ? warp_merge_equal(acc.stats()) : warp_merge(acc.stats());
    if (lane == 0) s.stats[warp] = st;
    __syncthreads();
    if (threadIdx.x == 0) {
        Stats r{0.0, 0.0, 0.0};
        for (int w = 0; w < nw; ++w) r = merge(r, s.stats[w]);
        double* o = a.stats_partials + (size_t)blockIdx.x * 3;
        o[0] = r.n; o[1] = r.mean; o[2] = r.m2;
    }
}

// A warp's 32 rows of NDIM doubles (one row per lane, in registers) -> global memory in ADDRESS order: the rows go
// through a skewed shared-memory tile and leave as 512 (256 for odd NDIM) contiguous bytes per store instruction.
// Rows written straight from registers (one 32-byte piece per lane and instruction, 32 different lines) top out at
// 4.8 TB/s on B200; warp-contiguous stores reach the copy peak (scripts/store_pattern_bench.cu).
// tile: 32 * NDIM doubles of this warp; nrows: valid rows (lanes 0 .. nrows-1); dst: address of row 0.
template <int NDIM>
__device__ __forceinline__ void warp_store_rows(double* tile, const double (&row)[NDIM], int lane, double* dst, int nrows) {
    if constexpr (NDIM % 2 == 0) {
        constexpr int P = NDIM / 2;                       // 16-byte pieces per row
        double2* t2 = reinterpret_cast<double2*>(tile);
#pragma unroll
        for (int j = 0; j < P; ++j) t2[lane * P + (j + lane) % P] = make_double2(row[2 * j], row[2 * j + 1]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int idx = j * 32 + lane;
            const int r = idx / P, c = idx % P;
            if (r < nrows) {
                const double2 w = t2[r * P + (c + r) % P];
                asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(dst + 2 * idx), "d"(w.x), "d"(w.y) : "memory");
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < NDIM; ++j) tile[lane * NDIM + (j + lane) % NDIM] = row[j];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NDIM; ++j) {
            const int idx = j * 32 + lane;
            const int r = idx / NDIM, c = idx % NDIM;
            if (r < nrows) dst[idx] = tile[r * NDIM + (c + r) % NDIM];
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------ plain MC
// NDIM_T: compile-time dimensionality (0 = runtime).  FAST: native RNG only, RNG = its compile-time
// mode; otherwise replay tapes and the runtime mode a.rng_mode are honoured.
template <int KIND, int NDIM_T, bool FAST, bool OUT = !FAST, int RNG = HEPMC_RNG_U52_R10>
__global__ void __launch_bounds__(256) plain_sample_kernel(const __grid_constant__ Dens dens,
                                                           const __grid_constant__ SampleArgs a) {
    __shared__ Stats s_stats[8];
    constexpr bool TILED = OUT && NDIM_T != 0;           // rows leave through warp_store_rows
    __shared__ double s_tile[TILED ? 8 * 32 * (NDIM_T ? NDIM_T : 1) : 1];
    const int ndim = NDIM_T ? NDIM_T : a.ndim;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool fused = KIND != HEPMC_NONE;
    const PhiloxKey& key = a.keys;   // expanded on the host: constant-bank operands of the rounds
    const bool replay = !FAST && a.u_replay != nullptr;
    const int mode = FAST ? RNG : a.rng_mode;
    const bool w32 = mode != HEPMC_RNG_U52_R10;
    constexpr int PERMAX = 4;
    const int per = w32 ? 4 : 2;
    const int64_t chunk0 = (int64_t)blockIdx.x * a.chunk;
    const int ppt = (int)(a.chunk / blockDim.x);
    const int ncalls = (ndim + per - 1) / per;
    ShiftedAcc acc;
    acc.init();
    for (int it = 0; it < ppt; ++it) {
        const int64_t li = chunk0 + (int64_t)it * blockDim.x + threadIdx.x;
        // TILED: the warp stays together (its lanes beyond n compute a point nobody stores); else a lane past n is done
        if (TILED ? (li - lane >= a.n) : (li >= a.n)) break;
        const bool ok = li < a.n;
        const uint64_t g = (uint64_t)(a.first_index + li);
        PdfAcc pa;
        pa.init();
        double xrow[NDIM_T ? NDIM_T : 1];
#pragma unroll
        for (int call = 0; call < ncalls; ++call) {
            {
                U4 r{0u, 0u, 0u, 0u};
                if (FAST || !replay) {
                    if constexpr (FAST) r = philox4x32_r<(RNG == HEPMC_RNG_U32_R7 ? 7 : 10)>((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)call, a.stream_id, key);
                    else r = philox4x32_mode(mode, (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)call, a.stream_id, key);
                }
                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int h = 0; h < PERMAX; ++h) {
                    const int d = per * call + h;
                    if (h < per && d < ndim) {
                        double x;
                        if (!FAST && replay) x = a.u_replay[li * ndim + d];
                        else if (w32) x = u32(rw[h]);
                        else x = u52(rw[(2 * h) & 3], rw[(2 * h + 1) & 3]);
                        // native draws lie strictly inside (0,1); replayed ones get the full check
                        if constexpr (fused) {
                            if constexpr (FAST) pdf_acc_dim<KIND, 0>(dens, pa, d, x);
                            else pdf_acc_dim<KIND, 2>(dens, pa, d, x);
                        }
                        if constexpr (OUT) {
                            // (N, ndim) row-major, integration.py:45.  Compile-time ndim: the row is staged in
                            // registers and leaves as 256- / 128-bit stores (one contiguous 8 ndim byte run per
                            // thread); scalar stores of a 128-byte-strided row reach a third of the bandwidth.
                            if constexpr (NDIM_T != 0) xrow[d] = x;
                            else if (a.x_out) a.x_out[li * ndim + d] = x;
                        }
                    }
                }
            }
        }
        if constexpr (TILED) {
            if (a.x_out) {
                const int64_t w0 = li - lane;                                       // first point of the warp
                const int nrows = (int)(a.n - w0 < 32 ? a.n - w0 : 32);
                warp_store_rows<NDIM_T>(s_tile + warp * 32 * NDIM_T, xrow, lane, a.x_out + w0 * NDIM_T, nrows);
            }
        }
        if constexpr (fused) {
            const double f = pdf_acc_finish<KIND>(dens, pa);
            if (ok) acc.add(f);
            if constexpr (OUT) {
                if (a.f_out && ok) a.f_out[li] = f;
            }
        }
    }
    if constexpr (!fused) return;
    // full chunk: equal counts in every lane -> the division-free butterfly (same bits, common.cuh)
    Stats st = (chunk0 + a.chunk <= a.n) ? warp_merge_equal(acc.stats()) : warp_merge(acc.stats());
    if (lane == 0) s_stats[warp] = st;
    __syncthreads();
    if (threadIdx.x == 0) {
        Stats r{0.0, 0.0, 0.0};
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = merge(r, s_stats[w]);
        double* o = a.stats_partials + (size_t)blockIdx.x * 3;
        o[0] = r.n; o[1] = r.mean; o[2] = r.m2;
    }
}

}  // namespace hepmc

#include "vegas_fast.cuh"

namespace hepmc {

// fast-path launchers, one translation unit per density kind (integrate_fast_*.cu)
// `out`: per-point outputs (x, f, w, idx) requested -> the OUT instance of the same kernel
int launch_fast_camel(bool vegas, bool out, int, const Dens&, const SampleArgs&, unsigned, int, size_t, cudaStream_t);
int launch_fast_ucamel(bool vegas, bool out, int, const Dens&, const SampleArgs&, unsigned, int, size_t, cudaStream_t);
int launch_fast_banana(bool vegas, bool out, int, const Dens&, const SampleArgs&, unsigned, int, size_t, cudaStream_t);
int launch_fast_gaussian(bool vegas, bool out, int, const Dens&, const SampleArgs&, unsigned, int, size_t, cudaStream_t);
int launch_fast_uniform(bool vegas, bool out, int, const Dens&, const SampleArgs&, unsigned, int, size_t, cudaStream_t);
int launch_fast_none(bool vegas, bool out, int, const Dens&, const SampleArgs&, unsigned, int, size_t, cudaStream_t);

#define HEPMC_DEFINE_FAST_LAUNCHER(NAME, KIND)                                                              \
    int NAME(bool vegas, bool out, int ndim, const Dens& dens, const SampleArgs& a, unsigned blocks,        \
             int threads, size_t smem, cudaStream_t st) {                                                   \
        switch (ndim) {                                                                                     \
            HEPMC_FAST_CASE(KIND, 1) HEPMC_FAST_CASE(KIND, 2) HEPMC_FAST_CASE(KIND, 3) HEPMC_FAST_CASE(KIND, 4)     \
            HEPMC_FAST_CASE(KIND, 5) HEPMC_FAST_CASE(KIND, 6) HEPMC_FAST_CASE(KIND, 7) HEPMC_FAST_CASE(KIND, 8)     \
            HEPMC_FAST_CASE(KIND, 9) HEPMC_FAST_CASE(KIND, 10) HEPMC_FAST_CASE(KIND, 11) HEPMC_FAST_CASE(KIND, 12)  \
            HEPMC_FAST_CASE(KIND, 13) HEPMC_FAST_CASE(KIND, 14) HEPMC_FAST_CASE(KIND, 15) HEPMC_FAST_CASE(KIND, 16) \
            default: return -100; /* no fast instance: caller falls back to the generic kernel */          \
        }                                                                                                   \
        return 0;                                                                                           \
    }
// the two function attributes are set once per instance (and again only if a call needs more shared memory): at the
// reference's own problem sizes the host side of a launch is a visible part of the iteration
#define HEPMC_VFAST_GO(KIND, ND, OUT, RNG, KP)                                                              \
    do {                                                                                                    \
        static size_t smem_set = 0;                                                                         \
        static int dev_set = -1;                                                                            \
        int dev_now = 0;                                                                                    \
        cudaGetDevice(&dev_now);                                                                            \
        if (smem > smem_set || dev_now != dev_set) {                                                        \
            HEPMC_CUDA(cudaFuncSetAttribute(vegas_fast_kernel<KIND, ND, RNG, OUT, KP>,                      \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
            HEPMC_CUDA(cudaFuncSetAttribute(vegas_fast_kernel<KIND, ND, RNG, OUT, KP>,                      \
                                            cudaFuncAttributePreferredSharedMemoryCarveout, 100));          \
            smem_set = smem;                                                                                \
            dev_set = dev_now;                                                                              \
        }                                                                                                   \
        vegas_fast_kernel<KIND, ND, RNG, OUT, KP><<<blocks, 256, smem, st>>>(dens, a);                      \
    } while (0)
#define HEPMC_VFAST_KP(KIND, ND, OUT, RNG)                                                                  \
    do {                                                                                                    \
        if (a.K <= 16) HEPMC_VFAST_GO(KIND, ND, OUT, RNG, 16);                                              \
        else HEPMC_VFAST_GO(KIND, ND, OUT, RNG, 0);                                                         \
    } while (0)
#define HEPMC_FAST_LAUNCH(KIND, ND, OUT)                                                                    \
    do {                                                                                                    \
        if (vegas) {                                                                                        \
            if (a.rng_mode == HEPMC_RNG_U32_R7) HEPMC_VFAST_KP(KIND, ND, OUT, HEPMC_RNG_U32_R7);            \
            else if (a.rng_mode == HEPMC_RNG_U32_R10) HEPMC_VFAST_KP(KIND, ND, OUT, HEPMC_RNG_U32_R10);     \
            else HEPMC_VFAST_KP(KIND, ND, OUT, HEPMC_RNG_U52_R10);                                          \
        } else {                                                                                            \
            if (a.rng_mode == HEPMC_RNG_U32_R7)                                                             \
                plain_sample_kernel<KIND, ND, true, OUT, HEPMC_RNG_U32_R7><<<blocks, 256, 0, st>>>(dens, a); \
            else if (a.rng_mode == HEPMC_RNG_U32_R10)                                                       \
                plain_sample_kernel<KIND, ND, true, OUT, HEPMC_RNG_U32_R10><<<blocks, 256, 0, st>>>(dens, a); \
            else                                                                                            \
                plain_sample_kernel<KIND, ND, true, OUT, HEPMC_RNG_U52_R10><<<blocks, 256, 0, st>>>(dens, a); \
        }                                                                                                   \
    } while (0)
#define HEPMC_FAST_CASE(KIND, ND)                                                                           \
    case ND:                                                                                                \
        if (out) HEPMC_FAST_LAUNCH(KIND, ND, true);                                                         \
        else if (KIND != HEPMC_NONE) HEPMC_FAST_LAUNCH(KIND, ND, false);                                    \
        else return -100;                                                                                   \
        break;

}  // namespace hepmc

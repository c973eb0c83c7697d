// Kernel templates of the integration path, shared by integrate.cu (generic instances:
// runtime ndim, replay tapes, per-point outputs) and integrate_fast_*.cu (throughput
// instances: compile-time ndim and density kind, native RNG; one instance without and one with
// per-point outputs -- keep_samples=True, the reference's default -- per (kind, ndim), and
// sample-only instances (kind NONE) for user integrands).
#pragma once
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
    // two-kernel path inputs (accumulate)
    const double* f_in;
    const double* w_in;
    const int64_t* idx_in;
};

// lane L < ngrp*ndim: axis d = L % ndim, walks points grp, grp+ngrp, ... of the warp tile.
__device__ __forceinline__ void warp_hist_update(double* hist_w, const double* wf_w, const uint32_t* bins_w,
                                                 int ndim, int lane) {
    const int ngrp = 32 / ndim;
    if (lane < ngrp * ndim) {
        const int d = lane % ndim, grp = lane / ndim;
        const int word = d >> 2, shift = 8 * (d & 3);
        for (int p = grp; p < 32; p += ngrp) {
            const double v = wf_w[p];
            const uint32_t bin = (bins_w[p * kBinWords + word] >> shift) & 0xffu;
            hist_w[bin * 32 + lane] += v;
        }
    }
}

#ifndef HEPMC_HIST_PAIR
#define HEPMC_HIST_PAIR 0
#endif
template <int NDIM>
__device__ __forceinline__ void warp_hist_update_n(double* hist_w, const double* wf_w, const uint32_t* bins_w,
                                                   int lane) {
    constexpr int ngrp = 32 / NDIM;
    constexpr int NUPD = (32 + ngrp - 1) / ngrp;  // updates per lane (points grp, grp+ngrp, ...)
    if (lane < ngrp * NDIM) {
        const int d = lane % NDIM, grp = lane / NDIM;
        const int word = d >> 2, shift = 8 * (d & 3);
        const uint32_t* bw = bins_w + grp * kBinWords + word;
        const double* wv = wf_w + grp;
        double* hcol = hist_w + lane;
        if constexpr (HEPMC_HIST_PAIR && (32 % ngrp == 0) && (NUPD % 2 == 0)) {
            // two read-modify-writes in flight: both column entries are loaded before the first
            // add; if the two bins coincide the second add takes the first result (same order of
            // additions as the sequential loop, so results are bit-identical)
#pragma unroll
            for (int i = 0; i < NUPD; i += 2) {
                const double v0 = wv[i * ngrp], v1 = wv[(i + 1) * ngrp];
                const uint32_t b0 = (bw[i * ngrp * kBinWords] >> shift) & 0xffu;
                const uint32_t b1 = (bw[(i + 1) * ngrp * kBinWords] >> shift) & 0xffu;
                const double h0 = hcol[b0 * 32], h1 = hcol[b1 * 32];
                const double t0 = h0 + v0;
                const double t1 = (b1 == b0 ? t0 : h1) + v1;
                hcol[b0 * 32] = t0;
                hcol[b1 * 32] = t1;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NUPD; ++i) {
                if (grp + i * ngrp < 32) {
                    const double v = wv[i * ngrp];
                    const uint32_t bin = (bw[i * ngrp * kBinWords] >> shift) & 0xffu;
                    hcol[bin * 32] += v;
                }
            }
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

// ------------------------------------------------------------------ VEGAS sample (+ fused integrand)
// KIND: integrand fused into the kernel (HEPMC_NONE: sample/weight only).
// NDIM_T: compile-time dimensionality (0 = runtime) -- full unroll, static bin packing.
// FAST: native RNG, no per-point outputs (the throughput path); otherwise replay tapes and
//       x / f / w / idx outputs are honoured through warp-uniform runtime flags.
// Shared grid layout: per axis K entries of (bound_k, size_k) as double2 -> one LDS.128.
template <int KIND, int NDIM_T, bool FAST, bool OUT = !FAST>
__global__ void __launch_bounds__(256, FAST ? kFastMinBlocks : 1) vegas_sample_kernel(const __grid_constant__ Dens dens,
                                                           const __grid_constant__ SampleArgs a) {
    extern __shared__ double smem_raw[];
    const int ndim = NDIM_T ? NDIM_T : a.ndim;
    const int K = a.K;
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SmemLayout s = carve(smem_raw, ndim, K, nw);
    constexpr bool fused = KIND != HEPMC_NONE;
    // points per thread and loop trip: two independent points in flight on the throughput path
    // (more ILP against the dependent Philox / FP64 chains); identical results either way
    constexpr int PPI = FAST ? kFastPPI : 1;
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

    const PhiloxKey key = philox_expand(a.seed);
    const bool replay = !FAST && a.u_replay != nullptr;
    const int64_t chunk0 = (int64_t)blockIdx.x * a.chunk;
    const int ppt = (int)(a.chunk / blockDim.x);
    const int ncalls = NDIM_T ? (NDIM_T + 1) / 2 : (ndim + 1) >> 1;
    constexpr int NPACK = NDIM_T ? (NDIM_T + 3) / 4 : kBinWords - 1;
    ShiftedAcc acc;
    acc.init();

    // one sample point: draws -> grid map -> integrand -> weight; returns f*w and the packed bins
    auto point = [&](int64_t li, double& wf, uint32_t (&packed)[NPACK]) {
        const uint64_t g = (uint64_t)(a.first_index + li);
        const bool ok = li < a.n;   // the FAST loop also runs the padding points of the last chunk: they store nothing
        double* xp = (OUT && a.x_out && ok) ? a.x_out + li : nullptr;
        int64_t* ip = (OUT && a.idx_out && ok) ? a.idx_out + li : nullptr;
        PdfAcc pa;
        pa.init();
        double w = 1.0;
#pragma unroll
        for (int call = 0; call < ncalls; ++call) {
            U4 r{0u, 0u, 0u, 0u};
            if (FAST || !replay) r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)call, a.stream_id, key);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int d = 2 * call + h;
                if (d < ndim) {
                    uint32_t bin;
                    double frac;
                    if (!FAST && replay) {
                        bin = (uint32_t)a.idx_replay[(int64_t)d * a.n + li];
                        frac = a.u_replay[(int64_t)d * a.n + li];
                    } else {
                        bin_frac(h ? r.z : r.x, h ? r.w : r.y, (uint32_t)K, bin, frac);
                    }
                    const double2 bs = s_grid2[d * K + bin];
                    const double x = fma(bs.y, frac, bs.x);  // stratified_volume.py:119-121
                    w *= bs.y;                               // vegas.py:69-70
                    if constexpr (fused) pdf_acc_dim<KIND, 1>(dens, pa, d, x);
                    if constexpr (NDIM_T != 0) {
                        packed[d >> 2] |= bin << (8 * (d & 3));
                    } else {
#pragma unroll
                        for (int q = 0; q < NPACK; ++q)
                            if (q == (d >> 2)) packed[q] |= bin << (8 * (d & 3));
                    }
                    if constexpr (OUT) {
                        // dim-major outputs: running pointers (one 64-bit add per axis instead of a
                        // 64-bit multiply-add per store)
                        if (xp) { *xp = x; xp += a.n; }
                        if (ip) { *ip = (int64_t)bin; ip += a.n; }
                    }
                }
            }
        }
        w *= a.kpow;  // vegas.py:72 (partition_count evaluated in floating point, App. B1)
        if constexpr (OUT) {
            if (a.w_out && ok) a.w_out[li] = w;
        }
        if constexpr (fused) {
            const double f = pdf_acc_finish<KIND>(dens, pa);
            wf = f * w;  // vegas.py:104
            if constexpr (OUT) {
                if (a.f_out && ok) a.f_out[li] = f;
            }
        }
    };

    for (int it = 0; it < ppt; it += PPI) {
        double wf[PPI];
        uint32_t packed[PPI][NPACK];
        bool valid[PPI];
#pragma unroll
        for (int u = 0; u < PPI; ++u) {
            const int64_t li = chunk0 + (int64_t)(it + u) * blockDim.x + threadIdx.x;
            valid[u] = li < a.n;
            wf[u] = 0.0;
#pragma unroll
            for (int q = 0; q < NPACK; ++q) packed[u][q] = 0u;
            if constexpr (FAST) {
                point(li, wf[u], packed[u]);  // branch-free: out-of-range points touch no memory and are masked below
            } else {
                if (valid[u]) point(li, wf[u], packed[u]);
            }
        }
        if constexpr (fused) {
#pragma unroll
            for (int u = 0; u < PPI; ++u) {
                if (valid[u]) acc.add(wf[u]);
                else wf[u] = 0.0;
                wf_w[u * 32 + lane] = wf[u];
#pragma unroll
                for (int q = 0; q < NPACK; ++q)
                    if (q * 4 < ndim) bins_w[(u * 32 + lane) * kBinWords + q] = packed[u][q];
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < PPI; ++u) {
                if constexpr (NDIM_T != 0) warp_hist_update_n<NDIM_T>(hist_w, wf_w + u * 32, bins_w + u * 32 * kBinWords, lane);
                else warp_hist_update(hist_w, wf_w + u * 32, bins_w + u * 32 * kBinWords, ndim, lane);
            }
            __syncwarp();
        }
    }
    if constexpr (!fused) return;
    __syncthreads();
    block_hist_flush(s.hist, ndim, K, nw, a.hist_partials + (size_t)blockIdx.x * ndim * K);
    // per-thread -> warp -> block (fixed order)
    Stats st = warp_merge(acc.stats());
    if (lane == 0) s.stats[warp] = st;
    __syncthreads();
    if (threadIdx.x == 0) {
        Stats r{0.0, 0.0, 0.0};
        for (int w = 0; w < nw; ++w) r = merge(r, s.stats[w]);
        double* o = a.stats_partials + (size_t)blockIdx.x * 3;
        o[0] = r.n; o[1] = r.mean; o[2] = r.m2;
    }
}

// ------------------------------------------------------------------ plain MC
template <int KIND, int NDIM_T, bool FAST, bool OUT = !FAST>
__global__ void __launch_bounds__(256) plain_sample_kernel(const __grid_constant__ Dens dens,
                                                           const __grid_constant__ SampleArgs a) {
    __shared__ Stats s_stats[8];
    const int ndim = NDIM_T ? NDIM_T : a.ndim;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool fused = KIND != HEPMC_NONE;
    const PhiloxKey key = philox_expand(a.seed);
    const bool replay = !FAST && a.u_replay != nullptr;
    const int64_t chunk0 = (int64_t)blockIdx.x * a.chunk;
    const int ppt = (int)(a.chunk / blockDim.x);
    const int ncalls = NDIM_T ? (NDIM_T + 1) / 2 : (ndim + 1) >> 1;
    ShiftedAcc acc;
    acc.init();
    for (int it = 0; it < ppt; ++it) {
        const int64_t li = chunk0 + (int64_t)it * blockDim.x + threadIdx.x;
        if (li >= a.n) break;
        const uint64_t g = (uint64_t)(a.first_index + li);
        PdfAcc pa;
        pa.init();
        double xrow[NDIM_T ? NDIM_T : 1];
#pragma unroll
        for (int call = 0; call < ncalls; ++call) {
            U4 r{0u, 0u, 0u, 0u};
            if (FAST || !replay) r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)call, a.stream_id, key);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int d = 2 * call + h;
                if (d < ndim) {
                    double x;
                    if (!FAST && replay) x = a.u_replay[li * ndim + d];
                    else x = u52(h ? r.z : r.x, h ? r.w : r.y);
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
        if constexpr (OUT && NDIM_T != 0) {
            if (a.x_out) {
                double* row = a.x_out + li * NDIM_T;
                if constexpr (NDIM_T % 4 == 0) {
#pragma unroll
                    for (int d = 0; d < NDIM_T; d += 4)
                        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(row + d), "d"(xrow[d]), "d"(xrow[d + 1]),
                                     "d"(xrow[d + 2]), "d"(xrow[d + 3]) : "memory");
                } else if constexpr (NDIM_T % 2 == 0) {
#pragma unroll
                    for (int d = 0; d < NDIM_T; d += 2) *reinterpret_cast<double2*>(row + d) = make_double2(xrow[d], xrow[d + 1]);
                } else {
#pragma unroll
                    for (int d = 0; d < NDIM_T; ++d) row[d] = xrow[d];
                }
            }
        }
        if constexpr (fused) {
            const double f = pdf_acc_finish<KIND>(dens, pa);
            acc.add(f);
            if constexpr (OUT) {
                if (a.f_out) a.f_out[li] = f;
            }
        }
    }
    if constexpr (!fused) return;
    Stats st = warp_merge(acc.stats());
    if (lane == 0) s_stats[warp] = st;
    __syncthreads();
    if (threadIdx.x == 0) {
        Stats r{0.0, 0.0, 0.0};
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = merge(r, s_stats[w]);
        double* o = a.stats_partials + (size_t)blockIdx.x * 3;
        o[0] = r.n; o[1] = r.mean; o[2] = r.m2;
    }
}


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
#define HEPMC_FAST_LAUNCH(KIND, ND, OUT)                                                                    \
    do {                                                                                                    \
        if (vegas) {                                                                                        \
            HEPMC_CUDA(cudaFuncSetAttribute(vegas_sample_kernel<KIND, ND, true, OUT>,                       \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
            HEPMC_CUDA(cudaFuncSetAttribute(vegas_sample_kernel<KIND, ND, true, OUT>,                       \
                                            cudaFuncAttributePreferredSharedMemoryCarveout, 100));          \
            vegas_sample_kernel<KIND, ND, true, OUT><<<blocks, threads, smem, st>>>(dens, a);               \
        } else {                                                                                            \
            plain_sample_kernel<KIND, ND, true, OUT><<<blocks, 256, 0, st>>>(dens, a);                      \
        }                                                                                                   \
    } while (0)
#define HEPMC_FAST_CASE(KIND, ND)                                                                           \
    case ND:                                                                                                \
        if (out) HEPMC_FAST_LAUNCH(KIND, ND, true);                                                         \
        else if (KIND != HEPMC_NONE) HEPMC_FAST_LAUNCH(KIND, ND, false);                                    \
        else return -100;                                                                                   \
        break;

}  // namespace hepmc

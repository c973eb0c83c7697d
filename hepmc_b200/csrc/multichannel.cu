// Multi-channel importance sampling (core/integration/multi_channel.py:58-176,
// importance.py:141-175): per point choose a channel with probability alpha_c, draw x from
// it, evaluate the mixture density  p(x) = sum_c alpha_c p_c(x)  and the integrand, and
// accumulate  mean(f/p)  plus per-channel  (count_c, sum_c (f/p)^2)  for the weight update.
// Channels are diagonal Gaussians or uniform boxes; the channel table (<= 16 channels) is
// staged in shared memory.  Same deterministic chunk/slot reduction as the other integrators:
// per-channel sums use a thread-private shared-memory column, no atomics.
#include "density.cuh"

namespace hepmc {

constexpr int kMaxCh = HEPMC_MAX_CHANNELS;
constexpr int kChDim = HEPMC_MAX_CHANNEL_DIM;
constexpr int kChRow = 8 + 3 * kChDim;  // [kind, alpha, lognorm|low, high, vol, -, -, -, mean[], scale[], sinv[]]

struct McArgs {
    const double* table;  // (nch, kChRow) device
    int nch, ndim;
    int64_t n, first_index, chunk;
    uint64_t seed;
    uint32_t stream_id;
    const double* xs_replay;    // (n, ndim)
    const int64_t* ch_replay;   // (n,)
    double* x_out;              // (n, ndim)
    double* f_out;              // (n,)
    double* p_out;              // (n,) mixture density
    int64_t* ch_out;            // (n,)
    double* stats_partials;     // (n_chunks, 3)
    double* ch_partials;        // (n_chunks, 2 * nch): counts then sums of (f/p)^2
    const double* f_in;         // accumulate path
    const double* p_in;
    const int64_t* ch_in;
};

__device__ __forceinline__ double channel_pdf(const double* row, int ndim, const double* x) {
    if ((int)row[0] == HEPMC_GAUSSIAN) {
        double m = 0.0;
        for (int d = 0; d < ndim; ++d) {
            const double v = (x[d] - row[8 + d]) * row[8 + 2 * kChDim + d];
            m = fma(v, v, m);
        }
        return exp(-0.5 * (row[2] + m));  // gaussian.py:30-33 via SciPy logpdf
    }
    bool inside = true;
    for (int d = 0; d < ndim; ++d) inside = inside && (x[d] > row[2]) && (x[d] < row[3]);
    return inside ? 1.0 / row[4] : 0.0;   // uniform.py:15-19
}

template <int NDIM>
__global__ void __launch_bounds__(128) mc_sample_kernel(const __grid_constant__ Dens dens,
                                                        const __grid_constant__ McArgs a) {
    extern __shared__ double sm[];
    double* s_tab = sm;                                // nch * kChRow
    double* s_cnt = s_tab + a.nch * kChRow;            // [nch][blockDim]
    double* s_sq = s_cnt + a.nch * blockDim.x;         // [nch][blockDim]
    __shared__ Stats s_stats[4];
    for (int i = threadIdx.x; i < a.nch * kChRow; i += blockDim.x) s_tab[i] = a.table[i];
    for (int i = threadIdx.x; i < 2 * a.nch * (int)blockDim.x; i += blockDim.x) s_cnt[i] = 0.0;
    __syncthreads();
    const bool fused = dens.kind != HEPMC_NONE;
    const PhiloxKey key = philox_expand(a.seed);
    const int64_t chunk0 = (int64_t)blockIdx.x * a.chunk;
    const int ppt = (int)(a.chunk / blockDim.x);
    constexpr int NCALL = (NDIM + 1) / 2;
    ShiftedAcc acc;
    acc.init();
    for (int it = 0; it < ppt; ++it) {
        const int64_t li = chunk0 + (int64_t)it * blockDim.x + threadIdx.x;
        if (li >= a.n) break;
        double x[NDIM];
        int c = 0;
        if (a.xs_replay) {
#pragma unroll
            for (int d = 0; d < NDIM; ++d) x[d] = a.xs_replay[li * NDIM + d];
            c = (int)a.ch_replay[li];
        } else {
            const uint64_t g = (uint64_t)(a.first_index + li);
            const U4 r0 = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0u, a.stream_id, key);
            const double uc = u52(r0.x, r0.y);
            // channel with cumulative weight first exceeding uc: i.i.d. categorical draws, so the
            // per-channel counts are multinomial like multi_channel.py:141
            double cum = 0.0;
            c = a.nch - 1;
            for (int k = 0; k < a.nch; ++k) {
                cum += s_tab[k * kChRow + 1];
                if (uc < cum) { c = k; break; }
            }
            const double* row = s_tab + c * kChRow;
            const bool gauss = (int)row[0] == HEPMC_GAUSSIAN;
#pragma unroll
            for (int k = 0; k < NCALL; ++k) {
                const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 1u + k, a.stream_id, key);
                double z0, z1;
                if (gauss) {
                    const double u1 = u52(r.x, r.y), u2 = u52(r.z, r.w);
                    const double rad = sqrt(-2.0 * fm::log(u1));
                    double sn, cs;
                    fm::sincos2pi(u2, &sn, &cs);
                    z0 = rad * cs;
                    z1 = rad * sn;
                    x[2 * k] = row[8 + 2 * k] + row[8 + kChDim + 2 * k] * z0;       // gaussian.py:49-51
                    if (2 * k + 1 < NDIM) x[(2 * k + 1) % NDIM] = row[8 + (2 * k + 1) % NDIM] + row[8 + kChDim + (2 * k + 1) % NDIM] * z1;
                } else {
                    z0 = u52(r.x, r.y);
                    z1 = u52(r.z, r.w);
                    x[2 * k] = row[2] + (row[3] - row[2]) * z0;                       // uniform.py:33-34
                    if (2 * k + 1 < NDIM) x[(2 * k + 1) % NDIM] = row[2] + (row[3] - row[2]) * z1;
                }
            }
        }
        double p = 0.0;
        for (int k = 0; k < a.nch; ++k) p += s_tab[k * kChRow + 1] * channel_pdf(s_tab + k * kChRow, NDIM, x);  // :111-114
        if (a.x_out) {
#pragma unroll
            for (int d = 0; d < NDIM; ++d) a.x_out[li * NDIM + d] = x[d];
        }
        if (a.p_out) a.p_out[li] = p;
        if (a.ch_out) a.ch_out[li] = c;
        if (fused) {
            const double f = density_pdf_n<NDIM>(dens, x);
            if (a.f_out) a.f_out[li] = f;
            const double wgt = f / p;                        // importance.py:163
            acc.add(wgt);
            s_cnt[c * blockDim.x + threadIdx.x] += 1.0;
            s_sq[c * blockDim.x + threadIdx.x] += wgt * wgt;  // :166-167
        }
    }
    if (!fused) return;
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * a.nch; j += blockDim.x) {
        double s = 0.0;
        for (int t = 0; t < (int)blockDim.x; ++t) s += s_cnt[j * blockDim.x + t];
        a.ch_partials[(size_t)blockIdx.x * 2 * a.nch + j] = s;
    }
    Stats st = warp_merge(acc.stats());
    if ((threadIdx.x & 31) == 0) s_stats[threadIdx.x >> 5] = st;
    __syncthreads();
    if (threadIdx.x == 0) {
        Stats r{0.0, 0.0, 0.0};
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = merge(r, s_stats[w]);
        double* o = a.stats_partials + (size_t)blockIdx.x * 3;
        o[0] = r.n; o[1] = r.mean; o[2] = r.m2;
    }
}

// two-kernel path: user integrand values f, mixture densities p and channel ids -> partials
__global__ void __launch_bounds__(128) mc_accumulate_kernel(const __grid_constant__ McArgs a) {
    extern __shared__ double sm[];
    double* s_cnt = sm;
    double* s_sq = s_cnt + a.nch * blockDim.x;
    __shared__ Stats s_stats[4];
    for (int i = threadIdx.x; i < 2 * a.nch * (int)blockDim.x; i += blockDim.x) s_cnt[i] = 0.0;
    __syncthreads();
    const int64_t chunk0 = (int64_t)blockIdx.x * a.chunk;
    const int ppt = (int)(a.chunk / blockDim.x);
    ShiftedAcc acc;
    acc.init();
    for (int it = 0; it < ppt; ++it) {
        const int64_t li = chunk0 + (int64_t)it * blockDim.x + threadIdx.x;
        if (li >= a.n) break;
        const double wgt = a.f_in[li] / a.p_in[li];
        const int c = (int)a.ch_in[li];
        acc.add(wgt);
        s_cnt[c * blockDim.x + threadIdx.x] += 1.0;
        s_sq[c * blockDim.x + threadIdx.x] += wgt * wgt;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * a.nch; j += blockDim.x) {
        double s = 0.0;
        for (int t = 0; t < (int)blockDim.x; ++t) s += s_cnt[j * blockDim.x + t];
        a.ch_partials[(size_t)blockIdx.x * 2 * a.nch + j] = s;
    }
    Stats st = warp_merge(acc.stats());
    if ((threadIdx.x & 31) == 0) s_stats[threadIdx.x >> 5] = st;
    __syncthreads();
    if (threadIdx.x == 0) {
        Stats r{0.0, 0.0, 0.0};
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = merge(r, s_stats[w]);
        double* o = a.stats_partials + (size_t)blockIdx.x * 3;
        o[0] = r.n; o[1] = r.mean; o[2] = r.m2;
    }
}

// out[i, :] = src[order[i], :]: the by-channel grouping of a multi-channel sample
// (multi_channel.py:141-150 returns the points channel by channel).  One thread per output
// element pair; reads are gathered, writes coalesced.
__global__ void __launch_bounds__(256) gather_rows_kernel(const double* __restrict__ src, const int64_t* __restrict__ order,
                                                          int64_t n, int row_len, double* __restrict__ out) {
    const int64_t total = n * row_len;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / row_len;
        const int k = (int)(e - i * row_len);
        out[e] = src[order[i] * row_len + k];
    }
}

}  // namespace hepmc

using namespace hepmc;

extern "C" {

int hepmc_multichannel_sample(const hepmc_density_t* dens, const double* channel_table_dev, int n_channels, int ndim,
                              int64_t n, int64_t first_index, int64_t chunk, uint64_t seed, uint32_t stream_id,
                              const double* xs_replay_dev, const int64_t* ch_replay_dev,
                              double* x_dev, double* f_dev, double* p_dev, int64_t* ch_dev,
                              double* stats_partials_dev, double* ch_partials_dev, void* stream) {
    HEPMC_REQUIRE(dens && channel_table_dev, HEPMC_E_ARG, "hepmc_multichannel_sample: NULL dens/table");
    HEPMC_REQUIRE(n_channels >= 1 && n_channels <= kMaxCh, HEPMC_E_UNSUPPORTED,
                  "hepmc_multichannel_sample: %d channels outside [1,%d]", n_channels, kMaxCh);
    HEPMC_REQUIRE(ndim >= 1 && ndim <= kChDim, HEPMC_E_UNSUPPORTED, "hepmc_multichannel_sample: ndim %d outside [1,%d]", ndim, kChDim);
    HEPMC_REQUIRE(n >= 0 && first_index >= 0 && chunk >= 128 && chunk % 128 == 0, HEPMC_E_ARG,
                  "hepmc_multichannel_sample: bad n/first_index/chunk");
    HEPMC_REQUIRE((xs_replay_dev == nullptr) == (ch_replay_dev == nullptr), HEPMC_E_ARG, "replay needs xs and channel ids");
    const bool fused = dens->kind != HEPMC_NONE;
    // partial rows follow the chunk grid; sample-only calls write none and may start anywhere
    HEPMC_REQUIRE(!(fused || stats_partials_dev || ch_partials_dev) || first_index % chunk == 0, HEPMC_E_ARG,
                  "hepmc_multichannel_sample: first_index must be a multiple of chunk when partials are written");
    HEPMC_REQUIRE(!fused || dens->ndim == ndim, HEPMC_E_ARG, "integrand ndim %d != channel ndim %d", dens->ndim, ndim);
    HEPMC_REQUIRE(!fused || (stats_partials_dev && ch_partials_dev), HEPMC_E_ARG, "partial buffers are NULL");
    HEPMC_REQUIRE(fused || (x_dev && p_dev && ch_dev), HEPMC_E_ARG, "two-kernel path needs x, p and channel outputs");
    if (n == 0) return 0;
    McArgs a{};
    a.table = channel_table_dev; a.nch = n_channels; a.ndim = ndim; a.n = n; a.first_index = first_index; a.chunk = chunk;
    a.seed = seed; a.stream_id = stream_id; a.xs_replay = xs_replay_dev; a.ch_replay = ch_replay_dev;
    a.x_out = x_dev; a.f_out = f_dev; a.p_out = p_dev; a.ch_out = ch_dev;
    a.stats_partials = stats_partials_dev; a.ch_partials = ch_partials_dev;
    const unsigned blocks = (unsigned)((n + chunk - 1) / chunk);
    const size_t smem = ((size_t)n_channels * kChRow + 2 * (size_t)n_channels * 128) * 8;
    cudaStream_t st = as_stream(stream);
#define MC_CASE(ND)                                                                                              \
    do {                                                                                                         \
        HEPMC_CUDA(cudaFuncSetAttribute(mc_sample_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        mc_sample_kernel<ND><<<blocks, 128, smem, st>>>(*dens, a);                                               \
    } while (0)
    switch (ndim) {
        case 1: MC_CASE(1); break;   case 2: MC_CASE(2); break;   case 3: MC_CASE(3); break;   case 4: MC_CASE(4); break;
        case 5: MC_CASE(5); break;   case 6: MC_CASE(6); break;   case 7: MC_CASE(7); break;   case 8: MC_CASE(8); break;
        case 9: MC_CASE(9); break;   case 10: MC_CASE(10); break; case 11: MC_CASE(11); break; case 12: MC_CASE(12); break;
        case 13: MC_CASE(13); break; case 14: MC_CASE(14); break; case 15: MC_CASE(15); break; case 16: MC_CASE(16); break;
        default: break;
    }
#undef MC_CASE
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_multichannel_accumulate(const double* f_dev, const double* p_dev, const int64_t* ch_dev, int n_channels,
                                  int64_t n, int64_t chunk, double* stats_partials_dev, double* ch_partials_dev,
                                  void* stream) {
    HEPMC_REQUIRE(f_dev && p_dev && ch_dev && stats_partials_dev && ch_partials_dev, HEPMC_E_ARG,
                  "hepmc_multichannel_accumulate: NULL buffer");
    HEPMC_REQUIRE(n_channels >= 1 && n_channels <= kMaxCh && n >= 0 && chunk >= 128 && chunk % 128 == 0, HEPMC_E_ARG,
                  "hepmc_multichannel_accumulate: bad sizes");
    if (n == 0) return 0;
    McArgs a{};
    a.nch = n_channels; a.n = n; a.chunk = chunk; a.f_in = f_dev; a.p_in = p_dev; a.ch_in = ch_dev;
    a.stats_partials = stats_partials_dev; a.ch_partials = ch_partials_dev;
    const unsigned blocks = (unsigned)((n + chunk - 1) / chunk);
    const size_t smem = 2 * (size_t)n_channels * 128 * 8;
    HEPMC_CUDA(cudaFuncSetAttribute(mc_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mc_accumulate_kernel<<<blocks, 128, smem, as_stream(stream)>>>(a);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_gather_rows(const double* src_dev, const int64_t* order_dev, int64_t n, int row_len, double* out_dev,
                      void* stream) {
    HEPMC_REQUIRE(n >= 0 && row_len >= 1, HEPMC_E_ARG, "hepmc_gather_rows: bad sizes");
    if (n == 0) return 0;
    HEPMC_REQUIRE(src_dev && order_dev && out_dev, HEPMC_E_ARG, "hepmc_gather_rows: NULL buffer");
    int64_t blocks = (n * row_len + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    gather_rows_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(src_dev, order_dev, n, row_len, out_dev);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

// Sample statistics ("next" row 3): lag autocovariance / effective sample size and the bin-wise
// chi^2 of a sample against its target (core/util/stat_tools.py:11-129).  These are the
// reductions Sample.__repr__ shows (core/sampling.py:79-94).  All sums are fixed-order
// (thread-strided partials -> shuffle butterfly -> warp partials in warp order): the results do
// not depend on scheduling.  Histogram counts use 64-bit integer atomics (order-independent).
#include "common.cuh"

namespace hepmc {

constexpr int kLagsPerPass = 8;
constexpr int kStatThreads = 256;
constexpr int kStatWarps = kStatThreads / 32;

// S[j] = sum_{t < T-(lag+j)} (x[t]-m) (x[t+lag+j]-m),  j = 0..7  (stat_tools.py:35-36 without
// the 1/size).  x has element stride `stride`.  Every thread of the CTA returns the same S.
__device__ __forceinline__ void lag_products(const double* __restrict__ x, int64_t stride, int64_t T, double m,
                                             int64_t lag, double (*sm)[kLagsPerPass], double S[kLagsPerPass]) {
    double s[kLagsPerPass];
#pragma unroll
    for (int j = 0; j < kLagsPerPass; ++j) s[j] = 0.0;
    for (int64_t t = threadIdx.x; t < T - lag; t += kStatThreads) {
        const double a = x[t * stride] - m;
#pragma unroll
        for (int j = 0; j < kLagsPerPass; ++j) {
            const int64_t u = t + lag + j;
            if (u < T) s[j] = fma(a, x[u * stride] - m, s[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < kLagsPerPass; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
    }
    const int warp = threadIdx.x >> 5;
    __syncthreads();  // previous pass has been read
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int j = 0; j < kLagsPerPass; ++j) sm[warp][j] = s[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kLagsPerPass; ++j) {
        double r = sm[0][j];
#pragma unroll
        for (int w = 1; w < kStatWarps; ++w) r += sm[w][j];
        S[j] = r;
    }
}

// acov[k, d] for k in [lag_begin, lag_begin + n_lags): one CTA per (group of 8 lags, series)
__global__ void __launch_bounds__(kStatThreads) autocov_kernel(const double* __restrict__ data, int64_t T, int ndim,
                                                               const double* __restrict__ mean,
                                                               const double* __restrict__ var, int64_t lag_begin,
                                                               int64_t n_lags, double* __restrict__ out) {
    __shared__ double sm[kStatWarps][kLagsPerPass];
    const int64_t series = blockIdx.y;           // chain * ndim + dim
    const int64_t c = series / ndim;
    const int d = (int)(series % ndim);
    const double* x = data + c * T * ndim + d;
    const int64_t lag0 = lag_begin + (int64_t)blockIdx.x * kLagsPerPass;
    double S[kLagsPerPass];
    lag_products(x, ndim, T, mean[d], lag0 == 0 ? 1 : lag0, sm, S);   // lag 0 is `var` (stat_tools.py:34)
    if (threadIdx.x < kLagsPerPass) {
        const int64_t k = lag0 + threadIdx.x;
        if (k < lag_begin + n_lags && k < T) {
            double v;
            if (k == 0) {
                v = var[d];
            } else {
                // the pass started at lag 1 when lag0 == 0: slot j holds lag 1 + j
                const int j = lag0 == 0 ? threadIdx.x - 1 : threadIdx.x;
                v = S[j] / (double)T;
            }
            out[(c * n_lags + (k - lag_begin)) * ndim + d] = v;
        }
    }
}

// effective_sample_size (stat_tools.py:46-70): one CTA per (chain, dim); sums
// (1 - lag/T) rho_lag from lag 1 while rho >= 0.05 and lag < T - 1.
__global__ void __launch_bounds__(kStatThreads) ess_kernel(const double* __restrict__ data, int64_t T, int ndim,
                                                           const double* __restrict__ mean,
                                                           const double* __restrict__ var, double* __restrict__ ess,
                                                           int64_t* __restrict__ lag_out) {
    __shared__ double sm[kStatWarps][kLagsPerPass];
    const int64_t series = blockIdx.x;
    const int64_t c = series / ndim;
    const int d = (int)(series % ndim);
    const double* x = data + c * T * ndim + d;
    const double m = mean[d], v = var[d];
    const double size = (double)T;
    double sum = 0.0;
    int64_t lag = 1, last = 1;
    bool done = false;
    while (!done) {
        double S[kLagsPerPass];
        lag_products(x, ndim, T, m, lag, sm, S);
        for (int j = 0; j < kLagsPerPass; ++j) {
            const int64_t k = lag + j;
            last = k;
            const double rho = (S[j] / size) / v * (size / (size - (double)k));   // acov/acov[0] * unbias
            if (!(rho >= 0.05 && k < T - 1)) {
                done = true;
                break;
            }
            sum = sum + (1.0 - (double)k / size) * rho;
        }
        lag += kLagsPerPass;
    }
    if (threadIdx.x == 0) {
        ess[series] = size / (1.0 + 2.0 * sum);
        if (lag_out) lag_out[series] = last;
    }
}

// ---------------------------------------------------------------- histogramdd
struct HistArgs {
    const double* data;   // (n, ndim)
    int64_t n;
    int ndim;
    const double* edges;  // concatenated per-dim edge arrays
    int edge_off[HEPMC_MAX_DIM];
    int nbins[HEPMC_MAX_DIM];
    int64_t stride[HEPMC_MAX_DIM];   // C order: last dim fastest
    unsigned long long* counts;
};

// bin of x in edges e[0..nb]: (#edges <= x) - 1, the last edge belongs to the last bin
// (np.histogramdd: searchsorted(..., 'right') and the on_edge correction); -1 = outside / NaN
__device__ __forceinline__ int find_bin(const double* __restrict__ e, int nb, double x) {
    const double e0 = e[0], eN = e[nb];
    if (!(x >= e0 && x <= eN)) return -1;
    int b = (int)((x - e0) / (eN - e0) * (double)nb);
    b = b < 0 ? 0 : (b > nb - 1 ? nb - 1 : b);
    while (b > 0 && x < e[b]) --b;
    while (b < nb - 1 && x >= e[b + 1]) ++b;
    return b;
}

__global__ void __launch_bounds__(256) histdd_kernel(const __grid_constant__ HistArgs a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t flat = 0;
        bool inside = true;
        for (int d = 0; d < a.ndim; ++d) {
            const int b = find_bin(a.edges + a.edge_off[d], a.nbins[d], a.data[i * a.ndim + d]);
            inside = inside && (b >= 0);
            flat += (int64_t)(b < 0 ? 0 : b) * a.stride[d];
        }
        if (inside) atomicAdd(a.counts + flat, 1ull);
    }
}

// int_steps uniform points in each listed bin (stat_tools.py:122-123):
// x = low + (high - low) * u, u from the replay tape (m, int_steps, ndim) or Philox
// (item = global point index, block = dim / 2).
struct BinPointArgs {
    const double* edges;
    int edge_off[HEPMC_MAX_DIM];
    int64_t stride[HEPMC_MAX_DIM];
    int nbins[HEPMC_MAX_DIM];
    int ndim;
    const int64_t* flat;   // (m,) flat bin indices
    int64_t m;
    int int_steps;
    uint64_t seed;
    uint32_t stream_id;
    const double* u_replay;
    double* out;           // (m * int_steps, ndim)
};

__global__ void __launch_bounds__(256) bin_points_kernel(const __grid_constant__ BinPointArgs a) {
    const PhiloxKey key = philox_expand(a.seed);
    const int64_t total = a.m * a.int_steps;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t rem = a.flat[i / a.int_steps];
        U4 r{0u, 0u, 0u, 0u};
        for (int d = 0; d < a.ndim; ++d) {
            const int b = (int)(rem / a.stride[d]);
            rem -= (int64_t)b * a.stride[d];
            const double* e = a.edges + a.edge_off[d];
            const double low = e[b], high = e[b + 1];
            double u;
            if (a.u_replay) {
                u = a.u_replay[i * a.ndim + d];
            } else {
                if ((d & 1) == 0) r = philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)(d >> 1), a.stream_id, key);
                u = (d & 1) ? u52(r.z, r.w) : u52(r.x, r.y);
            }
            a.out[i * a.ndim + d] = low + (high - low) * u;
        }
    }
}

// expected[j] = mean_s pdf[j, s] * size * vol  (stat_tools.py:122-123), one warp per bin
__global__ void __launch_bounds__(256) bin_expected_kernel(const double* __restrict__ pdf, int64_t m, int int_steps,
                                                           double size, double vol, double* __restrict__ expected) {
    const int lane = threadIdx.x & 31;
    for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < m; j += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        double s = 0.0;
        for (int k = lane; k < int_steps; k += 32) s += pdf[j * int_steps + k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) expected[j] = s / (double)int_steps * size * vol;
    }
}

// chi2 = sum_{expected >= min_count} (obs - expected)^2 / expected  (stat_tools.py:125-127 +
// scipy.stats.chisquare), single CTA, fixed order.  out = {chi2, number of bins used}.
__global__ void __launch_bounds__(kStatThreads) chi2_kernel(const int64_t* __restrict__ flat,
                                                            const unsigned long long* __restrict__ counts,
                                                            const double* __restrict__ expected, int64_t m,
                                                            double min_count, double* __restrict__ out) {
    __shared__ double sm[kStatWarps][2];
    double s = 0.0, n = 0.0;
    for (int64_t j = threadIdx.x; j < m; j += kStatThreads) {
        const double e = expected[j];
        if (e >= min_count) {
            const double o = (double)counts[flat[j]];
            const double diff = o - e;
            s += diff * diff / e;
            n += 1.0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sm[threadIdx.x >> 5][0] = s;
        sm[threadIdx.x >> 5][1] = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double rs = sm[0][0], rn = sm[0][1];
        for (int w = 1; w < kStatWarps; ++w) {
            rs += sm[w][0];
            rn += sm[w][1];
        }
        out[0] = rs;
        out[1] = rn;
    }
}

static int fill_grid(int ndim, const int* nbins_host, int* edge_off, int* nbins, int64_t* stride, int64_t* total,
                     const char* who) {
    HEPMC_REQUIRE(ndim >= 1 && ndim <= HEPMC_MAX_DIM, HEPMC_E_UNSUPPORTED, "%s: ndim %d outside [1,%d]", who, ndim,
                  HEPMC_MAX_DIM);
    HEPMC_REQUIRE(nbins_host != nullptr, HEPMC_E_ARG, "%s: nbins_host is NULL", who);
    int off = 0;
    for (int d = 0; d < ndim; ++d) {
        HEPMC_REQUIRE(nbins_host[d] >= 1, HEPMC_E_ARG, "%s: nbins[%d] = %d < 1", who, d, nbins_host[d]);
        edge_off[d] = off;
        nbins[d] = nbins_host[d];
        off += nbins_host[d] + 1;
    }
    int64_t s = 1;
    for (int d = ndim - 1; d >= 0; --d) {
        stride[d] = s;
        HEPMC_REQUIRE(s <= (int64_t)1 << 40, HEPMC_E_UNSUPPORTED, "%s: more than 2^40 bins", who);
        s *= nbins_host[d];
    }
    *total = s;
    return 0;
}

}  // namespace hepmc

using namespace hepmc;

extern "C" {

int hepmc_autocov(const double* data_dev, int64_t n_chains, int64_t n_steps, int ndim, const double* mean_dev,
                  const double* var_dev, int64_t lag_begin, int64_t n_lags, double* out_dev, void* stream) {
    HEPMC_REQUIRE(ndim >= 1 && n_chains >= 0 && n_steps >= 1 && lag_begin >= 0 && n_lags >= 0, HEPMC_E_ARG,
                  "hepmc_autocov: bad sizes");
    if (n_chains == 0 || n_lags == 0) return 0;
    HEPMC_REQUIRE(data_dev && mean_dev && var_dev && out_dev, HEPMC_E_ARG, "hepmc_autocov: NULL buffer");
    HEPMC_REQUIRE(n_chains * ndim <= 65535, HEPMC_E_UNSUPPORTED, "hepmc_autocov: more than 65535 series per call");
    dim3 grid((unsigned)((n_lags + kLagsPerPass - 1) / kLagsPerPass), (unsigned)(n_chains * ndim));
    autocov_kernel<<<grid, kStatThreads, 0, as_stream(stream)>>>(data_dev, n_steps, ndim, mean_dev, var_dev, lag_begin,
                                                                 n_lags, out_dev);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_effective_sample_size(const double* data_dev, int64_t n_chains, int64_t n_steps, int ndim,
                                const double* mean_dev, const double* var_dev, double* ess_dev, int64_t* lag_dev,
                                void* stream) {
    HEPMC_REQUIRE(ndim >= 1 && n_chains >= 0, HEPMC_E_ARG, "hepmc_effective_sample_size: bad sizes");
    HEPMC_REQUIRE(n_steps >= 2, HEPMC_E_ARG, "hepmc_effective_sample_size: needs at least 2 states per chain (got %lld)",
                  (long long)n_steps);
    if (n_chains == 0) return 0;
    HEPMC_REQUIRE(data_dev && mean_dev && var_dev && ess_dev, HEPMC_E_ARG, "hepmc_effective_sample_size: NULL buffer");
    HEPMC_REQUIRE(n_chains * ndim <= 0x7fffffffLL, HEPMC_E_UNSUPPORTED, "hepmc_effective_sample_size: too many series");
    ess_kernel<<<(unsigned)(n_chains * ndim), kStatThreads, 0, as_stream(stream)>>>(data_dev, n_steps, ndim, mean_dev,
                                                                                  var_dev, ess_dev, lag_dev);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_histogramdd(const double* data_dev, int64_t n, int ndim, const double* edges_dev, const int* nbins_host,
                      unsigned long long* counts_dev, void* stream) {
    HistArgs a;
    int64_t total = 0;
    if (int e = fill_grid(ndim, nbins_host, a.edge_off, a.nbins, a.stride, &total, "hepmc_histogramdd")) return e;
    HEPMC_REQUIRE(n >= 0, HEPMC_E_ARG, "hepmc_histogramdd: n < 0");
    HEPMC_REQUIRE(edges_dev && counts_dev, HEPMC_E_ARG, "hepmc_histogramdd: NULL buffer");
    HEPMC_CUDA(cudaMemsetAsync(counts_dev, 0, (size_t)total * sizeof(unsigned long long), as_stream(stream)));
    if (n == 0) return 0;
    HEPMC_REQUIRE(data_dev != nullptr, HEPMC_E_ARG, "hepmc_histogramdd: data_dev is NULL");
    a.data = data_dev;
    a.n = n;
    a.ndim = ndim;
    a.edges = edges_dev;
    a.counts = counts_dev;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    histdd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(a);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_bin_points(int ndim, const double* edges_dev, const int* nbins_host, const int64_t* flat_dev, int64_t m,
                     int int_steps, uint64_t seed, uint32_t stream_id, const double* u_replay_dev, double* out_dev,
                     void* stream) {
    BinPointArgs a;
    int64_t total = 0;
    if (int e = fill_grid(ndim, nbins_host, a.edge_off, a.nbins, a.stride, &total, "hepmc_bin_points")) return e;
    HEPMC_REQUIRE(m >= 0 && int_steps >= 1, HEPMC_E_ARG, "hepmc_bin_points: bad sizes");
    if (m == 0) return 0;
    HEPMC_REQUIRE(edges_dev && flat_dev && out_dev, HEPMC_E_ARG, "hepmc_bin_points: NULL buffer");
    a.edges = edges_dev;
    a.ndim = ndim;
    a.flat = flat_dev;
    a.m = m;
    a.int_steps = int_steps;
    a.seed = seed;
    a.stream_id = stream_id;
    a.u_replay = u_replay_dev;
    a.out = out_dev;
    int64_t blocks = (m * int_steps + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    bin_points_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(a);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_bin_chi2(const double* pdf_dev, const int64_t* flat_dev, const unsigned long long* counts_dev, int64_t m,
                   int int_steps, double sample_size, double bin_volume, double min_count, double* expected_dev,
                   double* out_dev, void* stream) {
    HEPMC_REQUIRE(m >= 0 && int_steps >= 1, HEPMC_E_ARG, "hepmc_bin_chi2: bad sizes");
    HEPMC_REQUIRE(out_dev != nullptr, HEPMC_E_ARG, "hepmc_bin_chi2: out_dev is NULL");
    if (m > 0) {
        HEPMC_REQUIRE(pdf_dev && flat_dev && counts_dev && expected_dev, HEPMC_E_ARG, "hepmc_bin_chi2: NULL buffer");
        int64_t blocks = (m * 32 + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        bin_expected_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(pdf_dev, m, int_steps, sample_size, bin_volume,
                                                                           expected_dev);
        HEPMC_LAUNCH_CHECK();
    }
    chi2_kernel<<<1, kStatThreads, 0, as_stream(stream)>>>(flat_dev, counts_dev, expected_dev, m, min_count, out_dev);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

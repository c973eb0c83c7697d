// Stand-alone ELM surrogate evaluation: FunctionBasis.eval, eval_gradient, output_matrix
// (surrogate/extreme_learning.py:18-29,120-146,190-230).  One thread per point, node
// tables broadcast from shared memory (see elm.cuh).
#include "chains.cuh"

namespace hepmc {

template <int NDIM, class T>
__global__ void __launch_bounds__(128) elm_eval_kernel(const __grid_constant__ hepmc_elm_t elm, int what,
                                                       const double* __restrict__ xs, int64_t n,
                                                       double* __restrict__ out) {
    extern __shared__ double smem_raw[];
    T* tab = reinterpret_cast<T*>(smem_raw);
    elm_stage_tables<T>(elm, NDIM, tab);
    __syncthreads();
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        double x[NDIM];
#pragma unroll
        for (int k = 0; k < NDIM; ++k) x[k] = xs[j * NDIM + k];
        if (what == 0) {
            out[j] = elm_value<NDIM, T>(elm.mode, elm.nodes, tab, x) + elm.out_bias;  // eval :26-29
        } else {
            double g[NDIM];
            elm_gradient<NDIM, T>(elm.mode, elm.nodes, tab, x, g);
#pragma unroll
            for (int k = 0; k < NDIM; ++k) out[j * NDIM + k] = g[k];
        }
    }
}

// output_matrix (n, nodes): thread per (point, node) pair, nodes fastest -> coalesced rows
__global__ void __launch_bounds__(256) elm_output_matrix_kernel(const __grid_constant__ hepmc_elm_t elm, int ndim,
                                                                const double* __restrict__ xs, int64_t n,
                                                                double* __restrict__ out) {
    const int64_t total = n * elm.nodes;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = t / elm.nodes;
        const int i = (int)(t % elm.nodes);
        const double* x = xs + j * ndim;
        const double* r = elm.p0_dev + (size_t)i * ndim;
        if (elm.mode == HEPMC_GRAD_ELM_GAUSS) {
            double ss = 0.0;
            for (int k = 0; k < ndim; ++k) {
                const double d = x[k] - r[k];
                ss = fma(d, d, ss);
            }
            const double v = sqrt(ss) / elm.p1_dev[i];  // :203
            out[t] = exp(-(v * v / 2.0));               // :204, GaussianBasis :289-290
        } else {
            double in = elm.p1_dev[i];
            for (int k = 0; k < ndim; ++k) in = fma(x[k], r[k], in);  // :126
            out[t] = cos(in);                                          // TrigBasis :253-254
        }
    }
}

}  // namespace hepmc

using namespace hepmc;

extern "C" int hepmc_elm_eval(const hepmc_elm_t* elm, int ndim, int what, int precision,
                              const double* xs_dev, int64_t n, double* out_dev, void* stream) {
    HEPMC_REQUIRE(elm != nullptr, HEPMC_E_ARG, "hepmc_elm_eval: elm is NULL");
    HEPMC_REQUIRE(elm->mode == HEPMC_GRAD_ELM_GAUSS || elm->mode == HEPMC_GRAD_ELM_TRIG, HEPMC_E_ARG,
                  "hepmc_elm_eval: bad basis mode %d", elm->mode);
    HEPMC_REQUIRE(elm->nodes >= 1 && elm->p0_dev && elm->p1_dev, HEPMC_E_ARG, "hepmc_elm_eval: incomplete parameters");
    HEPMC_REQUIRE(what >= 0 && what <= 2, HEPMC_E_ARG, "hepmc_elm_eval: what must be 0, 1 or 2");
    HEPMC_REQUIRE(what == 2 || elm->out_w_dev, HEPMC_E_ARG, "hepmc_elm_eval: out_w_dev is NULL");
    HEPMC_REQUIRE(precision >= 0 && precision <= 3, HEPMC_E_ARG, "hepmc_elm_eval: precision must be 0, 1, 2 or 3");
    if (precision >= 2) {
        HEPMC_REQUIRE(what == 1 && elm->mode == HEPMC_GRAD_ELM_GAUSS, HEPMC_E_UNSUPPORTED,
                      "hepmc_elm_eval: the tensor-core path implements eval_gradient of the Gaussian basis");
        if (n == 0) return 0;
        HEPMC_REQUIRE(xs_dev && out_dev, HEPMC_E_ARG, "hepmc_elm_eval: NULL buffer");
        return launch_elm_grad_tc(*elm, ndim, xs_dev, n, out_dev, precision == 3, as_stream(stream));
    }
    HEPMC_REQUIRE(n >= 0, HEPMC_E_ARG, "hepmc_elm_eval: n < 0");
    if (n == 0) return 0;
    HEPMC_REQUIRE(xs_dev && out_dev, HEPMC_E_ARG, "hepmc_elm_eval: NULL buffer");
    cudaStream_t st = as_stream(stream);
    if (what == 2) {
        HEPMC_REQUIRE(ndim >= 1 && ndim <= HEPMC_MAX_DIM, HEPMC_E_UNSUPPORTED, "hepmc_elm_eval: ndim %d", ndim);
        int64_t blocks = (n * elm->nodes + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        elm_output_matrix_kernel<<<(unsigned)blocks, 256, 0, st>>>(*elm, ndim, xs_dev, n, out_dev);
        HEPMC_LAUNCH_CHECK();
        return 0;
    }
    HEPMC_REQUIRE(ndim >= 1 && ndim <= HEPMC_MAX_CHAIN_DIM, HEPMC_E_UNSUPPORTED,
                  "hepmc_elm_eval: ndim %d outside [1,%d]", ndim, HEPMC_MAX_CHAIN_DIM);
    int dev = 0, maxs = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxs, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t smem = (size_t)elm->nodes * elm_row(ndim) * (precision == 0 ? 8 : 4);
    HEPMC_REQUIRE(smem <= (size_t)maxs, HEPMC_E_UNSUPPORTED, "hepmc_elm_eval: %d nodes do not fit shared memory", elm->nodes);
    int64_t blocks = (n + 127) / 128;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
#define ELM_CASE(ND)                                                                                          \
    do {                                                                                                      \
        if (precision == 0) {                                                                                 \
            HEPMC_CUDA(cudaFuncSetAttribute(elm_eval_kernel<ND, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            elm_eval_kernel<ND, double><<<(unsigned)blocks, 128, smem, st>>>(*elm, what, xs_dev, n, out_dev); \
        } else {                                                                                              \
            HEPMC_CUDA(cudaFuncSetAttribute(elm_eval_kernel<ND, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
            elm_eval_kernel<ND, float><<<(unsigned)blocks, 128, smem, st>>>(*elm, what, xs_dev, n, out_dev);  \
        }                                                                                                     \
    } while (0)
    HEPMC_DISPATCH_NDIM(ndim, ELM_CASE);
#undef ELM_CASE
    HEPMC_LAUNCH_CHECK();
    return 0;
}

// Extreme-learning-machine surrogate (surrogate/extreme_learning.py): SIMT evaluation of the
// value / gradient of  z(x) = bias + sum_i ow_i * g_i(x)  for the Gaussian radial basis and
// the trigonometric additive basis.  Node tables are staged once per CTA in shared memory
// in a transformed form (per node: ndim params, then two derived scalars), every thread of
// a warp reads the same node -> pure broadcast LDS.  T = double (parity) or float (fast).
#pragma once
#include "common.cuh"

namespace hepmc {

// per-node row layout: [p0_0 .. p0_{nd-1}, h, coef, ow]
//   GAUSS: p0 = centre, h = 1/(2 w^2), coef = ow / w^2      (extreme_learning.py:197-204,228)
//   TRIG : p0 = input weights, h = bias, coef = ow           (:126,143-145)
template <class T>
__device__ __forceinline__ void elm_stage_tables(const hepmc_elm_t& e, int ndim, T* tab) {
    const int row = ndim + 3;
    for (int i = threadIdx.x; i < e.nodes; i += blockDim.x) {
        for (int k = 0; k < ndim; ++k) tab[i * row + k] = (T)e.p0_dev[i * ndim + k];
        const double p1 = e.p1_dev[i], ow = e.out_w_dev[i];
        if (e.mode == HEPMC_GRAD_ELM_GAUSS) {
            tab[i * row + ndim] = (T)(1.0 / (2.0 * (p1 * p1)));
            tab[i * row + ndim + 1] = (T)(ow / (p1 * p1));
            tab[i * row + ndim + 2] = (T)ow;
        } else {
            tab[i * row + ndim] = (T)p1;
            tab[i * row + ndim + 1] = (T)ow;
            tab[i * row + ndim + 2] = (T)ow;
        }
    }
}
__host__ __device__ inline int elm_row(int ndim) { return ndim + 3; }

__device__ __forceinline__ double elm_exp(double x) { return exp(x); }
__device__ __forceinline__ float elm_exp(float x) { return __expf(x); }
__device__ __forceinline__ void elm_sincos(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ void elm_sincos(float x, float* s, float* c) { __sincosf(x, s, c); }

// gradient of the surrogate at x -> g (both double, NDIM compile-time)
// `first`, `stride`: this thread's share of the nodes (cooperative kernels: thread t of S takes
// nodes t, t + S, ...; the caller sums the partial gradients)
template <int NDIM, class T>
__device__ __forceinline__ void elm_gradient(int mode, int nodes, const T* __restrict__ tab,
                                             const double (&x)[NDIM], double (&g)[NDIM], int first = 0, int stride = 1) {
    constexpr int row = NDIM + 3;
    T xt[NDIM], acc[NDIM];
#pragma unroll
    for (int k = 0; k < NDIM; ++k) { xt[k] = (T)x[k]; acc[k] = (T)0; }
    if (mode == HEPMC_GRAD_ELM_GAUSS) {
        auto node = [&](int i) {
            const T* r = tab + i * row;
            T diff[NDIM];
            T ss0 = (T)0, ss1 = (T)0;
#pragma unroll
            for (int k = 0; k < NDIM; ++k) {
                diff[k] = xt[k] - r[k];
                if (sizeof(T) == 8 && (k & 1)) ss1 = fma(diff[k], diff[k], ss1);
                else ss0 = fma(diff[k], diff[k], ss0);
            }
            const T ss = sizeof(T) == 8 ? ss0 + ss1 : ss0;
            const T gi = -elm_exp(-(ss * r[NDIM])) * r[NDIM + 1];  // basis_function_gradient * ow/w^2
#pragma unroll
            for (int k = 0; k < NDIM; ++k) acc[k] = fma(gi, diff[k], acc[k]);
        };
        if constexpr (sizeof(T) == 8) {
            // float64: two nodes in flight and split square sums (the FP64 pipe is latency bound
            // at the occupancy 168 registers allow)
#pragma unroll 2
            for (int i = first; i < nodes; i += stride) node(i);
        } else {
            for (int i = first; i < nodes; i += stride) node(i);  // float32: compiler's own unrolling (issue bound)
        }
    } else {
        for (int i = first; i < nodes; i += stride) {
            const T* r = tab + i * row;
            T in = r[NDIM];
#pragma unroll
            for (int k = 0; k < NDIM; ++k) in = fma(xt[k], r[k], in);
            T sn, cs;
            elm_sincos(in, &sn, &cs);
            const T gi = -sn * r[NDIM + 1];
#pragma unroll
            for (int k = 0; k < NDIM; ++k) acc[k] = fma(gi, r[k], acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < NDIM; ++k) g[k] = (double)acc[k];
}

// surrogate value (without bias)
template <int NDIM, class T>
__device__ __forceinline__ double elm_value(int mode, int nodes, const T* __restrict__ tab,
                                            const double (&x)[NDIM]) {
    constexpr int row = NDIM + 3;
    T xt[NDIM];
#pragma unroll
    for (int k = 0; k < NDIM; ++k) xt[k] = (T)x[k];
    T acc = (T)0;
    if (mode == HEPMC_GRAD_ELM_GAUSS) {
        for (int i = 0; i < nodes; ++i) {
            const T* r = tab + i * row;
            T ss = (T)0;
#pragma unroll
            for (int k = 0; k < NDIM; ++k) {
                const T d = xt[k] - r[k];
                ss = fma(d, d, ss);
            }
            acc = fma(elm_exp(-(ss * r[NDIM])), r[NDIM + 2], acc);
        }
    } else {
        for (int i = 0; i < nodes; ++i) {
            const T* r = tab + i * row;
            T in = r[NDIM];
#pragma unroll
            for (int k = 0; k < NDIM; ++k) in = fma(xt[k], r[k], in);
            T sn, cs;
            elm_sincos(in, &sn, &cs);
            acc = fma(cs, r[NDIM + 1], acc);
        }
    }
    return (double)acc;
}

#define HEPMC_DISPATCH_NDIM(ND, MACRO)                                              \
    switch (ND) {                                                                   \
        case 1: MACRO(1); break;   case 2: MACRO(2); break;   case 3: MACRO(3); break;   \
        case 4: MACRO(4); break;   case 5: MACRO(5); break;   case 6: MACRO(6); break;   \
        case 7: MACRO(7); break;   case 8: MACRO(8); break;   case 9: MACRO(9); break;   \
        case 10: MACRO(10); break; case 11: MACRO(11); break; case 12: MACRO(12); break; \
        case 13: MACRO(13); break; case 14: MACRO(14); break; case 15: MACRO(15); break; \
        case 16: MACRO(16); break;                                                  \
        default: break;                                                             \
    }
#define HEPMC_MAX_CHAIN_DIM 16

}  // namespace hepmc

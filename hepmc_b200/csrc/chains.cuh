// Shared pieces of the chain-ensemble kernels (chains.cu: SIMT kernels; elm_tc.cu: HMC with the
// ELM gradient on tensor cores).
#pragma once
#include <stdlib.h>
#include "density.cuh"
#include "elm.cuh"

namespace hepmc {

struct ChainArgs {
    const double* x0;
    int64_t n_chains, first_chain, T, thin, n_kept;
    uint64_t seed;
    uint32_t stream_id;
    PhiloxKey keys;          // round keys of `seed`, expanded on the host: constant-bank operands (chains_rwm_fast.cu)
    const double* z_replay;  // rwm: normals (C, T-1, nd); hmc: momenta (C, T-1, nd)
    const double* u_replay;  // (C, T-1)
    double* chain;           // (C, n_kept, nd) or NULL
    double* last;            // (C, nd)
    int64_t* accepted;       // (C,)
    unsigned long long* total_accepted;
    double scale[HEPMC_MAX_CHAIN_DIM];  // rwm: sqrt(diag cov); hmc: sqrt(mass)
    int prop_kind;                      // rwm: 0 = Gaussian random walk, 1 = uniform independence on (lo, hi), 2 = UniformLocal (scale = delta, box [lo, hi])
    double prop_lo, prop_hi;
    // hmc
    double step_size;
    int steps;
    int precision;
};

// two standard normals from one Philox call (Box-Muller on two 52-bit uniforms)
__device__ __forceinline__ void normal_pair(const U4& r, double& z0, double& z1) {
    const double u1 = u52(r.x, r.y), u2 = u52(r.z, r.w);
    const double rad = sqrt(-2.0 * fm::log(u1));
    double sn, cs;
    fm::sincos2pi(u2, &sn, &cs);
    z0 = rad * cs;
    z1 = rad * sn;
}

template <int NDIM>
__device__ __forceinline__ void store_state(double* chain, int64_t c, int64_t n_kept, int64_t slot,
                                            const double (&x)[NDIM]) {
    double* o = chain + ((size_t)c * n_kept + slot) * NDIM;
#pragma unroll
    for (int d = 0; d < NDIM; ++d) o[d] = x[d];
}

__device__ __forceinline__ void finish_counts(const ChainArgs& a, bool active, int64_t c, int64_t acc) {
    if (active && a.accepted) a.accepted[c] = acc;
    const int64_t tot = warp_sum_i64(active ? acc : 0);
    if ((threadIdx.x & 31) == 0 && a.total_accepted && tot) atomicAdd(a.total_accepted, (unsigned long long)tot);
}


// COOP = true (small ensembles with a surrogate gradient): ONE CTA per chain.  Every thread carries
// the chain state redundantly (bit-identical: same draws, same arithmetic); the node sum of each
// gradient is split over the 128 threads and added up in a fixed order, so one chain no longer
// walks its 1024 nodes in a single thread.  Thread 0 writes the outputs.
constexpr int kCoopThreads = 128;

template <int NDIM>
__device__ __forceinline__ void coop_sum(double (&g)[NDIM], double* red) {
#pragma unroll
    for (int k = 0; k < NDIM; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g[k] += __shfl_xor_sync(0xffffffffu, g[k], o);
    }
    const int warp = threadIdx.x >> 5;
    __syncthreads();   // previous sum has been read
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < NDIM; ++k) red[warp * NDIM + k] = g[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NDIM; ++k) {
        double r = red[k];
#pragma unroll
        for (int w = 1; w < kCoopThreads / 32; ++w) r += red[w * NDIM + k];
        g[k] = r;
    }
}

// HamiltonLeapfrog.__call__, simulation.py:34-42, for one state held in registers: un-merged half steps,
// steps + 1 gradient evaluations; kin_gradient = Gaussian(mean, diag mass).pot_gradient = (p - mean)/mass
// (gaussian.py:45-47).  Shared by hmc_kernel, mc3_kernel's local HMC update and leapfrog_kernel
// (hepmc_leapfrog), so the three run the same arithmetic.
// `grad` on entry: the gradient at q (the caller has it from the previous trajectory); on exit: the gradient at the
// final q.  steps evaluations instead of steps + 1.
template <int NDIM, class GradFn>
__device__ __forceinline__ void leapfrog_steps_carry(GradFn& gradient, const Dens& pdist, double (&q)[NDIM], double (&p)[NDIM],
                                                     double (&grad)[NDIM], double step_size, int steps) {
    const double half_eps = step_size / 2;
    for (int s = 0; s < steps; ++s) {
#pragma unroll
        for (int d = 0; d < NDIM; ++d) p[d] -= half_eps * grad[d];
#pragma unroll
        for (int d = 0; d < NDIM; ++d) q[d] += step_size * (pdist.c[d] * (p[d] - pdist.a[d]));
        gradient(q, grad);
#pragma unroll
        for (int d = 0; d < NDIM; ++d) p[d] -= half_eps * grad[d];
    }
}

template <int NDIM, class GradFn>
__device__ __forceinline__ void leapfrog_steps(GradFn& gradient, const Dens& pdist, double (&q)[NDIM], double (&p)[NDIM],
                                               double step_size, int steps) {
    double grad[NDIM];
    gradient(q, grad);
    leapfrog_steps_carry<NDIM>(gradient, pdist, q, p, grad, step_size, steps);
}

// Ensembles up to this many chains per call run the cooperative HMC kernel.  Measured crossover of
// the total throughput (scripts/coop_perf.py, camel-8D, 20 leapfrog): ~12 000 chains at 1024 nodes,
// ~2000-3700 at 100 nodes; below it the cooperative kernel is up to 48x faster per chain
// (one chain, 1024 nodes, float64: 2.3e4 instead of 4.8e2 steps/s).  Development override:
// HEPMC_HMC_COOP_MAX.
// `coop` argument of the C ABI: 1 / 0 force the cooperative / the thread-per-chain kernel, -1 decides from
// the chain count OF THIS CALL.  The two kernels add the node terms of the surrogate gradient in
// different orders (last-bit differences that an accept/reject eventually amplifies), so a caller that
// shards one ensemble over several calls or GPUs must pass 0 or 1 -- the Python mirror decides from
// the GLOBAL ensemble size (core/markov/ensemble.py coop_flag).
inline bool use_coop(int coop, bool has_elm, int64_t n_chains, int nodes);

inline int64_t coop_max_chains(int nodes) {
    static int64_t forced = -2;
    if (forced == -2) {
        const char* e = getenv("HEPMC_HMC_COOP_MAX");
        forced = e ? atoll(e) : -1;
    }
    if (forced >= 0) return forced;
    const int64_t v = 8 * (int64_t)nodes;
    return v < 1024 ? 1024 : (v > 8192 ? 8192 : v);
}

inline bool use_coop(int coop, bool has_elm, int64_t n_chains, int nodes) {
    if (!has_elm) return false;
    if (coop >= 0) return coop != 0;
    return n_chains <= coop_max_chains(nodes);
}

// throughput instance of the random-walk Metropolis kernel (chains_rwm_fast.cu); -100 = no instance
int launch_rwm_fast(const Dens& dens, const ChainArgs& a, unsigned blocks, cudaStream_t st);

// tensor-core launchers (elm_tc.cu)
int launch_hmc_tc(const Dens& dens, const Dens& pdist, const hepmc_elm_t& elm, const ChainArgs& a, bool fma_g2,
                  cudaStream_t st);
int launch_elm_grad_tc(const hepmc_elm_t& elm, int ndim, const double* xs, int64_t n, double* out, bool fma_g2,
                       cudaStream_t st);

}  // namespace hepmc

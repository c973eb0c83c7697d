// Ensembles of independent Markov chains, one chain per thread, whole trajectory in-kernel:
//   * random-walk Metropolis with a diagonal Gaussian proposal (rwm_kernel)
//   * Hamiltonian Monte Carlo, leapfrog with un-merged half steps, exact or ELM-surrogate
//     gradient, Metropolis test on  pdf(q) * N(p)  (hmc_kernel)
// State (x, pdf, counters) lives in registers; randomness is Philox4x32-10 keyed by the
// GLOBAL chain index and the step, so any split of the chain range over GPUs reproduces
// the same chains.  Acceptance counts: per chain + warp-shuffle reduced grid total.
#include <stdlib.h>
#include "chains.cuh"
#include "tabmath.cuh"

namespace hepmc {

// ------------------------------------------------------------------ random-walk Metropolis
// base.py:56-97, metropolis.py:76-95,122-124, proposals/gaussian.py:25-27
template <int NDIM>
__global__ void __launch_bounds__(128) rwm_kernel(const __grid_constant__ Dens dens,
                                                  const __grid_constant__ ChainArgs a) {
    // Box-Muller and the target's exponential through the table-driven functions of tabmath.cuh (built once per
    // CTA, used T times per thread): 412 -> ~350 instructions per step
    __shared__ tm::Tables tabs;
    tm::build(tabs);
    struct TabExp {
        const tm::Tables& t;
        __device__ __forceinline__ double operator()(double v) const { return tm::exp(v, t); }
    };
    const TabExp texp{tabs};
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = c < a.n_chains;
    const PhiloxKey& key = a.keys;
    constexpr int NCALL = (NDIM + 1) / 2;  // Philox calls for the proposal normals
    constexpr uint32_t CPS = NCALL + 1;    // + one for the accept uniform
    int64_t acc = 0;
    if (active) {
        const uint64_t g = (uint64_t)(a.first_chain + c);
        double x[NDIM];
#pragma unroll
        for (int d = 0; d < NDIM; ++d) x[d] = a.x0[c * NDIM + d];
        double pdf = density_pdf_n<NDIM>(dens, x, texp);  // init_state, metropolis.py:68-74
        if (a.chain) store_state<NDIM>(a.chain, c, a.n_kept, 0, x);
        for (int64_t t = 1; t < a.T; ++t) {
            double cand[NDIM];
            double u_spare = 0.5;
            double z[NDIM];  // standard normals (random walk) or uniforms (independence proposal)
            if (a.z_replay) {
                const double* zt = a.z_replay + ((size_t)c * (a.T - 1) + (t - 1)) * NDIM;
#pragma unroll
                for (int d = 0; d < NDIM; ++d) z[d] = zt[d];
            } else {
#pragma unroll
                for (int k = 0; k < NCALL; ++k) {
                    const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(t - 1) * CPS + k,
                                               a.stream_id, key);
                    if (k == 0) u_spare = u24_spare(r.x, r.z);
                    double z0, z1;
                    if (a.prop_kind == 0) {
                        // Box-Muller on two 52-bit uniforms (same convention as normal_pair, table-driven functions)
                        const double u1 = u52(r.x, r.y), u2 = u52(r.z, r.w);
                        const double rad = sqrt(-2.0 * tm::log(u1, tabs));
                        double sn, cs;
                        tm::sincos2pi_post(tm::sincos2pi_pre(u2, tabs), sn, cs);
                        z0 = rad * cs;
                        z1 = rad * sn;
                    } else {
                        z0 = u52(r.x, r.y);
                        z1 = u52(r.z, r.w);
                    }
                    z[2 * k] = z0;
                    if (2 * k + 1 < NDIM) z[(2 * k + 1) % NDIM] = z1;
                }
            }
            if (a.prop_kind == 2) {
                // proposals/uniform_local.py:27-29: a box of edge delta around the state, pushed back
                // into [lo, hi]; ONE uniform per step moves every coordinate (np.random.rand() is scalar)
#pragma unroll
                for (int d = 0; d < NDIM; ++d) {
                    const double base = fmax(a.prop_lo, x[d] - a.scale[d] / 2);
                    cand[d] = fmin(base, a.prop_hi - a.scale[d]) + z[0] * a.scale[d];
                }
            } else {
#pragma unroll
                for (int d = 0; d < NDIM; ++d)
                    cand[d] = a.prop_kind == 0 ? x[d] + a.scale[d] * z[d]            // proposals/gaussian.py:25-27
                                               : a.prop_lo + (a.prop_hi - a.prop_lo) * z[d];  // uniform.py:33-34
            }
            const double cpdf = density_pdf_n<NDIM>(dens, cand, texp);
            const double ratio = cpdf / pdf;  // metropolis.py:50
            // native mode: the accept uniform comes from the spare bits of the proposal draw (common.cuh): 24 bits,
            // (k + 1/2) 2^-24.  The realised acceptance probability is the ratio rounded to a multiple of 2^-24 and
            // moves with ratio < 3e-8 are never accepted (documented in DefaultMetropolis).  A full-resolution
            // uniform for small ratios was measured: with wide proposals MOST warps hold a lane in that branch, the
            // second Philox call then costs 12 % of the kernel (7.6e10 -> 6.6e10 chain-steps/s) for a bias that is
            // below 2^-24 absolute per step.
            const double u = a.u_replay ? a.u_replay[(size_t)c * (a.T - 1) + (t - 1)] : u_spare;
            if (ratio >= 1.0 || u < ratio) {  // metropolis.py:87 (NaN -> reject)
                bool changed = false;
#pragma unroll
                for (int d = 0; d < NDIM; ++d) {
                    changed = changed || (cand[d] != x[d]);
                    x[d] = cand[d];
                }
                pdf = cpdf;
                acc += changed ? 1 : 0;  // base.py:72-73
            }
            if (a.chain && (t % a.thin) == 0) store_state<NDIM>(a.chain, c, a.n_kept, t / a.thin, x);
        }
        if (a.last) {
#pragma unroll
            for (int d = 0; d < NDIM; ++d) a.last[c * NDIM + d] = x[d];
        }
    }
    finish_counts(a, active, c, acc);
}

// ------------------------------------------------------------------ HMC
#ifndef HEPMC_HMC_CARRY
#define HEPMC_HMC_CARRY 1
#endif
// hmc.py:54-88, simulation.py:34-42.  pdist = Gaussian(0, diag(mass)) as a Dens block.
template <int NDIM, class T, bool COOP = false>
__global__ void __launch_bounds__(128) hmc_kernel(const __grid_constant__ Dens dens,
                                                  const __grid_constant__ Dens pdist,
                                                  const __grid_constant__ hepmc_elm_t elm,
                                                  const __grid_constant__ ChainArgs a, int tab_in_smem) {
    extern __shared__ double smem_raw[];
    __shared__ double s_red[COOP ? (kCoopThreads / 32) * NDIM : 1];
    T* tab = reinterpret_cast<T*>(smem_raw);
    const bool use_elm = elm.mode != HEPMC_GRAD_EXACT;
    if (use_elm && tab_in_smem) {
        elm_stage_tables<T>(elm, NDIM, tab);
        __syncthreads();
    }
    const int64_t c = COOP ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = c < a.n_chains;
    const bool writer = !COOP || threadIdx.x == 0;
    auto gradient = [&](const double (&q)[NDIM], double (&grad)[NDIM]) {
        if (use_elm) {
            if constexpr (COOP) {
                elm_gradient<NDIM, T>(elm.mode, elm.nodes, tab, q, grad, (int)threadIdx.x, kCoopThreads);
                coop_sum<NDIM>(grad, s_red);
            } else {
                elm_gradient<NDIM, T>(elm.mode, elm.nodes, tab, q, grad);
            }
        } else {
            density_pot_gradient_n<NDIM>(dens, q, grad);
        }
    };
    const PhiloxKey key = philox_expand(a.seed);
    constexpr int NCALL = (NDIM + 1) / 2;
    constexpr uint32_t CPS = NCALL + 1;
    int64_t acc = 0;
    if (active) {
        const uint64_t g = (uint64_t)(a.first_chain + c);
        double x[NDIM];
#pragma unroll
        for (int d = 0; d < NDIM; ++d) x[d] = a.x0[c * NDIM + d];
        double pdf = density_pdf_n<NDIM>(dens, x);
        if (a.chain && writer) store_state<NDIM>(a.chain, c, a.n_kept, 0, x);
        // The gradient at a transition's start state is the last one of the previous trajectory (accepted) or the
        // previous start gradient (rejected): carried instead of evaluated again -- steps instead of the reference's
        // steps + 1 evaluations per transition, the same values.  Up to 8 dimensions (registers).
        constexpr bool CARRY = HEPMC_HMC_CARRY && NDIM <= 8 && sizeof(T) == 8;   // float32 instance: +21 registers, -6 % measured
        double gx[CARRY ? NDIM : 1];
        if constexpr (CARRY) gradient(x, gx);
        for (int64_t t = 1; t < a.T; ++t) {
            double p0[NDIM], q[NDIM], p[NDIM];
            if (a.z_replay) {
                const double* z = a.z_replay + ((size_t)c * (a.T - 1) + (t - 1)) * NDIM;
#pragma unroll
                for (int d = 0; d < NDIM; ++d) p0[d] = z[d];
            } else {
#pragma unroll
                for (int k = 0; k < NCALL; ++k) {
                    const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(t - 1) * CPS + k,
                                               a.stream_id, key);
                    double z0, z1;
                    normal_pair(r, z0, z1);
                    p0[2 * k] = a.scale[2 * k] * z0;  // gaussian.py:49-51 with mean 0
                    if (2 * k + 1 < NDIM) p0[(2 * k + 1) % NDIM] = a.scale[(2 * k + 1) % NDIM] * z1;
                }
            }
#pragma unroll
            for (int d = 0; d < NDIM; ++d) { q[d] = x[d]; p[d] = p0[d]; }
            double grad[CARRY ? NDIM : 1];
            if constexpr (CARRY) {
#pragma unroll
                for (int d = 0; d < NDIM; ++d) grad[d] = gx[d];
                leapfrog_steps_carry<NDIM>(gradient, pdist, q, p, grad, a.step_size, a.steps);
            } else {
                leapfrog_steps<NDIM>(gradient, pdist, q, p, a.step_size, a.steps);   // simulation.py:34-42
            }
            const double cpdf = density_pdf_n<NDIM>(dens, q);              // hmc.py:80
            const double np1 = density_pdf_n<NDIM>(pdist, p);
            const double np0 = density_pdf_n<NDIM>(pdist, p0);
            double prob = cpdf * np1 / pdf / np0;                          // hmc.py:56-57
            if (isnan(prob)) prob = 0.0;                                   // hmc.py:58-59
            double u;
            if (a.u_replay) {
                u = a.u_replay[(size_t)c * (a.T - 1) + (t - 1)];
            } else {
                const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(t - 1) * CPS + NCALL,
                                           a.stream_id, key);
                u = u52(r.x, r.y);
            }
            if (prob >= 1.0 || u < prob) {
                bool changed = false;
#pragma unroll
                for (int d = 0; d < NDIM; ++d) {
                    changed = changed || (q[d] != x[d]);
                    x[d] = q[d];
                    if constexpr (CARRY) gx[d] = grad[d];
                }
                pdf = cpdf;
                acc += changed ? 1 : 0;
            }
            if (a.chain && writer && (t % a.thin) == 0) store_state<NDIM>(a.chain, c, a.n_kept, t / a.thin, x);
        }
        if (a.last && writer) {
#pragma unroll
            for (int d = 0; d < NDIM; ++d) a.last[c * NDIM + d] = x[d];
        }
    }
    finish_counts(a, active && writer, c, acc);
}

// ------------------------------------------------------------------ HamiltonLeapfrog.__call__, batched
// simulation.py:26-46 for n independent (q, p) states: thread-per-state, or ONE CTA per state (COOP) with the
// node sum of a surrogate gradient split over the threads -- the same two layouts as hmc_kernel.
struct LeapArgs {
    const double* q_in;
    const double* p_in;
    double* q_out;
    double* p_out;
    int64_t n;
    double step_size;
    int steps;
};

template <int NDIM, class T, bool COOP = false>
__global__ void __launch_bounds__(128) leapfrog_kernel(const __grid_constant__ Dens dens,
                                                       const __grid_constant__ Dens pdist,
                                                       const __grid_constant__ hepmc_elm_t elm,
                                                       const __grid_constant__ LeapArgs a) {
    extern __shared__ double smem_raw[];
    __shared__ double s_red[COOP ? (kCoopThreads / 32) * NDIM : 1];
    T* tab = reinterpret_cast<T*>(smem_raw);
    const bool use_elm = elm.mode != HEPMC_GRAD_EXACT;
    if (use_elm) {
        elm_stage_tables<T>(elm, NDIM, tab);
        __syncthreads();
    }
    const int64_t c = COOP ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n) return;        // COOP: the whole CTA leaves together
    auto gradient = [&](const double (&q)[NDIM], double (&grad)[NDIM]) {
        if (use_elm) {
            if constexpr (COOP) {
                elm_gradient<NDIM, T>(elm.mode, elm.nodes, tab, q, grad, (int)threadIdx.x, kCoopThreads);
                coop_sum<NDIM>(grad, s_red);
            } else {
                elm_gradient<NDIM, T>(elm.mode, elm.nodes, tab, q, grad);
            }
        } else {
            density_pot_gradient_n<NDIM>(dens, q, grad);
        }
    };
    double q[NDIM], p[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) { q[d] = a.q_in[c * NDIM + d]; p[d] = a.p_in[c * NDIM + d]; }
    leapfrog_steps<NDIM>(gradient, pdist, q, p, a.step_size, a.steps);
    if (!COOP || threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < NDIM; ++d) { a.q_out[c * NDIM + d] = q[d]; a.p_out[c * NDIM + d] = p[d]; }
    }
}

// momentum distribution Gaussian(0, diag(mass)) as a density block: gaussian.py:30-51,73-76
static int fill_momentum_dist(Dens& pd, int nd, const double* mass_host, const char* who) {
    pd = Dens{};
    pd.kind = HEPMC_GAUSSIAN; pd.ndim = nd;
    double logdet = 0.0;
    for (int d = 0; d < nd; ++d) {
        HEPMC_REQUIRE(mass_host[d] > 0.0, HEPMC_E_ARG, "%s: mass must be positive", who);
        pd.a[d] = 0.0; pd.b[d] = sqrt(1.0 / mass_host[d]); pd.c[d] = 1.0 / mass_host[d];
        logdet += log(mass_host[d]);
    }
    pd.s1 = nd * log(2.0 * 3.141592653589793) + logdet;
    return 0;
}

static int check_elm(hepmc_elm_t& el, const hepmc_elm_t* elm, const Dens& dens, const char* who) {
    el = hepmc_elm_t{};
    el.mode = HEPMC_GRAD_EXACT;
    if (elm != nullptr && elm->mode != HEPMC_GRAD_EXACT) {
        HEPMC_REQUIRE(elm->mode == HEPMC_GRAD_ELM_GAUSS || elm->mode == HEPMC_GRAD_ELM_TRIG, HEPMC_E_ARG,
                      "%s: bad surrogate mode %d", who, elm->mode);
        HEPMC_REQUIRE(elm->nodes >= 1 && elm->p0_dev && elm->p1_dev && elm->out_w_dev, HEPMC_E_ARG,
                      "%s: incomplete surrogate parameters", who);
        el = *elm;
    } else {
        HEPMC_REQUIRE(!(dens.kind == HEPMC_BANANA && dens.ndim != 2), HEPMC_E_UNSUPPORTED,
                      "Banana.pot_gradient exists for ndim == 2 only (banana.py:44-51)");
    }
    return 0;
}

static int fill_chain_args(ChainArgs& a, const char* who, int ndim, const double* x0_dev, const double* scale_host,
                           int scale_len, bool sqrt_scale, int64_t n_chains, int64_t first_chain, int64_t sample_size, int64_t thin,
                           uint64_t seed, uint32_t stream_id, const double* z, const double* u, double* chain_dev,
                           double* last_dev, int64_t* accepted_dev, int64_t* total_dev) {
    HEPMC_REQUIRE(ndim >= 1 && ndim <= HEPMC_MAX_CHAIN_DIM, HEPMC_E_UNSUPPORTED, "%s: ndim %d outside [1,%d]", who, ndim,
                  HEPMC_MAX_CHAIN_DIM);
    HEPMC_REQUIRE(n_chains >= 0 && first_chain >= 0 && sample_size >= 1 && thin >= 1, HEPMC_E_ARG,
                  "%s: bad n_chains/sample_size/thin", who);
    HEPMC_REQUIRE(sample_size < ((int64_t)1 << 28), HEPMC_E_UNSUPPORTED, "%s: sample_size too large for the counter layout", who);
    HEPMC_REQUIRE(x0_dev && scale_host, HEPMC_E_ARG, "%s: NULL x0/scale", who);
    HEPMC_REQUIRE((z == nullptr) == (u == nullptr), HEPMC_E_ARG, "%s: replay needs both tapes", who);
    a.x0 = x0_dev; a.n_chains = n_chains; a.first_chain = first_chain; a.T = sample_size; a.thin = thin;
    a.n_kept = (sample_size + thin - 1) / thin;
    a.seed = seed; a.stream_id = stream_id; a.z_replay = z; a.u_replay = u;
    a.keys = philox_expand(seed);
    a.chain = chain_dev; a.last = last_dev; a.accepted = accepted_dev;
    a.total_accepted = reinterpret_cast<unsigned long long*>(total_dev);
    for (int d = 0; d < HEPMC_MAX_CHAIN_DIM; ++d) a.scale[d] = 0.0;
    if (scale_len > ndim) scale_len = ndim;
    for (int d = 0; d < scale_len; ++d) a.scale[d] = sqrt_scale ? sqrt(scale_host[d]) : scale_host[d];
    return 0;
}

}  // namespace hepmc

using namespace hepmc;

extern "C" {

int hepmc_rwm_chains(const hepmc_density_t* dens, const double* x0_dev, int proposal_kind,
                     const double* prop_scale_host, int64_t n_chains, int64_t first_chain, int64_t sample_size, int64_t thin,
                     uint64_t seed, uint32_t stream_id,
                     const double* z_replay_dev, const double* u_replay_dev,
                     double* chain_dev, double* last_dev, int64_t* accepted_dev,
                     int64_t* total_accepted_dev, void* stream) {
    HEPMC_REQUIRE(dens != nullptr, HEPMC_E_ARG, "hepmc_rwm_chains: dens is NULL");
    HEPMC_REQUIRE(dens->kind >= HEPMC_CAMEL && dens->kind <= HEPMC_UNIFORM, HEPMC_E_ARG,
                  "hepmc_rwm_chains: target must be a built-in density (kind %d)", dens->kind);
    ChainArgs a{};
    int e = fill_chain_args(a, "hepmc_rwm_chains", dens->ndim, x0_dev, prop_scale_host,
                            proposal_kind == 1 ? 0 : dens->ndim, false, n_chains, first_chain,
                            sample_size, thin, seed, stream_id, z_replay_dev, u_replay_dev, chain_dev, last_dev,
                            accepted_dev, total_accepted_dev);
    if (e) return e;
    HEPMC_REQUIRE(proposal_kind >= 0 && proposal_kind <= 2, HEPMC_E_ARG, "hepmc_rwm_chains: proposal_kind must be 0, 1 or 2");
    a.prop_kind = proposal_kind;
    if (proposal_kind == 1) {
        a.prop_lo = prop_scale_host[0];
        a.prop_hi = prop_scale_host[1];
    } else if (proposal_kind == 2) {
        a.prop_lo = prop_scale_host[dens->ndim];
        a.prop_hi = prop_scale_host[dens->ndim + 1];
    }
    if (n_chains == 0) return 0;
    const int threads = 128;
    const unsigned blocks = (unsigned)((n_chains + threads - 1) / threads);
    cudaStream_t st = as_stream(stream);
    // native draws + Gaussian random walk: the kind-specialised throughput kernel (same chains).  The choice depends
    // on the call's mode only, never on the ensemble size or its sharding.
    if (proposal_kind == 0 && z_replay_dev == nullptr && u_replay_dev == nullptr) {
        if (launch_rwm_fast(*dens, a, blocks, st) != -100) {
            HEPMC_LAUNCH_CHECK();
            return 0;
        }
    }
#define RWM_CASE(ND) rwm_kernel<ND><<<blocks, threads, 0, st>>>(*dens, a)
    HEPMC_DISPATCH_NDIM(dens->ndim, RWM_CASE);
#undef RWM_CASE
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_hmc_chains(const hepmc_density_t* dens, const hepmc_elm_t* elm, int precision, int coop_flag,
                     const double* x0_dev, const double* mass_host, double step_size, int steps,
                     int64_t n_chains, int64_t first_chain, int64_t sample_size, int64_t thin,
                     uint64_t seed, uint32_t stream_id,
                     const double* p_replay_dev, const double* u_replay_dev,
                     double* chain_dev, double* last_dev, int64_t* accepted_dev,
                     int64_t* total_accepted_dev, void* stream) {
    HEPMC_REQUIRE(dens != nullptr, HEPMC_E_ARG, "hepmc_hmc_chains: dens is NULL");
    HEPMC_REQUIRE(dens->kind >= HEPMC_CAMEL && dens->kind <= HEPMC_UNIFORM, HEPMC_E_ARG,
                  "hepmc_hmc_chains: target must be a built-in density (kind %d)", dens->kind);
    HEPMC_REQUIRE(steps >= 0, HEPMC_E_ARG, "hepmc_hmc_chains: steps < 0");
    HEPMC_REQUIRE(precision >= 0 && precision <= 3, HEPMC_E_ARG, "hepmc_hmc_chains: precision must be 0, 1, 2 or 3");
    const int nd = dens->ndim;
    ChainArgs a{};
    int e = fill_chain_args(a, "hepmc_hmc_chains", nd, x0_dev, mass_host, nd, true, n_chains, first_chain, sample_size, thin,
                            seed, stream_id, p_replay_dev, u_replay_dev, chain_dev, last_dev, accepted_dev,
                            total_accepted_dev);
    if (e) return e;
    a.step_size = step_size; a.steps = steps; a.precision = precision;
    hepmc_elm_t el;
    e = check_elm(el, elm, *dens, "hepmc_hmc_chains");
    if (e) return e;
    Dens pd;
    e = fill_momentum_dist(pd, nd, mass_host, "hepmc_hmc_chains");
    if (e) return e;
    if (n_chains == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (precision >= 2 && el.mode != HEPMC_GRAD_EXACT) {
        HEPMC_REQUIRE(el.mode == HEPMC_GRAD_ELM_GAUSS, HEPMC_E_UNSUPPORTED,
                      "hepmc_hmc_chains: the tensor-core path implements the Gaussian basis");
        return launch_hmc_tc(*dens, pd, el, a, precision == 3, st);
    }
    if (precision >= 2) precision = 0;  // exact gradient: nothing to put on tensor cores
    const int threads = 128;
    const unsigned blocks = (unsigned)((n_chains + threads - 1) / threads);
    int dev = 0, maxs = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxs, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t esz = precision == 0 ? 8 : 4;
    size_t smem = el.mode == HEPMC_GRAD_EXACT ? 0 : (size_t)el.nodes * elm_row(nd) * esz;
    HEPMC_REQUIRE(smem <= (size_t)maxs, HEPMC_E_UNSUPPORTED,
                  "hepmc_hmc_chains: %d surrogate nodes do not fit shared memory (%zu > %d bytes)", el.nodes, smem, maxs);
    // small ensembles with a surrogate gradient: one CTA per chain (see hmc_kernel<.., COOP>); the caller
    // chooses (coop = 0 / 1) or leaves it to the chain count of this call (coop = -1), see use_coop
    const bool coop = use_coop(coop_flag, el.mode != HEPMC_GRAD_EXACT, n_chains, el.nodes);
    const unsigned cblocks = (unsigned)n_chains;
#define HMC_LAUNCH(ND, TT, CO, NB)                                                                         \
    do {                                                                                                   \
        HEPMC_CUDA(cudaFuncSetAttribute(hmc_kernel<ND, TT, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        hmc_kernel<ND, TT, CO><<<NB, threads, smem, st>>>(*dens, pd, el, a, 1);                            \
    } while (0)
#define HMC_CASE(ND)                                                                                       \
    do {                                                                                                   \
        if (precision == 0) {                                                                              \
            if (coop) HMC_LAUNCH(ND, double, true, cblocks); else HMC_LAUNCH(ND, double, false, blocks);   \
        } else {                                                                                           \
            if (coop) HMC_LAUNCH(ND, float, true, cblocks); else HMC_LAUNCH(ND, float, false, blocks);     \
        }                                                                                                  \
    } while (0)
    HEPMC_DISPATCH_NDIM(nd, HMC_CASE);
#undef HMC_CASE
#undef HMC_LAUNCH
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_leapfrog(const hepmc_density_t* dens, const hepmc_elm_t* elm, int precision, int coop_flag,
                   const double* mass_host, double step_size, int steps, int64_t n,
                   const double* q_dev, const double* p_dev, double* q_out_dev, double* p_out_dev, void* stream) {
    HEPMC_REQUIRE(dens != nullptr && mass_host != nullptr, HEPMC_E_ARG, "hepmc_leapfrog: NULL dens/mass");
    HEPMC_REQUIRE(dens->kind >= HEPMC_CAMEL && dens->kind <= HEPMC_UNIFORM, HEPMC_E_ARG,
                  "hepmc_leapfrog: the potential must belong to a built-in density (kind %d)", dens->kind);
    const int nd = dens->ndim;
    HEPMC_REQUIRE(nd >= 1 && nd <= HEPMC_MAX_CHAIN_DIM, HEPMC_E_UNSUPPORTED, "hepmc_leapfrog: ndim %d outside [1,%d]", nd,
                  HEPMC_MAX_CHAIN_DIM);
    HEPMC_REQUIRE(steps >= 0 && n >= 0, HEPMC_E_ARG, "hepmc_leapfrog: negative steps / n");
    HEPMC_REQUIRE(precision == 0 || precision == 1, HEPMC_E_UNSUPPORTED,
                  "hepmc_leapfrog: precision must be 0 (float64) or 1 (float32 surrogate arithmetic)");
    hepmc_elm_t el;
    int e = check_elm(el, elm, *dens, "hepmc_leapfrog");
    if (e) return e;
    Dens pd;
    e = fill_momentum_dist(pd, nd, mass_host, "hepmc_leapfrog");
    if (e) return e;
    if (n == 0) return 0;
    HEPMC_REQUIRE(q_dev && p_dev && q_out_dev && p_out_dev, HEPMC_E_ARG, "hepmc_leapfrog: NULL state buffer");
    if (el.mode == HEPMC_GRAD_EXACT) precision = 0;
    LeapArgs a{q_dev, p_dev, q_out_dev, p_out_dev, n, step_size, steps};
    int dev = 0, maxs = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxs, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t esz = precision == 0 ? 8 : 4;
    const size_t smem = el.mode == HEPMC_GRAD_EXACT ? 0 : (size_t)el.nodes * elm_row(nd) * esz;
    HEPMC_REQUIRE(smem <= (size_t)maxs, HEPMC_E_UNSUPPORTED,
                  "hepmc_leapfrog: %d surrogate nodes do not fit shared memory (%zu > %d bytes)", el.nodes, smem, maxs);
    const bool coop = use_coop(coop_flag, el.mode != HEPMC_GRAD_EXACT, n, el.nodes);
    const unsigned blocks = coop ? (unsigned)n : (unsigned)((n + 127) / 128);
    cudaStream_t st = as_stream(stream);
#define LEAP_LAUNCH(ND, TT, CO)                                                                            \
    do {                                                                                                   \
        HEPMC_CUDA(cudaFuncSetAttribute(leapfrog_kernel<ND, TT, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        leapfrog_kernel<ND, TT, CO><<<blocks, 128, smem, st>>>(*dens, pd, el, a);                          \
    } while (0)
#define LEAP_CASE(ND)                                                                                      \
    do {                                                                                                   \
        if (precision == 0) {                                                                              \
            if (coop) LEAP_LAUNCH(ND, double, true); else LEAP_LAUNCH(ND, double, false);                  \
        } else {                                                                                           \
            if (coop) LEAP_LAUNCH(ND, float, true); else LEAP_LAUNCH(ND, float, false);                    \
        }                                                                                                  \
    } while (0)
    HEPMC_DISPATCH_NDIM(nd, LEAP_CASE);
#undef LEAP_CASE
#undef LEAP_LAUNCH
    HEPMC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

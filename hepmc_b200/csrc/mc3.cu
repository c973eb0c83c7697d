// (MC)^3 mixing chains ("next" row 3): per step and chain choose, with probability beta, an
// importance-sampling Metropolis-Hastings update whose independence proposal is the multi-channel
// mixture (core/markov/metropolis.py:37-50 with core/integration/multi_channel.py as proposal), and
// otherwise a local update: Gaussian random walk or HMC (exact / ELM surrogate gradient).
// Mirrors MixingMarkovUpdate.next_state (core/markov/base.py:139-179) for an ensemble, one chain
// per thread; the per-thread branch on the update kind costs warp divergence only inside a step.
#include "chains.cuh"

namespace hepmc {

constexpr int kMcRow = HEPMC_CHANNEL_ROW;
constexpr int kMcDim = HEPMC_MAX_CHANNEL_DIM;

struct Mc3Args {
    const double* table;  // (nch, kMcRow) channel table, see hepmc_multichannel_sample
    int nch;
    double beta;          // probability of the importance-sampling update
    int local_kind;       // 0 = Gaussian random walk (scale = sqrt(diag cov)), 1 = HMC (scale = sqrt(mass)),
                          // 2 = UniformLocal (scale = delta, box [prop_lo, prop_hi])
    const int64_t* sel_replay;  // (C, T-1): 0 = IS update, 1 = local update
    unsigned long long* local_steps;   // optional: number of local updates taken, summed over the chains
};

__device__ __forceinline__ double mixture_pdf(const double* tab, int nch, int ndim, const double* x) {
    double p = 0.0;
    for (int k = 0; k < nch; ++k) {
        const double* row = tab + k * kMcRow;
        double pk;
        if ((int)row[0] == HEPMC_GAUSSIAN) {
            double m = 0.0;
            for (int d = 0; d < ndim; ++d) {
                const double v = (x[d] - row[8 + d]) * row[8 + 2 * kMcDim + d];
                m = fma(v, v, m);
            }
            pk = exp(-0.5 * (row[2] + m));
        } else {
            bool inside = true;
            for (int d = 0; d < ndim; ++d) inside = inside && (x[d] > row[2]) && (x[d] < row[3]);
            pk = inside ? 1.0 / row[4] : 0.0;
        }
        p += row[1] * pk;  // multi_channel.py:111-114
    }
    return p;
}

template <int NDIM, class T, bool COOP = false>
__global__ void __launch_bounds__(128) mc3_kernel(const __grid_constant__ Dens dens, const __grid_constant__ Dens pdist,
                                                  const __grid_constant__ hepmc_elm_t elm,
                                                  const __grid_constant__ ChainArgs a,
                                                  const __grid_constant__ Mc3Args m) {
    extern __shared__ double smem_raw[];
    __shared__ double s_red[COOP ? (kCoopThreads / 32) * NDIM : 1];
    double* s_tab = smem_raw;  // channel table
    T* tab = reinterpret_cast<T*>(s_tab + m.nch * kMcRow);  // ELM node table
    for (int i = threadIdx.x; i < m.nch * kMcRow; i += blockDim.x) s_tab[i] = m.table[i];
    const bool use_elm = elm.mode != HEPMC_GRAD_EXACT;
    if (use_elm && m.local_kind == 1) elm_stage_tables<T>(elm, NDIM, tab);
    __syncthreads();
    // COOP: one CTA per chain, all threads carry the chain redundantly, the surrogate's node sum is
    // split over the threads (see hmc_kernel in chains.cu)
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
    const PhiloxKey& key = a.keys;   // expanded on the host: constant-bank operands
    constexpr int NCALL = (NDIM + 1) / 2;
    constexpr uint32_t CPS = NCALL + 2;  // [selector + channel pick | NCALL sample draws | accept uniform]
    int64_t acc = 0, nloc = 0;
    if (active) {
        const uint64_t g = (uint64_t)(a.first_chain + c);
        double x[NDIM];
#pragma unroll
        for (int d = 0; d < NDIM; ++d) x[d] = a.x0[c * NDIM + d];
        double pdf = density_pdf_n<NDIM>(dens, x);
        // mixture density of the CURRENT state (the Hastings ratio needs it at every importance-sampling step): kept
        // until the state changes instead of being re-evaluated -- the same value, one mixture evaluation less per step
        double q_x = 0.0;
        bool q_x_valid = false;
        if (a.chain && writer) store_state<NDIM>(a.chain, c, a.n_kept, 0, x);
        for (int64_t t = 1; t < a.T; ++t) {
            const uint32_t base = (uint32_t)(t - 1) * CPS;
            const size_t tape = (size_t)c * (a.T - 1) + (t - 1);
            const bool replay = a.z_replay != nullptr;
            double draw[NDIM];   // replay: candidate (IS) / momentum (HMC) / normals (random walk)
            bool is_step;
            double u_ch = 0.0;
            if (replay) {
                is_step = m.sel_replay[tape] == 0;
#pragma unroll
                for (int d = 0; d < NDIM; ++d) draw[d] = a.z_replay[tape * NDIM + d];
            } else {
                const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), base, a.stream_id, key);
                is_step = u52(r.x, r.y) < m.beta;  // np.random.choice(2, p=[beta, 1-beta]), base.py:170
                u_ch = u52(r.z, r.w);
            }
            double cand[NDIM], cpdf, prob, q_cand = 0.0;
            nloc += is_step ? 0 : 1;
            if (is_step) {
                // ---- importance-sampling Metropolis-Hastings: candidate ~ mixture (independent of x)
                if (replay) {
#pragma unroll
                    for (int d = 0; d < NDIM; ++d) cand[d] = draw[d];
                } else {
                    double cum = 0.0;
                    int ch = m.nch - 1;
                    for (int k = 0; k < m.nch; ++k) {
                        cum += s_tab[k * kMcRow + 1];
                        if (u_ch < cum) { ch = k; break; }
                    }
                    const double* row = s_tab + ch * kMcRow;
                    const bool gauss = (int)row[0] == HEPMC_GAUSSIAN;
#pragma unroll
                    for (int k = 0; k < NCALL; ++k) {
                        const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), base + 1u + k, a.stream_id, key);
                        double z0, z1;
                        if (gauss) {
                            normal_pair(r, z0, z1);
                            cand[2 * k] = row[8 + 2 * k] + row[8 + kMcDim + 2 * k] * z0;
                            if (2 * k + 1 < NDIM)
                                cand[(2 * k + 1) % NDIM] = row[8 + (2 * k + 1) % NDIM] + row[8 + kMcDim + (2 * k + 1) % NDIM] * z1;
                        } else {
                            cand[2 * k] = row[2] + (row[3] - row[2]) * u52(r.x, r.y);
                            if (2 * k + 1 < NDIM) cand[(2 * k + 1) % NDIM] = row[2] + (row[3] - row[2]) * u52(r.z, r.w);
                        }
                    }
                }
                cpdf = density_pdf_n<NDIM>(dens, cand);
                if (!q_x_valid) { q_x = mixture_pdf(s_tab, m.nch, NDIM, x); q_x_valid = true; }
                const double q_state = q_x;
                q_cand = mixture_pdf(s_tab, m.nch, NDIM, cand);
                prob = cpdf * q_state / pdf / q_cand;  // metropolis.py:46-47 (Hastings ratio)
            } else if (m.local_kind == 0) {
                // ---- local Gaussian random walk (proposals/gaussian.py:25-27)
                double z[NDIM];
                if (replay) {
#pragma unroll
                    for (int d = 0; d < NDIM; ++d) z[d] = draw[d];
                } else {
#pragma unroll
                    for (int k = 0; k < NCALL; ++k) {
                        const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), base + 1u + k, a.stream_id, key);
                        double z0, z1;
                        normal_pair(r, z0, z1);
                        z[2 * k] = z0;
                        if (2 * k + 1 < NDIM) z[(2 * k + 1) % NDIM] = z1;
                    }
                }
#pragma unroll
                for (int d = 0; d < NDIM; ++d) cand[d] = x[d] + a.scale[d] * z[d];
                cpdf = density_pdf_n<NDIM>(dens, cand);
                prob = cpdf / pdf;  // metropolis.py:50
            } else if (m.local_kind == 2) {
                // ---- local UniformLocal Metropolis (proposals/uniform_local.py:27-29; one uniform per step)
                double u0;
                if (replay) {
                    u0 = draw[0];
                } else {
                    const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), base + 1u, a.stream_id, key);
                    u0 = u52(r.x, r.y);
                }
#pragma unroll
                for (int d = 0; d < NDIM; ++d) {
                    const double lo = fmax(a.prop_lo, x[d] - a.scale[d] / 2);
                    cand[d] = fmin(lo, a.prop_hi - a.scale[d]) + u0 * a.scale[d];
                }
                cpdf = density_pdf_n<NDIM>(dens, cand);
                prob = cpdf / pdf;  // metropolis.py:50 (the reference treats the proposal as symmetric)
            } else {
                // ---- local HMC (hmc.py:54-88, simulation.py:34-42)
                double p0[NDIM], p[NDIM];
                if (replay) {
#pragma unroll
                    for (int d = 0; d < NDIM; ++d) p0[d] = draw[d];
                } else {
#pragma unroll
                    for (int k = 0; k < NCALL; ++k) {
                        const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), base + 1u + k, a.stream_id, key);
                        double z0, z1;
                        normal_pair(r, z0, z1);
                        p0[2 * k] = a.scale[2 * k] * z0;
                        if (2 * k + 1 < NDIM) p0[(2 * k + 1) % NDIM] = a.scale[(2 * k + 1) % NDIM] * z1;
                    }
                }
#pragma unroll
                for (int d = 0; d < NDIM; ++d) { cand[d] = x[d]; p[d] = p0[d]; }
                leapfrog_steps<NDIM>(gradient, pdist, cand, p, a.step_size, a.steps);
                cpdf = density_pdf_n<NDIM>(dens, cand);
                prob = cpdf * density_pdf_n<NDIM>(pdist, p) / pdf / density_pdf_n<NDIM>(pdist, p0);
                if (isnan(prob)) prob = 0.0;  // hmc.py:58-59
            }
            double u;
            if (a.u_replay) {
                u = a.u_replay[tape];
            } else {
                const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), base + 1u + NCALL, a.stream_id, key);
                u = u52(r.x, r.y);
            }
            if (prob >= 1.0 || u < prob) {  // metropolis.py:87
                bool changed = false;
#pragma unroll
                for (int d = 0; d < NDIM; ++d) {
                    changed = changed || (cand[d] != x[d]);
                    x[d] = cand[d];
                }
                pdf = cpdf;
                acc += changed ? 1 : 0;
                q_x = q_cand;            // importance-sampling step: known; local step: evaluated when next needed
                q_x_valid = is_step;
            }
            if (a.chain && writer && (t % a.thin) == 0) store_state<NDIM>(a.chain, c, a.n_kept, t / a.thin, x);
        }
        if (a.last && writer) {
#pragma unroll
            for (int d = 0; d < NDIM; ++d) a.last[c * NDIM + d] = x[d];
        }
    }
    finish_counts(a, active && writer, c, acc);
    if (m.local_steps) {
        const int64_t tot = warp_sum_i64((active && writer) ? nloc : 0);
        if ((threadIdx.x & 31) == 0 && tot) atomicAdd(m.local_steps, (unsigned long long)tot);
    }
}

}  // namespace hepmc

using namespace hepmc;

extern "C" int hepmc_mc3_chains(const hepmc_density_t* dens, const double* channel_table_dev, int n_channels, double beta,
                                int local_kind, const double* local_scale_host, double step_size, int steps,
                                const hepmc_elm_t* elm, int precision, int coop_flag, const double* x0_dev,
                                int64_t n_chains, int64_t first_chain, int64_t sample_size, int64_t thin,
                                uint64_t seed, uint32_t stream_id,
                                const int64_t* sel_replay_dev, const double* draw_replay_dev, const double* u_replay_dev,
                                double* chain_dev, double* last_dev, int64_t* accepted_dev, int64_t* total_accepted_dev,
                                int64_t* local_steps_dev, void* stream) {
    HEPMC_REQUIRE(dens && channel_table_dev && x0_dev, HEPMC_E_ARG, "hepmc_mc3_chains: NULL dens/table/x0");
    HEPMC_REQUIRE(dens->kind >= HEPMC_CAMEL && dens->kind <= HEPMC_UNIFORM, HEPMC_E_ARG,
                  "hepmc_mc3_chains: target must be a built-in density");
    const int nd = dens->ndim;
    HEPMC_REQUIRE(nd >= 1 && nd <= HEPMC_MAX_CHAIN_DIM, HEPMC_E_UNSUPPORTED, "hepmc_mc3_chains: ndim %d outside [1,%d]", nd,
                  HEPMC_MAX_CHAIN_DIM);
    HEPMC_REQUIRE(n_channels >= 1 && n_channels <= HEPMC_MAX_CHANNELS, HEPMC_E_UNSUPPORTED, "hepmc_mc3_chains: bad channel count");
    HEPMC_REQUIRE(beta >= 0.0 && beta <= 1.0, HEPMC_E_ARG, "hepmc_mc3_chains: beta outside [0,1]");
    HEPMC_REQUIRE(local_kind >= 0 && local_kind <= 2, HEPMC_E_ARG, "hepmc_mc3_chains: local_kind must be 0, 1 or 2");
    HEPMC_REQUIRE(beta >= 1.0 || local_scale_host, HEPMC_E_ARG, "hepmc_mc3_chains: local update parameters missing");
    HEPMC_REQUIRE(precision == 0 || precision == 1, HEPMC_E_UNSUPPORTED, "hepmc_mc3_chains: precision must be 0 or 1");
    HEPMC_REQUIRE(n_chains >= 0 && first_chain >= 0 && sample_size >= 1 && thin >= 1 && steps >= 0, HEPMC_E_ARG,
                  "hepmc_mc3_chains: bad sizes");
    const bool any_replay = sel_replay_dev || draw_replay_dev || u_replay_dev;
    HEPMC_REQUIRE(!any_replay || (sel_replay_dev && draw_replay_dev && u_replay_dev), HEPMC_E_ARG,
                  "hepmc_mc3_chains: replay needs the selector, draw and uniform tapes");
    ChainArgs a{};
    a.x0 = x0_dev; a.n_chains = n_chains; a.first_chain = first_chain; a.T = sample_size; a.thin = thin;
    a.n_kept = (sample_size + thin - 1) / thin;
    a.seed = seed; a.stream_id = stream_id; a.z_replay = draw_replay_dev; a.u_replay = u_replay_dev;
    a.keys = philox_expand(seed);
    a.chain = chain_dev; a.last = last_dev; a.accepted = accepted_dev;
    a.total_accepted = reinterpret_cast<unsigned long long*>(total_accepted_dev);
    a.step_size = step_size; a.steps = steps;
    Dens pd{};
    pd.kind = HEPMC_GAUSSIAN; pd.ndim = nd;
    double logdet = 0.0;
    for (int d = 0; d < nd; ++d) {
        const double v = local_scale_host ? local_scale_host[d] : 1.0;  // diag cov (random walk) or mass (HMC)
        HEPMC_REQUIRE(v > 0.0, HEPMC_E_ARG, "hepmc_mc3_chains: local covariance / mass must be positive");
        a.scale[d] = local_kind == 2 ? v : sqrt(v);   // UniformLocal: delta itself
        pd.a[d] = 0.0; pd.b[d] = sqrt(1.0 / v); pd.c[d] = 1.0 / v;
        logdet += log(v);
    }
    if (local_kind == 2) {
        HEPMC_REQUIRE(local_scale_host != nullptr, HEPMC_E_ARG, "hepmc_mc3_chains: UniformLocal needs delta and its range");
        a.prop_lo = local_scale_host[nd];
        a.prop_hi = local_scale_host[nd + 1];
    }
    pd.s1 = nd * log(2.0 * 3.141592653589793) + logdet;
    hepmc_elm_t el{};
    el.mode = HEPMC_GRAD_EXACT;
    if (local_kind == 1 && elm && elm->mode != HEPMC_GRAD_EXACT) {
        HEPMC_REQUIRE(elm->nodes >= 1 && elm->p0_dev && elm->p1_dev && elm->out_w_dev, HEPMC_E_ARG,
                      "hepmc_mc3_chains: incomplete surrogate parameters");
        el = *elm;
    } else if (local_kind == 1) {
        HEPMC_REQUIRE(!(dens->kind == HEPMC_BANANA && nd != 2), HEPMC_E_UNSUPPORTED,
                      "Banana.pot_gradient exists for ndim == 2 only");
    }
    if (n_chains == 0) return 0;
    Mc3Args m{channel_table_dev, n_channels, beta, local_kind, sel_replay_dev,
              reinterpret_cast<unsigned long long*>(local_steps_dev)};
    int dev = 0, maxs = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxs, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t esz = precision == 0 ? 8 : 4;
    size_t smem = (size_t)n_channels * kMcRow * 8 + (el.mode == HEPMC_GRAD_EXACT ? 0 : (size_t)el.nodes * elm_row(nd) * esz);
    HEPMC_REQUIRE(smem <= (size_t)maxs, HEPMC_E_UNSUPPORTED, "hepmc_mc3_chains: tables need %zu bytes of shared memory", smem);
    const bool coop = use_coop(coop_flag, el.mode != HEPMC_GRAD_EXACT, n_chains, el.nodes);
    const unsigned blocks = coop ? (unsigned)n_chains : (unsigned)((n_chains + 127) / 128);
    cudaStream_t st = as_stream(stream);
#define MC3_LAUNCH(ND, TT, CO)                                                                                  \
    do {                                                                                                        \
        HEPMC_CUDA(cudaFuncSetAttribute(mc3_kernel<ND, TT, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        mc3_kernel<ND, TT, CO><<<blocks, 128, smem, st>>>(*dens, pd, el, a, m);                                 \
    } while (0)
#define MC3_CASE(ND)                                                                                            \
    do {                                                                                                        \
        if (precision == 0) {                                                                                   \
            if (coop) MC3_LAUNCH(ND, double, true); else MC3_LAUNCH(ND, double, false);                         \
        } else {                                                                                                \
            if (coop) MC3_LAUNCH(ND, float, true); else MC3_LAUNCH(ND, float, false);                           \
        }                                                                                                       \
    } while (0)
    HEPMC_DISPATCH_NDIM(nd, MC3_CASE);
#undef MC3_CASE
#undef MC3_LAUNCH
    HEPMC_LAUNCH_CHECK();
    return 0;
}

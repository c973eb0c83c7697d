// Throughput instance of the random-walk Metropolis ensemble (base.py:56-97, metropolis.py:76-95,122-124,
// proposals/gaussian.py:25-27): native draws, Gaussian proposal, the target's kind and dimensionality at compile
// time.  Same chains as rwm_kernel (chains.cu) -- same Philox counters, same Box-Muller, same table-driven
// functions -- with the per-step overhead of the generic kernel removed (ncu, 282 instructions per step):
//   * Philox round keys as constant-bank operands (expanded on the host into ChainArgs::keys) instead of 20
//     uniform loads per call;
//   * no run-time switches inside the step (replay tapes, proposal kind, density kind);
//   * the Metropolis test u < cpdf / pdf as u * pdf < cpdf: no FP64 division (~28 instructions).  Same decision
//     except when u * pdf and cpdf agree to the last bit (probability ~1e-16 per step); the replay path keeps the
//     reference's division.
#include "chains.cuh"
#include "tabmath.cuh"

namespace hepmc {

template <int NDIM, int KIND>
__global__ void __launch_bounds__(128) rwm_fast_kernel(const __grid_constant__ Dens dens,
                                                       const __grid_constant__ ChainArgs a) {
    __shared__ tm::Tables tabs;
    tm::build(tabs);
    struct TabExp {
        const tm::Tables& t;
        __device__ __forceinline__ double operator()(double v) const { return tm::exp(v, t); }
    };
    const TabExp texp{tabs};
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = c < a.n_chains;
    constexpr int NCALL = (NDIM + 1) / 2;  // Philox calls for the proposal normals
    constexpr uint32_t CPS = NCALL + 1;    // counter stride per step (rwm_kernel's layout)
    int64_t acc = 0;
    if (active) {
        const uint64_t g = (uint64_t)(a.first_chain + c);
        const uint32_t g0 = (uint32_t)g, g1 = (uint32_t)(g >> 32);
        double x[NDIM];
#pragma unroll
        for (int d = 0; d < NDIM; ++d) x[d] = a.x0[c * NDIM + d];
        double pdf = pdf_kn<KIND, NDIM>(dens, x, texp);  // init_state, metropolis.py:68-74
        if (a.chain) store_state<NDIM>(a.chain, c, a.n_kept, 0, x);
        const int T = (int)a.T;
        const int thin = (int)a.thin;
        int until_store = thin;
        uint32_t ctr = 0;
        for (int t = 1; t < T; ++t) {
            double cand[NDIM];
            double u = 0.5;
#pragma unroll
            for (int k = 0; k < NCALL; ++k) {
                const U4 r = philox4x32_10(g0, g1, ctr + k, a.stream_id, a.keys);
                if (k == 0) u = u24_spare(r.x, r.z);
                const double u1 = u52(r.x, r.y), u2 = u52(r.z, r.w);
                const double rad = sqrt(-2.0 * tm::log(u1, tabs));
                double sn, cs;
                tm::sincos2pi_post(tm::sincos2pi_pre(u2, tabs), sn, cs);
                cand[2 * k] = x[2 * k] + a.scale[2 * k] * (rad * cs);               // proposals/gaussian.py:25-27
                if (2 * k + 1 < NDIM) cand[(2 * k + 1) % NDIM] = x[(2 * k + 1) % NDIM] + a.scale[(2 * k + 1) % NDIM] * (rad * sn);
            }
            ctr += CPS;
            const double cpdf = pdf_kn<KIND, NDIM>(dens, cand, texp);
            // metropolis.py:50,87: accept if ratio >= 1 or u < ratio, ratio = cpdf / pdf (0/0 and NaN reject)
            const bool accept = cpdf >= pdf ? cpdf > 0.0 : u * pdf < cpdf;
            if (accept) {
                bool changed = false;
#pragma unroll
                for (int d = 0; d < NDIM; ++d) {
                    changed = changed || (cand[d] != x[d]);
                    x[d] = cand[d];
                }
                pdf = cpdf;
                acc += changed ? 1 : 0;  // base.py:72-73
            }
            if (a.chain) {
                if (--until_store == 0) {
                    until_store = thin;
                    store_state<NDIM>(a.chain, c, a.n_kept, t / thin, x);
                }
            }
        }
        if (a.last) {
#pragma unroll
            for (int d = 0; d < NDIM; ++d) a.last[c * NDIM + d] = x[d];
        }
    }
    finish_counts(a, active, c, acc);
}

template <int KIND>
static int launch_kind(const Dens& dens, const ChainArgs& a, unsigned blocks, cudaStream_t st) {
#define RWMF_CASE(ND) rwm_fast_kernel<ND, KIND><<<blocks, 128, 0, st>>>(dens, a)
    HEPMC_DISPATCH_NDIM(dens.ndim, RWMF_CASE);
#undef RWMF_CASE
    return 0;
}

// native draws + Gaussian random-walk proposal only (the caller checks); -100: no instance for this kind
int launch_rwm_fast(const Dens& dens, const ChainArgs& a, unsigned blocks, cudaStream_t st) {
    switch (dens.kind) {
        case HEPMC_CAMEL: return launch_kind<HEPMC_CAMEL>(dens, a, blocks, st);
        case HEPMC_UCAMEL: return launch_kind<HEPMC_UCAMEL>(dens, a, blocks, st);
        case HEPMC_BANANA: return launch_kind<HEPMC_BANANA>(dens, a, blocks, st);
        case HEPMC_GAUSSIAN: return launch_kind<HEPMC_GAUSSIAN>(dens, a, blocks, st);
        default: return -100;
    }
}

}  // namespace hepmc

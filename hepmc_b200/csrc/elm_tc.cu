// Kernels that run the ELM surrogate gradient on tcgen05 tensor cores (see elm_tc.cuh):
//   elm_grad_tc_kernel : stand-alone eval_gradient (hepmc_elm_eval, precision = 2)
//   hmc_tc_kernel      : HMC ensembles (hepmc_hmc_chains, precision = 2): momentum draw,
//                        leapfrog and the float64 Metropolis test stay per-thread exactly as in
//                        chains.cu; every gradient evaluation is a CTA-collective GEMM pipeline.
// Block = two independent groups, each 128 owner threads (one chain = one TMEM lane) + one
// MMA-issuing warp (see elm_tc.cuh); every group is persistent over its own tiles of 128 chains;
// node tables are built once per CTA and shared.
#include "chains.cuh"
#include "elm_tc.cuh"

namespace hepmc {

using namespace tc;

struct TcSetup {
    Shared s;
    uint32_t bars, tmem_base, b1_saddr, b2_saddr;
    int n_node_tiles;
};

template <int D, bool FMA_G2>
__device__ __forceinline__ TcSetup tc_prologue(unsigned char* smem, const hepmc_elm_t& elm) {
    TcSetup u;
    u.n_node_tiles = (elm.nodes + TILE_N - 1) / TILE_N;
    u.s = carve_tc(smem, u.n_node_tiles);
    u.bars = smem_u32(u.s.bars);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
    if (warp == MMA_WARP) tmem_alloc(smem_u32(u.s.tmem_ptr), TMEM_COLS);
    if (threadIdx.x == 0) {
        for (int g = 0; g < N_GROUPS; ++g) {
            const uint32_t gb = u.bars + 8 * N_BARS * g;
            mbar_init(gb + 8 * BAR_A, N_OWNER / 32);
            mbar_init(gb + 8 * BAR_D2, N_ISS);
            for (int b = 0; b < N_SBUF; ++b) {
                mbar_init(gb + 8 * (BAR_S0 + b), 1);
                mbar_init(gb + 8 * (BAR_P0 + b), N_OWNER / 32);
                mbar_init(gb + 8 * (BAR_G0 + b), N_ISS > 1 ? N_ISS - 1 : 1);
            }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    build_tables<D, FMA_G2>(u.s, elm.p0_dev, elm.p1_dev, elm.out_w_dev, elm.nodes, u.n_node_tiles);
    fence_async_smem();   // tables were written through the generic proxy, the MMA reads them via the async proxy
    fence_before();
    __syncthreads();
    fence_after();
    u.tmem_base = *u.s.tmem_ptr;
    u.b1_saddr = smem_u32(u.s.b1);
    u.b2_saddr = smem_u32(u.s.b2);
    return u;
}

__device__ __forceinline__ void tc_epilogue(const TcSetup& u) {
    fence_before();
    __syncthreads();
    if (__shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) == MMA_WARP) tmem_dealloc(u.tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ stand-alone gradient
template <int D, bool FMA_G2>
__global__ void __launch_bounds__(N_THREADS, 1) elm_grad_tc_kernel(const __grid_constant__ hepmc_elm_t elm,
                                                                   const double* __restrict__ xs, int64_t n,
                                                                   double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_tc[];
    const TcSetup u = tc_prologue<D, FMA_G2>(smem_tc, elm);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
    const int64_t n_tiles = (n + TILE_M - 1) / TILE_M;
    // group of this warp; its barriers, TMEM columns and tiles (group g of CTA k: tiles 2k + g, stride 2 gridDim)
    const int grp = warp < MMA_WARP ? warp / (N_OWNER / 32) : (warp - MMA_WARP) / N_ISS;
    const int iss_q = warp < MMA_WARP ? 0 : (warp - MMA_WARP) % N_ISS;
    const uint32_t gbars = u.bars + 8 * N_BARS * grp;
    const uint32_t gtmem = u.tmem_base + GROUP_COLS * grp;
    const int64_t tile0 = (int64_t)blockIdx.x * N_GROUPS + grp, tile_step = (int64_t)gridDim.x * N_GROUPS;
    const int dbg = HEPMC_TC_DBG();
    if (warp < MMA_WARP) {
        ComputeState st;
        const uint32_t lane_base = gtmem + ((uint32_t)((warp & 3) * 32) << 16);
        for (int64_t tile = tile0; tile < n_tiles; tile += tile_step) {
            const int64_t j = tile * TILE_M + (threadIdx.x & (N_OWNER - 1));
            double x[D], g[D];
#pragma unroll
            for (int k = 0; k < D; ++k) x[k] = j < n ? xs[j * D + k] : 0.5;
            compute_eval<D, FMA_G2>(st, gbars, lane_base, elm.nodes, u.n_node_tiles, dbg, x, g, u.s.b2);
            if (j < n) {
#pragma unroll
                for (int k = 0; k < D; ++k) out[j * D + k] = g[k];
            }
        }
    } else {   // issuing warps: all lanes run the loop, one elected lane issues each instruction
        IssuerState st;
        for (int64_t tile = tile0; tile < n_tiles; tile += tile_step) {
            if constexpr (FMA_G2) {
                if (iss_q == 0) issuer_eval_g1(st, gbars, gtmem, u.b1_saddr, elm.nodes, u.n_node_tiles, dbg);
            } else {
                if (iss_q == 0) issuer_eval(st, gbars, gtmem, u.b1_saddr, u.b2_saddr, elm.nodes, u.n_node_tiles, dbg);
                else aux_issuer_eval(st, gbars, gtmem, u.b2_saddr, elm.nodes, u.n_node_tiles, iss_q, dbg);
            }
        }
    }
    tc_epilogue(u);
}

// ------------------------------------------------------------------ HMC with the tensor-core gradient
// hmc.py:54-88, simulation.py:34-42 -- identical control flow to hmc_kernel in chains.cu
template <int NDIM, bool FMA_G2>
__global__ void __launch_bounds__(N_THREADS, 1) hmc_tc_kernel(const __grid_constant__ Dens dens,
                                                              const __grid_constant__ Dens pdist,
                                                              const __grid_constant__ hepmc_elm_t elm,
                                                              const __grid_constant__ ChainArgs a) {
    extern __shared__ __align__(128) unsigned char smem_tc[];
    const TcSetup u = tc_prologue<NDIM, FMA_G2>(smem_tc, elm);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
    const int64_t n_tiles = (a.n_chains + TILE_M - 1) / TILE_M;
    // gradient evaluations per tile of chains: one at the start state, then `steps` per transition -- the gradient at a
    // transition's start state is the last one of the previous trajectory (accepted) or the previous start gradient
    // (rejected), so it is carried in registers instead of being evaluated again (the reference evaluates steps + 1)
    // Only in the both-GEMMs instance: the FFMA instance (tc32) has no registers to spare (measured -28 % with the
    // carried gradient spilled).
    constexpr bool CARRY = !FMA_G2;
    const int64_t evals_per_tile = CARRY ? 1 + (a.T - 1) * (int64_t)a.steps : (a.T - 1) * (int64_t)(a.steps + 1);
    const int grp = warp < MMA_WARP ? warp / (N_OWNER / 32) : (warp - MMA_WARP) / N_ISS;
    const int iss_q = warp < MMA_WARP ? 0 : (warp - MMA_WARP) % N_ISS;
    const uint32_t gbars = u.bars + 8 * N_BARS * grp;
    const uint32_t gtmem = u.tmem_base + GROUP_COLS * grp;
    int64_t tile0 = (int64_t)blockIdx.x * N_GROUPS + grp, tile_step = (int64_t)gridDim.x * N_GROUPS;
    const int dbg = HEPMC_TC_DBG();
    if ((dbg & 32) && grp == 1) tile0 = n_tiles;   // development knob: group 1 idle
    if (warp < MMA_WARP) {
        ComputeState cs;
#ifdef HEPMC_TC_TRACE
        if (threadIdx.x == 0 && blockIdx.x == 0) cs.trp = g_tc_trace;
#endif
        const uint32_t lane_base = gtmem + ((uint32_t)((warp & 3) * 32) << 16);
        const PhiloxKey key = philox_expand(a.seed);
        constexpr int NCALL = (NDIM + 1) / 2;
        constexpr uint32_t CPS = NCALL + 1;
        const double half_eps = a.step_size / 2;
        for (int64_t tile = tile0; tile < n_tiles; tile += tile_step) {
            const int64_t c = tile * TILE_M + (threadIdx.x & (N_OWNER - 1));
            const bool active = c < a.n_chains;
            const uint64_t g = (uint64_t)(a.first_chain + c);
            int64_t acc = 0;
            double x[NDIM];
#pragma unroll
            for (int d = 0; d < NDIM; ++d) x[d] = active ? a.x0[c * NDIM + d] : 0.5;
            double pdf = density_pdf_n<NDIM>(dens, x);
            if (active && a.chain) store_state<NDIM>(a.chain, c, a.n_kept, 0, x);
            double gx[CARRY ? NDIM : 1];                       // surrogate gradient at the current state x
            if constexpr (CARRY) compute_eval<NDIM, FMA_G2>(cs, gbars, lane_base, elm.nodes, u.n_node_tiles, dbg, x, gx, u.s.b2);
            for (int64_t t = 1; t < a.T; ++t) {
                double p0[NDIM], q[NDIM], p[NDIM], grad[NDIM];
                if (a.z_replay && active) {
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
                        p0[2 * k] = a.scale[2 * k] * z0;
                        if (2 * k + 1 < NDIM) p0[(2 * k + 1) % NDIM] = a.scale[(2 * k + 1) % NDIM] * z1;
                    }
                }
#pragma unroll
                for (int d = 0; d < NDIM; ++d) { q[d] = x[d]; p[d] = p0[d]; }
                if constexpr (CARRY) {
#pragma unroll
                    for (int d = 0; d < NDIM; ++d) grad[d] = gx[d];
                } else {
                    compute_eval<NDIM, FMA_G2>(cs, gbars, lane_base, elm.nodes, u.n_node_tiles, dbg, q, grad, u.s.b2);
                }
                for (int s = 0; s < a.steps; ++s) {
#pragma unroll
                    for (int d = 0; d < NDIM; ++d) p[d] -= half_eps * grad[d];
#pragma unroll
                    for (int d = 0; d < NDIM; ++d) q[d] += a.step_size * (pdist.c[d] * (p[d] - pdist.a[d]));
                    compute_eval<NDIM, FMA_G2>(cs, gbars, lane_base, elm.nodes, u.n_node_tiles, dbg, q, grad, u.s.b2);
#pragma unroll
                    for (int d = 0; d < NDIM; ++d) p[d] -= half_eps * grad[d];
                }
                const double cpdf = density_pdf_n<NDIM>(dens, q);
                const double np1 = density_pdf_n<NDIM>(pdist, p);
                const double np0 = density_pdf_n<NDIM>(pdist, p0);
                double prob = cpdf * np1 / pdf / np0;
                if (isnan(prob)) prob = 0.0;
                double uacc;
                if (a.u_replay && active) {
                    uacc = a.u_replay[(size_t)c * (a.T - 1) + (t - 1)];
                } else {
                    const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(t - 1) * CPS + NCALL,
                                               a.stream_id, key);
                    uacc = u52(r.x, r.y);
                }
                if (prob >= 1.0 || uacc < prob) {
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
                if (active && a.chain && (t % a.thin) == 0) store_state<NDIM>(a.chain, c, a.n_kept, t / a.thin, x);
            }
            if (active && a.last) {
#pragma unroll
                for (int d = 0; d < NDIM; ++d) a.last[c * NDIM + d] = x[d];
            }
            finish_counts(a, active, c, acc);
        }
    } else {   // issuing warps: all lanes run the loop, one elected lane issues each instruction
        IssuerState st;
#ifdef HEPMC_TC_TRACE
        if (grp == 0 && iss_q == 0 && blockIdx.x == 0) st.trp = g_tc_trace;
#endif
        for (int64_t tile = tile0; tile < n_tiles; tile += tile_step)
            for (int64_t e = 0; e < evals_per_tile; ++e) {
                if constexpr (FMA_G2) {
                    if (iss_q == 0) issuer_eval_g1(st, gbars, gtmem, u.b1_saddr, elm.nodes, u.n_node_tiles, dbg);
                } else {
                    if (iss_q == 0) issuer_eval(st, gbars, gtmem, u.b1_saddr, u.b2_saddr, elm.nodes, u.n_node_tiles, dbg);
                    else aux_issuer_eval(st, gbars, gtmem, u.b2_saddr, elm.nodes, u.n_node_tiles, iss_q, dbg);
                }
            }
    }
    tc_epilogue(u);
}

static int tc_common_checks(const hepmc_elm_t& elm, int ndim, size_t& smem) {
    HEPMC_REQUIRE(ndim >= 1 && ndim <= tc::MAX_D, HEPMC_E_UNSUPPORTED,
                  "tensor-core ELM path: ndim %d outside [1,%d]", ndim, tc::MAX_D);
    int dev = 0, maxs = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxs, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    smem = tc::smem_bytes(elm.nodes);
    HEPMC_REQUIRE(smem <= (size_t)maxs, HEPMC_E_UNSUPPORTED,
                  "tensor-core ELM path: %d nodes need %zu bytes of shared memory (> %d)", elm.nodes, smem, maxs);
    return 0;
}

#define HEPMC_DISPATCH_TC_D(ND, MACRO)                                                          \
    switch (ND) {                                                                               \
        case 1: MACRO(1); break; case 2: MACRO(2); break; case 3: MACRO(3); break; case 4: MACRO(4); break; \
        case 5: MACRO(5); break; case 6: MACRO(6); break; case 7: MACRO(7); break; case 8: MACRO(8); break; \
        default: break;                                                                         \
    }

#ifdef HEPMC_TC_TRACE
extern "C" int hepmc_tc_trace(long long* buf_dev) {
    return (int)cudaMemcpyToSymbol(tc::g_tc_trace, &buf_dev, sizeof(buf_dev));
}
#endif

// development builds only (scripts/tc_debug*.py); the product library has no knob: -1
extern "C" int hepmc_tc_debug(int mode) {
#ifdef HEPMC_TC_DEBUG
    return (int)cudaMemcpyToSymbol(tc::g_tc_debug, &mode, sizeof(int));
#else
    (void)mode;
    return -1;
#endif
}

int launch_elm_grad_tc(const hepmc_elm_t& elm, int ndim, const double* xs, int64_t n, double* out, bool fma_g2,
                       cudaStream_t st) {
    size_t smem = 0;
    int e = tc_common_checks(elm, ndim, smem);
    if (e) return e;
    const int64_t tiles = (n + tc::TILE_M - 1) / tc::TILE_M;
    const int64_t pairs = (tiles + tc::N_GROUPS - 1) / tc::N_GROUPS;
    const unsigned blocks = (unsigned)(pairs < sm_count() ? pairs : sm_count());
#define GRAD_TC_CASE(ND)                                                                                           \
    do {                                                                                                           \
        if (fma_g2) {                                                                                              \
            HEPMC_CUDA(cudaFuncSetAttribute(elm_grad_tc_kernel<ND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            elm_grad_tc_kernel<ND, true><<<blocks, tc::N_THREADS, smem, st>>>(elm, xs, n, out);                    \
        } else {                                                                                                   \
            HEPMC_CUDA(cudaFuncSetAttribute(elm_grad_tc_kernel<ND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            elm_grad_tc_kernel<ND, false><<<blocks, tc::N_THREADS, smem, st>>>(elm, xs, n, out);                   \
        }                                                                                                          \
    } while (0)
    HEPMC_DISPATCH_TC_D(ndim, GRAD_TC_CASE);
#undef GRAD_TC_CASE
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int launch_hmc_tc(const Dens& dens, const Dens& pdist, const hepmc_elm_t& elm, const ChainArgs& a, bool fma_g2,
                  cudaStream_t st) {
    size_t smem = 0;
    int e = tc_common_checks(elm, dens.ndim, smem);
    if (e) return e;
    const int64_t tiles = (a.n_chains + tc::TILE_M - 1) / tc::TILE_M;
    const int64_t pairs = (tiles + tc::N_GROUPS - 1) / tc::N_GROUPS;
    const unsigned blocks = (unsigned)(pairs < sm_count() ? pairs : sm_count());
#define HMC_TC_CASE(ND)                                                                                            \
    do {                                                                                                           \
        if (fma_g2) {                                                                                              \
            HEPMC_CUDA(cudaFuncSetAttribute(hmc_tc_kernel<ND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            hmc_tc_kernel<ND, true><<<blocks, tc::N_THREADS, smem, st>>>(dens, pdist, elm, a);                     \
        } else {                                                                                                   \
            HEPMC_CUDA(cudaFuncSetAttribute(hmc_tc_kernel<ND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            hmc_tc_kernel<ND, false><<<blocks, tc::N_THREADS, smem, st>>>(dens, pdist, elm, a);                    \
        }                                                                                                          \
    } while (0)
    HEPMC_DISPATCH_TC_D(dens.ndim, HMC_TC_CASE);
#undef HMC_TC_CASE
    HEPMC_LAUNCH_CHECK();
    return 0;
}

}  // namespace hepmc

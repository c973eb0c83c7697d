// RAMBO massless phase-space generator: 4n uniforms -> n four-momenta summing to (E,0,0,0).
// One thread per phase-space point; the 4n uniforms come from Philox4x32 calls held in registers
// (or from the reference's own draws in replay mode); the 4n output doubles of a point are one
// contiguous 32n-byte run; a warp transposes its 32 rows through shared memory and writes 512
// contiguous bytes per store instruction.  HBM-write roofline: 32n B per point.
//
// Round 2: the round-1 kernel ran 1473 instructions per 2->4 point (ncu: FP64 pipe 47 %, issue 69 %,
// 41 % of the HBM write roofline): 547 FP64 (library-grade log and sincos polynomials, a double
// division inside the log), 320 Philox.  Native mode now
//   * draws in the integrators' RNG modes (hepmc_rng_mode): in the 32-bit modes one Philox call per
//     particle (u0..u3 = its four words) instead of two;
//   * -log(u2 u3) and (sin, cos)(2 pi u1) by the table-driven functions of tabmath.cuh (128-entry table of
//     (1/c_j, log c_j) + degree-6 log1p, no division; 64-entry table of (cos, sin)(2 pi k/64) + degree-7/8
//     polynomials + one rotation; abs. errors < 2e-16);
// The tables are built by the CTA itself with the library functions (grid-stride CTAs: amortised over
// hundreds of points per thread).  Replay mode (the reference's own uniforms) uses the same arithmetic:
// parity with the reference is 1e-12 either way (tests/test_gpu_parity.py::test_rambo_replay).
#include "integrate_kernels.cuh"
#include "tabmath.cuh"

namespace hepmc {

struct RamboArgs {
    double e_cm;
    int64_t n, first_index;
    uint64_t seed;
    uint32_t stream_id;
    int rng_mode;
    PhiloxKey keys;
    const double* xs;  // replay: (n, 4*NP)
    double* out;       // (n, 4*NP)
};

__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// map_fourvector_rambo, rambo.py:162-173, in two stages so that the table look-ups of all particles of a
// point are in flight before the first polynomial starts (the look-up latency was 28 % of the stall samples):
// stage 1: indices, reduced arguments and the two table entries; stage 2: polynomials, rotation, q.
struct RamboPre {
    double c;            // cos theta = 2 u0 - 1
    tm::TrigPre trig;    // phi = 2 pi u1
    tm::LogPre lg;       // log(u2 u3)
};

__device__ __forceinline__ void rambo_stage1(double u0, double u1, double u2, double u3, const tm::Tables& tb, RamboPre& r) {
    r.c = fma(2.0, u0, -1.0);
    r.trig = tm::sincos2pi_pre(u1, tb);
    r.lg = tm::log_pre(u2 * u3, tb);                            // normal: >= 2^-106
}

__device__ __forceinline__ void rambo_stage2(const RamboPre& r, double q[4]) {
    const double q0 = -tm::log_post(r.lg);
    double sn, cs;
    tm::sincos2pi_post(r.trig, sn, cs);
    // 1 - c*c with the reference's two roundings (rambo.py:166): near c = +-1 the rounding of c*c IS the
    // value; a fused multiply-add would be more accurate and differ from the reference by up to 1e-11
    const double st = q0 * sqrt(1.0 - r.c * r.c);
    q[0] = q0;
    q[1] = st * cs;
    q[2] = st * sn;
    q[3] = q0 * r.c;
}

__device__ __forceinline__ void rambo_q(double u0, double u1, double u2, double u3, const tm::Tables& tb, double q[4]) {
    RamboPre r;
    rambo_stage1(u0, u1, u2, u3, tb, r);
    rambo_stage2(r, q);
}

// boost + scale of one particle (rambo.py:54-67), fused multiply-adds
__device__ __forceinline__ void rambo_boost(const double q[4], double b1, double b2, double b3, double x, double gamma,
                                            double aa, double (&o)[4]) {
    const double bdotq = fma(b3, q[3], fma(b2, q[2], b1 * q[1]));   // :54
    const double e = x * fma(gamma, q[0], bdotq);                    // :59
    const double ab = aa * bdotq;
    const double px = x * fma(ab, b1, fma(b1, q[0], q[1]));          // :66-67
    const double py = x * fma(ab, b2, fma(b2, q[0], q[2]));
    const double pz = x * fma(ab, b3, fma(b3, q[0], q[3]));
    o[0] = e; o[1] = px; o[2] = py; o[3] = pz;
}

// the four uniforms of particle p of point g: u52 mode = two Philox calls (round-1 convention), 32-bit
// modes = the four words of ONE call
template <int RNG>
__device__ __forceinline__ void rambo_draw(uint64_t g, int p, uint32_t stream_id, const PhiloxKey& key, double& u0,
                                           double& u1, double& u2, double& u3) {
    if constexpr (RNG == HEPMC_RNG_U52_R10) {
        const U4 r0 = philox4x32_r<10>((uint32_t)g, (uint32_t)(g >> 32), 2u * p, stream_id, key);
        const U4 r1 = philox4x32_r<10>((uint32_t)g, (uint32_t)(g >> 32), 2u * p + 1u, stream_id, key);
        u0 = u52(r0.x, r0.y); u1 = u52(r0.z, r0.w);
        u2 = u52(r1.x, r1.y); u3 = u52(r1.z, r1.w);
    } else {
        const U4 r = philox4x32_r<(RNG == HEPMC_RNG_U32_R7 ? 7 : 10)>((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)p,
                                                                      stream_id, key);
        u0 = u32(r.x); u1 = u32(r.y); u2 = u32(r.z); u3 = u32(r.w);
    }
}

#ifndef HEPMC_RAMBO_MINBLOCKS
#define HEPMC_RAMBO_MINBLOCKS 1
#endif
#ifndef HEPMC_RAMBO_GROUP
#define HEPMC_RAMBO_GROUP 8      // particles whose table look-ups are issued together (stage 1) before their polynomials
#endif

// Output through a shared-memory transpose: a thread owns one point = one row of 32 NP bytes, but a warp that
// writes its rows straight from registers (32 B per lane and instruction, 32 different 128-byte lines) tops out
// at 4.8 TB/s on B200, while 512 contiguous bytes per store instruction reach the copy peak (6.5-6.8 TB/s;
// scripts/store_pattern_bench.cu).  Rows go into a skewed tile (16-byte pieces, conflict-free both ways), the
// warp reads it back in address order and writes with st.global.v2.f64.
template <int NP, int RNG, bool REPLAY>
__global__ void __launch_bounds__(128, HEPMC_RAMBO_MINBLOCKS) rambo_kernel(const __grid_constant__ RamboArgs a) {
    constexpr int P = 2 * NP;                      // 16-byte pieces per point
    __shared__ tm::Tables tabs;
    __shared__ double2 tile[4][32 * P];
    tm::build(tabs, false);
    const PhiloxKey& key = a.keys;
    const double e_cm = a.e_cm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* tw = tile[warp];
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + warp * 32; i0 < a.n; i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + lane;
        if (i < a.n) {
            double q[NP][4];
            double Q0 = 0.0, Q1 = 0.0, Q2 = 0.0, Q3 = 0.0;
            const uint64_t g = (uint64_t)(a.first_index + i);
            constexpr int GRP = HEPMC_RAMBO_GROUP < NP ? HEPMC_RAMBO_GROUP : NP;
#pragma unroll
            for (int p0 = 0; p0 < NP; p0 += GRP) {
                RamboPre pre[GRP];
#pragma unroll
                for (int j = 0; j < GRP; ++j) {
                    const int p = p0 + j;
                    if (p < NP) {
                        double u0, u1, u2, u3;
                        if constexpr (REPLAY) ld256(a.xs + (i * NP + p) * 4, u0, u1, u2, u3);
                        else rambo_draw<RNG>(g, p, a.stream_id, key, u0, u1, u2, u3);
                        rambo_stage1(u0, u1, u2, u3, tabs, pre[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < GRP; ++j) {
                    const int p = p0 + j;
                    if (p < NP) {
                        rambo_stage2(pre[j], q[p]);
                        Q0 += q[p][0]; Q1 += q[p][1]; Q2 += q[p][2]; Q3 += q[p][3];  // rambo.py:46
                    }
                }
            }
            // 1/M once (rsqrt, <= 1 ulp) instead of sqrt + five divisions: rambo.py:48-52
            const double invM = rsqrt(fma(-Q3, Q3, fma(-Q2, Q2, fma(-Q1, Q1, Q0 * Q0))));
            const double b1 = -Q1 * invM, b2 = -Q2 * invM, b3 = -Q3 * invM;
            const double x = e_cm * invM;
            const double gamma = Q0 * invM;
            const double aa = 1.0 / (1.0 + gamma);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                double o[4];
                rambo_boost(q[p], b1, b2, b3, x, gamma, aa, o);
                tw[lane * P + ((2 * p + lane) % P)] = make_double2(o[0], o[1]);
                tw[lane * P + ((2 * p + 1 + lane) % P)] = make_double2(o[2], o[3]);
            }
        }
        __syncwarp();
        const int64_t pieces = (a.n - i0 < 32 ? a.n - i0 : 32) * P;     // 16-byte pieces this warp owns
        double* dst = a.out + i0 * (2 * P);
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int idx = j * 32 + lane;
            if (idx < pieces) {
                const int row = idx / P, col = idx % P;
                const double2 w = tw[row * P + ((col + row) % P)];
                asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(dst + 2 * idx), "d"(w.x), "d"(w.y) : "memory");
            }
        }
        __syncwarp();
    }
}

// any particle count: q parked in the output row (same thread re-reads what it wrote)
__global__ void __launch_bounds__(128) rambo_kernel_generic(const __grid_constant__ RamboArgs a, int np) {
    __shared__ tm::Tables tabs;
    tm::build(tabs, false);
    const PhiloxKey& key = a.keys;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double Q0 = 0.0, Q1 = 0.0, Q2 = 0.0, Q3 = 0.0;
        const uint64_t g = (uint64_t)(a.first_index + i);
        for (int p = 0; p < np; ++p) {
            double u0, u1, u2, u3, q[4];
            if (a.xs) ld256(a.xs + (i * np + p) * 4, u0, u1, u2, u3);
            else if (a.rng_mode == HEPMC_RNG_U52_R10) rambo_draw<HEPMC_RNG_U52_R10>(g, p, a.stream_id, key, u0, u1, u2, u3);
            else if (a.rng_mode == HEPMC_RNG_U32_R10) rambo_draw<HEPMC_RNG_U32_R10>(g, p, a.stream_id, key, u0, u1, u2, u3);
            else rambo_draw<HEPMC_RNG_U32_R7>(g, p, a.stream_id, key, u0, u1, u2, u3);
            rambo_q(u0, u1, u2, u3, tabs, q);
            Q0 += q[0]; Q1 += q[1]; Q2 += q[2]; Q3 += q[3];
            st256(a.out + (i * np + p) * 4, q[0], q[1], q[2], q[3]);
        }
        const double invM = rsqrt(fma(-Q3, Q3, fma(-Q2, Q2, fma(-Q1, Q1, Q0 * Q0))));
        const double b1 = -Q1 * invM, b2 = -Q2 * invM, b3 = -Q3 * invM;
        const double x = a.e_cm * invM;
        const double gamma = Q0 * invM;
        const double aa = 1.0 / (1.0 + gamma);
        for (int p = 0; p < np; ++p) {
            double* o = a.out + (i * np + p) * 4;
            const double qq[4] = {o[0], o[1], o[2], o[3]};
            double r[4];
            rambo_boost(qq, b1, b2, b3, x, gamma, aa, r);
            st256(o, r[0], r[1], r[2], r[3]);
        }
    }
}

}  // namespace hepmc

using namespace hepmc;

extern "C" {

double hepmc_rambo_pdf(int nparticles, double e_cm) {
    // rambo.py:33-36
    const double pi = 3.141592653589793;
    const double vol = pow(pi / 2.0, (double)(nparticles - 1)) * pow(e_cm, (double)(2 * nparticles - 4)) /
                       (tgamma((double)nparticles) * tgamma((double)(nparticles - 1)));
    return 1.0 / vol;
}

int hepmc_rambo_map(int nparticles, double e_cm, int64_t n, int64_t first_index, uint64_t seed,
                    uint32_t stream_id, int rng_mode, const double* xs_replay_dev, double* out_dev, void* stream) {
    HEPMC_REQUIRE(rng_mode >= 0 && rng_mode < HEPMC_RNG_MODES, HEPMC_E_ARG, "hepmc_rambo_map: unknown rng_mode %d", rng_mode);
    HEPMC_REQUIRE(nparticles >= 2 && nparticles <= 1024, HEPMC_E_UNSUPPORTED,
                  "hepmc_rambo_map: nparticles %d outside [2,1024]", nparticles);
    HEPMC_REQUIRE(n >= 0 && first_index >= 0, HEPMC_E_ARG, "hepmc_rambo_map: negative n/first_index");
    if (n == 0) return 0;
    HEPMC_REQUIRE(out_dev != nullptr, HEPMC_E_ARG, "hepmc_rambo_map: out_dev is NULL");
    HEPMC_REQUIRE(((uintptr_t)out_dev & 31) == 0 && ((uintptr_t)xs_replay_dev & 31) == 0, HEPMC_E_ARG,
                  "hepmc_rambo_map: buffers must be 32-byte aligned");
    RamboArgs a{e_cm, n, first_index, seed, stream_id, rng_mode, philox_expand(seed), xs_replay_dev, out_dev};
    const int threads = 128;
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = as_stream(stream);
#define RAMBO_GO(NP)                                                                                       \
    do {                                                                                                   \
        if (xs_replay_dev) rambo_kernel<NP, HEPMC_RNG_U52_R10, true><<<(unsigned)blocks, threads, 0, st>>>(a);          \
        else if (rng_mode == HEPMC_RNG_U52_R10) rambo_kernel<NP, HEPMC_RNG_U52_R10, false><<<(unsigned)blocks, threads, 0, st>>>(a); \
        else if (rng_mode == HEPMC_RNG_U32_R10) rambo_kernel<NP, HEPMC_RNG_U32_R10, false><<<(unsigned)blocks, threads, 0, st>>>(a); \
        else rambo_kernel<NP, HEPMC_RNG_U32_R7, false><<<(unsigned)blocks, threads, 0, st>>>(a);           \
    } while (0)
    switch (nparticles) {
        case 2: RAMBO_GO(2); break;
        case 3: RAMBO_GO(3); break;
        case 4: RAMBO_GO(4); break;
        case 5: RAMBO_GO(5); break;
        case 6: RAMBO_GO(6); break;
        case 7: RAMBO_GO(7); break;
        case 8: RAMBO_GO(8); break;
        default: rambo_kernel_generic<<<(unsigned)blocks, threads, 0, st>>>(a, nparticles); break;
    }
#undef RAMBO_GO
    HEPMC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

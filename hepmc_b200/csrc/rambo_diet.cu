// RamboOnDiet ("next" row 4): the invertible 3n-4 dimensional massless phase-space map and its
// inverse (core/phase_space/rambo.py:72-149,176-189), one thread per point; the per-point
// brentq root of the reference becomes a safeguarded Newton iteration on
//   g(y) = (k+1) y^k - k y^(k+1) - r,  y = u^2 in (0,1)   (monotone: g' = k(k+1) y^(k-1) (1-y) >= 0),
// plus the accept/reject test of AcceptRejectSampler (core/sampling.py:149-204).
#include "common.cuh"

namespace hepmc {

struct DietArgs {
    double e_cm;
    int np;
    int64_t n, first_index;
    uint64_t seed;
    uint32_t stream_id;
    const double* in;   // map: xs (n, 3np-4) or NULL (native RNG); inverse: p (n, 4np)
    double* out;        // map: p (n, 4np);       inverse: xs (n, 3np-4)
};

__device__ __forceinline__ double ipow(double x, int k) {
    double r = 1.0;
    for (int i = 0; i < k; ++i) r *= x;
    return r;
}

// root u in (0,1) of (k+1) u^(2k) - k u^(2k+2) = r   (rambo.py:96-103), as y = u^2 with g(y) = (k+1) y^k - k y^(k+1).
// k = 1 has the closed form y = 1 - sqrt(1 - r) = r / (1 + sqrt(1 - r)).  Otherwise Newton starts inside a bracket that
// both ends of the unit interval supply:  g(y) <= (k+1) y^k  gives  y >= (r/(k+1))^(1/k),  and  g(1-e) >= 1 - k(k+1)/2 e^2
// gives  y <= 1 - sqrt(2 (1-r) / (k(k+1))).  Started from y = 1/2 the iteration crawled towards small roots by a factor
// (k-1)/k per step (10-30 steps per root, the whole cost of the map kernel); from the nearer bracket end it needs 3-6.
__device__ __forceinline__ double diet_root(int k, double r) {
    if (k == 1) return sqrt(r / (1.0 + sqrt(1.0 - r)));
    const double kk = (double)k * (k + 1);
    double lo = k == 2 ? sqrt(r / 3.0) : exp(log(r / (k + 1)) / k);
    double hi = 1.0 - sqrt(2.0 * (1.0 - r) / kk);
    if (!(lo > 0.0)) lo = 0.0;
    if (!(hi < 1.0)) hi = 1.0;
    if (!(hi > lo)) { lo = 0.0; hi = 1.0; }                 // rounding at the ends: fall back to the unit interval
    lo *= 1.0 - 1e-14;                                      // the bounds hold in exact arithmetic; leave room for rounding
    hi = fmin(1.0, hi * (1.0 + 1e-14));
    double y = r < 0.5 ? lo : hi;
    if (!(y > 0.0 && y < 1.0)) y = 0.5 * (lo + hi);
    for (int it = 0; it < 100; ++it) {
        const double yk1 = ipow(y, k - 1);
        const double g = (k + 1) * yk1 * y - k * yk1 * y * y - r;
        if (g > 0.0) hi = y; else lo = y;
        const double dg = kk * yk1 * (1.0 - y);
        double yn = y - g / dg;
        if (!(yn > lo && yn < hi)) yn = 0.5 * (lo + hi);   // Newton left the bracket: bisect
        if (fabs(yn - y) <= 4.0e-16 * yn) { y = yn; break; }
        y = yn;
    }
    return sqrt(y);
}

// boost(q, ph), rambo.py:180-189 (note the plain, non-Minkowski dot product in p0)
__device__ __forceinline__ void diet_boost(const double q[4], const double ph[4], double p[4]) {
    const double rsq = sqrt(((q[0] * q[0] - q[1] * q[1]) - q[2] * q[2]) - q[3] * q[3]);
    p[0] = (((q[0] * ph[0] + q[1] * ph[1]) + q[2] * ph[2]) + q[3] * ph[3]) / rsq;
    const double c1 = (ph[0] + p[0]) / (rsq + q[0]);
    p[1] = ph[1] + c1 * q[1];
    p[2] = ph[2] + c1 * q[2];
    p[3] = ph[3] + c1 * q[3];
}

__global__ void __launch_bounds__(128) rambo_diet_map_kernel(const __grid_constant__ DietArgs a) {
    const PhiloxKey key = philox_expand(a.seed);
    const int np = a.np, nx = 3 * np - 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const double* xs = a.in ? a.in + i * nx : nullptr;
        const uint64_t g = (uint64_t)(a.first_index + i);
        auto x_at = [&](int j) -> double {
            if (xs) return xs[j];
            const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(j >> 1), a.stream_id, key);
            return (j & 1) ? u52(r.z, r.w) : u52(r.x, r.y);
        };
        double Q[4] = {a.e_cm, 0.0, 0.0, 0.0};
        double M_prev = a.e_cm, prod_u = 1.0;
        double* out = a.out + i * 4 * np;
        for (int ip = 2; ip <= np; ++ip) {
            double M_cur = 0.0;
            if (ip != np) {
                const double u = diet_root(np - ip, x_at(ip - 2));
                prod_u *= u;
                M_cur = prod_u * a.e_cm;                                     // :104
            }
            const double cos_t = 2.0 * x_at(np - 6 + 2 * ip) - 1.0;         // :106
            double sn, cs;
            fm::sincos2pi(x_at(np - 5 + 2 * ip), &sn, &cs);                // :107
            const double m2 = M_prev * M_prev;
            const double q = 4.0 * M_prev * (1.0 / (8.0 * m2) * sqrt((m2 - M_cur * M_cur) * (m2 - M_cur * M_cur)));  // :108-109,176-177
            const double sin_t = sqrt(1.0 - cos_t * cos_t);
            double ph[4] = {q, q * cs * sin_t, q * sn * sin_t, q * cos_t};  // :111-114
            double Qn[4] = {sqrt(q * q + M_cur * M_cur), -ph[1], -ph[2], -ph[3]};  // :115-116
            double pb[4], Qb[4];
            diet_boost(Q, ph, pb);                                           // :117
            diet_boost(Q, Qn, Qb);                                           // :118
            out[(ip - 2) * 4 + 0] = pb[0]; out[(ip - 2) * 4 + 1] = pb[1];
            out[(ip - 2) * 4 + 2] = pb[2]; out[(ip - 2) * 4 + 3] = pb[3];
            Q[0] = Qb[0]; Q[1] = Qb[1]; Q[2] = Qb[2]; Q[3] = Qb[3];
            M_prev = M_cur;
        }
        out[(np - 1) * 4 + 0] = Q[0]; out[(np - 1) * 4 + 1] = Q[1];         // :120
        out[(np - 1) * 4 + 2] = Q[2]; out[(np - 1) * 4 + 3] = Q[3];
    }
}

// Compile-time particle count (3..8): the same arithmetic as rambo_diet_map_kernel, line for line, with
//   * the loops over particles and the powers of the root iteration unrolled,
//   * the uniforms of a point drawn once (the generic kernel re-derives the Philox call of every coordinate),
//   * rows moved through shared memory: a warp reads its 32 input rows as one contiguous run and writes its 32 output
//     rows as contiguous 16-byte pieces (per-thread row accesses touch 32 different lines per instruction).
template <int NP>
__global__ void __launch_bounds__(128) rambo_diet_map_fast_kernel(const __grid_constant__ DietArgs a) {
    constexpr int NX = 3 * NP - 4, NO = 4 * NP, P = NO / 2, XS = NX | 1;   // odd input row stride: conflict-free rows
    __shared__ double2 tile[4][32 * P];
    const PhiloxKey key = philox_expand(a.seed);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* tw = tile[warp];
    double* tin = reinterpret_cast<double*>(tw);
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + warp * 32; i0 < a.n; i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + lane;
        const int rows = (int)(a.n - i0 < 32 ? a.n - i0 : 32);
        double x[NX];
        if (a.in) {
            const double* src = a.in + i0 * NX;
            for (int j = lane; j < rows * NX; j += 32) tin[(j / NX) * XS + (j % NX)] = src[j];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < NX; ++k) x[k] = tin[lane * XS + k];
            __syncwarp();
        } else {
            const uint64_t g = (uint64_t)(a.first_index + i);
#pragma unroll
            for (int c = 0; c < (NX + 1) / 2; ++c) {
                const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)c, a.stream_id, key);
                x[2 * c] = u52(r.x, r.y);
                if (2 * c + 1 < NX) x[(2 * c + 1) % NX] = u52(r.z, r.w);
            }
        }
        if (i < a.n) {
            double Q[4] = {a.e_cm, 0.0, 0.0, 0.0};
            double M_prev = a.e_cm, prod_u = 1.0;
#pragma unroll
            for (int ip = 2; ip <= NP; ++ip) {
                double M_cur = 0.0;
                if (ip != NP) {
                    const double u = diet_root(NP - ip, x[ip - 2]);
                    prod_u *= u;
                    M_cur = prod_u * a.e_cm;                                     // :104
                }
                const double cos_t = 2.0 * x[NP - 6 + 2 * ip] - 1.0;            // :106
                double sn, cs;
                fm::sincos2pi(x[NP - 5 + 2 * ip], &sn, &cs);                    // :107
                const double m2 = M_prev * M_prev;
                const double q = 4.0 * M_prev * (1.0 / (8.0 * m2) * sqrt((m2 - M_cur * M_cur) * (m2 - M_cur * M_cur)));  // :108-109,176-177
                const double sin_t = sqrt(1.0 - cos_t * cos_t);
                double ph[4] = {q, q * cs * sin_t, q * sn * sin_t, q * cos_t};  // :111-114
                double Qn[4] = {sqrt(q * q + M_cur * M_cur), -ph[1], -ph[2], -ph[3]};  // :115-116
                double pb[4], Qb[4];
                diet_boost(Q, ph, pb);                                           // :117
                diet_boost(Q, Qn, Qb);                                           // :118
                tw[lane * P + ((2 * (ip - 2) + lane) % P)] = make_double2(pb[0], pb[1]);
                tw[lane * P + ((2 * (ip - 2) + 1 + lane) % P)] = make_double2(pb[2], pb[3]);
                Q[0] = Qb[0]; Q[1] = Qb[1]; Q[2] = Qb[2]; Q[3] = Qb[3];
                M_prev = M_cur;
            }
            tw[lane * P + ((2 * (NP - 1) + lane) % P)] = make_double2(Q[0], Q[1]);          // :120
            tw[lane * P + ((2 * (NP - 1) + 1 + lane) % P)] = make_double2(Q[2], Q[3]);
        }
        __syncwarp();
        double* dst = a.out + i0 * NO;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int idx = j * 32 + lane;
            if (idx < rows * P) {
                const int row = idx / P, col = idx % P;
                const double2 w = tw[row * P + ((col + row) % P)];
                asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(dst + 2 * idx), "d"(w.x), "d"(w.y) : "memory");
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128) rambo_diet_inverse_kernel(const __grid_constant__ DietArgs a) {
    const int np = a.np, nx = 3 * np - 4;
    const double two_pi = 6.283185307179586;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const double* p = a.in + i * 4 * np;
        double* r = a.out + i * nx;
        double Q[4] = {p[(np - 1) * 4], p[(np - 1) * 4 + 1], p[(np - 1) * 4 + 2], p[(np - 1) * 4 + 3]};  // :136
        double P[4] = {Q[0], Q[1], Q[2], Q[3]};   // running sum p[i-2:]
        double M = 0.0;
        for (int ip = np; ip >= 2; --ip) {
            const double M_prev = M;
            const double* pi = p + (ip - 2) * 4;
            P[0] += pi[0]; P[1] += pi[1]; P[2] += pi[2]; P[3] += pi[3];                  // :140
            M = sqrt(((P[0] * P[0] - P[1] * P[1]) - P[2] * P[2]) - P[3] * P[3]);        // :141
            if (ip != np) {
                const double u = M_prev / M;
                const int k = np - ip;
                r[ip - 2] = (k + 1) * ipow(u, 2 * k) - k * ipow(u, 2 * (k + 1));           // :144-145
            }
            Q[0] += pi[0]; Q[1] += pi[1]; Q[2] += pi[2]; Q[3] += pi[3];                  // :147
            const double Qm[4] = {Q[0], -Q[1], -Q[2], -Q[3]};
            const double ph[4] = {pi[0], pi[1], pi[2], pi[3]};
            double pp[4];
            diet_boost(Qm, ph, pp);                                                       // :148
            r[np - 6 + 2 * ip] = 0.5 * (pp[3] / sqrt((pp[1] * pp[1] + pp[2] * pp[2]) + pp[3] * pp[3]) + 1.0);  // :149
            const double phi = atan2(pp[2], pp[1]);
            r[np - 5 + 2 * ip] = phi / two_pi + (phi < 0.0 ? 1.0 : 0.0);                  // :150-151
        }
    }
}

// AcceptRejectSampler.sample, sampling.py:196-201: accept[i] = u_i * c * q_i <= f_i
__global__ void __launch_bounds__(256) accept_flags_kernel(const double* __restrict__ f, const double* __restrict__ q,
                                                           double q_scalar, double c, int64_t n, int64_t first_index,
                                                           uint64_t seed, uint32_t stream_id,
                                                           unsigned char* __restrict__ flags) {
    const PhiloxKey key = philox_expand(seed);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t g = (uint64_t)(first_index + i);
        const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0u, stream_id, key);
        const double u = u52(r.x, r.y);
        flags[i] = (u * c * (q ? q[i] : q_scalar) <= f[i]) ? 1 : 0;
    }
}

}  // namespace hepmc

using namespace hepmc;

extern "C" {

int hepmc_rambo_diet_map(int nparticles, double e_cm, int64_t n, int64_t first_index, uint64_t seed, uint32_t stream_id,
                         const double* xs_replay_dev, double* out_dev, void* stream) {
    HEPMC_REQUIRE(nparticles >= 2 && nparticles <= 64, HEPMC_E_UNSUPPORTED, "hepmc_rambo_diet_map: nparticles %d outside [2,64]", nparticles);
    HEPMC_REQUIRE(n >= 0 && first_index >= 0, HEPMC_E_ARG, "hepmc_rambo_diet_map: negative n/first_index");
    if (n == 0) return 0;
    HEPMC_REQUIRE(out_dev != nullptr, HEPMC_E_ARG, "hepmc_rambo_diet_map: out_dev is NULL");
    DietArgs a{e_cm, nparticles, n, first_index, seed, stream_id, xs_replay_dev, out_dev};
    int64_t blocks = (n + 127) / 128;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const bool aligned = ((uintptr_t)out_dev & 15) == 0;   // st.global.v2.f64
    switch (aligned ? nparticles : 0) {
        case 3: rambo_diet_map_fast_kernel<3><<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(a); break;
        case 4: rambo_diet_map_fast_kernel<4><<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(a); break;
        case 5: rambo_diet_map_fast_kernel<5><<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(a); break;
        case 6: rambo_diet_map_fast_kernel<6><<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(a); break;
        case 7: rambo_diet_map_fast_kernel<7><<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(a); break;
        case 8: rambo_diet_map_fast_kernel<8><<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(a); break;
        default: rambo_diet_map_kernel<<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(a); break;
    }
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_rambo_diet_inverse(int nparticles, double e_cm, int64_t n, const double* p_dev, double* xs_out_dev, void* stream) {
    HEPMC_REQUIRE(nparticles >= 2 && nparticles <= 64 && n >= 0, HEPMC_E_ARG, "hepmc_rambo_diet_inverse: bad sizes");
    if (n == 0) return 0;
    HEPMC_REQUIRE(p_dev && xs_out_dev, HEPMC_E_ARG, "hepmc_rambo_diet_inverse: NULL buffer");
    DietArgs a{e_cm, nparticles, n, 0, 0, 0, p_dev, xs_out_dev};
    int64_t blocks = (n + 127) / 128;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    rambo_diet_inverse_kernel<<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(a);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_accept_flags(const double* f_dev, const double* q_dev, double q_scalar, double bound, int64_t n,
                       int64_t first_index, uint64_t seed, uint32_t stream_id, unsigned char* flags_dev, void* stream) {
    HEPMC_REQUIRE(n >= 0, HEPMC_E_ARG, "hepmc_accept_flags: n < 0");
    if (n == 0) return 0;
    HEPMC_REQUIRE(f_dev && flags_dev, HEPMC_E_ARG, "hepmc_accept_flags: NULL buffer");
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    accept_flags_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(f_dev, q_dev, q_scalar, bound, n, first_index, seed,
                                                                      stream_id, flags_dev);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

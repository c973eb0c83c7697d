// RAMBO massless phase-space generator: 4n uniforms -> n four-momenta summing to (E,0,0,0).
// One thread per phase-space point; the 4n uniforms come from 2n Philox4x32-10 calls held in
// registers (or from the reference's own draws in replay mode); the 4n output doubles of a
// point are one contiguous 32n-byte run written with 256-bit stores (STG.E.ENL2.256), so a
// warp writes 32 full 128-byte lines per particle pair.  HBM-write bound: 32n B per point.
#include "common.cuh"

namespace hepmc {

struct RamboArgs {
    double e_cm;
    int64_t n, first_index;
    uint64_t seed;
    uint32_t stream_id;
    const double* xs;  // replay: (n, 4*NP)
    double* out;       // (n, 4*NP)
};

__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// map_fourvector_rambo, rambo.py:162-173
__device__ __forceinline__ void rambo_q(double u0, double u1, double u2, double u3, double q[4]) {
    const double c = 2.0 * u0 - 1.0;
    double sn, cs;
    fm::sincos2pi(u1, &sn, &cs);  // phi = 2 pi u1
    const double q0 = -fm::log(u2 * u3);
    const double st = q0 * sqrt(1.0 - c * c);
    q[0] = q0;
    q[1] = st * cs;
    q[2] = st * sn;
    q[3] = q0 * c;
}

template <int NP>
__global__ void __launch_bounds__(128) rambo_kernel(const __grid_constant__ RamboArgs a) {
    const PhiloxKey key = philox_expand(a.seed);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double q[NP][4];
        double Q0 = 0.0, Q1 = 0.0, Q2 = 0.0, Q3 = 0.0;
        const uint64_t g = (uint64_t)(a.first_index + i);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            double u0, u1, u2, u3;
            if (a.xs) {
                ld256(a.xs + (i * NP + p) * 4, u0, u1, u2, u3);
            } else {
                const U4 r0 = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 2u * p, a.stream_id, key);
                const U4 r1 = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 2u * p + 1u, a.stream_id, key);
                u0 = u52(r0.x, r0.y); u1 = u52(r0.z, r0.w);
                u2 = u52(r1.x, r1.y); u3 = u52(r1.z, r1.w);
            }
            rambo_q(u0, u1, u2, u3, q[p]);
            Q0 += q[p][0]; Q1 += q[p][1]; Q2 += q[p][2]; Q3 += q[p][3];  // rambo.py:46
        }
        // 1/M once (rsqrt, <= 1 ulp) instead of sqrt + five divisions: rambo.py:48-52
        const double invM = rsqrt(((Q0 * Q0 - Q1 * Q1) - Q2 * Q2) - Q3 * Q3);
        const double b1 = -Q1 * invM, b2 = -Q2 * invM, b3 = -Q3 * invM;
        const double x = a.e_cm * invM;
        const double gamma = Q0 * invM;
        const double aa = 1.0 / (1.0 + gamma);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const double bdotq = (b1 * q[p][1] + b2 * q[p][2]) + b3 * q[p][3];  // :54
            const double e = x * (gamma * q[p][0] + bdotq);                     // :59
            const double ab = aa * bdotq;
            const double px = x * ((q[p][1] + b1 * q[p][0]) + ab * b1);         // :66-67
            const double py = x * ((q[p][2] + b2 * q[p][0]) + ab * b2);
            const double pz = x * ((q[p][3] + b3 * q[p][0]) + ab * b3);
            st256(a.out + (i * NP + p) * 4, e, px, py, pz);
        }
    }
}

// any particle count: q parked in the output row (same thread re-reads what it wrote)
__global__ void __launch_bounds__(128) rambo_kernel_generic(const __grid_constant__ RamboArgs a, int np) {
    const PhiloxKey key = philox_expand(a.seed);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double Q0 = 0.0, Q1 = 0.0, Q2 = 0.0, Q3 = 0.0;
        const uint64_t g = (uint64_t)(a.first_index + i);
        for (int p = 0; p < np; ++p) {
            double u0, u1, u2, u3, q[4];
            if (a.xs) {
                ld256(a.xs + (i * np + p) * 4, u0, u1, u2, u3);
            } else {
                const U4 r0 = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 2u * p, a.stream_id, key);
                const U4 r1 = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 2u * p + 1u, a.stream_id, key);
                u0 = u52(r0.x, r0.y); u1 = u52(r0.z, r0.w);
                u2 = u52(r1.x, r1.y); u3 = u52(r1.z, r1.w);
            }
            rambo_q(u0, u1, u2, u3, q);
            Q0 += q[0]; Q1 += q[1]; Q2 += q[2]; Q3 += q[3];
            st256(a.out + (i * np + p) * 4, q[0], q[1], q[2], q[3]);
        }
        const double invM = rsqrt(((Q0 * Q0 - Q1 * Q1) - Q2 * Q2) - Q3 * Q3);
        const double b1 = -Q1 * invM, b2 = -Q2 * invM, b3 = -Q3 * invM;
        const double x = a.e_cm * invM;
        const double gamma = Q0 * invM;
        const double aa = 1.0 / (1.0 + gamma);
        for (int p = 0; p < np; ++p) {
            double* o = a.out + (i * np + p) * 4;
            const double q0 = o[0], q1 = o[1], q2 = o[2], q3 = o[3];
            const double bdotq = (b1 * q1 + b2 * q2) + b3 * q3;
            const double ab = aa * bdotq;
            st256(o, x * (gamma * q0 + bdotq), x * ((q1 + b1 * q0) + ab * b1), x * ((q2 + b2 * q0) + ab * b2),
                  x * ((q3 + b3 * q0) + ab * b3));
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
                    uint32_t stream_id, const double* xs_replay_dev, double* out_dev, void* stream) {
    HEPMC_REQUIRE(nparticles >= 2 && nparticles <= 1024, HEPMC_E_UNSUPPORTED,
                  "hepmc_rambo_map: nparticles %d outside [2,1024]", nparticles);
    HEPMC_REQUIRE(n >= 0 && first_index >= 0, HEPMC_E_ARG, "hepmc_rambo_map: negative n/first_index");
    if (n == 0) return 0;
    HEPMC_REQUIRE(out_dev != nullptr, HEPMC_E_ARG, "hepmc_rambo_map: out_dev is NULL");
    HEPMC_REQUIRE(((uintptr_t)out_dev & 31) == 0 && ((uintptr_t)xs_replay_dev & 31) == 0, HEPMC_E_ARG,
                  "hepmc_rambo_map: buffers must be 32-byte aligned");
    RamboArgs a{e_cm, n, first_index, seed, stream_id, xs_replay_dev, out_dev};
    const int threads = 128;
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = as_stream(stream);
    switch (nparticles) {
        case 2: rambo_kernel<2><<<(unsigned)blocks, threads, 0, st>>>(a); break;
        case 3: rambo_kernel<3><<<(unsigned)blocks, threads, 0, st>>>(a); break;
        case 4: rambo_kernel<4><<<(unsigned)blocks, threads, 0, st>>>(a); break;
        case 5: rambo_kernel<5><<<(unsigned)blocks, threads, 0, st>>>(a); break;
        case 6: rambo_kernel<6><<<(unsigned)blocks, threads, 0, st>>>(a); break;
        case 7: rambo_kernel<7><<<(unsigned)blocks, threads, 0, st>>>(a); break;
        case 8: rambo_kernel<8><<<(unsigned)blocks, threads, 0, st>>>(a); break;
        default: rambo_kernel_generic<<<(unsigned)blocks, threads, 0, st>>>(a, nparticles); break;
    }
    HEPMC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

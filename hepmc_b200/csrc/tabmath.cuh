// Table-driven double-precision log, sincos(2 pi u) and exp for the draw transforms and densities of the
// throughput kernels (RAMBO, random-walk Metropolis): small tables in shared memory, short polynomials, no
// division.  The library-grade versions cost 40 / 45 / 29 instructions (a double division inside the log);
// these cost about 20 / 25 / 17 with absolute errors below 2e-16 on the ranges the kernels use.
//   log:    x = 2^e m, m in [1,2);  j = top 7 mantissa bits, c_j = 1 + (j + 1/2)/128;
//           log x = e ln2 + log c_j + log1p(m / c_j - 1),  |m / c_j - 1| <= 2^-8, degree-6 log1p
//   sincos: 2 pi u = 2 pi k/64 + x, |x| <= pi/64: table (cos, sin)(2 pi k/64), degree-7/8 polynomials, rotation
//   exp:    Tang's method, x = (64 k + j) ln2/64 + r: 2^k * 2^(j/64) * e^r, degree 5; |x| < 708, else the library
// The tables are built by the CTA itself with the library functions (tm::build); callers amortise that over
// hundreds of evaluations per thread.
#pragma once
#include "common.cuh"

namespace hepmc {
namespace tm {

constexpr int kLogTab = 128, kTrigTab = 64, kExpTab = 64;

struct Tables {
    double2 logt[kLogTab];    // (1/c_j, log c_j)
    double2 trig[kTrigTab];   // (cos, sin)(2 pi k / 64)
    double expt[kExpTab];     // 2^(j/64)
};

// polynomial and scaling constants in the constant bank: direct DFMA / DMUL operands (64-bit literals would be
// re-materialised with UMOV pairs at every use)
static __constant__ double kC[20] = {
    -1.0 / 6.0, 0.2, 1.0 / 3.0,                       // 0..2   log1p
    0.6931471803691238, 1.9082149292705877e-10,       // 3,4    ln2 hi, lo
    6755399441055744.0,                               // 5      1.5 * 2^52
    6.283185307179586 / kTrigTab,                     // 6      2 pi / 64
    -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0,           // 7..9   sin
    1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0,          // 10..12 cos
    (double)kTrigTab,                                 // 13
    92.332482616893657,                               // 14     64 / ln2
    0.01083042469326756, 2.9815858269852933e-12,      // 15,16  ln2/64: top 32 mantissa bits, rest
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};              // 17..19 exp

__device__ __forceinline__ void build(Tables& t, bool with_exp = true) {
    for (int j = threadIdx.x; j < kLogTab; j += blockDim.x) {
        const double c = 1.0 + (j + 0.5) * (1.0 / kLogTab);
        t.logt[j] = make_double2(1.0 / c, ::log(c));
    }
    for (int k = threadIdx.x; k < kTrigTab; k += blockDim.x) {
        double sn, cs;
        sincospi((double)k * (2.0 / kTrigTab), &sn, &cs);
        t.trig[k] = make_double2(cs, sn);
    }
    if (with_exp)
        for (int j = threadIdx.x; j < kExpTab; j += blockDim.x) t.expt[j] = exp2((double)j * (1.0 / kExpTab));
    __syncthreads();
}

// stage 1 of the logarithm: exponent, mantissa in [1,2) and the table entry (the load can be issued early)
struct LogPre {
    double m, ed;
    double2 t;
};
__device__ __forceinline__ LogPre log_pre(double x, const Tables& tb) {   // x normal, > 0
    const int hi = __double2hiint(x);
    LogPre r;
    r.ed = (double)((hi >> 20) - 1023);
    r.m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    r.t = tb.logt[(hi >> 13) & (kLogTab - 1)];
    return r;
}
__device__ __forceinline__ double log_post(const LogPre& r) {
    const double t = fma(r.m, r.t.x, -1.0);                     // m / c_j - 1
    double p = fma(t, kC[0], kC[1]);
    p = fma(t, p, -0.25);
    p = fma(t, p, kC[2]);
    p = fma(t, p, -0.5);
    p = fma(t * t, p, t);                                       // log1p
    return fma(r.ed, kC[3], r.t.y) + fma(r.ed, kC[4], p);
}
__device__ __forceinline__ double log(double x, const Tables& tb) { return log_post(log_pre(x, tb)); }

// stage 1 of sincos(2 pi u), u in [0,1]: reduced angle |x| <= pi/64 and the table entry
struct TrigPre {
    double x;
    double2 t;
};
__device__ __forceinline__ TrigPre sincos2pi_pre(double u, const Tables& tb) {
    const double tk = fma(u, kC[13], kC[5]);
    const int k = __double2loint(tk);                           // nearest multiple of 1/64: 0 .. 64
    const double f = fma(u, kC[13], -(tk - kC[5]));             // in [-1/2, 1/2], exact
    TrigPre r;
    r.x = f * kC[6];
    r.t = tb.trig[k & (kTrigTab - 1)];
    return r;
}
__device__ __forceinline__ void sincos2pi_post(const TrigPre& r, double& sn, double& cs) {
    const double x2 = r.x * r.x;
    double ps = fma(x2, kC[7], kC[8]);
    ps = fma(x2, ps, kC[9]);
    const double s = fma(r.x * x2, ps, r.x);
    double pc = fma(x2, kC[10], kC[11]);
    pc = fma(x2, pc, kC[12]);
    pc = fma(x2, pc, -0.5);
    const double c = fma(x2, pc, 1.0);
    sn = fma(r.t.y, c, r.t.x * s);
    cs = fma(r.t.x, c, -(r.t.y * s));
}

// |x| < 708 (false for NaN and inf): one mask and one integer compare on the high word
__device__ __forceinline__ bool exp_ok(double x) { return (__double2hiint(x) & 0x7fffffff) < 0x40862000; }

__device__ __forceinline__ double exp_fast(double x, const Tables& tb) {   // |x| < 708
    const double t = fma(x, kC[14], kC[5]);
    const int n = __double2loint(t);
    const double nd = t - kC[5];
    double r = fma(nd, -kC[15], x);                             // nd * hi is exact
    r = fma(nd, -kC[16], r);
    double p = fma(r, kC[17], kC[18]);
    p = fma(r, p, kC[19]);
    p = fma(r, p, 0.5);
    p = fma(r * r, p, r);                                       // e^r - 1
    const double tj = tb.expt[n & (kExpTab - 1)];
    const double v = fma(tj, p, tj);
    return __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
}
__device__ __forceinline__ double exp(double x, const Tables& tb) {
    return __builtin_expect(exp_ok(x), 1) ? exp_fast(x, tb) : ::exp(x);
}

}  // namespace tm
}  // namespace hepmc

// Double-precision log / sincos(2 pi u) for the draw transforms (Box-Muller, RAMBO), ~1 ulp.
// The CUDA math library materialises its 64-bit polynomial coefficients with UMOV pairs at every
// use (ncu: 14 % of the issue slots of the RAMBO kernel, 18 % of the RWM kernel); these versions
// keep the coefficients in constant memory and use the restricted argument ranges of the kernels
// (uniforms in (0,1)).  Measured: RAMBO +7 %, RWM +4 %.  A matching exp() was tried as well and
// was SLOWER than the library's in the float64 ELM loop (-23 %, ptxas emits LDC loads instead of
// folding the constant-bank operands), so exp stays with the library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hepmc {
namespace fm {

// 2/(2k+1), k = 1..10 (log: atanh series in z = s^2)
static __constant__ double kLog[10] = {
    0.6666666666666666, 0.4, 0.2857142857142857, 0.2222222222222222, 0.18181818181818182,
    0.15384615384615385, 0.13333333333333333, 0.11764705882352941, 0.10526315789473684, 0.09523809523809523};
// sin Taylor: (-1)^k / (2k+1)!, k = 1..8 ; cos Taylor: (-1)^k / (2k)!, k = 1..8
static __constant__ double kSin[8] = {
    -0.16666666666666666, 0.008333333333333333, -1.984126984126984e-04, 2.7557319223985893e-06,
    -2.505210838544172e-08, 1.6059043836821613e-10, -7.647163731819816e-13, 2.8114572543455206e-15};
static __constant__ double kCos[8] = {
    -0.5, 0.041666666666666664, -0.001388888888888889, 2.48015873015873e-05, -2.755731922398589e-07,
    2.08767569878681e-09, -1.1470745597729725e-11, 4.779477332387385e-14};

// log(x) for normal x > 0 (the kernels call it on uniforms in (0,1) and their products >= 2^-106).
// x <= 0, subnormal, inf or NaN fall back to the library.
__device__ __forceinline__ double log(double x) {
    int hi = __double2hiint(x);
    if (hi < 0x00100000 || hi >= 0x7ff00000) return ::log(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;               // mantissa in [1, 2)
    double m = __hiloint2double(hi, __double2loint(x));
    if (hi > 0x3ff6a09e) {                             // m > sqrt(2): halve (exact)
        m *= 0.5;
        e += 1;
    }
    const double f = m - 1.0;                          // exact
    const double s = f / (2.0 + f);
    const double z = s * s;
    double p = kLog[9];
#pragma unroll
    for (int k = 8; k >= 0; --k) p = fma(p, z, kLog[k]);
    // log(m) = 2s + s*z*p ;  2s = f - s*f exactly in real arithmetic -> better conditioned sum
    const double ef = (double)e;
    const double lm = fma(s, z * p, -s * f) + f;       // f - s f + s z p
    return fma(ef, 0.6931471803691238, fma(ef, 1.9082149292705877e-10, lm));
}

// (sin, cos)(2 pi u) for u in [0, 1]
__device__ __forceinline__ void sincos2pi(double u, double* sn, double* cs) {
    const double t = 4.0 * u;                          // quarter turns
    const double MAGIC = 6755399441055744.0;
    const double tq = t + MAGIC;
    const int q = __double2loint(tq);                  // nearest quarter turn, 0..4
    const double r = t - (tq - MAGIC);                 // in [-1/2, 1/2] quarter turns (exact)
    const double x = r * 1.5707963267948966;           // angle in [-pi/4, pi/4]
    const double x2 = x * x;
    double ps = kSin[7], pc = kCos[7];
#pragma unroll
    for (int k = 6; k >= 0; --k) {
        ps = fma(ps, x2, kSin[k]);
        pc = fma(pc, x2, kCos[k]);
    }
    const double s = fma(ps * x2, x, x);
    const double c = fma(pc, x2, 1.0);
    // rotate by q quarter turns
    const double a = (q & 1) ? c : s;
    const double b = (q & 1) ? s : c;
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}

}  // namespace fm
}  // namespace hepmc

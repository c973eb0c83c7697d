// Device-side analytic densities (camel, banana, gaussian, uniform): pdf and pot-gradient.
//   logpdf = -0.5 * (lognorm + sum_d ((x_d - mu_d) * s_d)^2),  s_d = sqrt(1/cov_d)
// (SciPy multivariate_normal).  For the isotropic camel modes the
// common factor s^2 is applied once to the accumulated squares (2 FP64 instructions per mode
// and dimension instead of 3); this moves the result by a few ulp -- replay parity with the
// reference stays ~1e-15 relative, far inside the 1e-12 gate.
// Everything is templated on the density kind so hot loops carry no kind dispatch; the
// hepmc_density_t block travels as a __grid_constant__ kernel parameter (constant bank).
#pragma once
#include "common.cuh"

namespace hepmc {

typedef hepmc_density_t Dens;

struct PdfAcc {
    double m0, m1;  // accumulated squares (camel a/b) or Mahalanobis sum (others)
    double t0;      // banana: x0
    bool inside;    // camel: open unit cube; uniform: open box
    __device__ __forceinline__ void init() { m0 = 0.0; m1 = 0.0; t0 = 0.0; inside = true; }
};

// BOUNDS: 0 = no support check needed (caller guarantees 0 < x < 1), 1 = only x < 1 can fail
// (VEGAS map: x = b + s*frac > 0 by construction), 2 = full check.
template <int KIND, int BOUNDS = 2>
__device__ __forceinline__ void pdf_acc_dim(const Dens& p, PdfAcc& acc, int d, double x) {
    if constexpr (KIND == HEPMC_CAMEL || KIND == HEPMC_UCAMEL) {
        if constexpr (KIND == HEPMC_CAMEL) {
            if constexpr (BOUNDS == 2) acc.inside = acc.inside && (0.0 < x) && (x < 1.0);
            if constexpr (BOUNDS == 1) acc.inside = acc.inside && (x < 1.0);
        }
        const double ta = x - p.a[d];
        const double tb = x - p.b[d];
        acc.m0 = fma(ta, ta, acc.m0);
        acc.m1 = fma(tb, tb, acc.m1);
    } else if constexpr (KIND == HEPMC_BANANA) {
        double phi = x;
        if (d == 0) acc.t0 = x;
        if (d == 1) phi = x + p.s0 * (acc.t0 * acc.t0) - 100.0 * p.s0;  // banana.py:21
        const double v = phi * p.a[d];
        acc.m0 = fma(v, v, acc.m0);
    } else if constexpr (KIND == HEPMC_GAUSSIAN) {
        const double v = (x - p.a[d]) * p.b[d];
        acc.m0 = fma(v, v, acc.m0);
    } else if constexpr (KIND == HEPMC_UNIFORM) {
        acc.inside = acc.inside && (x > p.s0) && (x < p.s1);
    }
}

// EXP: the exponential to use -- the library's (default) or a table-driven one (tabmath.cuh) in throughput kernels
struct LibExp {
    __device__ __forceinline__ double operator()(double x) const { return exp(x); }
};

template <int KIND, class EXP = LibExp>
__device__ __forceinline__ double pdf_acc_finish(const Dens& p, const PdfAcc& acc, const EXP& ex = EXP()) {
    if constexpr (KIND == HEPMC_CAMEL || KIND == HEPMC_UCAMEL) {
        if (KIND == HEPMC_CAMEL && !acc.inside) return 0.0;
        const double s2 = p.s0 * p.s0;  // (sqrt(1/cov))^2
        return (ex(-0.5 * (p.s1 + s2 * acc.m0)) + ex(-0.5 * (p.s1 + s2 * acc.m1))) / 2.0;  // camel.py:31
    } else if constexpr (KIND == HEPMC_BANANA || KIND == HEPMC_GAUSSIAN) {
        return ex(-0.5 * (p.s1 + acc.m0));
    } else if constexpr (KIND == HEPMC_UNIFORM) {
        return acc.inside ? 1.0 / p.s2 : 0.0;
    } else {
        return 0.0;
    }
}

#define HEPMC_KIND_SWITCH(kind, STMT)                                   \
    switch (kind) {                                                     \
        case HEPMC_CAMEL: { constexpr int KIND = HEPMC_CAMEL; STMT; break; }       \
        case HEPMC_UCAMEL: { constexpr int KIND = HEPMC_UCAMEL; STMT; break; }     \
        case HEPMC_BANANA: { constexpr int KIND = HEPMC_BANANA; STMT; break; }     \
        case HEPMC_GAUSSIAN: { constexpr int KIND = HEPMC_GAUSSIAN; STMT; break; } \
        case HEPMC_UNIFORM: { constexpr int KIND = HEPMC_UNIFORM; STMT; break; }   \
        default: break;                                                 \
    }

// ---- kind-templated whole-point functions (X: any indexable point) ------------------
template <int KIND, class X>
__device__ __forceinline__ double pdf_k(const Dens& p, const X& x, int ndim) {
    PdfAcc acc;
    acc.init();
    for (int d = 0; d < ndim; ++d) pdf_acc_dim<KIND>(p, acc, d, x[d]);
    return pdf_acc_finish<KIND>(p, acc);
}

template <int KIND, int NDIM, class X, class EXP = LibExp>
__device__ __forceinline__ double pdf_kn(const Dens& p, const X& x, const EXP& ex = EXP()) {
    PdfAcc acc;
    acc.init();
#pragma unroll
    for (int d = 0; d < NDIM; ++d) pdf_acc_dim<KIND>(p, acc, d, x[d]);
    return pdf_acc_finish<KIND, EXP>(p, acc, ex);
}

// runtime-kind wrappers: ONE dispatch per point, loops inside the case
template <class X>
__device__ __forceinline__ double density_pdf(const Dens& p, const X& x) {
    double r = 0.0;
    HEPMC_KIND_SWITCH(p.kind, r = (pdf_k<KIND>(p, x, p.ndim)));
    return r;
}

template <int NDIM, class X, class EXP = LibExp>
__device__ __forceinline__ double density_pdf_n(const Dens& p, const X& x, const EXP& ex = EXP()) {
    double r = 0.0;
    HEPMC_KIND_SWITCH(p.kind, r = (pdf_kn<KIND, NDIM, X, EXP>(p, x, ex)));
    return r;
}

// pot = -log pdf with NaN -> inf (density.py:48-53); Gaussian overrides with -logpdf
// (gaussian.py:40-43); Uniform has its own (uniform.py:24-28); Banana returns +log pdf
// written out in closed form (banana.py:25-35, reference sign quirk B3).
template <class X>
__device__ __forceinline__ double density_pot(const Dens& p, const X& x) {
    if (p.kind == HEPMC_GAUSSIAN) {
        PdfAcc acc;
        acc.init();
        for (int d = 0; d < p.ndim; ++d) pdf_acc_dim<HEPMC_GAUSSIAN>(p, acc, d, x[d]);
        return 0.5 * (p.s1 + acc.m0);
    }
    if (p.kind == HEPMC_UNIFORM) {
        PdfAcc acc;
        acc.init();
        for (int d = 0; d < p.ndim; ++d) pdf_acc_dim<HEPMC_UNIFORM>(p, acc, d, x[d]);
        return acc.inside ? -log(p.s2) : -0.0;
    }
    if (p.kind == HEPMC_BANANA) {
        const double x0 = x[0];
        double rest = 0.0;
        for (int d = 1; d < p.ndim; ++d) {
            double phi = x[d];
            if (d == 1) phi = phi + p.s0 * (x0 * x0) - 100.0 * p.s0;
            rest += phi * phi;
        }
        return p.s3 - 0.5 * (x0 * x0 / 100.0 + rest);  // s3 = -log(sqrt((2pi)^n * 100))
    }
    const double pdf = density_pdf(p, x);
    double pot = -log(pdf);
    if (isnan(pot)) pot = INFINITY;
    return pot;
}

struct CamelModes {
    double na, nb;
    bool inside;
};

template <class X>
__device__ __forceinline__ CamelModes camel_modes(const Dens& p, const X& x, int ndim) {
    PdfAcc acc;
    acc.init();
    for (int d = 0; d < ndim; ++d) pdf_acc_dim<HEPMC_CAMEL>(p, acc, d, x[d]);
    CamelModes m;
    m.inside = acc.inside;
    const double s2 = p.s0 * p.s0;
    m.na = exp(-0.5 * (p.s1 + s2 * acc.m0));
    m.nb = exp(-0.5 * (p.s1 + s2 * acc.m1));
    return m;
}

// pot_gradient into g[0..ndim): density.py:55-68 for camel (inf where pdf == 0),
// banana.py:44-48 (ndim == 2), gaussian.py:45-47, uniform.py:30-31.  NDIM = 0: runtime ndim.
template <int NDIM, class X, class G>
__device__ __forceinline__ void density_pot_gradient_n(const Dens& p, const X& x, G& g) {
    const int nd = NDIM ? NDIM : p.ndim;
    switch (p.kind) {
        case HEPMC_CAMEL:
        case HEPMC_UCAMEL: {
            const CamelModes m = camel_modes(p, x, nd);
            const bool inside = (p.kind == HEPMC_UCAMEL) || m.inside;
            const double pdf = inside ? (m.na + m.nb) / 2.0 : 0.0;
#pragma unroll
            for (int d = 0; d < nd; ++d) {
                if (pdf != 0.0) {
                    const double grad = ((-x[d] + p.a[d]) / p.s2 * m.na + (-x[d] + p.b[d]) / p.s2 * m.nb) / 2.0;
                    g[d] = -grad / pdf;
                } else {
                    g[d] = INFINITY;
                }
            }
            break;
        }
        case HEPMC_BANANA: {
            const double b = p.s0;
            const double x0 = x[0], x1 = x[nd > 1 ? 1 : 0];
            const double y0 = -1.0 / 100.0 * x0 - 2.0 * b * x0 * (x1 + b * (x0 * x0) - 100.0 * b);
            const double y1 = 100.0 * b - x1 - b * (x0 * x0);
            g[0] = -y0;
            if (nd > 1) g[nd > 1 ? 1 : 0] = -y1;
#pragma unroll
            for (int d = 2; d < nd; ++d) g[d] = x[d];  // extension: unit-Gaussian tail dims
            break;
        }
        case HEPMC_GAUSSIAN:
#pragma unroll
            for (int d = 0; d < nd; ++d) g[d] = p.c[d] * (x[d] - p.a[d]);
            break;
        default:
#pragma unroll
            for (int d = 0; d < nd; ++d) g[d] = 0.0;
            break;
    }
}

template <class X, class G>
__device__ __forceinline__ void density_pot_gradient(const Dens& p, const X& x, G& g) {
    density_pot_gradient_n<0>(p, x, g);
}

}  // namespace hepmc

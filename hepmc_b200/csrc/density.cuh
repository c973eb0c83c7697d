// Device-side analytic densities (camel, banana, gaussian, uniform): pdf and pot-gradient,
// written to reproduce the reference's arithmetic (SciPy multivariate_normal, SURVEY.md
// App. A.1-A.3) so replay-mode outputs agree to ~1e-15 relative:
//   logpdf = -0.5 * (lognorm + sum_d ((x_d - mu_d) * s_d)^2),  s_d = sqrt(1/cov_d).
// The hepmc_density_t block travels as a __grid_constant__ kernel parameter, so every
// parameter read is a uniform constant-bank load.
#pragma once
#include "common.cuh"

namespace hepmc {

typedef hepmc_density_t Dens;

// Streaming accumulator so kernels that produce x one dimension at a time (VEGAS, plain
// MC) never hold the point in registers.
struct PdfAcc {
    double m0, m1;  // two Mahalanobis sums (camel a/b) or one (others)
    double t0;      // banana: x0 ; uniform: unused
    bool inside;    // camel: open unit cube; uniform: open box
    __device__ __forceinline__ void init() { m0 = 0.0; m1 = 0.0; t0 = 0.0; inside = true; }
};

__device__ __forceinline__ void pdf_acc_dim(const Dens& p, PdfAcc& acc, int d, double x) {
    switch (p.kind) {
        case HEPMC_CAMEL:
            acc.inside = acc.inside && (0.0 < x) && (x < 1.0);
            // fallthrough
        case HEPMC_UCAMEL: {
            const double da = (x - p.a[d]) * p.s0;
            const double db = (x - p.b[d]) * p.s0;
            acc.m0 = fma(da, da, acc.m0);
            acc.m1 = fma(db, db, acc.m1);
            break;
        }
        case HEPMC_BANANA: {
            double phi = x;
            if (d == 0) acc.t0 = x;
            if (d == 1) phi = x + p.s0 * (acc.t0 * acc.t0) - 100.0 * p.s0;  // banana.py:21
            const double v = phi * p.a[d];
            acc.m0 = fma(v, v, acc.m0);
            break;
        }
        case HEPMC_GAUSSIAN: {
            const double v = (x - p.a[d]) * p.b[d];
            acc.m0 = fma(v, v, acc.m0);
            break;
        }
        case HEPMC_UNIFORM:
            acc.inside = acc.inside && (x > p.s0) && (x < p.s1);
            break;
        default:
            break;
    }
}

__device__ __forceinline__ double pdf_acc_finish(const Dens& p, const PdfAcc& acc) {
    switch (p.kind) {
        case HEPMC_CAMEL:
            if (!acc.inside) return 0.0;
            // fallthrough
        case HEPMC_UCAMEL:
            return (exp(-0.5 * (p.s1 + acc.m0)) + exp(-0.5 * (p.s1 + acc.m1))) / 2.0;  // camel.py:31
        case HEPMC_BANANA:
        case HEPMC_GAUSSIAN:
            return exp(-0.5 * (p.s1 + acc.m0));
        case HEPMC_UNIFORM:
            return acc.inside ? 1.0 / p.s2 : 0.0;
        default:
            return 0.0;
    }
}

// pdf of a point held in memory / registers (any indexable x)
template <class X>
__device__ __forceinline__ double density_pdf(const Dens& p, const X& x) {
    PdfAcc acc;
    acc.init();
    for (int d = 0; d < p.ndim; ++d) pdf_acc_dim(p, acc, d, x[d]);
    return pdf_acc_finish(p, acc);
}

template <int NDIM, class X>
__device__ __forceinline__ double density_pdf_n(const Dens& p, const X& x) {
    PdfAcc acc;
    acc.init();
#pragma unroll
    for (int d = 0; d < NDIM; ++d) pdf_acc_dim(p, acc, d, x[d]);
    return pdf_acc_finish(p, acc);
}

// pot = -log pdf with NaN -> inf (density.py:48-53); Gaussian overrides with -logpdf
// (gaussian.py:40-43); Uniform has its own (uniform.py:24-28); Banana returns +log pdf
// written out in closed form (banana.py:25-35, reference sign quirk B3).
template <class X>
__device__ __forceinline__ double density_pot(const Dens& p, const X& x) {
    if (p.kind == HEPMC_GAUSSIAN) {
        PdfAcc acc;
        acc.init();
        for (int d = 0; d < p.ndim; ++d) pdf_acc_dim(p, acc, d, x[d]);
        return 0.5 * (p.s1 + acc.m0);
    }
    if (p.kind == HEPMC_UNIFORM) {
        PdfAcc acc;
        acc.init();
        for (int d = 0; d < p.ndim; ++d) pdf_acc_dim(p, acc, d, x[d]);
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

// pdf_gradient (camel.py:33-45,91-100; corrected Gaussian, App. B6) component d, given
// the per-mode pdfs.  Helper for the generic -grad/pdf rule.
struct CamelModes {
    double na, nb;
    bool inside;
};

template <class X>
__device__ __forceinline__ CamelModes camel_modes(const Dens& p, const X& x) {
    PdfAcc acc;
    acc.init();
    for (int d = 0; d < p.ndim; ++d) pdf_acc_dim(p, acc, d, x[d]);
    CamelModes m;
    m.inside = acc.inside;
    m.na = exp(-0.5 * (p.s1 + acc.m0));
    m.nb = exp(-0.5 * (p.s1 + acc.m1));
    return m;
}

// pot_gradient into g[0..ndim): density.py:55-68 for camel (inf where pdf == 0),
// banana.py:44-48 (ndim == 2), gaussian.py:45-47, uniform.py:30-31.
template <class X, class G>
__device__ __forceinline__ void density_pot_gradient(const Dens& p, const X& x, G& g) {
    switch (p.kind) {
        case HEPMC_CAMEL:
        case HEPMC_UCAMEL: {
            const CamelModes m = camel_modes(p, x);
            const bool inside = (p.kind == HEPMC_UCAMEL) || m.inside;
            const double pdf = inside ? (m.na + m.nb) / 2.0 : 0.0;
            for (int d = 0; d < p.ndim; ++d) {
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
            const double x0 = x[0], x1 = x[1];
            const double y0 = -1.0 / 100.0 * x0 - 2.0 * b * x0 * (x1 + b * (x0 * x0) - 100.0 * b);
            const double y1 = 100.0 * b - x1 - b * (x0 * x0);
            g[0] = -y0;
            g[1] = -y1;
            for (int d = 2; d < p.ndim; ++d) g[d] = x[d];  // extension: unit Gaussian tail dims
            break;
        }
        case HEPMC_GAUSSIAN:
            for (int d = 0; d < p.ndim; ++d) g[d] = p.c[d] * (x[d] - p.a[d]);
            break;
        default:
            for (int d = 0; d < p.ndim; ++d) g[d] = 0.0;
            break;
    }
}

template <int NDIM, class X, class G>
__device__ __forceinline__ void density_pot_gradient_n(const Dens& p, const X& x, G& g) {
    switch (p.kind) {
        case HEPMC_CAMEL:
        case HEPMC_UCAMEL: {
            PdfAcc acc;
            acc.init();
#pragma unroll
            for (int d = 0; d < NDIM; ++d) pdf_acc_dim(p, acc, d, x[d]);
            const double na = exp(-0.5 * (p.s1 + acc.m0));
            const double nb = exp(-0.5 * (p.s1 + acc.m1));
            const bool inside = (p.kind == HEPMC_UCAMEL) || acc.inside;
            const double pdf = inside ? (na + nb) / 2.0 : 0.0;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) {
                if (pdf != 0.0) {
                    const double grad = ((-x[d] + p.a[d]) / p.s2 * na + (-x[d] + p.b[d]) / p.s2 * nb) / 2.0;
                    g[d] = -grad / pdf;
                } else {
                    g[d] = INFINITY;
                }
            }
            break;
        }
        case HEPMC_BANANA: {
            const double b = p.s0;
            const double x0 = x[0], x1 = x[NDIM > 1 ? 1 : 0];
            const double y0 = -1.0 / 100.0 * x0 - 2.0 * b * x0 * (x1 + b * (x0 * x0) - 100.0 * b);
            const double y1 = 100.0 * b - x1 - b * (x0 * x0);
            g[0] = -y0;
            if (NDIM > 1) g[NDIM > 1 ? 1 : 0] = -y1;
#pragma unroll
            for (int d = 2; d < NDIM; ++d) g[d] = x[d];
            break;
        }
        case HEPMC_GAUSSIAN:
#pragma unroll
            for (int d = 0; d < NDIM; ++d) g[d] = p.c[d] * (x[d] - p.a[d]);
            break;
        default:
#pragma unroll
            for (int d = 0; d < NDIM; ++d) g[d] = 0.0;
            break;
    }
}

}  // namespace hepmc

// Density.pdf / pot / pot_gradient / pdf_gradient for the built-in analytic densities.
// One thread per point; a point's ndim doubles are read straight from its (n, ndim) row
// (a warp sweeps 32 consecutive rows, so every fetched line is fully consumed via L1).
#include "density.cuh"

namespace hepmc {

struct RowView {
    const double* p;
    __device__ __forceinline__ double operator[](int d) const { return __ldg(p + d); }
};
struct RowOut {
    double* p;
    __device__ __forceinline__ double& operator[](int d) { return p[d]; }
};

__global__ void __launch_bounds__(256) density_eval_kernel(const __grid_constant__ Dens dens, int what,
                                                           const double* __restrict__ xs, int64_t n,
                                                           double* __restrict__ out) {
    const int nd = dens.ndim;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        RowView x{xs + i * nd};
        if (what == HEPMC_PDF) {
            out[i] = density_pdf(dens, x);
        } else if (what == HEPMC_POT) {
            out[i] = density_pot(dens, x);
        } else if (what == HEPMC_POT_GRADIENT) {
            RowOut g{out + i * nd};
            density_pot_gradient(dens, x, g);
        } else {  // HEPMC_PDF_GRADIENT
            double* g = out + i * nd;
            if (dens.kind == HEPMC_CAMEL || dens.kind == HEPMC_UCAMEL) {
                const CamelModes m = camel_modes(dens, x, nd);
                const bool inside = (dens.kind == HEPMC_UCAMEL) || m.inside;
                for (int d = 0; d < nd; ++d)
                    g[d] = inside ? ((-x[d] + dens.a[d]) / dens.s2 * m.na + (-x[d] + dens.b[d]) / dens.s2 * m.nb) / 2.0
                                  : 0.0;
            } else if (dens.kind == HEPMC_GAUSSIAN) {
                const double pdf = density_pdf(dens, x);
                for (int d = 0; d < nd; ++d) g[d] = -(dens.c[d] * (x[d] - dens.a[d])) * pdf;
            } else {
                for (int d = 0; d < nd; ++d) g[d] = 0.0;
            }
        }
    }
}

}  // namespace hepmc

extern "C" int hepmc_density_eval(const hepmc_density_t* dens, int what, const double* xs_dev,
                                  int64_t n, double* out_dev, void* stream) {
    using namespace hepmc;
    HEPMC_REQUIRE(dens != nullptr, HEPMC_E_ARG, "hepmc_density_eval: dens is NULL");
    HEPMC_REQUIRE(dens->ndim >= 1 && dens->ndim <= HEPMC_MAX_DIM, HEPMC_E_UNSUPPORTED,
                  "hepmc_density_eval: ndim %d outside [1, %d]", dens->ndim, HEPMC_MAX_DIM);
    HEPMC_REQUIRE(dens->kind >= HEPMC_CAMEL && dens->kind <= HEPMC_UNIFORM, HEPMC_E_ARG,
                  "hepmc_density_eval: unknown density kind %d", dens->kind);
    HEPMC_REQUIRE(what >= HEPMC_PDF && what <= HEPMC_PDF_GRADIENT, HEPMC_E_ARG, "hepmc_density_eval: bad what=%d", what);
    HEPMC_REQUIRE(!(dens->kind == HEPMC_BANANA && what == HEPMC_PDF_GRADIENT), HEPMC_E_UNSUPPORTED,
                  "Banana.pdf_gradient is not implemented by the reference (density.py:73)");
    HEPMC_REQUIRE(!(dens->kind == HEPMC_BANANA && what == HEPMC_POT_GRADIENT && dens->ndim != 2), HEPMC_E_UNSUPPORTED,
                  "Banana.pot_gradient exists for ndim == 2 only (banana.py:44-51)");
    HEPMC_REQUIRE(n >= 0, HEPMC_E_ARG, "hepmc_density_eval: n < 0");
    if (n == 0) return 0;
    HEPMC_REQUIRE(xs_dev && out_dev, HEPMC_E_ARG, "hepmc_density_eval: NULL buffer");
    const int threads = 256;
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    density_eval_kernel<<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(*dens, what, xs_dev, n, out_dev);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

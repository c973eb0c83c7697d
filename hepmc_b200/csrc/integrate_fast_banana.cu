// Throughput instances of the fused plain-MC / VEGAS kernels for the banana integrand:
// compile-time dimensionality 1..16, native Philox draws, no per-point outputs.
#include "integrate_kernels.cuh"

namespace hepmc {
HEPMC_DEFINE_FAST_LAUNCHER(launch_fast_banana, HEPMC_BANANA)
}  // namespace hepmc

// Sample-only instances (no fused integrand) of the plain-MC / VEGAS kernels for user integrands:
// compile-time dimensionality 1..16, native Philox draws, points / weights / bin indices written out
// for the two-kernel path (sample -> user fn on CUDA tensors -> accumulate).
#include "integrate_kernels.cuh"

namespace hepmc {
HEPMC_DEFINE_FAST_LAUNCHER(launch_fast_none, HEPMC_NONE)
}  // namespace hepmc

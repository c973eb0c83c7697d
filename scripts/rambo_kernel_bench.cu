// Development harness: hepmc_rambo_map (2->4) alone, three RNG modes, CUDA events; build variants with -D.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hepmc_b200/csrc/rambo.cu"
namespace hepmc {
void set_error(const char*, ...) {}
int check_cuda(cudaError_t e, const char* w) { if (e != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e), w); exit(1); } return 0; }
int sm_count() { int n = 148; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, 0); return n; }
}
int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? (int64_t)atof(argv[1]) : (1LL << 27);
    double* out; cudaMalloc(&out, (size_t)n * 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; ++mode) {
        float sum = 0;
        for (int i = 0; i < 6; ++i) {
            cudaEventRecord(e0);
            int rc = hepmc_rambo_map(4, 100.0, n, 0, 1234, 7 + i, mode, nullptr, out, nullptr);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            if (rc) { printf("rc %d\n", rc); return 1; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (i >= 2) sum += ms;
        }
        std::vector<double> h(64); cudaMemcpy(h.data(), out + 16 * 12345, 64 * 8, cudaMemcpyDeviceToHost);
        double cs = 0; for (double v : h) cs += v;
        cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, hepmc::rambo_kernel<4, 1, false>);
        printf("mode %d  %.3f ms  %.4g pts/s  %.0f GB/s  regs %d  checksum %.15g\n", mode, sum / 4, n / (sum / 4 * 1e-3), n * 128.0 / (sum / 4 * 1e-3) / 1e9, fa.numRegs, cs);
    }
    return 0;
}

// Library plumbing: error string, device info, FP64 DFMA probe (roofline denominator).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace hepmc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// 8 independent DFMA chains per thread; 16 flops per inner iteration per thread.
__global__ void __launch_bounds__(256) fp64_probe_kernel(int64_t iters, double seed, double* sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0;
    double a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 0.999999, c = 1e-9;
    for (int64_t i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

// 8 independent ex2.approx chains per thread (MUFU.EX2): the special-function roofline of the float32 /
// tensor-core surrogate kernels, whose binding unit is one exponential per (chain, node, gradient).
__global__ void __launch_bounds__(256) mufu_probe_kernel(int64_t iters, float seed, float* sink) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + 0.001f * (threadIdx.x + i);
    for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float r;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a[i]));
            a[i] = r - 1.0f;          // keeps the argument in (-1, 1): one FADD per exponential
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
}

}  // namespace hepmc

extern "C" {

int hepmc_abi_version(void) { return HEPMC_B200_ABI_VERSION; }

const char* hepmc_last_error(void) { return hepmc::g_err; }

int hepmc_device_info(int device, int64_t* out5) {
    HEPMC_REQUIRE(out5 != nullptr, HEPMC_E_ARG, "hepmc_device_info: out5 is NULL");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        hepmc::set_error("no CUDA device");
        return HEPMC_E_NODEVICE;
    }
    int v = 0;
    HEPMC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)); out5[0] = v;
    HEPMC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device)); out5[1] = v;
    HEPMC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device)); out5[2] = v;
    HEPMC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)); out5[3] = v;
    HEPMC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device)); out5[4] = v;
    return 0;
}

int hepmc_fp64_probe(int64_t iters, double* ms_out, double* flops_out, void* stream) {
    HEPMC_REQUIRE(ms_out && flops_out && iters > 0, HEPMC_E_ARG, "hepmc_fp64_probe: bad arguments");
    cudaStream_t st = hepmc::as_stream(stream);
    const int blocks = hepmc::sm_count() * 8;
    double* sink = nullptr;
    HEPMC_CUDA(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    HEPMC_CUDA(cudaEventCreate(&e0));
    HEPMC_CUDA(cudaEventCreate(&e1));
    hepmc::fp64_probe_kernel<<<blocks, 256, 0, st>>>(iters / 8 + 1, 1.0, sink);  // warm-up
    HEPMC_CUDA(cudaEventRecord(e0, st));
    hepmc::fp64_probe_kernel<<<blocks, 256, 0, st>>>(iters, 1.0, sink);
    HEPMC_CUDA(cudaEventRecord(e1, st));
    HEPMC_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    HEPMC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms;
    *flops_out = 16.0 * (double)iters * 256.0 * (double)blocks;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_mufu_probe(int64_t iters, double* ms_out, double* ops_out, void* stream) {
    HEPMC_REQUIRE(ms_out && ops_out && iters > 0, HEPMC_E_ARG, "hepmc_mufu_probe: bad arguments");
    cudaStream_t st = hepmc::as_stream(stream);
    const int blocks = hepmc::sm_count() * 8;
    float* sink = nullptr;
    HEPMC_CUDA(cudaMalloc(&sink, sizeof(float)));
    cudaEvent_t e0, e1;
    HEPMC_CUDA(cudaEventCreate(&e0));
    HEPMC_CUDA(cudaEventCreate(&e1));
    hepmc::mufu_probe_kernel<<<blocks, 256, 0, st>>>(iters / 8 + 1, 0.5f, sink);  // warm-up
    HEPMC_CUDA(cudaEventRecord(e0, st));
    hepmc::mufu_probe_kernel<<<blocks, 256, 0, st>>>(iters, 0.5f, sink);
    HEPMC_CUDA(cudaEventRecord(e1, st));
    HEPMC_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    HEPMC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms;
    *ops_out = 8.0 * (double)iters * 256.0 * (double)blocks;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

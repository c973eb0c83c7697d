// exp stage in isolation: TMEM ld -> 16 x ex2 -> TMEM st, per warp, software pipelined (development tool)
#include <cstdio>
#include "../hepmc_b200/csrc/elm_tc.cuh"
using namespace hepmc::tc;

// variant 0: ld16 / ex2 / st16 double-buffered (the kernel's loop); 1: no st; 2: no ld (regs only) + st; 3: ex2 only
// variant 4: ld32 / ex2 / st32 double buffered
__global__ void __launch_bounds__(256) k(int variant, int chunks, int warps_active, long long* out) {
    __shared__ uint32_t tmem_ptr;
    if (threadIdx.x / 32 == 0) tmem_alloc(smem_u32(&tmem_ptr), 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tb = tmem_ptr;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_base = tb + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    long long t0 = 0, t1 = 0;
    if (warp < warps_active) {
        uint32_t va[16], vb[16];
        for (int j = 0; j < 16; ++j) { va[j] = __float_as_uint(-0.01f * j); vb[j] = va[j]; }
        for (int c = 0; c < 256; c += 16) tmem_st16(lane_base + c, va);
        tmem_wait_st();
        __syncwarp();
        t0 = clock64();
        if (variant == 4) {
            uint32_t wa[32], wb[32];
            tmem_ld32(lane_base, wa);
            for (int c = 0; c < chunks / 2; c += 2) {
                tmem_wait_ld();
                tmem_ld32(lane_base + ((c + 1) % 8) * 32, wb);
#pragma unroll
                for (int j = 0; j < 32; ++j) wa[j] = __float_as_uint(ex2_approx(__uint_as_float(wa[j])));
                tmem_st32(lane_base + (c % 8) * 32, wa);
                tmem_wait_ld();
                tmem_ld32(lane_base + ((c + 2) % 8) * 32, wa);
#pragma unroll
                for (int j = 0; j < 32; ++j) wb[j] = __float_as_uint(ex2_approx(__uint_as_float(wb[j])));
                tmem_st32(lane_base + ((c + 1) % 8) * 32, wb);
            }
            tmem_wait_ld();
            tmem_wait_st();
            if (wa[0] == 0x1234567u) out[3] = wb[1];
        } else if (variant == 5 || variant == 6) {
            tmem_ld16(lane_base, va);
            for (int c = 0; c < chunks; c += 2) {
                tmem_wait_ld(); tmem_ld16(lane_base + ((c + 1) % 16) * 16, vb);
                if (variant == 5) ex2_chunk(va);
                else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) va[j] = __float_as_uint(ex2_poly(__uint_as_float(va[j])));
                }
                tmem_st16(lane_base + (c % 16) * 16, va);
                tmem_wait_ld(); tmem_ld16(lane_base + ((c + 2) % 16) * 16, va);
                if (variant == 5) ex2_chunk(vb);
                else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) vb[j] = __float_as_uint(ex2_poly(__uint_as_float(vb[j])));
                }
                tmem_st16(lane_base + ((c + 1) % 16) * 16, vb);
            }
            tmem_wait_ld();
            tmem_wait_st();
            if (va[0] == 0x1234567u) out[3] = vb[1];
        } else {
            if (variant == 0 || variant == 1) tmem_ld16(lane_base, va);
            for (int c = 0; c < chunks; c += 2) {
                if (variant == 0 || variant == 1) { tmem_wait_ld(); tmem_ld16(lane_base + ((c + 1) % 16) * 16, vb); }
#pragma unroll
                for (int j = 0; j < 16; ++j) va[j] = __float_as_uint(ex2_approx(__uint_as_float(va[j])));
                if (variant == 0 || variant == 2) tmem_st16(lane_base + (c % 16) * 16, va);
                if (variant == 0 || variant == 1) { tmem_wait_ld(); tmem_ld16(lane_base + ((c + 2) % 16) * 16, va); }
#pragma unroll
                for (int j = 0; j < 16; ++j) vb[j] = __float_as_uint(ex2_approx(__uint_as_float(vb[j])));
                if (variant == 0 || variant == 2) tmem_st16(lane_base + ((c + 1) % 16) * 16, vb);
            }
            tmem_wait_ld();
            tmem_wait_st();
            if (va[0] == 0x1234567u) out[3] = vb[1];
        }
        t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (threadIdx.x / 32 == 0) tmem_dealloc(tb, 512);
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 64);
    const char* names[] = {"ld16/ex2/st16", "ld16/ex2", "ex2/st16", "ex2 only", "ld32/ex2/st32", "ld16/mixed/st16", "ld16/poly/st16"};
    for (int wa : {4, 8}) {
        for (int v : {0, 5, 6}) {
            const int chunks = 4096;
            for (int it = 0; it < 2; ++it) {
                k<<<148, 256>>>(v, chunks, wa, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
            }
            printf("warps %d  %-16s %7.1f cycles per 16-column chunk per warp (MUFU floor: %d)\n", wa, names[v],
                   (double)out[0] / chunks, wa == 4 ? 128 : 256);
        }
    }
    return 0;
}

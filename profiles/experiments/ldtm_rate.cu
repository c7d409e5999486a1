// ldtm_rate.cu — TMEM -> register bandwidth of one SM: W warps (W = 4: one warpgroup, 8: two) each read their 32-lane slice
// of TMEM with tcgen05.ld.32x32b.x32 (4 KB per warp instruction) in a loop, wait::ld after each or after every other one.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/experiments/ldtm_rate profiles/experiments/ldtm_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../ggml-cuda-experiments_b200/csrc/sm100_ptx.cuh"
using namespace b200fa::ptx;

template <int X16>
__global__ void __launch_bounds__(256, 1) rate(int iters, int batch, long long* out) {
    __shared__ uint32_t tm;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(&tm, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        uint32_t r[2][32];
        if (X16) {
            uint32_t q[4][16];
            tmem_ld16(base, q[0]); tmem_ld16(base + 16, q[1]); tmem_ld16(base + 32, q[2]); tmem_ld16(base + 48, q[3]);
            tmem_wait_ld();
#pragma unroll
            for (int b = 0; b < 4; b++)
#pragma unroll
                for (int i = 0; i < 16; i++) acc ^= q[b][i];
        } else {
            tmem_ld32(base, r[0]);
            if (batch == 2) tmem_ld32(base + 32, r[1]);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i++) acc ^= r[0][i];
            if (batch == 2) {
#pragma unroll
                for (int i = 0; i < 32; i++) acc ^= r[1][i];
            }
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) out[1] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    const int iters = 2000;
    for (int warps : {1, 4, 8}) for (int batch : {1, 2}) {
        rate<0><<<148, warps * 32>>>(iters, batch, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)iters * batch * warps * 4096;
        printf("x32 loads, %d warps, %d per wait: %7.1f cycles per round, %6.1f B/clk/SM %s\n", warps, batch, (double)h / iters, bytes / h, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    for (int warps : {4, 8}) {
        rate<1><<<148, warps * 32>>>(iters, 2, d);
        cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("x16 loads, %d warps, 4 per wait: %7.1f cycles per round, %6.1f B/clk/SM\n", warps, (double)h / iters, (double)iters * 4 * warps * 2048 / h);
    }
    return 0;
}

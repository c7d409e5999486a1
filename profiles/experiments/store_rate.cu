// store_rate.cu — how fast can one SM push a finished 128 x 128 f32 tile (64 KB) to global memory?
// Measures cycles for each warp of a CTA to store 16 KB (32 rows x 512 B, row stride 16 KB like dst[q][head][D]) by
//   mode 0: STG.128 straight from registers, 4 rows x 128 B per instruction
//   mode 1: the same through a shared-memory transposition (STS, LDS, STG), as the prefill epilogue does
//   mode 2: cp.async.bulk (1-D, 512 B per row) from shared memory, one elected lane, wait_group.read before reuse
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a store_rate.cu -o store_rate && ./store_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256) k_store(char* out, long long* cycles, int n_warps, int64_t row_stride) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= n_warps) return;
    char* base = out + ((int64_t)blockIdx.x * 8 + warp) * 32 * row_stride;   // this warp's 32 rows
    uint4* stg = reinterpret_cast<uint4*>(smem + warp * 16384);
    uint4 v = make_uint4(lane, warp, blockIdx.x, 7);
    __syncwarp();
    const long long t0 = clock64();
    if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int i = q * 4 + (lane >> 3);
                *reinterpret_cast<uint4*>(base + i * row_stride + c * 128 + (lane & 7) * 16) = v;
            }
    } else if (MODE == 1) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int j = 0; j < 8; j++) stg[lane * 8 + (j ^ (lane & 7))] = v;
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int i = q * 4 + (lane >> 3), j = lane & 7;
                const uint4 x = stg[i * 8 + (j ^ (i & 7))];
                *reinterpret_cast<uint4*>(base + i * row_stride + c * 128 + j * 16) = x;
            }
            __syncwarp();
        }
    } else {
        // whole 16 KB staged once ([row][512 B]), then 32 bulk copies of 512 B
#pragma unroll
        for (int j = 0; j < 32; j++) stg[lane * 32 + (j ^ lane)] = v;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            for (int i = 0; i < 32; i++)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(base + i * row_stride), "r"(smem_u32(stg + i * 32)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * 8 + warp] = t1 - t0;
}

int main() {
    const int64_t row_stride = 16384;
    const int G = 148;
    char* out; cudaMalloc(&out, (size_t)G * 8 * 32 * row_stride);
    long long* cyc; cudaMalloc(&cyc, G * 8 * 8);
    cudaFuncSetAttribute(k_store<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384);
    cudaFuncSetAttribute(k_store<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384);
    cudaFuncSetAttribute(k_store<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384);
    for (int mode = 0; mode < 3; mode++)
        for (int grid : {1, 148})
            for (int nw : {1, 4, 8}) {
                long long h[G * 8];
                for (int rep = 0; rep < 3; rep++) {
                    cudaMemset(cyc, 0, G * 8 * 8);
                    if (mode == 0) k_store<0><<<grid, 256, 8 * 16384>>>(out, cyc, nw, row_stride);
                    if (mode == 1) k_store<1><<<grid, 256, 8 * 16384>>>(out, cyc, nw, row_stride);
                    if (mode == 2) k_store<2><<<grid, 256, 8 * 16384>>>(out, cyc, nw, row_stride);
                    cudaDeviceSynchronize();
                }
                cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
                long long mx = 0, sum = 0; int n = 0;
                for (int i = 0; i < grid * 8; i++) if (h[i]) { if (h[i] > mx) mx = h[i]; sum += h[i]; n++; }
                printf("mode %d grid %3d warps %d: %6lld cycles avg, %6lld max per warp for 16 KB -> %.1f B/clk/SM\n", mode, grid, nw, sum / (n ? n : 1), mx, nw * 16384.0 / mx);
            }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// umma_rate.cu — how many cycles does one tcgen05.mma (kind::f16, M=128, K=16) take on one SM, as a function of where its
// operands come from?  SS = A and B from shared memory (128B-swizzled K-major tiles), TS = A from TMEM.  N = 64/128/256.
// One CTA per SM issues `iters` x 8 MMAs back to back (the 8 k-steps of a 128-wide tile), commits, waits; clock64 around.
// build + run: python profiles/experiments/umma_rate.py
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../ggml-cuda-experiments_b200/csrc/sm100_ptx.cuh"
using namespace b200fa::ptx;

struct __align__(1024) Sm { uint8_t a[32768]; uint8_t b[65536]; uint64_t bar; uint32_t tmem; };

template <int N, bool TS, bool BMN>
__global__ void __launch_bounds__(128, 1) rate(int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    Sm& sm = *reinterpret_cast<Sm*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(sm.a)[i] = 0x3c003c00u;  // f16 1.0
    if (threadIdx.x == 0) { mbar_init(&sm.bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&sm.tmem, 512); tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = sm.tmem;
    if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_f16(128, N, 0, BMN ? 1 : 0);
        const uint64_t da = make_smem_desc_sw128(smem_u32(sm.a), 16, 1024);
        const uint64_t db = BMN ? make_smem_desc_sw128(smem_u32(sm.b), 16384, 1024) : make_smem_desc_sw128(smem_u32(sm.b), 16, 1024);
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; rep++) {
            t0 = clock64();
            if (elect_one()) {
                for (int it = 0; it < iters; it++) {
#pragma unroll
                    for (int ks = 0; ks < 8; ks++) {
                        const uint64_t off = BMN ? (uint64_t)(ks * 2048 >> 4) : (uint64_t)(((ks >> 2) * 16384 + (ks & 3) * 32) >> 4);
                        const uint64_t offa = (uint64_t)(((ks >> 2) * 16384 + (ks & 3) * 32) >> 4);
                        if (TS) mma_ts(tmem + 256, tmem + ks * 8, db + off, idesc, 1u);
                        else mma_ss(tmem + 256, da + offa, db + off, idesc, 1u);
                    }
                }
                tc_commit(&sm.bar);
            }
            __syncwarp();
            mbar_wait(&sm.bar, rep & 1);
            t1 = clock64();
        }
        if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool TS, bool BMN>
static void run(const char* name, long long* d_out) {
    const int iters = 64;
    const int smem = sizeof(Sm) + 1024;
    cudaFuncSetAttribute(rate<N, TS, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate<N, TS, BMN><<<148, 128, smem>>>(iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %7.1f cycles per MMA (ideal %d)  %s\n", name, (double)h / (iters * 8), 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

extern "C" int umma_rate_main() {
    long long* d; cudaMalloc(&d, 64);
    run<128, false, false>("SS  M128 N128 (A,B K-major: Q K^T)", d);
    run<256, false, false>("SS  M128 N256 (A,B K-major)", d);
    run<64, false, false>("SS  M128 N64", d);
    run<128, true, true>("TS  M128 N128 (A tmem, B MN-major: P V)", d);
    run<128, false, true>("SS  M128 N128 (B MN-major)", d);
    run<256, true, true>("TS  M128 N256 (B MN-major)", d);
    cudaFree(d);
    return 0;
}

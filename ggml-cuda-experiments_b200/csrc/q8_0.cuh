// q8_0.cuh — ggml block_q8_0 {f16 d; int8 qs[32]} quantise / dequantise kernels.
// Not in the reference; format and rounding are ggml's published ones, pinned against
// gguf.quants through the oracle.  One warp per block, lane = element.
#pragma once
#include "common.cuh"

namespace b200fa {

template <typename T>
__global__ void q8_0_quantize_kernel(const T* __restrict__ x, uint8_t* __restrict__ y, int64_t n_blocks) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= n_blocks) return;
    float v;
    if constexpr (sizeof(T) == 2) v = __half2float(x[b * 32 + lane]);
    else v = x[b * 32 + lane];
    float amax = fabsf(v);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, s));
    const float d = __fdiv_rn(amax, 127.0f);
    const float id = d ? __fdiv_rn(1.0f, d) : 0.0f;
    const float q = roundf(__fmul_rn(v, id));  // round half away from zero, like ggml's roundf
    uint8_t* blk = y + b * kQ8BlockBytes;
    if (lane == 0) *reinterpret_cast<__half*>(blk) = __float2half_rn(d);
    reinterpret_cast<int8_t*>(blk + 2)[lane] = (int8_t)q;
}

__global__ void q8_0_dequantize_kernel(const uint8_t* __restrict__ x, float* __restrict__ y, int64_t n_blocks) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= n_blocks) return;
    const uint8_t* blk = x + b * kQ8BlockBytes;
    const float d = __half2float(*reinterpret_cast<const __half*>(blk));
    y[b * 32 + lane] = __fmul_rn(d, (float)reinterpret_cast<const int8_t*>(blk + 2)[lane]);
}

}  // namespace b200fa

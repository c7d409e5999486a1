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

// KV-cache append: rows of new K (or V) vectors -> the cache tensor at position n_past, converting on
// the way: f32/f16 -> f16, or -> ggml q8_0 blocks (same rounding as q8_0_quantize_kernel).  This is the step the
// reference's driver does by hand on the host (flash-matrix.cu:130-165).  One warp per (token, head, batch) row.
template <typename T>
__global__ void kv_append_kernel(const char* __restrict__ src, char* __restrict__ cache, int cache_type, int D, int n_tokens,
                                 int n_head_kv, int64_t n_rows, int64_t snb1, int64_t snb2, int64_t snb3, int64_t cnb1, int64_t cnb2,
                                 int64_t cnb3, int64_t n_past) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int tok = (int)(row % n_tokens), head = (int)((row / n_tokens) % n_head_kv);
    const int64_t b = row / ((int64_t)n_tokens * n_head_kv);
    const T* x = reinterpret_cast<const T*>(src + tok * snb1 + head * snb2 + b * snb3);
    char* dst = cache + (n_past + tok) * cnb1 + head * cnb2 + b * cnb3;
    for (int blk = 0; blk < (D + 31) / 32; blk++) {
        const bool in = blk * 32 + lane < D;  // f16 caches take any head size (multiple of 8); q8_0 rows are whole 32-element blocks
        float v = 0.f;
        if (in) {
            if constexpr (sizeof(T) == 2) v = __half2float(x[blk * 32 + lane]);
            else v = x[blk * 32 + lane];
        }
        if (cache_type == B200FA_TYPE_F16) {
            if (in) reinterpret_cast<__half*>(dst)[blk * 32 + lane] = __float2half_rn(v);
        } else {
            float amax = fabsf(v);
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, s));
            const float d = __fdiv_rn(amax, 127.0f);
            const float id = d ? __fdiv_rn(1.0f, d) : 0.0f;
            uint8_t* qb = reinterpret_cast<uint8_t*>(dst) + blk * kQ8BlockBytes;
            if (lane == 0) *reinterpret_cast<__half*>(qb) = __float2half_rn(d);
            reinterpret_cast<int8_t*>(qb + 2)[lane] = (int8_t)roundf(__fmul_rn(v, id));
        }
    }
}

// q8_0 K/V rows (any ne/nb strides) -> dense f16 [batch][head][row][D] with y = RN_f16(f32(d) * q): the pre-pass that lets the
// tcgen05 prefill kernel run on a quantised KV cache.  One thread per 8 elements (five independent 16-bit loads — blocks are only
// 2-byte aligned — and one 16-byte store); a warp-per-block version was latency-bound at 0.5 TB/s.
__global__ void __launch_bounds__(256) q8_rows_to_f16_kernel(const char* __restrict__ src, __half* __restrict__ dst, int D, int n_rows,
                                                             int n_head, int64_t n_oct, int64_t nb1, int64_t nb2, int64_t nb3) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // octet index over the dense output
    if (o >= n_oct) return;
    const int opr = D / 8;                             // octets per row
    const int c8 = (int)(o % opr);
    const int64_t r = o / opr;                         // dense row index ((b * n_head + h) * n_rows + row)
    const int row = (int)(r % n_rows), h = (int)((r / n_rows) % n_head);
    const int64_t b = r / ((int64_t)n_rows * n_head);
    const uint16_t* blk = reinterpret_cast<const uint16_t*>(src + row * nb1 + h * nb2 + b * nb3 + (c8 >> 2) * kQ8BlockBytes);
    const uint16_t* qs = blk + 1 + (c8 & 3) * 4;
    const uint32_t w0 = __ldg(qs), w1 = __ldg(qs + 1), w2 = __ldg(qs + 2), w3 = __ldg(qs + 3);
    const float d = __half2float(__ushort_as_half(__ldg(blk)));
    const uint32_t w[4] = {w0, w1, w2, w3};
    uint32_t out[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float lo = __fmul_rn(d, (float)(int8_t)(w[i] & 0xffu)), hi = __fmul_rn(d, (float)(int8_t)((w[i] >> 8) & 0xffu));
        const __half2 hh = __halves2half2(__float2half_rn(lo), __float2half_rn(hi));
        out[i] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    *reinterpret_cast<uint4*>(dst + r * D + c8 * 8) = make_uint4(out[0], out[1], out[2], out[3]);
}

}  // namespace b200fa

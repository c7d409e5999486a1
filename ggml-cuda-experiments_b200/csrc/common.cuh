// common.cuh — shared definitions for the b200fa kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200fa.h"

#include <stdlib.h>

namespace b200fa {

// Tuning knobs are environment variables that exist ONLY in -DB200FA_TUNING builds (ggml-cuda-experiments_b200/build.py
// build(tuning=True) -> libb200fa_tuning.so, used by the profiles/ tools).  In the shipped library tune_env() is the constant
// nullptr: no getenv, no process-global switches, and the variants they select are folded away.
#ifdef B200FA_TUNING
inline const char* tune_env(const char* name) { return getenv(name); }
#else
inline const char* tune_env(const char*) { return nullptr; }
#endif

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kQ8BlockBytes = 34;
constexpr int kQ8BlockElems = 32;

// Everything a kernel needs to know about one attention call.  Mirrors the reference kernel's
// argument list (flash-llama.h:6-32) plus the split-KV bookkeeping of flash_attn_row/fa_reduce.
struct FaParams {
    const char* q;
    const char* k;
    const char* v;
    const char* mask;  // f16, may be null
    void* dst;         // final output (f32/f16), or null when only partials are wanted
    float* part;       // split-KV partial triples: [split][row][D+2]  (O~[D], m, l)
    float* part_out;   // sequence-split entry: one merged triple per row goes here instead of dst
    unsigned int* counters;  // one arrival counter per (row group, kv head, batch); zero between calls
    float scale;       // as passed by the caller
    float scale_log2;  // scale * log2(e)
    int q_type, kv_type, dst_type;
    int D, n_q, n_head, n_batch;  // D: the head size the kernel is BUILT for (64 or 128; the data is zero-padded up to it on the fly)
    int Dr;            // the real head size ne00 (multiple of 8, <= D): extent of Q/K/V/dst rows in memory
    int n_kv, n_head_kv, n_batch_kv;
    int gqa;           // rk2 = n_head / n_head_kv   (flash-llama.h:128)
    int kv_div;        // stream decode with 17..128 rows per KV head: n_head_kv and gqa describe VIRTUAL kv heads (a real head split into
                       // kv_div groups of q heads, <= 16 rows each); K/V of virtual head v are those of real head v / kv_div.  1 otherwise
    int rk3;           // n_batch / n_batch_kv       (flash-llama.h:129)
    int64_t nb01, nb02, nb03;
    int64_t nb11, nb12, nb13;
    int64_t nb21, nb22, nb23;
    int64_t nb31;
    // mask slices (b200fa_flash_attn_ext2, upstream ggml's ne32/ne33 broadcast): m_ne2 in {1, n_head}, m_ne3 in {1, n_batch}; the mask row
    // of (query iq1, head iq2, batch iq3) starts at mask + iq1*nb31 + (iq2 % m_ne2)*nb32 + (iq3 % m_ne3)*nb33.  The reference shares one
    // mask between heads and batches (flash-llama.h:151,194): m_ne2 = m_ne3 = 1.
    int m_ne2, m_ne3;
    int64_t nb32, nb33;
    int causal;          // B200FA_FLAG_CAUSAL
    int64_t kv_pos0;     // global position of this slice's first key (sequence-split); 0 otherwise
    int64_t causal_off;  // a query at iq1 sees global kv positions <= iq1 + causal_off  (n_kv_total - n_q)
    int n_splits;        // KV splits across CTAs
    int split_len;       // keys per split (multiple of 16)
    int64_t total_rows;  // n_batch * n_q * n_head
    int dbg_mode;        // tuning only (env B200FA_DBG_MODE): 1 = stream K/V but skip the tile maths
    // b200fa_flash_attn_ext2 extensions (upstream ggml semantics; all zero / off on the reference's entry)
    float alibi_m0l, alibi_m1l;  // log2 of the ALiBi bases: slope(h) = 2^(m0l*(h+1)) for h < n_head_log2, else 2^(m1l*(2(h-n_head_log2)+1))
    int alibi_nhl2;              // n_head_log2; 0 = no ALiBi (slope 1)
    float cap_in;                // logit soft-cap: s_log2 = tanh(qk * cap_in) * cap_out; cap_in = scale / cap; 0 = off
    float cap_out;               //                 cap * log2(e)
    float cap_raw;               // cap / scale: tanh(qk * cap_in) * cap_raw is the capped score in RAW (unscaled) units
};

// byte offset of the mask slice of (head iq2, batch iq3)
__device__ __forceinline__ int64_t fa_mask_slice_off(const FaParams& p, int iq2, int iq3) {
    return (p.m_ne2 > 1 ? (int64_t)iq2 * p.nb32 : 0) + (p.m_ne3 > 1 ? (int64_t)iq3 * p.nb33 : 0);
}
// index of that slice in per-slice tables (mask tile classes of the prefill kernel)
__device__ __forceinline__ int fa_mask_slice(const FaParams& p, int iq2, int iq3) {
    return (p.m_ne2 > 1 ? iq2 : 0) + p.m_ne2 * (p.m_ne3 > 1 ? iq3 : 0);
}

// ALiBi slope of query head h (1 when off)
__device__ __forceinline__ float fa_slope(const FaParams& p, int h) {
    if (p.alibi_nhl2 == 0) return 1.f;
    return exp2f(h < p.alibi_nhl2 ? p.alibi_m0l * (float)(h + 1) : p.alibi_m1l * (float)(2 * (h - p.alibi_nhl2) + 1));
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// tanh through ex2 + rcp (tanh.approx is only good to 2^-11, which a soft-cap of 30-50 would turn into visible score errors)
__device__ __forceinline__ float fa_tanh(float x) {
    const float e = fast_exp2(x * 2.88539008f);  // e^(2x); inf for large x -> 1, 0 for very negative x -> -1
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return fmaf(-2.f, r, 1.f);
}

__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// D = A * B with a zero C operand: no accumulator registers to clear beforehand (the C quad becomes RZ)
__device__ __forceinline__ void mma_16816_zc(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}

__device__ __forceinline__ float ld_mask(const char* mask_row, int64_t col) {
    return __half2float(*reinterpret_cast<const __half*>(mask_row + col * 2));
}

}  // namespace b200fa

// decode_mma.cuh — split-KV "rows16" kernel: the bandwidth-bound decode path, and the general path for
// any shape the tcgen05 prefill kernel does not take.
//
// Replaces the reference's flash_attn_row<128,8,2,256> / flash_attn_row_fast (flash_row_float.h:4-413) AND its
// fa_reduce<128,nw> (flash_row_float.h:415-472): the last CTA to finish a row group merges the splits, so a
// decode call is ONE kernel launch (the reference needs two plus a cudaMalloc).
//
// One CTA = one KV split of one (kv head, batch) for one group of up to 16 output rows, where a "row" is a
// (query position, q head of the GQA group) pair — all rows of a group share the K/V stream, so K/V is
// read from HBM once per GQA group (the reference re-reads it per q head, flash_row_float.h:19,58).
//
// Data path: each warp streams 16-key tiles straight from global memory into mma.sync fragments with
// 128-bit ld.global.nc.L1::no_allocate loads, double-buffered in registers (tile i+1 is in flight while
// tile i is computed) — no shared-memory staging, no shuffles for operands.
// The trick is that the contraction index of an MMA may be permuted freely as long as both operands use
// the same permutation, so every lane loads 16 contiguous bytes of the row it needs:
//   QK^T : B-fragment lane (g,t) loads K[row rho(g)][8*(t+4c) .. +7]; the matching A-fragment lane loads the
//          same 16-byte chunk of Q rows g and g+8.  rho maps fragment column n to key 4*(n/2)+2*nt+(n%2)
//          so that a lane ends up holding scores of keys 4t..4t+3 — exactly the P·V A-fragment it needs.
//   P·V  : lane (g,t) loads V[key 4t+i][64c+8g .. +7] for i=0..3 and byte-permutes pairs of rows into
//          B-fragments; output column n of MMA (c,j) is head dim 64c+8n+j.
// Scores, softmax state and the output accumulate in fp32 (the reference keeps them in f16,
// flash_row_float.h:51,93,159).  exp is exp2 with scale*log2(e) folded into one FMA.
//
// q8_0 K/V: int8 -> f16 is exact (magic-number trick), K block scales are applied in fp32 to per-block
// partial dot products (bit-equivalent to dotting with f32(d)*q), V is dequantised to f16 = RN(d*q).
#pragma once
#include "common.cuh"

namespace b200fa {

constexpr int kRows = 16;     // output rows per CTA (MMA M)
constexpr int kTileKV = 16;   // keys per warp iteration
constexpr int kDecodeWarps = 4;
constexpr int kStages = 3;    // per-warp cp.async FIFO depth (f16 K/V): two tiles in flight behind the one being computed

// Per-warp FIFO stage: every lane owns one 16-byte slot per chunk (it copies the chunk in with cp.async and
// reads the same slot back into its MMA fragment, so no cross-lane synchronisation is needed), plus 8-byte mask slots.
template <int D>
struct FifoGeom {
    static constexpr int kChunks = (D / 32) * 2 + (D / 64) * 4;      // K: 2 n-tiles x D/32, V: 4 keys x D/64
    static constexpr int kStageBytes = kChunks * 512 + 2 * 256;      // + mask slots for up to 2 rows
    static constexpr int kWarpBytes = kStages * kStageBytes;
    static constexpr int kCtaBytes = kDecodeWarps * kWarpBytes;
};

__device__ __forceinline__ void cp_async16(uint32_t smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}
// src_bytes = 16 copies, 0 writes 16 zero bytes without touching gmem (columns past the real head size)
__device__ __forceinline__ void cp_async16z(uint32_t smem, const void* gmem, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
    return r;
}

// 4 int8 (packed in w) -> 4 f16 holding the same integers, exactly.  0x6400|u is 1024+u in f16.
__device__ __forceinline__ void q8x4_to_h2(uint32_t w, uint32_t& lo, uint32_t& hi) {
    const uint32_t u = w ^ 0x80808080u;  // int8 -> biased uint8
    uint32_t l = prmt(u, 0x64646464u, 0x4140);
    uint32_t h = prmt(u, 0x64646464u, 0x4342);
    const uint32_t bias = 0x64806480u;   // (1152, 1152) = 1024 + 128
    __half2 lh = __hsub2(*reinterpret_cast<__half2*>(&l), *reinterpret_cast<const __half2*>(&bias));
    __half2 hh = __hsub2(*reinterpret_cast<__half2*>(&h), *reinterpret_cast<const __half2*>(&bias));
    lo = *reinterpret_cast<uint32_t*>(&lh);
    hi = *reinterpret_cast<uint32_t*>(&hh);
}

// 8 int8 at a 2-byte-aligned address.  `al8`: the row base is 8-byte aligned, so the 8 bytes sit inside two
// aligned 8-byte words that are loaded whole and funnel-shifted; otherwise four 16-bit loads.
__device__ __forceinline__ uint2 ld_q8x8(const char* p, bool al8) {
    if (al8) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const uint32_t sh = (uint32_t)(a & 7);  // 0, 2, 4 or 6
        const uint2* base = reinterpret_cast<const uint2*>(a - sh);
        const uint2 lo = __ldg(base);
        if (sh == 0) return lo;
        const uint2 hi = __ldg(base + 1);
        if (sh == 4) return make_uint2(lo.y, hi.x);
        if (sh == 2) return make_uint2(__funnelshift_r(lo.x, lo.y, 16), __funnelshift_r(lo.y, hi.x, 16));
        return make_uint2(__funnelshift_r(lo.y, hi.x, 16), __funnelshift_r(hi.x, hi.y, 16));
    }
    const uint16_t* s = reinterpret_cast<const uint16_t*>(p);
    const uint32_t a = __ldg(s), b = __ldg(s + 1), c = __ldg(s + 2), d = __ldg(s + 3);
    return make_uint2(a | (b << 16), c | (d << 16));
}

__device__ __forceinline__ uint16_t ld_u16(const char* p) { return __ldg(reinterpret_cast<const uint16_t*>(p)); }
__device__ __forceinline__ float h_bits_to_f(uint32_t b) { return __half2float(__ushort_as_half((unsigned short)b)); }

// m16n8k16 with only rows 0-7 of A/C live (RH == 1): rows 8-15 are fed zeros and their outputs are dropped.
__device__ __forceinline__ void mma_16816_top(float& c0, float& c1, uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    float d2, d3;
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%10,%10};"
        : "+f"(c0), "+f"(c1), "=f"(d2), "=f"(d3)
        : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1), "f"(0.f));
}

template <int D, bool Q8>
struct KVTile {
    static constexpr int NC4 = D / 32, NCV = D / 64;
    // f16: kf/vf hold ready fragments.  q8_0: raw int8 words (2 per 8 elements) + f16 scale bits.
    uint32_t kf[2][NC4][Q8 ? 2 : 4];
    uint32_t vf[4][NCV][Q8 ? 2 : 4];
    uint32_t kd[Q8 ? 4 : 1][Q8 ? NC4 : 1];   // K block scales (f16 bits) of this lane's 4 score columns
    uint32_t vd[Q8 ? 4 : 1][Q8 ? NCV : 1];   // V block scale (f16 bits) of (key 4t+i, the block holding dims 64c+8g..)
    uint2 mk[2];                             // mask halves of keys kv0+4t..+3 for the lane's row(s)
};

// RH = 1: the group has at most 8 live rows (only fragment rows g), RH = 2: up to 16 (rows g and g+8)
template <int D, int KV_TYPE, int RH, bool FIFO = true, bool EXT = false>
__global__ void __launch_bounds__(kDecodeWarps * 32, 2)
fa_rows16_splitkv(const __grid_constant__ FaParams p) {
    static_assert(D == 64 || D == 128 || D == 256, "head size");
    constexpr int NC4 = D / 32;   // 16-byte chunks per lane per K row (c loop) == q8_0 blocks per row
    constexpr int NCV = D / 64;   // 64-wide halves of a V row
    constexpr int NT = D / 8;     // output n-tiles
    constexpr bool Q8 = (KV_TYPE == B200FA_TYPE_Q8_0);
    constexpr int RLIVE = 8 * RH;
    using Tile = KVTile<D, Q8>;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int split = blockIdx.x, grp = blockIdx.y;
    const int ik2 = blockIdx.z % p.n_head_kv, iq3 = blockIdx.z / p.n_head_kv;
    const int ik3 = iq3 / p.rk3;
    const int rows_total = p.n_q * p.gqa;  // rows sharing this kv head

    // ---- the rows this lane owns (g, and g+8 when RH == 2) ----
    int iq1r[RH], iq2r[RH];
    bool rvalid[RH];
#pragma unroll
    for (int h = 0; h < RH; h++) {
        const int R = grp * kRows + g + 8 * h;
        rvalid[h] = R < rows_total;
        const int Rc = rvalid[h] ? R : 0;
        iq1r[h] = Rc / p.gqa;
        iq2r[h] = ik2 * p.gqa + Rc % p.gqa;
    }

    float mslope[RH];  // log2(e) x ALiBi slope of the row's head: the factor on raw mask values
#pragma unroll
    for (int h = 0; h < RH; h++) mslope[h] = EXT ? kLog2e * fa_slope(p, iq2r[h]) : kLog2e;

    // ---- Q fragments (f16; an f32 Q is rounded like the reference does, flash-llama.h:80) ----
    uint32_t qa[NC4][RH][4];
#pragma unroll
    for (int h = 0; h < RH; h++) {
        const char* qrow = p.q + iq1r[h] * p.nb01 + iq2r[h] * p.nb02 + (int64_t)iq3 * p.nb03;
#pragma unroll
        for (int c = 0; c < NC4; c++) {
            const int e0 = 8 * (t + 4 * c);
            if (!rvalid[h] || e0 >= p.Dr) {  // dead row, or a column past the real head size (zero padding)
                qa[c][h][0] = qa[c][h][1] = qa[c][h][2] = qa[c][h][3] = 0u;
            } else if (p.q_type == B200FA_TYPE_F16) {
                const uint4 x = *reinterpret_cast<const uint4*>(qrow + e0 * 2);
                qa[c][h][0] = x.x; qa[c][h][1] = x.y; qa[c][h][2] = x.z; qa[c][h][3] = x.w;
            } else {
                const float4 x = *reinterpret_cast<const float4*>(qrow + e0 * 4);
                const float4 y = *reinterpret_cast<const float4*>(qrow + e0 * 4 + 16);
                qa[c][h][0] = pack_half2(x.x, x.y); qa[c][h][1] = pack_half2(x.z, x.w);
                qa[c][h][2] = pack_half2(y.x, y.y); qa[c][h][3] = pack_half2(y.z, y.w);
            }
        }
    }

    // ---- KV range of this split, clipped by causality for the whole row group ----
    const int kv_begin = split * p.split_len;
    int kv_end = min(p.n_kv, kv_begin + p.split_len);
    if (p.causal) {
        const int last_row = min(rows_total, (grp + 1) * kRows) - 1;
        const int64_t lim = (int64_t)(last_row / p.gqa) + p.causal_off - p.kv_pos0 + 1;  // local keys < lim visible
        kv_end = (int)max((int64_t)kv_begin, min((int64_t)kv_end, lim));
    }
    int64_t vis[RH];  // per-row local visibility limit (exclusive) under the causal flag
#pragma unroll
    for (int h = 0; h < RH; h++) vis[h] = p.causal ? (int64_t)iq1r[h] + p.causal_off - p.kv_pos0 + 1 : (int64_t)p.n_kv;
    int lim[RH];  // keys >= lim[h] are invisible to row h (split end, sequence end, causality)
#pragma unroll
    for (int h = 0; h < RH; h++) lim[h] = (int)max((int64_t)0, min((int64_t)kv_end, vis[h]));

    const char* kbase = p.k + (int64_t)ik2 * p.nb12 + (int64_t)ik3 * p.nb13;
    const char* vbase = p.v + (int64_t)ik2 * p.nb22 + (int64_t)ik3 * p.nb23;
    const char* mrow[RH];
#pragma unroll
    for (int h = 0; h < RH; h++) mrow[h] = p.mask ? p.mask + iq1r[h] * p.nb31 + (EXT ? fa_mask_slice_off(p, iq2r[h], iq3) : 0) : nullptr;
    const bool al8 = Q8 && ((((uintptr_t)p.k | (uintptr_t)p.v | (uintptr_t)p.nb11 | (uintptr_t)p.nb12 | (uintptr_t)p.nb13 |
                              (uintptr_t)p.nb21 | (uintptr_t)p.nb22 | (uintptr_t)p.nb23) & 7) == 0);
    const bool mask_al8 = p.mask != nullptr && ((((uintptr_t)p.mask | (uintptr_t)p.nb31 | (uintptr_t)p.nb32 | (uintptr_t)p.nb33) & 7) == 0);
    const int last = p.n_kv - 1;
    // Per-lane running row pointers: a full tile costs one 64-bit add per stream, no multiplies.
    constexpr int kStrideKV = kDecodeWarps * kTileKV;
    const int kv_first = kv_begin + warp * kTileKV;
    const char* klane[2];
    const char* vlane[4];
    const char* mlane[RH];
    const char* kdlane[Q8 ? 4 : 1];  // q8_0: K rows of this lane's 4 score columns (block scales)
    if constexpr (Q8) {
#pragma unroll
        for (int i = 0; i < 4; i++) kdlane[i] = kbase + (int64_t)(kv_first + 4 * t + i) * p.nb11;
    }
#pragma unroll
    for (int nt = 0; nt < 2; nt++) klane[nt] = kbase + (int64_t)(kv_first + 4 * (g >> 1) + 2 * nt + (g & 1)) * p.nb11;
#pragma unroll
    for (int i = 0; i < 4; i++) vlane[i] = vbase + (int64_t)(kv_first + 4 * t + i) * p.nb21;
#pragma unroll
    for (int h = 0; h < RH; h++) mlane[h] = p.mask ? mrow[h] + (int64_t)(kv_first + 4 * t) * 2 : nullptr;
    const int64_t kstep = (int64_t)kStrideKV * p.nb11, vstep = (int64_t)kStrideKV * p.nb21;

    // Head sizes below D (p.Dr, f16 K/V only): this lane's 16-byte chunks that start past the real row end are never read from
    // memory; they are fed as zeros (K: a zero column adds nothing to the score; V: those output columns are never stored).
    bool kcol[NC4], vcol[NCV];
#pragma unroll
    for (int c = 0; c < NC4; c++) kcol[c] = Q8 || 8 * (t + 4 * c) < p.Dr;
#pragma unroll
    for (int c = 0; c < NCV; c++) vcol[c] = Q8 || (64 * c + 8 * g) < p.Dr;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

    // full tile (all 16 keys exist): loads through the running pointers, then advances them
    auto load_tile = [&](Tile& T) {
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
#pragma unroll
            for (int c = 0; c < NC4; c++) {
                if constexpr (!Q8) {
                    const uint4 x = kcol[c] ? ld_nc_v4(klane[nt] + (t + 4 * c) * 16) : zero4;
                    T.kf[nt][c][0] = x.x; T.kf[nt][c][1] = x.y; T.kf[nt][c][2] = x.z; T.kf[nt][c][3] = x.w;
                } else {
                    const uint2 w = ld_q8x8(klane[nt] + c * kQ8BlockBytes + 2 + 8 * t, al8);
                    T.kf[nt][c][0] = w.x; T.kf[nt][c][1] = w.y;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int c = 0; c < NCV; c++) {
                if constexpr (!Q8) {
                    const uint4 x = vcol[c] ? ld_nc_v4(vlane[i] + (64 * c + 8 * g) * 2) : zero4;
                    T.vf[i][c][0] = x.x; T.vf[i][c][1] = x.y; T.vf[i][c][2] = x.z; T.vf[i][c][3] = x.w;
                } else {
                    const char* b = vlane[i] + (2 * c + (g >> 2)) * kQ8BlockBytes;
                    const uint2 w = ld_q8x8(b + 2 + 8 * (g & 3), al8);
                    T.vf[i][c][0] = w.x; T.vf[i][c][1] = w.y;
                    T.vd[i][c] = ld_u16(b);
                }
            }
            if constexpr (Q8) {
#pragma unroll
                for (int c = 0; c < NC4; c++) T.kd[i][c] = ld_u16(kdlane[i] + c * kQ8BlockBytes);
                kdlane[i] += kstep;
            }
        }
        if (p.mask != nullptr) {
#pragma unroll
            for (int h = 0; h < RH; h++) {
                if (mask_al8) {
                    T.mk[h] = __ldg(reinterpret_cast<const uint2*>(mlane[h]));
                } else {
                    const uint32_t w0 = ld_u16(mlane[h]), w1 = ld_u16(mlane[h] + 2), w2 = ld_u16(mlane[h] + 4), w3 = ld_u16(mlane[h] + 6);
                    T.mk[h] = make_uint2(w0 | (w1 << 16), w2 | (w3 << 16));
                }
                mlane[h] += kStrideKV * 2;
            }
        }
#pragma unroll
        for (int nt = 0; nt < 2; nt++) klane[nt] += kstep;
#pragma unroll
        for (int i = 0; i < 4; i++) vlane[i] += vstep;
    };

    // ---- f16 K/V: per-warp cp.async FIFO in shared memory ----
    extern __shared__ __align__(16) unsigned char dsm[];
    using Geo = FifoGeom<D>;
    const uint32_t fifo_lane = (uint32_t)__cvta_generic_to_shared(dsm) + warp * Geo::kWarpBytes + lane * 16;
    auto issue_tile = [&](int stage) {  // copies one full tile into `stage` through the running pointers
        const uint32_t sb = fifo_lane + stage * Geo::kStageBytes;
        int j = 0;
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int c = 0; c < NC4; c++) cp_async16z(sb + (j++) * 512, klane[nt] + (t + 4 * c) * 16, kcol[c] ? 16u : 0u);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int c = 0; c < NCV; c++) cp_async16z(sb + (j++) * 512, vlane[i] + (64 * c + 8 * g) * 2, vcol[c] ? 16u : 0u);
        if (mask_al8) {
#pragma unroll
            for (int h = 0; h < RH; h++) {
                cp_async8(sb - lane * 8 + Geo::kChunks * 512 + h * 256, mlane[h]);
                mlane[h] += kStrideKV * 2;
            }
        }
#pragma unroll
        for (int nt = 0; nt < 2; nt++) klane[nt] += kstep;
#pragma unroll
        for (int i = 0; i < 4; i++) vlane[i] += vstep;
    };
    auto read_tile = [&](Tile& T, int stage, int kv0) {
        const uint32_t sb = fifo_lane + stage * Geo::kStageBytes;
        int j = 0;
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int c = 0; c < NC4; c++) {
                const uint4 x = lds128(sb + (j++) * 512);
                T.kf[nt][c][0] = x.x; T.kf[nt][c][1] = x.y; T.kf[nt][c][2] = x.z; T.kf[nt][c][3] = x.w;
            }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int c = 0; c < NCV; c++) {
                const uint4 x = lds128(sb + (j++) * 512);
                T.vf[i][c][0] = x.x; T.vf[i][c][1] = x.y; T.vf[i][c][2] = x.z; T.vf[i][c][3] = x.w;
            }
        if (p.mask != nullptr) {
#pragma unroll
            for (int h = 0; h < RH; h++) {
                if (mask_al8) {
                    T.mk[h] = lds64(sb - lane * 8 + Geo::kChunks * 512 + h * 256);
                } else {  // odd mask alignment: plain loads at consume time
                    const char* mp = mrow[h] + (int64_t)(kv0 + 4 * t) * 2;
                    const uint32_t w0 = ld_u16(mp), w1 = ld_u16(mp + 2), w2 = ld_u16(mp + 4), w3 = ld_u16(mp + 6);
                    T.mk[h] = make_uint2(w0 | (w1 << 16), w2 | (w3 << 16));
                }
            }
        }
    };

    // ragged last tile of the sequence: rows past the end are clamped (their scores are masked to -inf anyway)
    auto load_tile_clamped = [&](Tile& T, int kv0) {
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            const char* kr = kbase + (int64_t)min(kv0 + 4 * (g >> 1) + 2 * nt + (g & 1), last) * p.nb11;
#pragma unroll
            for (int c = 0; c < NC4; c++) {
                if constexpr (!Q8) {
                    const uint4 x = kcol[c] ? ld_nc_v4(kr + (t + 4 * c) * 16) : zero4;
                    T.kf[nt][c][0] = x.x; T.kf[nt][c][1] = x.y; T.kf[nt][c][2] = x.z; T.kf[nt][c][3] = x.w;
                } else {
                    const uint2 w = ld_q8x8(kr + c * kQ8BlockBytes + 2 + 8 * t, false);
                    T.kf[nt][c][0] = w.x; T.kf[nt][c][1] = w.y;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int64_t row = min(kv0 + 4 * t + i, last);
            const char* vr = vbase + row * p.nb21;
#pragma unroll
            for (int c = 0; c < NCV; c++) {
                if constexpr (!Q8) {
                    const uint4 x = vcol[c] ? ld_nc_v4(vr + (64 * c + 8 * g) * 2) : zero4;
                    T.vf[i][c][0] = x.x; T.vf[i][c][1] = x.y; T.vf[i][c][2] = x.z; T.vf[i][c][3] = x.w;
                } else {
                    const char* b = vr + (2 * c + (g >> 2)) * kQ8BlockBytes;
                    const uint2 w = ld_q8x8(b + 2 + 8 * (g & 3), false);
                    T.vf[i][c][0] = w.x; T.vf[i][c][1] = w.y;
                    T.vd[i][c] = ld_u16(b);
                }
            }
            if constexpr (Q8) {
                const char* kr = kbase + row * p.nb11;
#pragma unroll
                for (int c = 0; c < NC4; c++) T.kd[i][c] = ld_u16(kr + c * kQ8BlockBytes);
            }
        }
        if (p.mask != nullptr) {
#pragma unroll
            for (int h = 0; h < RH; h++) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; j++) w[j] = (kv0 + 4 * t + j < p.n_kv) ? ld_u16(mrow[h] + (int64_t)(kv0 + 4 * t + j) * 2) : 0u;
                T.mk[h] = make_uint2(w[0] | (w[1] << 16), w[2] | (w[3] << 16));
            }
        }
    };

    float o[NT][2 * RH];
#pragma unroll
    for (int i = 0; i < NT; i++)
#pragma unroll
        for (int e = 0; e < 2 * RH; e++) o[i][e] = 0.f;
    float m_run[RH], l_run[RH];  // l: per-lane partial sums (reduced over the quad at the end)
#pragma unroll
    for (int h = 0; h < RH; h++) { m_run[h] = -INFINITY; l_run[h] = 0.f; }

    auto compute_tile = [&](const Tile& T, int kv0) {
        if (p.dbg_mode == 1) {  // tuning aid: consume the loads, skip the maths
            uint32_t x = 0;
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int c = 0; c < NC4; c++)
#pragma unroll
                    for (int u = 0; u < (Q8 ? 2 : 4); u++) x ^= T.kf[nt][c][u];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int c = 0; c < NCV; c++)
#pragma unroll
                    for (int u = 0; u < (Q8 ? 2 : 4); u++) x ^= T.vf[i][c][u];
            o[0][0] += __uint_as_float(x & 0x3f800000u);
            return;
        }
        // ---------- S = Q K^T  (fp32) ----------
        float s[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
            for (int c = 0; c < NC4; c++) {
                uint32_t k0, k1, k2, k3;
                if constexpr (!Q8) {
                    k0 = T.kf[nt][c][0]; k1 = T.kf[nt][c][1]; k2 = T.kf[nt][c][2]; k3 = T.kf[nt][c][3];
                } else {
                    q8x4_to_h2(T.kf[nt][c][0], k0, k1);
                    q8x4_to_h2(T.kf[nt][c][1], k2, k3);
                }
                float a[4] = {0.f, 0.f, 0.f, 0.f};
                float(&acc)[4] = Q8 ? a : s[nt];
                if constexpr (RH == 2) {
                    mma_16816(acc, qa[c][0][0], qa[c][1][0], qa[c][0][1], qa[c][1][1], k0, k1);
                    mma_16816(acc, qa[c][0][2], qa[c][1][2], qa[c][0][3], qa[c][1][3], k2, k3);
                } else {
                    mma_16816_top(acc[0], acc[1], qa[c][0][0], qa[c][0][1], k0, k1);
                    mma_16816_top(acc[0], acc[1], qa[c][0][2], qa[c][0][3], k2, k3);
                }
                if constexpr (Q8) {
                    const float d0 = h_bits_to_f(T.kd[2 * nt][c]), d1 = h_bits_to_f(T.kd[2 * nt + 1][c]);
                    s[nt][0] += a[0] * d0; s[nt][1] += a[1] * d1;
                    if constexpr (RH == 2) { s[nt][2] += a[2] * d0; s[nt][3] += a[3] * d1; }
                }
            }
        }

        // ---------- scale, mask, online softmax.  Lane holds keys kv0+4t+j (j=0..3) of its rows ----------
        float pr[RH][4];
#pragma unroll
        for (int h = 0; h < RH; h++) {
            float tmax = -INFINITY;
            float mv[4] = {0.f, 0.f, 0.f, 0.f};
            if (p.mask != nullptr) {
                const float2 m01 = __half22float2(*reinterpret_cast<const __half2*>(&T.mk[h].x));
                const float2 m23 = __half22float2(*reinterpret_cast<const __half2*>(&T.mk[h].y));
                const float ms = EXT ? mslope[h] : kLog2e;
                mv[0] = m01.x * ms; mv[1] = m01.y * ms; mv[2] = m23.x * ms; mv[3] = m23.y * ms;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int kv = kv0 + 4 * t + j;
                const float sv = s[j >> 1][2 * h + (j & 1)];
                float x;
                if (EXT && p.cap_in != 0.f) x = fmaf(fa_tanh(sv * p.cap_in), p.cap_out, mv[j]);
                else x = fmaf(sv, p.scale_log2, mv[j]);
                if (kv >= lim[h]) x = -INFINITY;
                pr[h][j] = x;
                tmax = fmaxf(tmax, x);
            }
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
            const float m_new = fmaxf(m_run[h], tmax);
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            const float alpha = fast_exp2(m_run[h] - m_use);  // m_run = -inf -> 0
            float psum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                pr[h][j] = fast_exp2(pr[h][j] - m_use);
                psum += pr[h][j];
            }
            l_run[h] = l_run[h] * alpha + psum;
            m_run[h] = m_new;
            if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
                for (int i = 0; i < NT; i++) { o[i][2 * h] *= alpha; o[i][2 * h + 1] *= alpha; }
            }
        }
        const uint32_t pa0 = pack_half2(pr[0][0], pr[0][1]), pa2 = pack_half2(pr[0][2], pr[0][3]);
        uint32_t pa1 = 0u, pa3 = 0u;
        if constexpr (RH == 2) { pa1 = pack_half2(pr[1][0], pr[1][1]); pa3 = pack_half2(pr[1][2], pr[1][3]); }

        // ---------- O += P V ----------
#pragma unroll
        for (int c = 0; c < NCV; c++) {
            uint32_t vv[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if constexpr (!Q8) {
                    vv[i][0] = T.vf[i][c][0]; vv[i][1] = T.vf[i][c][1]; vv[i][2] = T.vf[i][c][2]; vv[i][3] = T.vf[i][c][3];
                } else {
                    q8x4_to_h2(T.vf[i][c][0], vv[i][0], vv[i][1]);
                    q8x4_to_h2(T.vf[i][c][1], vv[i][2], vv[i][3]);
                    const __half2 d2 = __half2half2(__ushort_as_half((unsigned short)T.vd[i][c]));
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        __half2 x = __hmul2(*reinterpret_cast<__half2*>(&vv[i][u]), d2);  // RN(d*q) per element
                        vv[i][u] = *reinterpret_cast<uint32_t*>(&x);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
                const uint32_t b0 = prmt(vv[0][j >> 1], vv[1][j >> 1], sel);
                const uint32_t b1 = prmt(vv[2][j >> 1], vv[3][j >> 1], sel);
                if constexpr (RH == 2) mma_16816(o[c * 8 + j], pa0, pa1, pa2, pa3, b0, b1);
                else mma_16816_top(o[c * 8 + j][0], o[c * 8 + j][1], pa0, pa2, b0, b1);
            }
        }
    };

    // ---- main loop ----
    {
        const int kv_fast_end = min(kv_end, p.n_kv - kTileKV + 1);  // tiles starting below this are complete in memory
        int kv0 = kv_first;
        if constexpr (!Q8 && FIFO) {
            // f16: kStages-deep cp.async FIFO; tiles i+1 and i+2 are in flight while tile i is computed
            const int n_tiles = kv_fast_end > kv_first ? (kv_fast_end - kv_first + kStrideKV - 1) / kStrideKV : 0;
#pragma unroll
            for (int s0 = 0; s0 < kStages - 1; s0++) {
                if (s0 < n_tiles) issue_tile(s0);
                cp_async_commit();
            }
            int stage = 0, stage_in = kStages - 1;
            for (int i = 0; i < n_tiles; i++) {
                if (i + kStages - 1 < n_tiles) issue_tile(stage_in);
                cp_async_commit();
                cp_async_wait<kStages - 1>();
                Tile T;
                read_tile(T, stage, kv0);
                compute_tile(T, kv0);
                kv0 += kStrideKV;
                stage = (stage + 1 == kStages) ? 0 : stage + 1;
                stage_in = (stage_in + 1 == kStages) ? 0 : stage_in + 1;
            }
            cp_async_wait<0>();
        } else {
            // q8_0: rows are only 2-byte aligned, so no cp.async: register double buffer
            Tile cur;
            bool have = kv0 < kv_fast_end;
            if (have) load_tile(cur);
            while (have) {
                const bool have_next = kv0 + kStrideKV < kv_fast_end;
                Tile nxt;
                if (have_next) load_tile(nxt);
                compute_tile(cur, kv0);
                cur = nxt;
                kv0 += kStrideKV;
                have = have_next;
            }
        }
        if (kv0 < kv_end) {  // the one ragged tile of the sequence, if this warp owns it
            Tile T;
            load_tile_clamped(T, kv0);
            compute_tile(T, kv0);
        }
    }

    // ---- quad-reduce l, then merge the warps' (m, l, O) through shared memory ----
#pragma unroll
    for (int h = 0; h < RH; h++) {
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
    }
    // merge buffers live in the dynamic shared memory: sO[w] sits inside warp w's own (drained) FIFO region
    static_assert(RLIVE * (D + 4) * 4 + 2 * RLIVE * 4 <= FifoGeom<D>::kWarpBytes, "merge buffers must fit a warp's FIFO region");
    auto sO = [&](int w, int r, int d) -> float& {
        return reinterpret_cast<float*>(dsm + w * FifoGeom<D>::kWarpBytes)[r * (D + 4) + d];
    };
    auto sM = [&](int w, int r) -> float& {
        return reinterpret_cast<float*>(dsm + w * FifoGeom<D>::kWarpBytes)[RLIVE * (D + 4) + r];
    };
    auto sL = [&](int w, int r) -> float& {
        return reinterpret_cast<float*>(dsm + w * FifoGeom<D>::kWarpBytes)[RLIVE * (D + 4) + RLIVE + r];
    };
    __shared__ int s_last;
#pragma unroll
    for (int c = 0; c < NCV; c++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int d0 = 64 * c + 16 * t + j;
            sO(warp, g, d0) = o[c * 8 + j][0];      sO(warp, g, d0 + 8) = o[c * 8 + j][1];
            if constexpr (RH == 2) { sO(warp, g + 8, d0) = o[c * 8 + j][2];  sO(warp, g + 8, d0 + 8) = o[c * 8 + j][3]; }
        }
    if (t == 0) {
#pragma unroll
        for (int h = 0; h < RH; h++) { sM(warp, g + 8 * h) = m_run[h]; sL(warp, g + 8 * h) = l_run[h]; }
    }
    __syncthreads();

    for (int idx = threadIdx.x; idx < RLIVE * D; idx += kDecodeWarps * 32) {
        const int r = idx / D, d = idx % D;
        const int R = grp * kRows + r;
        if (R >= rows_total) continue;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < kDecodeWarps; w++) M = fmaxf(M, sM(w, r));
        const float Mu = (M == -INFINITY) ? 0.f : M;
        float L = 0.f, acc = 0.f;
#pragma unroll
        for (int w = 0; w < kDecodeWarps; w++) {
            const float wt = fast_exp2(sM(w, r) - Mu);
            L += sL(w, r) * wt;
            acc += sO(w, r, d) * wt;
        }
        const int iq1 = R / p.gqa, iq2 = ik2 * p.gqa + R % p.gqa;
        const int64_t orow = ((int64_t)iq3 * p.n_q + iq1) * p.n_head + iq2;  // flash-llama.h:434
        if (p.n_splits == 1) {
            if (p.dst != nullptr) {
                const float y = L > 0.f ? acc / L : 0.f;
                if (d < p.Dr) {
                    if (p.dst_type == B200FA_TYPE_F16) reinterpret_cast<__half*>(p.dst)[orow * p.Dr + d] = __float2half_rn(y);
                    else reinterpret_cast<float*>(p.dst)[orow * p.Dr + d] = y;
                }
            } else {
                float* rec = p.part_out + orow * (D + 2);
                rec[d] = acc;
                if (d == 0) { rec[D] = M * kLn2; rec[D + 1] = L; }
            }
        } else {
            float* rec = p.part + ((int64_t)split * p.total_rows + orow) * (D + 2);
            rec[d] = acc;
            if (d == 0) { rec[D] = M * kLn2; rec[D + 1] = L; }  // m in natural-log units
        }
    }
    if (p.n_splits == 1) return;

    // ---- the last CTA of this row group to arrive merges the splits (fa_reduce, flash_row_float.h:415-472:
    //      M = max m_i, L = sum l_i e^(m_i-M), O = sum O~_i e^(m_i-M) / L — here in one parallel fp32 pass) ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int* ctr = p.counters + ((int64_t)blockIdx.z * gridDim.y + grp);
        const unsigned int old = atomicInc(ctr, (unsigned int)p.n_splits - 1);  // wraps back to 0: self-resetting
        s_last = (old == (unsigned int)p.n_splits - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int idx = threadIdx.x; idx < RLIVE * D; idx += kDecodeWarps * 32) {
        const int r = idx / D, d = idx % D;
        const int R = grp * kRows + r;
        if (R >= rows_total) continue;
        const int iq1 = R / p.gqa, iq2 = ik2 * p.gqa + R % p.gqa;
        const int64_t orow = ((int64_t)iq3 * p.n_q + iq1) * p.n_head + iq2;
        const float* rec0 = p.part + orow * (D + 2);
        const int64_t sstride = p.total_rows * (D + 2);
        float M = -INFINITY;
        for (int s = 0; s < p.n_splits; s++) M = fmaxf(M, __ldcg(rec0 + s * sstride + D));
        const float Mu = (M == -INFINITY) ? 0.f : M;
        float L = 0.f, acc = 0.f;
#pragma unroll 4
        for (int s = 0; s < p.n_splits; s++) {
            const float* rec = rec0 + s * sstride;
            const float wt = __expf(__ldcg(rec + D) - Mu);
            L += __ldcg(rec + D + 1) * wt;
            acc += __ldcg(rec + d) * wt;
        }
        if (p.dst != nullptr) {
            const float y = L > 0.f ? acc / L : 0.f;
            if (d < p.Dr) {
                if (p.dst_type == B200FA_TYPE_F16) reinterpret_cast<__half*>(p.dst)[orow * p.Dr + d] = __float2half_rn(y);
                else reinterpret_cast<float*>(p.dst)[orow * p.Dr + d] = y;
            }
        } else {
            float* out = p.part_out + orow * (D + 2);
            out[d] = acc;
            if (d == 0) { out[D] = M; out[D + 1] = L; }
        }
    }
}

// fa_combine — merges partial triples that come from OUTSIDE one kernel (the per-GPU results of a
// sequence-split call, after the all-gather).  Replaces fa_reduce<128,nw> (flash_row_float.h:415-472).
// One CTA of D threads per output row; partials laid out [part][row][D+2].
template <int D>
__global__ void __launch_bounds__(D) fa_combine(const float* __restrict__ part, int n_parts, int64_t n_rows,
                                                void* __restrict__ dst, int dst_type) {
    const int64_t row = blockIdx.x;
    const int d = threadIdx.x;
    float M = -INFINITY;
    for (int s = 0; s < n_parts; s++) M = fmaxf(M, part[((int64_t)s * n_rows + row) * (D + 2) + D]);
    const float Mu = (M == -INFINITY) ? 0.f : M;
    float L = 0.f, acc = 0.f;
#pragma unroll 4
    for (int s = 0; s < n_parts; s++) {
        const float* rec = part + ((int64_t)s * n_rows + row) * (D + 2);
        const float wt = __expf(rec[D] - Mu);
        L += rec[D + 1] * wt;
        acc += rec[d] * wt;
    }
    const float y = L > 0.f ? acc / L : 0.f;
    if (dst_type == B200FA_TYPE_F16) reinterpret_cast<__half*>(dst)[row * D + d] = __float2half_rn(y);
    else reinterpret_cast<float*>(dst)[row * D + d] = y;
}

// The same merge for the split-KV prefill: records are [part][row][D + 4] (16-byte aligned rows written by TMA stores: O~[D], m, l,
// 2 unused), dst rows are Dr wide.
template <int D>
__global__ void __launch_bounds__(D) fa_combine_pad(const float* __restrict__ part, int n_parts, int64_t n_rows, void* __restrict__ dst,
                                                    int dst_type, int Dr) {
    const int64_t row = blockIdx.x;
    const int d = threadIdx.x;
    float M = -INFINITY;
    for (int s = 0; s < n_parts; s++) M = fmaxf(M, __ldcg(part + ((int64_t)s * n_rows + row) * (D + 4) + D));
    const float Mu = (M == -INFINITY) ? 0.f : M;
    float L = 0.f, acc = 0.f;
#pragma unroll 4
    for (int s = 0; s < n_parts; s++) {
        const float* rec = part + ((int64_t)s * n_rows + row) * (D + 4);
        const float wt = __expf(__ldcg(rec + D) - Mu);
        L += __ldcg(rec + D + 1) * wt;
        acc += __ldcg(rec + d) * wt;
    }
    const float y = L > 0.f ? acc / L : 0.f;
    if (d >= Dr) return;
    if (dst_type == B200FA_TYPE_F16) reinterpret_cast<__half*>(dst)[row * Dr + d] = __float2half_rn(y);
    else reinterpret_cast<float*>(dst)[row * Dr + d] = y;
}

// ---- cross-GPU combine over peer-mapped memory (NVLink stores), no NCCL on the path ----
// Exchange buffer, identical on every rank (zero-filled once):
//   header (256 B): [0] arrival counter (monotonic: += 1 per rank per step, written by every rank)
//                   [16] scatter block counter, [17] merge block counter (self-resetting), [32] steps completed by THIS rank
//   staging   [row][D + 2] f32 : this rank's triples of the current step (written by the attention kernel)
//   gathered  [gen][rank][row][D + 2] f32 : two generations (step parity)
//   gathered, flag-in-data [gen][rank][row][D + 2] x {f32 value, u32 tag} : the fused one-kernel step (b200fa_flash_attn_seqpar)
//       publishes every float as ONE 8-byte store carrying the value and the step's tag; a reader polls the element itself until
//       the tag matches.  No fence, no arrival counter, no second NVLink round trip (the idea of NCCL's LL protocol).
// The step number lives on the device, so a step is a fixed sequence of launches that can be captured in a CUDA graph.
constexpr int kXchgHeader = 256;
// byte offset of the flag-in-data area
__host__ __device__ inline int64_t xchg_ll_offset(int world, int64_t n_floats) { return (int64_t)kXchgHeader + (1 + 2 * (int64_t)world) * n_floats * 4; }
// header words (u32): [0] arrivals of all ranks (monotonic), [16] [17] local block counters, [32] steps completed on this rank,
// [33] error flag: 1 = a wait for the peers timed out (the step's output was NOT written; b200fa_peer_status / b200fa_peer_reset),
// [34] timeout of those waits in milliseconds (0 = kXchgDefaultTimeoutMs; b200fa_peer_set_timeout)
// [35] epoch: incremented by b200fa_peer_reset (never cleared), part of the flag-in-data tag so that values of an abandoned step
//      sequence can never be taken for current ones
constexpr int kXchgErrWord = 33, kXchgTimeoutWord = 34, kXchgEpochWord = 35;
// tag of a step: never 0 (the buffer starts zero-filled), never equal to the tag two steps earlier (the slot's previous contents)
__device__ __forceinline__ unsigned int xchg_ll_tag(const unsigned int* hdr, unsigned int step) { return 0x80000000u | ((hdr[kXchgEpochWord] & 0x7fu) << 24) | (step & 0xffffffu); }
// one aligned 64-bit access per element: value in the low word, tag in the high word (single-copy atomic)
__device__ __forceinline__ void st_ll(void* p, float v, unsigned int tag) {
    const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const void* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return make_uint2((unsigned int)w, (unsigned int)(w >> 32));
}
constexpr unsigned int kXchgDefaultTimeoutMs = 4000;

__device__ __forceinline__ unsigned long long xchg_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// One thread waits (system-scope acquire loads) until all ranks' arrivals of step `want / n` are in.  Returns false after the
// exchange's timeout — a missing, late or failed rank — having raised the header's error flag: the caller then SKIPS its merge and
// the kernel ends normally (no trap: a trap leaves a sticky context error on every rank for what may be a few seconds of skew).
__device__ __forceinline__ bool xchg_wait_arrivals(unsigned int* hdr, unsigned int want);

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ bool xchg_wait_arrivals(unsigned int* hdr, unsigned int want) {
    if ((int)(ld_acquire_sys(hdr) - want) >= 0) return true;
    const unsigned int ms = hdr[kXchgTimeoutWord] ? hdr[kXchgTimeoutWord] : kXchgDefaultTimeoutMs;
    const unsigned long long t0 = xchg_now_ns(), limit = (unsigned long long)ms * 1000000ull;
    while ((int)(ld_acquire_sys(hdr) - want) < 0) {
        if (xchg_now_ns() - t0 > limit) {
            atomicExch(hdr + kXchgErrWord, 1u);
            return false;
        }
    }
    return true;
}
__global__ void fa_xchg_set_word(char* xchg, int word, unsigned int value) { reinterpret_cast<unsigned int*>(xchg)[word] = value; }
// header back to its initial state (steps, arrivals, counters, error flag), keeping the configured timeout
__global__ void fa_xchg_reset(char* xchg) {
    unsigned int* hdr = reinterpret_cast<unsigned int*>(xchg);
    if (threadIdx.x < kXchgHeader / 4 && threadIdx.x != kXchgTimeoutWord && threadIdx.x != kXchgEpochWord) hdr[threadIdx.x] = 0u;
    if (threadIdx.x == kXchgEpochWord) hdr[kXchgEpochWord] = (hdr[kXchgEpochWord] + 1u) & 0x7fu;
}

// Copies this rank's staged triples into slot `rank` of every rank's gathered area (its own included) with plain stores —
// over NVLink for the peers — then publishes them: system-scope fence + one atomic increment of every arrival counter.
__global__ void __launch_bounds__(256) fa_scatter_signal(char* const* __restrict__ peers, int rank, int world, int64_t n_floats) {
    char* own = peers[rank];
    unsigned int* hdr = reinterpret_cast<unsigned int*>(own);
    const unsigned int step = hdr[32] + 1;                       // this rank's step in progress
    const int64_t gen_off = (int64_t)(step & 1u) * world * n_floats;
    const float* src = reinterpret_cast<const float*>(own + kXchgHeader);
    for (int pr = 0; pr < world; pr++) {
        float* dst = reinterpret_cast<float*>(peers[pr] + kXchgHeader) + n_floats + gen_off + (int64_t)rank * n_floats;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_floats; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicInc(hdr + 16, gridDim.x - 1) == gridDim.x - 1) {  // the last block to finish signals every rank
            __threadfence_system();
            for (int pr = 0; pr < world; pr++) atomicAdd_system(reinterpret_cast<unsigned int*>(peers[pr]), 1u);
        }
    }
}

// fa_combine that first waits until all `n_parts` ranks have published this rank's current step into its exchange buffer.
template <int D>
__global__ void __launch_bounds__(D) fa_combine_wait(char* __restrict__ xchg, int n_parts, int64_t n_rows, void* __restrict__ dst, int dst_type) {
    unsigned int* hdr = reinterpret_cast<unsigned int*>(xchg);
    const unsigned int step = hdr[32] + 1;
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = xchg_wait_arrivals(hdr, step * (unsigned int)n_parts) ? 1 : 0;
    __syncthreads();
    if (!s_ok) return;  // timed out: error flag raised, dst untouched, the step is not counted (b200fa_peer_reset)
    const int64_t n_floats = n_rows * (D + 2);
    const float* part = reinterpret_cast<const float*>(xchg + kXchgHeader) + n_floats + (int64_t)(step & 1u) * n_parts * n_floats;
    const int64_t row = blockIdx.x;
    const int d = threadIdx.x;
    float M = -INFINITY;
    for (int s = 0; s < n_parts; s++) M = fmaxf(M, __ldcv(part + ((int64_t)s * n_rows + row) * (D + 2) + D));
    const float Mu = (M == -INFINITY) ? 0.f : M;
    float L = 0.f, acc = 0.f;
#pragma unroll 4
    for (int s = 0; s < n_parts; s++) {
        const float* rec = part + ((int64_t)s * n_rows + row) * (D + 2);
        const float wt = __expf(__ldcv(rec + D) - Mu);
        L += __ldcv(rec + D + 1) * wt;
        acc += __ldcv(rec + d) * wt;
    }
    const float y = L > 0.f ? acc / L : 0.f;
    if (dst_type == B200FA_TYPE_F16) reinterpret_cast<__half*>(dst)[row * D + d] = __float2half_rn(y);
    else reinterpret_cast<float*>(dst)[row * D + d] = y;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicInc(hdr + 17, gridDim.x - 1) == gridDim.x - 1) hdr[32] = step;  // last block: the step is complete on this rank
    }
}

}  // namespace b200fa

// decode_mma.cuh — split-KV "rows16" kernel: the bandwidth-bound decode path, and the general path for
// any shape the tcgen05 prefill kernel does not take.
//
// Replaces the reference's flash_attn_row<128,8,2,256> / flash_attn_row_fast (flash_row_float.h:4-413) and,
// with n_splits > 1, is followed by fa_combine (fa_reduce, flash_row_float.h:415-472).
//
// One CTA = one KV split of one (kv head, batch) for one group of up to 16 output rows, where a "row" is a
// (query position, q head of the GQA group) pair — all rows of a group share the K/V stream, so K/V is
// read from HBM once per GQA group (the reference re-reads it per q head, flash_row_float.h:19,58).
//
// Data path: each warp streams 16-key tiles straight from global memory into mma.sync fragments with
// 128-bit ld.global.nc.L1::no_allocate loads — no shared-memory staging, no shuffles for operands.
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

// 8 int8 at a 2-byte-aligned address -> two 32-bit words
__device__ __forceinline__ uint2 ld_q8x8(const char* p) {
    const uint16_t* s = reinterpret_cast<const uint16_t*>(p);
    uint32_t a = __ldg(s), b = __ldg(s + 1), c = __ldg(s + 2), d = __ldg(s + 3);
    return make_uint2(a | (b << 16), c | (d << 16));
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

__device__ __forceinline__ float ld_h(const char* p) { return __half2float(__ldg(reinterpret_cast<const __half*>(p))); }

template <int D, int KV_TYPE>
__global__ void __launch_bounds__(kDecodeWarps * 32)
fa_rows16_splitkv(const __grid_constant__ FaParams p) {
    static_assert(D == 64 || D == 128, "head size");
    constexpr int NC4 = D / 32;   // 16-byte chunks per lane per K row (c' loop) == q8_0 blocks per row
    constexpr int NCV = D / 64;   // 64-wide halves of a V row
    constexpr int NT = D / 8;     // output n-tiles
    constexpr bool Q8 = (KV_TYPE == B200FA_TYPE_Q8_0);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int split = blockIdx.x, grp = blockIdx.y;
    const int ik2 = blockIdx.z % p.n_head_kv, iq3 = blockIdx.z / p.n_head_kv;
    const int ik3 = iq3 / p.rk3;
    const int rows_total = p.n_q * p.gqa;  // rows sharing this kv head

    // ---- the two rows this lane owns (g and g+8) ----
    int iq1r[2], iq2r[2];
    bool rvalid[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int R = grp * kRows + g + 8 * h;
        rvalid[h] = R < rows_total;
        const int Rc = rvalid[h] ? R : 0;
        iq1r[h] = Rc / p.gqa;
        iq2r[h] = ik2 * p.gqa + Rc % p.gqa;
    }

    // ---- Q fragments (f16; an f32 Q is rounded like the reference does, flash-llama.h:80) ----
    uint32_t qa[NC4][2][4];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const char* qrow = p.q + iq1r[h] * p.nb01 + iq2r[h] * p.nb02 + (int64_t)iq3 * p.nb03;
#pragma unroll
        for (int c = 0; c < NC4; c++) {
            const int e0 = 8 * (t + 4 * c);
            if (!rvalid[h]) {
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
    int64_t vis[2];  // per-row local visibility limit (exclusive) under the causal flag
#pragma unroll
    for (int h = 0; h < 2; h++) vis[h] = p.causal ? (int64_t)iq1r[h] + p.causal_off - p.kv_pos0 + 1 : (int64_t)p.n_kv;

    const char* kbase = p.k + (int64_t)ik2 * p.nb12 + (int64_t)ik3 * p.nb13;
    const char* vbase = p.v + (int64_t)ik2 * p.nb22 + (int64_t)ik3 * p.nb23;
    const char* mrow[2] = {p.mask ? p.mask + iq1r[0] * p.nb31 : nullptr, p.mask ? p.mask + iq1r[1] * p.nb31 : nullptr};

    float o[NT][4];
#pragma unroll
    for (int i = 0; i < NT; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};  // per-lane partial sums (reduced over the quad at the end)

    for (int kv0 = kv_begin + warp * kTileKV; kv0 < kv_end; kv0 += kDecodeWarps * kTileKV) {
        const int last = p.n_kv - 1;
        // ---------- K: two n8-tiles, lane loads the row rho(g) ----------
        uint32_t kf[2][NC4][4];
        float kd[4][NC4];  // q8_0: block scales of this lane's 4 score columns (keys kv0+4t+j)
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            const int row = min(kv0 + 4 * (g >> 1) + 2 * nt + (g & 1), last);
            const char* kr = kbase + (int64_t)row * p.nb11;
#pragma unroll
            for (int c = 0; c < NC4; c++) {
                if constexpr (!Q8) {
                    const uint4 x = ld_nc_v4(kr + (t + 4 * c) * 16);
                    kf[nt][c][0] = x.x; kf[nt][c][1] = x.y; kf[nt][c][2] = x.z; kf[nt][c][3] = x.w;
                } else {
                    const uint2 w = ld_q8x8(kr + c * kQ8BlockBytes + 2 + 8 * t);
                    q8x4_to_h2(w.x, kf[nt][c][0], kf[nt][c][1]);
                    q8x4_to_h2(w.y, kf[nt][c][2], kf[nt][c][3]);
                }
            }
        }
        if constexpr (Q8) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const char* kr = kbase + (int64_t)min(kv0 + 4 * t + j, last) * p.nb11;
#pragma unroll
                for (int c = 0; c < NC4; c++) kd[j][c] = ld_h(kr + c * kQ8BlockBytes);
            }
        }
        // ---------- V: lane loads keys kv0+4t+i, head dims 64c+8g.. ----------
        uint32_t vf[4][NCV][4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const char* vr = vbase + (int64_t)min(kv0 + 4 * t + i, last) * p.nb21;
#pragma unroll
            for (int c = 0; c < NCV; c++) {
                if constexpr (!Q8) {
                    const uint4 x = ld_nc_v4(vr + (64 * c + 8 * g) * 2);
                    vf[i][c][0] = x.x; vf[i][c][1] = x.y; vf[i][c][2] = x.z; vf[i][c][3] = x.w;
                } else {
                    const int blk = 2 * c + (g >> 2);
                    const char* b = vr + blk * kQ8BlockBytes;
                    const uint2 w = ld_q8x8(b + 2 + 8 * (g & 3));
                    const __half dh = __ldg(reinterpret_cast<const __half*>(b));
                    const __half2 d2 = __half2half2(dh);
                    uint32_t r[4];
                    q8x4_to_h2(w.x, r[0], r[1]);
                    q8x4_to_h2(w.y, r[2], r[3]);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        __half2 x = __hmul2(*reinterpret_cast<__half2*>(&r[u]), d2);  // RN(d*q) per element
                        vf[i][c][u] = *reinterpret_cast<uint32_t*>(&x);
                    }
                }
            }
        }

        // ---------- S = Q K^T  (fp32) ----------
        float s[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
            for (int c = 0; c < NC4; c++) {
                if constexpr (!Q8) {
                    mma_16816(s[nt], qa[c][0][0], qa[c][1][0], qa[c][0][1], qa[c][1][1], kf[nt][c][0], kf[nt][c][1]);
                    mma_16816(s[nt], qa[c][0][2], qa[c][1][2], qa[c][0][3], qa[c][1][3], kf[nt][c][2], kf[nt][c][3]);
                } else {
                    float a[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_16816(a, qa[c][0][0], qa[c][1][0], qa[c][0][1], qa[c][1][1], kf[nt][c][0], kf[nt][c][1]);
                    mma_16816(a, qa[c][0][2], qa[c][1][2], qa[c][0][3], qa[c][1][3], kf[nt][c][2], kf[nt][c][3]);
                    s[nt][0] += a[0] * kd[2 * nt][c];     s[nt][1] += a[1] * kd[2 * nt + 1][c];
                    s[nt][2] += a[2] * kd[2 * nt][c];     s[nt][3] += a[3] * kd[2 * nt + 1][c];
                }
            }
        }

        // ---------- scale, mask, online softmax.  Lane holds keys kv0+4t+j (j=0..3) of rows g, g+8 ----------
        float pr[2][4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float tmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int kv = kv0 + 4 * t + j;
                float x = s[j >> 1][2 * h + (j & 1)] * p.scale_log2;
                if (mrow[h] != nullptr && kv < p.n_kv) x += ld_mask(mrow[h], kv) * kLog2e;
                if (kv >= kv_end || (int64_t)kv >= vis[h]) x = -INFINITY;
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
        const uint32_t pa0 = pack_half2(pr[0][0], pr[0][1]), pa1 = pack_half2(pr[1][0], pr[1][1]);
        const uint32_t pa2 = pack_half2(pr[0][2], pr[0][3]), pa3 = pack_half2(pr[1][2], pr[1][3]);

        // ---------- O += P V ----------
#pragma unroll
        for (int c = 0; c < NCV; c++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
                const uint32_t b0 = prmt(vf[0][c][j >> 1], vf[1][c][j >> 1], sel);
                const uint32_t b1 = prmt(vf[2][c][j >> 1], vf[3][c][j >> 1], sel);
                mma_16816(o[c * 8 + j], pa0, pa1, pa2, pa3, b0, b1);
            }
        }
    }

    // ---- quad-reduce l, then merge the warps' (m, l, O) through shared memory ----
#pragma unroll
    for (int h = 0; h < 2; h++) {
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
    }
    __shared__ float sO[kDecodeWarps][kRows][D + 4];
    __shared__ float sM[kDecodeWarps][kRows], sL[kDecodeWarps][kRows];
#pragma unroll
    for (int c = 0; c < NCV; c++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int d0 = 64 * c + 16 * t + j;
            sO[warp][g][d0] = o[c * 8 + j][0];      sO[warp][g][d0 + 8] = o[c * 8 + j][1];
            sO[warp][g + 8][d0] = o[c * 8 + j][2];  sO[warp][g + 8][d0 + 8] = o[c * 8 + j][3];
        }
    if (t == 0) {
        sM[warp][g] = m_run[0]; sM[warp][g + 8] = m_run[1];
        sL[warp][g] = l_run[0]; sL[warp][g + 8] = l_run[1];
    }
    __syncthreads();

    for (int idx = threadIdx.x; idx < kRows * D; idx += kDecodeWarps * 32) {
        const int r = idx / D, d = idx % D;
        const int R = grp * kRows + r;
        if (R >= rows_total) continue;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < kDecodeWarps; w++) M = fmaxf(M, sM[w][r]);
        const float Mu = (M == -INFINITY) ? 0.f : M;
        float L = 0.f, acc = 0.f;
#pragma unroll
        for (int w = 0; w < kDecodeWarps; w++) {
            const float wt = fast_exp2(sM[w][r] - Mu);
            L += sL[w][r] * wt;
            acc += sO[w][r][d] * wt;
        }
        const int iq1 = R / p.gqa, iq2 = ik2 * p.gqa + R % p.gqa;
        const int64_t orow = ((int64_t)iq3 * p.n_q + iq1) * p.n_head + iq2;  // flash-llama.h:434
        if (p.write_final) {
            const float y = L > 0.f ? acc / L : 0.f;
            if (p.dst_type == B200FA_TYPE_F16) reinterpret_cast<__half*>(p.dst)[orow * D + d] = __float2half_rn(y);
            else reinterpret_cast<float*>(p.dst)[orow * D + d] = y;
        } else {
            float* rec = p.part + ((int64_t)split * p.total_rows + orow) * (D + 2);
            rec[d] = acc;
            if (d == 0) { rec[D] = M * kLn2; rec[D + 1] = L; }  // m in natural-log units
        }
    }
}

// fa_combine — merges split-KV partial triples.  Replaces fa_reduce<128,nw> (flash_row_float.h:415-472):
// M = max m_i, L = sum l_i e^(m_i-M), O = sum O~_i e^(m_i-M) / L, but in one parallel pass over fp32 state
// (the reference scans blocks serially in thread 0 and folds each head dim serially, in f16).
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

// Same merge, but the result stays a partial triple (used when a sequence-split rank itself ran several
// CTA-level splits and must hand ONE triple per row to the cross-GPU combine).
template <int D>
__global__ void __launch_bounds__(D) fa_combine_to_partial(const float* __restrict__ part, int n_parts, int64_t n_rows,
                                                           float* __restrict__ out) {
    const int64_t row = blockIdx.x;
    const int d = threadIdx.x;
    float M = -INFINITY;
    for (int s = 0; s < n_parts; s++) M = fmaxf(M, part[((int64_t)s * n_rows + row) * (D + 2) + D]);
    const float Mu = (M == -INFINITY) ? 0.f : M;
    float L = 0.f, acc = 0.f;
    for (int s = 0; s < n_parts; s++) {
        const float* rec = part + ((int64_t)s * n_rows + row) * (D + 2);
        const float wt = __expf(rec[D] - Mu);
        L += rec[D + 1] * wt;
        acc += rec[d] * wt;
    }
    float* o = out + row * (D + 2);
    o[d] = acc;
    if (d == 0) { o[D] = M; o[D + 1] = L; }
}

}  // namespace b200fa

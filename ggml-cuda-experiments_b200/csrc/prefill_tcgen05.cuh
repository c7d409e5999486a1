// prefill_tcgen05.cuh — the tensor-bound prefill path: QK^T and P·V as tcgen05.mma tiles accumulating in
// TMEM, K/V streamed by TMA through an mbarrier pipeline, online softmax one thread per query row.
//
// Replaces the reference's flash_attn_ext_f16<128,16,128> (flash-llama.h:5-438; WMMA 16x16x16, f16
// accumulators, per-warp K/V fragment loads from global, smem round-trips per S tile).
//
// CTA = one 128-row query tile of one head (M = 128).  192 threads:
//   warps 0-3  softmax + correction + epilogue: thread r owns query row r (TMEM lane r), so row max / sum
//              need no shuffles at all;
//   warp  4    TMA producer (one elected lane): Q once, then K_j / V_j tiles into a 2-stage ring;
//   warp  5    MMA issuer (one elected lane): S_j = Q K_j^T (SS, both K-major), O += P_j V_j (TS: P from
//              TMEM, V MN-major from smem).
// TMEM (512 columns): S0 [0,128) S1 [128,256) O [256,384).  P_j (f16) overwrites the first 64 columns
// of its S buffer.  S is double-buffered so S_{j+1} is computed while the softmax of tile j runs.
// O is rescaled lazily: only when a row max grows by more than 2^8 (P stays within f16 range).
// KV tiles are classified full / mixed / skip — from the causal flag arithmetically, or from a one-pass
// scan of the mask tensor (the reference detects all -inf blocks at run time, flash-llama.h:276-278) —
// so the mask is only read on mixed (diagonal) tiles and masked tiles cost nothing.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace b200fa {

constexpr int PF_BM = 128, PF_BN = 128, PF_D = 128;
constexpr int PF_THREADS = 192;
constexpr uint32_t PF_TILE_BYTES = 128 * 128 * 2;  // one 128x128 f16 tile
constexpr uint32_t PF_TMEM_COLS = 512;
constexpr uint32_t PF_TM_S = 0, PF_TM_O = 256;     // S buffer b at PF_TM_S + 128*b
constexpr float PF_RESCALE_THRESHOLD = 8.0f;       // log2 units

struct __align__(1024) PfShared {
    uint8_t q[PF_TILE_BYTES];     // [2 k-blocks][128 rows][64 d]  128B-swizzled, K-major
    uint8_t k[2][PF_TILE_BYTES];  // same layout, rows = keys
    uint8_t v[2][PF_TILE_BYTES];  // [2 d-halves][128 keys][64 d]  128B-swizzled, MN-major B operand
    uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full[2], pv_done;
    uint32_t tmem_base;
};

struct PfArgs {
    const uint8_t* cls;        // [n_q_tiles][n_kv_tiles] tile classes from the mask scan, or null
    int n_q_tiles, n_kv_tiles;
    float inv_scale;           // 1/scale (mask values are folded into raw scores)
    unsigned long long* dbg;   // timeout codes (mapped host memory), may be null
    float* dump;               // diagnostics: S of the first tile + final O of CTA `dump_cta`, may be null
    int dump_cta;
};

// 0 = every element visible, 1 = mixed (mask / causal edge / ragged tail), 2 = nothing visible
__device__ __forceinline__ int pf_tile_class(const FaParams& p, const PfArgs& a, int qt, int j) {
    const int kv0 = j * PF_BN;
    if (kv0 >= p.n_kv) return 2;
    int c = (kv0 + PF_BN > p.n_kv) ? 1 : 0;
    if (p.causal) {
        const int64_t q0 = (int64_t)qt * PF_BM;
        const int64_t first_lim = q0 + p.causal_off;
        const int64_t last_lim = min(q0 + PF_BM - 1, (int64_t)p.n_q - 1) + p.causal_off;
        if (kv0 > last_lim) return 2;
        if (kv0 + PF_BN - 1 > first_lim) c = 1;
    } else if (a.cls != nullptr) {
        const int u = a.cls[(int64_t)qt * a.n_kv_tiles + j];
        if (u == 2) return 2;
        if (u == 1) c = 1;
    } else if (p.mask != nullptr) {
        c = 1;
    }
    return c;
}
__device__ __forceinline__ int pf_next_tile(const FaParams& p, const PfArgs& a, int qt, int j) {
    for (; j < a.n_kv_tiles; j++)
        if (pf_tile_class(p, a, qt, j) != 2) return j;
    return -1;
}

__global__ void __launch_bounds__(PF_THREADS, 1)
fa_prefill_tcgen05(const __grid_constant__ FaParams p, const __grid_constant__ PfArgs a,
                   const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV) {
    using namespace ptx;
    extern __shared__ uint8_t pf_smem_raw[];
    PfShared& sm = *reinterpret_cast<PfShared*>((reinterpret_cast<uintptr_t>(pf_smem_raw) + 1023) & ~(uintptr_t)1023);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // heavy (late) query tiles first: under a causal mask tile qt costs qt+1 KV tiles
    const int per_tile = p.n_head * p.n_batch;
    const int qt = a.n_q_tiles - 1 - (int)(blockIdx.x / per_tile);
    const int iq2 = (int)(blockIdx.x % per_tile) % p.n_head;
    const int iq3 = (int)(blockIdx.x % per_tile) / p.n_head;
    const int ik2 = iq2 / p.gqa, ik3 = iq3 / p.rk3;
    const int q0 = qt * PF_BM;

    if (threadIdx.x == 0) {
        mbar_init(&sm.q_full, 1);
        for (int s = 0; s < 2; s++) {
            mbar_init(&sm.k_full[s], 1); mbar_init(&sm.k_empty[s], 1);
            mbar_init(&sm.v_full[s], 1); mbar_init(&sm.v_empty[s], 1);
            mbar_init(&sm.s_full[s], 1); mbar_init(&sm.p_full[s], 128);
        }
        mbar_init(&sm.pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 4) {
        tmem_alloc(&sm.tmem_base, PF_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV);
            mbar_arrive_expect_tx(&sm.q_full, PF_TILE_BYTES);
            tma_load_4d(sm.q, &tmQ, &sm.q_full, 0, q0, iq2, iq3);
            tma_load_4d(sm.q + PF_TILE_BYTES / 2, &tmQ, &sm.q_full, 64, q0, iq2, iq3);
            int it = 0;
            for (int j = pf_next_tile(p, a, qt, 0); j >= 0; j = pf_next_tile(p, a, qt, j + 1), it++) {
                const int st = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&sm.k_empty[st], ph ^ 1, a.dbg, 1);
                mbar_arrive_expect_tx(&sm.k_full[st], PF_TILE_BYTES);
                tma_load_4d(sm.k[st], &tmK, &sm.k_full[st], 0, j * PF_BN, ik2, ik3);
                tma_load_4d(sm.k[st] + PF_TILE_BYTES / 2, &tmK, &sm.k_full[st], 64, j * PF_BN, ik2, ik3);
                mbar_wait(&sm.v_empty[st], ph ^ 1, a.dbg, 2);
                mbar_arrive_expect_tx(&sm.v_full[st], PF_TILE_BYTES);
                tma_load_4d(sm.v[st], &tmV, &sm.v_full[st], 0, j * PF_BN, ik2, ik3);
                tma_load_4d(sm.v[st] + PF_TILE_BYTES / 2, &tmV, &sm.v_full[st], 64, j * PF_BN, ik2, ik3);
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_f16(PF_BM, PF_BN, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_f16(PF_BM, PF_D, 0, 1);
            const uint32_t q_addr = smem_u32(sm.q);
            auto issue_qk = [&](int it_s) {
                const int st = it_s & 1;
                mbar_wait(&sm.k_full[st], (it_s >> 1) & 1, a.dbg, 3);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sm.k[st]);
                const uint32_t d_tmem = tmem + PF_TM_S + 128u * (it_s & 1);
#pragma unroll
                for (int ks = 0; ks < 8; ks++) {
                    const uint32_t off = (ks >> 2) * (PF_TILE_BYTES / 2) + (ks & 3) * 32;
                    mma_ss(d_tmem, make_smem_desc_sw128(q_addr + off, 16, 1024), make_smem_desc_sw128(k_addr + off, 16, 1024),
                           idesc_qk, ks > 0);
                }
                tc_commit(&sm.k_empty[st]);
                tc_commit(&sm.s_full[it_s & 1]);
            };
            mbar_wait(&sm.q_full, 0, a.dbg, 4);
            int it = 0;
            int j = pf_next_tile(p, a, qt, 0);
            if (j >= 0) issue_qk(0);
            while (j >= 0) {
                const int jn = pf_next_tile(p, a, qt, j + 1);
                if (jn >= 0) issue_qk(it + 1);  // S_{it+1} runs on the tensor pipe while softmax(it) is in flight
                const int b = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&sm.p_full[b], ph, a.dbg, 5);
                mbar_wait(&sm.v_full[b], ph, a.dbg, 6);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(sm.v[b]);
                const uint32_t p_tmem = tmem + PF_TM_S + 128u * b;
#pragma unroll
                for (int ks = 0; ks < 8; ks++)
                    mma_ts(tmem + PF_TM_O, p_tmem + ks * 8, make_smem_desc_sw128(v_addr + ks * 2048, PF_TILE_BYTES / 2, 1024),
                           idesc_pv, (it > 0 || ks > 0) ? 1u : 0u);
                tc_commit(&sm.v_empty[b]);
                tc_commit(&sm.pv_done);
                j = jn;
                it++;
            }
        }
    } else {
        // ===================== softmax / correction / epilogue: thread = query row =====================
        const int r = threadIdx.x;  // 0..127
        const int qrow = q0 + r;
        const bool row_valid = qrow < p.n_q;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        const float c = p.scale_log2;
        // rows past n_q read the last real mask row (their results are never stored): every lane of a warp takes the
        // same path, so the .sync.aligned tcgen05 instructions below always see a converged warp
        const char* mrow = (p.mask != nullptr && !p.causal) ? p.mask + (int64_t)min(qrow, p.n_q - 1) * p.nb31 : nullptr;
        const bool mask_vec = (((uintptr_t)p.mask | (uintptr_t)p.nb31) & 15) == 0;
        const int64_t vis = p.causal ? (int64_t)qrow + p.causal_off : (int64_t)p.n_kv;  // last visible key (inclusive)
        const bool dumping = a.dump != nullptr && (int)blockIdx.x == a.dump_cta;

        float m_ref = -INFINITY, l = 0.f;
        int it = 0;
        for (int j = pf_next_tile(p, a, qt, 0); j >= 0; j = pf_next_tile(p, a, qt, j + 1), it++) {
            const int cls = pf_tile_class(p, a, qt, j);
            const int b = it & 1;
            mbar_wait(&sm.s_full[b], (it >> 1) & 1, a.dbg, 7);
                __syncwarp();
            tc_fence_after();
            uint32_t s[4][32];
#pragma unroll
            for (int q4 = 0; q4 < 4; q4++) tmem_ld32(trow + PF_TM_S + 128u * b + 32u * q4, s[q4]);
            tmem_wait_ld();
            if (dumping && it == 0) {
#pragma unroll
                for (int q4 = 0; q4 < 4; q4++)
#pragma unroll
                    for (int i = 0; i < 32; i++) a.dump[r * 128 + q4 * 32 + i] = __uint_as_float(s[q4][i]);
            }
            if (cls == 1) {
                const int kv0 = j * PF_BN;
#pragma unroll
                for (int q4 = 0; q4 < 4; q4++) {
                    if (mrow != nullptr) {
                        if (mask_vec && kv0 + PF_BN <= p.n_kv) {
#pragma unroll
                            for (int v8 = 0; v8 < 4; v8++) {
                                const uint4 mv = *reinterpret_cast<const uint4*>(mrow + (int64_t)(kv0 + q4 * 32 + v8 * 8) * 2);
                                const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                                for (int u = 0; u < 4; u++) {
                                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[u]));
                                    s[q4][v8 * 8 + 2 * u] = __float_as_uint(__uint_as_float(s[q4][v8 * 8 + 2 * u]) + f.x * a.inv_scale);
                                    s[q4][v8 * 8 + 2 * u + 1] = __float_as_uint(__uint_as_float(s[q4][v8 * 8 + 2 * u + 1]) + f.y * a.inv_scale);
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; i++) {
                                const int kv = kv0 + q4 * 32 + i;
                                if (kv < p.n_kv) s[q4][i] = __float_as_uint(__uint_as_float(s[q4][i]) + ld_mask(mrow, kv) * a.inv_scale);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        const int kv = kv0 + q4 * 32 + i;
                        if (kv >= p.n_kv || (int64_t)kv > vis) s[q4][i] = 0xff800000u;  // -inf
                    }
                }
            }
            __syncwarp();
            // ---- row max (raw scores; scale > 0 on this path) ----
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int q4 = 0; q4 < 4; q4++)
#pragma unroll
                for (int i = 0; i < 32; i++) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(s[q4][i]));
            const float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * c;
            const bool need = m_tile > m_ref + PF_RESCALE_THRESHOLD;  // also true for the first finite max
            // Every waiter must observe EVERY phase of an mbarrier in order (a parity wait that skips a phase is satisfied by
            // the phase before it), so pv_done(it-1) is waited for in every iteration: here when O has to be rescaled,
            // otherwise just before this tile's P is published.
            bool saw_pv = false;
            if (it > 0 && __any_sync(0xffffffffu, need)) {
                // O currently holds sum_{tiles < it}; PV_{it-1} must have landed before we touch it
                mbar_wait(&sm.pv_done, (it - 1) & 1, a.dbg, 8);
                saw_pv = true;
                __syncwarp();
                tc_fence_after();
                const float alpha = need ? fast_exp2(m_ref - m_tile) : 1.f;
                l *= alpha;
#pragma unroll
                for (int q4 = 0; q4 < 4; q4++) {
                    uint32_t o[32];
                    tmem_ld32(trow + PF_TM_O + 32u * q4, o);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st32(trow + PF_TM_O + 32u * q4, o);
                }
                tmem_wait_st();
            }
            if (need) m_ref = m_tile;
            const float m_eff = (m_ref == -INFINITY) ? 0.f : m_ref;
            // ---- P = exp2(s*c - m), row sum, pack to f16, store over S ----
            float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const int q4 = h * 2 + (i >> 4), e = (i & 15) * 2;
                    const float p0 = fast_exp2(fmaf(__uint_as_float(s[q4][e]), c, -m_eff));
                    const float p1 = fast_exp2(fmaf(__uint_as_float(s[q4][e + 1]), c, -m_eff));
                    ls[i & 3] += p0 + p1;
                    pk[i] = pack_half2(p0, p1);
                }
                tmem_st32(trow + PF_TM_S + 128u * b + 32u * h, pk);
            }
            l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
            if (it > 0 && !saw_pv) {
                mbar_wait(&sm.pv_done, (it - 1) & 1, a.dbg, 10);
                __syncwarp();
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&sm.p_full[b]);
        }

        // ---- epilogue: O / l -> dst[(iq3*n_q + q)*n_head + head][D]   (flash-llama.h:434) ----
        const int64_t orow = ((int64_t)iq3 * p.n_q + qrow) * p.n_head + iq2;
        if (it > 0) {
            mbar_wait(&sm.pv_done, (it - 1) & 1, a.dbg, 9);
                __syncwarp();
            tc_fence_after();
        }
        const float inv_l = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
        for (int q4 = 0; q4 < 4; q4++) {
            uint32_t o[32];
            if (it > 0) {
                tmem_ld32(trow + PF_TM_O + 32u * q4, o);
                tmem_wait_ld();
            } else {
#pragma unroll
                for (int i = 0; i < 32; i++) o[i] = 0u;
            }
            if (dumping) {
#pragma unroll
                for (int i = 0; i < 32; i++) a.dump[128 * 128 + r * 128 + q4 * 32 + i] = __uint_as_float(o[i]);
                if (q4 == 0) { a.dump[2 * 128 * 128 + r] = l; a.dump[2 * 128 * 128 + 128 + r] = m_ref; }
            }
            if (row_valid) {
                if (p.dst_type == B200FA_TYPE_F32) {
                    float4* d4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dst) + orow * PF_D + q4 * 32);
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        d4[i] = make_float4(__uint_as_float(o[4 * i]) * inv_l, __uint_as_float(o[4 * i + 1]) * inv_l,
                                            __uint_as_float(o[4 * i + 2]) * inv_l, __uint_as_float(o[4 * i + 3]) * inv_l);
                } else {
                    uint4* d4 = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.dst) + orow * PF_D + q4 * 32);
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        d4[i] = make_uint4(pack_half2(__uint_as_float(o[8 * i]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l),
                                           pack_half2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l),
                                           pack_half2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l),
                                           pack_half2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem, PF_TMEM_COLS);
    }
}

// One pass over the mask: class of every 128x128 tile (0 all zero, 2 all -inf, 1 anything else).
__global__ void __launch_bounds__(256) fa_mask_classify(const char* __restrict__ mask, int64_t nb31, int n_q, int n_kv,
                                                        int n_kv_tiles, uint8_t* __restrict__ cls) {
    const int j = blockIdx.x, qt = blockIdx.y;
    int has_zero = 0, has_ninf = 0, has_other = 0;
    for (int idx = threadIdx.x; idx < PF_BM * (PF_BN / 2); idx += blockDim.x) {
        const int r = qt * PF_BM + idx / (PF_BN / 2), col = j * PF_BN + (idx % (PF_BN / 2)) * 2;
        if (r >= n_q) continue;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            if (col + e >= n_kv) continue;
            const uint16_t bits = *reinterpret_cast<const uint16_t*>(mask + (int64_t)r * nb31 + (int64_t)(col + e) * 2);
            if ((bits & 0x7fffu) == 0) has_zero = 1;
            else if (bits == 0xfc00u) has_ninf = 1;
            else has_other = 1;
        }
    }
    has_zero = __syncthreads_or(has_zero);
    has_ninf = __syncthreads_or(has_ninf);
    has_other = __syncthreads_or(has_other);
    if (threadIdx.x == 0) cls[(int64_t)qt * n_kv_tiles + j] = (has_other || (has_zero && has_ninf)) ? 1 : (has_ninf ? 2 : 0);
}

// f32 Q (any ggml strides) -> dense f16 [batch][head][q][D]; same rounding as the reference (flash-llama.h:80)
__global__ void __launch_bounds__(256) fa_q_to_f16(const char* __restrict__ q, __half* __restrict__ out, int D, int n_q, int n_head,
                                                   int64_t total_rows, int64_t nb01, int64_t nb02, int64_t nb03) {
    const int chunks = D / 8;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_rows * chunks) return;
    const int64_t row = idx / chunks;
    const int ch = (int)(idx % chunks);
    const int iq1 = (int)(row % n_q), iq2 = (int)((row / n_q) % n_head);
    const int64_t iq3 = row / ((int64_t)n_q * n_head);
    const float4* src = reinterpret_cast<const float4*>(q + iq1 * nb01 + iq2 * nb02 + iq3 * nb03 + ch * 32);
    const float4 x = src[0], y = src[1];
    *reinterpret_cast<uint4*>(out + row * D + ch * 8) =
        make_uint4(pack_half2(x.x, x.y), pack_half2(x.z, x.w), pack_half2(y.x, y.y), pack_half2(y.z, y.w));
}

// ---------------- host side ----------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// f16 tensor [head_dim][rows][heads][batch] with byte strides nb1..nb3; box = 64 x box_rows x 1 x 1, 128B swizzle
inline bool make_tile_map(CUtensorMap* m, const void* base, int64_t rows, int64_t heads, int64_t batch, int64_t nb1, int64_t nb2,
                          int64_t nb3, int box_rows = 128, int head_dim = PF_D) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)head_dim, (cuuint64_t)rows, (cuuint64_t)heads, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)nb1, (cuuint64_t)nb2, (cuuint64_t)nb3};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct PfDebug {
    unsigned long long* dbg = nullptr;
    float* dump = nullptr;
    int dump_cta = 0;
};
inline PfDebug& pf_debug() {
    static PfDebug d;
    return d;
}

inline int launch_prefill_tcgen05(const FaParams& p, char* ws, size_t qf16_bytes, size_t cls_bytes, int sm_count,
                                  cudaStream_t st, int* launches) {
    (void)sm_count; (void)cls_bytes;
    if (p.D != PF_D || p.kv_type != B200FA_TYPE_F16 || !(p.scale > 0.f)) return B200FA_ERR_UNSUPPORTED;
    int n = 0;
    const void* qbase = p.q;
    int64_t qnb1 = p.nb01, qnb2 = p.nb02, qnb3 = p.nb03;
    if (p.q_type == B200FA_TYPE_F32) {
        __half* q16 = reinterpret_cast<__half*>(ws);
        const int64_t work = p.total_rows * (PF_D / 8);
        fa_q_to_f16<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(p.q, q16, PF_D, p.n_q, p.n_head, p.total_rows, p.nb01, p.nb02,
                                                                   p.nb03);
        n++;
        qbase = q16;
        qnb1 = PF_D * 2; qnb2 = (int64_t)p.n_q * PF_D * 2; qnb3 = (int64_t)p.n_head * p.n_q * PF_D * 2;
    }
    PfArgs a{};
    a.n_q_tiles = (p.n_q + PF_BM - 1) / PF_BM;
    a.n_kv_tiles = (p.n_kv + PF_BN - 1) / PF_BN;
    a.inv_scale = 1.0f / p.scale;
    a.dbg = pf_debug().dbg; a.dump = pf_debug().dump; a.dump_cta = pf_debug().dump_cta;
    if (p.mask != nullptr && !p.causal) {
        uint8_t* cls = reinterpret_cast<uint8_t*>(ws + qf16_bytes);
        fa_mask_classify<<<dim3(a.n_kv_tiles, a.n_q_tiles), 256, 0, st>>>(p.mask, p.nb31, p.n_q, p.n_kv, a.n_kv_tiles, cls);
        n++;
        a.cls = cls;
    }
    CUtensorMap tq, tk, tv;
    if (!make_tile_map(&tq, qbase, p.n_q, p.n_head, p.n_batch, qnb1, qnb2, qnb3)) return B200FA_ERR_CUDA;
    if (!make_tile_map(&tk, p.k, p.n_kv, p.n_head_kv, p.n_batch_kv, p.nb11, p.nb12, p.nb13)) return B200FA_ERR_CUDA;
    if (!make_tile_map(&tv, p.v, p.n_kv, p.n_head_kv, p.n_batch_kv, p.nb21, p.nb22, p.nb23)) return B200FA_ERR_CUDA;
    constexpr size_t smem_bytes = sizeof(PfShared) + 1024;
    static thread_local bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        if (cudaFuncSetAttribute(fa_prefill_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess)
            return B200FA_ERR_CUDA;
        attr_set[dev] = true;
    }
    const unsigned grid = (unsigned)((int64_t)a.n_q_tiles * p.n_head * p.n_batch);
    fa_prefill_tcgen05<<<grid, PF_THREADS, smem_bytes, st>>>(p, a, tq, tk, tv);
    n++;
    if (launches) *launches = n;
    return cudaGetLastError() == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}

}  // namespace b200fa

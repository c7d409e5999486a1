// prefill_persistent2.cuh — the persistent prefill kernel with TWO softmax threads per query row.
// MEASURED NEGATIVE RESULT (round 1): 417 vs 787 TFLOP/s on C3, 560 vs 1064 on 8K causal.  Not the default; selected with
// B200FA_PREFILL=p2.  The launcher of both persistent variants lives at the end of this file.
//
// The loop of prefill_persistent.cuh is softmax-paced: one thread per row means two softmax warps per SM sub-partition,
// too few to hide the latencies of TMEM loads, exp chains and barrier waits (issue slots ~45 % used).  Here every 64-key
// half tile of a row is split between two threads (32 columns each) that live in different warps (a warp can only reach
// its own 32-lane quarter of TMEM, so the pair shares the quarter, not the warp): 16 softmax warps, four per
// sub-partition.  The two threads of a row exchange their partial row maxima through shared memory (one named barrier per
// query tile and half tile) and keep partial row sums that are added in the epilogue; everything else — the producer
// warp, the two single-thread MMA issuers, the barriers' phase rules — is shared with prefill_persistent.cuh.
// CTA = 640 threads: warps 0-15 softmax (tile = w >> 3, column side = (w >> 2) & 1, TMEM quarter = w & 3),
// warp 16 producer, warps 17/18 issuers, warp 19 idle (completes the warpgroup that donates registers).
#pragma once
#include "prefill_persistent.cuh"

namespace b200fa {

constexpr int PP2_THREADS = 640;
// setmaxnreg only redistributes the CTA's own launch allocation (640 x 96): the donor warpgroup frees (96-40)*128 = 7168
// registers, the four softmax warpgroups take (104-96)*512 = 4096 of them.
constexpr int PP2_REGS_SOFTMAX = 104, PP2_REGS_OTHER = 40;

__device__ __forceinline__ void bar_tile(int t) { asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory"); }

template <int POLY>
__global__ void __launch_bounds__(PP2_THREADS, 1)
fa_prefill_persistent2(const __grid_constant__ FaParams p, const __grid_constant__ PpArgs pa,
                       const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO) {
    using namespace ptx;
    extern __shared__ __align__(1024) uint8_t pp_smem_raw[];  // 128B-swizzled TMA tiles need 1024-byte alignment; no slack to spare
    if ((smem_u32(pp_smem_raw) & 1023u) != 0) __trap();
    PpShared& sm = *reinterpret_cast<PpShared*>(pp_smem_raw);
    const PfArgs& a = pa.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&sm.q_full[s], 1); mbar_init(&sm.q_empty[s], 1);
            mbar_init(&sm.k_full[s], 1); mbar_init(&sm.k_empty[s], 2);  // one arrival per MMA issuer
            mbar_init(&sm.v_full[s], 1); mbar_init(&sm.v_empty[s], 2);
            mbar_init(&sm.pv_done[s], 1); mbar_init(&sm.o_free[s], 8);  // o_free, p_full: one arrival per softmax warp of the tile
            for (int h = 0; h < 2; h++) { mbar_init(&sm.s_full[s][h], 1); mbar_init(&sm.p_full[s][h], 8); }
            mbar_init(&sm.item_full[s], 1); mbar_init(&sm.item_empty[s], 18);  // 2 issuers + 16 softmax warps
        }
        fence_barrier_init();
    }
    if (warp == 16) {
        tmem_alloc(&sm.tmem_base, PF_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp >= 16) {
#ifndef PP2_NO_SETMAXNREG
        reg_dec<PP2_REGS_OTHER>();
#endif
        if (warp == 16) pp_producer_role(p, pa, sm, lane, tmQ, tmK, tmV);
        else if (warp <= 18) pp_issuer_role(p, pa, sm, tmem, warp - 17);
    } else {
        // ===================== softmax / correction / epilogue: two threads per query row =====================
#ifndef PP2_NO_SETMAXNREG
        reg_inc<PP2_REGS_SOFTMAX>();
#endif
        const int t = warp >> 3;                 // query tile
        const int side = (warp >> 2) & 1;        // which 32 of the 64 columns of a half tile / which 64 of the 128 output columns
        const int qd = warp & 3;                 // TMEM quarter = rows 32*qd .. +31 of the tile
        const int r = qd * 32 + lane;            // row within the tile = TMEM lane
        const uint32_t trow = tmem + ((uint32_t)(qd * 32) << 16);
        const uint32_t tS = trow + PF_TM_S + 128u * t, tO = trow + PF_TM_O + 128u * t + 64u * side;
        const float c = p.scale_log2;
        const bool mask_vec = (((uintptr_t)p.mask | (uintptr_t)p.nb31) & 15) == 0;
        // The pair's exchange area lives in the staging buffer of the side-0 warp: [parity][side][32 rows] floats = 512 bytes.
        // It is idle whenever the staging buffer is in use (tile barriers separate the two uses).
        float* xchg = reinterpret_cast<float*>(sm.stage[t * 4 + qd][0]);
        uint4* stg = reinterpret_cast<uint4*>(sm.stage[t * 4 + qd][side]);  // this warp's 2 KB of epilogue staging
        const bool causal = p.causal != 0 || (pa.detect_causal != 0 && __ldcg(pa.counters + 2) == 0u);
        int it_tot = 0;   // tiles done over all items (phase counter of s_full[t][h])
        int g_tot = 0;    // half tiles done over all items (phase counter of pv_done[t])
        for (int k = 0;; k++) {
            const int slot = k & 1;
            mbar_wait(&sm.item_full[slot], (k >> 1) & 1, a.dbg, 24);
            __syncwarp();
            const int w = sm.it_w[slot];
            if (w < 0) break;
            const int j_lo = sm.it_jlo[slot], j_hi = sm.it_jhi[slot];
            int qt0, iq2, iq3;
            pp_decode_item(p, a, w, qt0, iq2, iq3);
            const int qt = qt0 + t;
            const int q0 = qt * PF_BM;
            const int qrow = q0 + r;
            const char* mrow = (p.mask != nullptr && !causal) ? p.mask + (int64_t)min(qrow, p.n_q - 1) * p.nb31 : nullptr;
            const int64_t vis = causal ? (int64_t)qrow + p.causal_off : (int64_t)p.n_kv;  // last visible key (inclusive)
            float m_ref = -INFINITY, l = 0.f;   // l: this thread's half of the row sum
            int g = 0;
            bar_tile(t);  // both sides have left the previous item's epilogue: the exchange area is free again
            for (int j = j_lo; j < j_hi; j++) {
                const int cls = (sm.cls2[slot][j] >> (2 * t)) & 3;
                if (cls == 2) continue;
#pragma unroll 1
                for (int h = 0; h < 2; h++, g++, g_tot++) {
                    mbar_wait(&sm.s_full[t][h], it_tot & 1, a.dbg, 7);
                    __syncwarp();
                    tc_fence_after();
                    const uint32_t tSh = tS + 64u * h;
                    uint32_t s[32];
                    tmem_ld32(tSh + 32u * side, s);
                    tmem_wait_ld();
                    if (cls == 1) {
                        const int kv0 = j * PF_BN + 64 * h + 32 * side;
                        const int lim = (int)min((int64_t)(p.n_kv - 1), vis) - kv0;  // last visible column of this block for this row
                        if (mrow != nullptr) {
                            if (mask_vec && kv0 + 32 <= p.n_kv) {
#pragma unroll
                                for (int v8 = 0; v8 < 4; v8++) {
                                    const uint4 mv = *reinterpret_cast<const uint4*>(mrow + (int64_t)(kv0 + v8 * 8) * 2);
                                    const uint32_t wd[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                                    for (int e = 0; e < 4; e++) {
                                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wd[e]));
                                        s[v8 * 8 + 2 * e] = __float_as_uint(__uint_as_float(s[v8 * 8 + 2 * e]) + f.x * a.inv_scale);
                                        s[v8 * 8 + 2 * e + 1] = __float_as_uint(__uint_as_float(s[v8 * 8 + 2 * e + 1]) + f.y * a.inv_scale);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; i++) {
                                    const int kv = kv0 + i;
                                    if (kv < p.n_kv) s[i] = __float_as_uint(__uint_as_float(s[i]) + ld_mask(mrow, kv) * a.inv_scale);
                                }
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 32; i++)
                            if (i > lim) s[i] = 0xff800000u;  // -inf: past the sequence end or the causal limit
                        __syncwarp();
                    }
                    // ---- row max: own 32 columns, then the partner's through shared memory ----
                    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                    for (int i = 0; i < 32; i++) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(s[i]));
                    const float m_own = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
                    float* xb = xchg + (g & 1) * 64;
                    xb[side * 32 + lane] = m_own;
                    bar_tile(t);
                    const float m_tile = fmaxf(m_own, xb[(side ^ 1) * 32 + lane]) * c;
                    const bool need = m_tile > m_ref + PF_RESCALE_THRESHOLD;  // identical in both threads of the row
                    bool saw_pv = (g == 0);
                    if (g > 0 && __any_sync(0xffffffffu, need)) {
                        mbar_wait(&sm.pv_done[t], (g_tot - 1) & 1, a.dbg, 8);
                        saw_pv = true;
                        __syncwarp();
                        tc_fence_after();
                        const float alpha = need ? fast_exp2(m_ref - m_tile) : 1.f;
                        l *= alpha;
#pragma unroll
                        for (int q2 = 0; q2 < 2; q2++) {  // this thread's 64 of the 128 output columns
                            uint32_t o[32];
                            tmem_ld32(tO + 32u * q2, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st32(tO + 32u * q2, o);
                        }
                        tmem_wait_st();
                    }
                    if (need) m_ref = m_tile;
                    const float m_eff = (m_ref == -INFINITY) ? 0.f : m_ref;
                    // ---- P = exp2(s*c - m), partial row sum, pack to f16, store over this thread's 16 columns of S^h ----
                    const uint64_t cc = pack2(c, c), nm = pack2(-m_eff, -m_eff);
                    uint64_t ls2[4] = {0ull, 0ull, 0ull, 0ull};
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        float x0, x1;
                        const uint64_t x2 = fma2(pack2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), cc, nm);
                        if (POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == 1) {
                            exp2_poly2(x2, x0, x1);
                        } else {
                            unpack2(x2, x0, x1);
                            x0 = fast_exp2(x0); x1 = fast_exp2(x1);
                        }
                        s[2 * i] = __float_as_uint(x0);
                        s[2 * i + 1] = __float_as_uint(x1);
                    }
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float p0 = __uint_as_float(s[2 * i]), p1 = __uint_as_float(s[2 * i + 1]);
                        ls2[i & 3] = add2(ls2[i & 3], pack2(p0, p1));
                        pk[i] = pack_half2(p0, p1);
                    }
                    tmem_st16(tSh + 16u * side, pk);
                    {
                        float a0, a1, b0, b1;
                        unpack2(add2(ls2[0], ls2[1]), a0, a1); unpack2(add2(ls2[2], ls2[3]), b0, b1);
                        l += (a0 + a1) + (b0 + b1);
                    }
                    if (!saw_pv) {
                        mbar_wait(&sm.pv_done[t], (g_tot - 1) & 1, a.dbg, 10);
                        __syncwarp();
                    }
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.p_full[t][h]);
                }
                it_tot++;
            }

            // ---- epilogue: O / l -> dst[(iq3*n_q + q)*n_head + head][D]   (flash-llama.h:434); this thread: 64 columns ----
            uint32_t o[2][32];
            if (g > 0) {
                mbar_wait(&sm.pv_done[t], (g_tot - 1) & 1, a.dbg, 9);
                __syncwarp();
                tc_fence_after();
                tmem_ld32(tO, o[0]);
                tmem_ld32(tO + 32u, o[1]);
                tmem_wait_ld();
            } else {
#pragma unroll
                for (int i = 0; i < 32; i++) { o[0][i] = 0u; o[1][i] = 0u; }
            }
            tc_fence_before();
            __syncwarp();
            if (g > 0 && lane == 0) mbar_arrive(&sm.o_free[t]);  // O_t is in registers: the next item may accumulate into it
            // the row sum is the pair's: exchange the halves (the area is free: the last max exchange was read before the
            // barrier of the half tile after it, or — for g == 1 — is in the other parity slot)
            float* xb = xchg + 128;  // a third slot, not used by the max exchange
            xb[side * 32 + lane] = l;
            bar_tile(t);
            const float l_tot = l + xb[(side ^ 1) * 32 + lane];
            bar_tile(t);  // everyone has read the exchange area: side 0's staging (which contains it) may be written
            const float inv_l = l_tot > 0.f ? 1.f / l_tot : 0.f;
            if (qt < a.n_q_tiles) {
                // 2 KB of staging per warp: 32 rows x 64 bytes, 16-byte chunk c of row i at chunk position c ^ ((i >> 1) & 3)
                // (= the TMA 64-byte swizzle); one lane hands each box to the TMA engine, which clips rows past n_q.
                const bool f32out = p.dst_type == B200FA_TYPE_F32;
                const int row0 = q0 + qd * 32;
                const int n_pass = f32out ? 4 : 2;        // this thread's 64 columns: 16 f32 or 32 f16 columns = 64 bytes per pass
#pragma unroll
                for (int pass = 0; pass < 4; pass++) {
                    if (pass >= n_pass) break;
                    uint4 piece[4];
                    if (f32out) {
                        const int q2 = pass >> 1, b = (pass & 1) * 16;
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            piece[i] = make_uint4(__float_as_uint(__uint_as_float(o[q2][b + 4 * i]) * inv_l), __float_as_uint(__uint_as_float(o[q2][b + 4 * i + 1]) * inv_l),
                                                  __float_as_uint(__uint_as_float(o[q2][b + 4 * i + 2]) * inv_l), __float_as_uint(__uint_as_float(o[q2][b + 4 * i + 3]) * inv_l));
                    } else {
                        const int q2 = pass;
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            piece[i] = make_uint4(pack_half2(__uint_as_float(o[q2][8 * i]) * inv_l, __uint_as_float(o[q2][8 * i + 1]) * inv_l),
                                                  pack_half2(__uint_as_float(o[q2][8 * i + 2]) * inv_l, __uint_as_float(o[q2][8 * i + 3]) * inv_l),
                                                  pack_half2(__uint_as_float(o[q2][8 * i + 4]) * inv_l, __uint_as_float(o[q2][8 * i + 5]) * inv_l),
                                                  pack_half2(__uint_as_float(o[q2][8 * i + 6]) * inv_l, __uint_as_float(o[q2][8 * i + 7]) * inv_l));
                    }
                    if (lane == 0) bulk_wait_read0();   // the previous pass's box has left the staging buffer
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; i++) stg[lane * 4 + (i ^ ((lane >> 1) & 3))] = piece[i];
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(&tmO, stg, side * 64 + pass * (f32out ? 16 : 32), iq2, row0, iq3);
                        bulk_commit();
                    }
                }
                if (lane == 0) bulk_wait_read0();
                __syncwarp();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.item_empty[slot]);
        }
    }

    if (warp < 16 && lane == 0) bulk_wait_all0();  // this lane's output boxes have been written
    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        tmem_dealloc(tmem, PF_TMEM_COLS);
    }
    if (threadIdx.x == 0) {  // the last CTA out leaves both counters at zero for the next call
        __threadfence();
        if (atomicAdd(pa.counters + 1, 1u) == gridDim.x - 1) {
            pa.counters[0] = 0u;
            pa.counters[1] = 0u;
            pa.counters[2] = 0u;
            __threadfence();
        }
    }
}

// n_seg > 1: split-KV prefill — `part` ([n_seg][total_rows][D + 4] f32, in the workspace) receives the segments' partial rows and
// fa_combine_pad merges them into dst in a second launch.
inline int launch_prefill_persistent(const FaParams& p, char* ws, size_t qf16_bytes, unsigned int* counters, int sm_count,
                                     cudaStream_t st, int* launches, int n_seg = 1, float* part = nullptr) {
    if (p.D != PF_D || p.kv_type != B200FA_TYPE_F16 || !(p.scale > 0.f) || p.n_kv > PP_MAX_KV_TILES * PF_BN) return B200FA_ERR_UNSUPPORTED;
    // Head sizes below 128 (Dr, a multiple of 8) run on the same 128-wide kernel: the tensor maps describe rows of Dr elements,
    // so TMA zero-fills columns Dr..127 of every Q/K/V tile on the way in (zeros add nothing to Q.K^T, and the P.V columns they
    // produce are never stored: the dst tensor map clips them on the way out).
    const int Dr = p.Dr;
    int n = 0;
    const void* qbase = p.q;
    int64_t qnb1 = p.nb01, qnb2 = p.nb02, qnb3 = p.nb03;
    if (p.q_type == B200FA_TYPE_F32) {
        __half* q16 = reinterpret_cast<__half*>(ws);
        const int64_t work = p.total_rows * (Dr / 8);
        fa_q_to_f16<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(p.q, q16, Dr, p.n_q, p.n_head, p.total_rows, p.nb01, p.nb02,
                                                                   p.nb03);
        n++;
        qbase = q16;
        qnb1 = Dr * 2; qnb2 = (int64_t)p.n_q * Dr * 2; qnb3 = (int64_t)p.n_head * p.n_q * Dr * 2;
    }
    PpArgs pa{};
    PfArgs& a = pa.f;
    a.n_q_tiles = (p.n_q + PF_BM - 1) / PF_BM;
    a.n_kv_tiles = (p.n_kv + PF_BN - 1) / PF_BN;
    a.n_q_pairs = (a.n_q_tiles + 1) / 2;
    a.inv_scale = 1.0f / p.scale;
    a.dbg = pf_debug().dbg; a.dump = pf_debug().dump; a.dump_cta = pf_debug().dump_cta;
    if (p.mask != nullptr && !p.causal) {
        uint8_t* cls = reinterpret_cast<uint8_t*>(ws + qf16_bytes);
        fa_mask_classify<<<dim3(a.n_kv_tiles, a.n_q_tiles), 256, 0, st>>>(p.mask, p.nb31, p.n_q, p.n_kv, a.n_kv_tiles, cls, counters + 2);
        n++;
        a.cls = cls;
        pa.detect_causal = 1;
    }
    if (n_seg < 1 || (n_seg > 1 && (part == nullptr || p.Dr != PF_D))) return B200FA_ERR_INVALID;
    pa.n_seg = n_seg;
    pa.n_items = a.n_q_pairs * p.n_head * p.n_batch * n_seg;
    pa.counters = counters;
    CUtensorMap tq, tk, tv;
    if (!make_tile_map(&tq, qbase, p.n_q, p.n_head, p.n_batch, qnb1, qnb2, qnb3, 128, Dr)) return B200FA_ERR_CUDA;
    if (!make_tile_map(&tk, p.k, p.n_kv, p.n_head_kv, p.n_batch_kv, p.nb11, p.nb12, p.nb13, 128, Dr)) return B200FA_ERR_CUDA;
    if (!make_tile_map(&tv, p.v, p.n_kv, p.n_head_kv, p.n_batch_kv, p.nb21, p.nb22, p.nb23, 128, Dr)) return B200FA_ERR_CUDA;
    CUtensorMap to;
    {   // dst [batch][n_q][n_head][D]: box = 64 bytes x 1 head x 32 rows, 64-byte swizzle (the epilogue's staging layout)
        PFN_encodeTiled enc = get_encode_tiled();
        if (!enc) return B200FA_ERR_CUDA;
        const bool f32o = n_seg > 1 || p.dst_type == B200FA_TYPE_F32;
        const cuuint64_t es = f32o ? 4 : 2;
        const cuuint64_t rowlen = n_seg > 1 ? PF_D + 4 : Dr;  // partial records: O~[D], m, l, 2 unused; dim 3 = segment * n_batch + batch
        cuuint64_t dims[4] = {rowlen, (cuuint64_t)p.n_head, (cuuint64_t)p.n_q, (cuuint64_t)p.n_batch * n_seg};
        cuuint64_t strides[3] = {rowlen * es, (cuuint64_t)p.n_head * rowlen * es, (cuuint64_t)p.n_q * p.n_head * rowlen * es};
        cuuint32_t box[4] = {(cuuint32_t)(64 / es), 1, 32, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (enc(&to, f32o ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, n_seg > 1 ? (void*)part : p.dst, dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return B200FA_ERR_CUDA;
    }
    constexpr size_t smem_bytes = sizeof(PpShared);
    static_assert(smem_bytes <= 227 * 1024, "prefill shared memory budget");
    static const int poly = getenv("B200FA_POLY") ? atoi(getenv("B200FA_POLY")) : 2;  // default: every 2nd pair on the FMA pipes
    // two softmax threads per row: an experiment that measured HALF the speed of one thread per row (the per-half-tile
    // named barrier + shared-memory max exchange costs more than the extra warps hide); kept selectable for comparison
    static const bool two_env = getenv("B200FA_PREFILL") && !strcmp(getenv("B200FA_PREFILL"), "p2");
    const bool ext = p.cap_in != 0.f || p.alibi_nhl2 != 0;  // ext2 score modifiers: their own instantiation
    const bool two = two_env && Dr == PF_D && !ext && n_seg == 1;
    auto kern = ext ? fa_prefill_persistent<2, true>
              : two ? (poly == 0 ? fa_prefill_persistent2<0> : fa_prefill_persistent2<2>)
                    : (poly == 0 ? fa_prefill_persistent<0> : (poly == 3 ? fa_prefill_persistent<3> : (poly == 4 ? fa_prefill_persistent<4> : fa_prefill_persistent<2>)));
    static thread_local bool attr_set[64][7] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int ai = ext ? 6 : two ? (poly == 0 ? 4 : 5) : (poly == 0 ? 0 : (poly == 3 ? 2 : (poly == 4 ? 3 : 1)));
    if (dev >= 0 && dev < 64 && !attr_set[dev][ai]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess)
            return B200FA_ERR_CUDA;
        attr_set[dev][ai] = true;
    }
    const unsigned grid = (unsigned)(pa.n_items < sm_count ? pa.n_items : sm_count);
    // programmatic dependent launch: the prologue (barrier init, TMEM allocation) may run while the previous kernel of the stream —
    // the mask classifier, the Q conversion, or the caller's own kernel — is still draining; the kernel waits before its first
    // global access.  The experimental two-thread variant keeps the plain launch.
    static const bool no_pdl = getenv("B200FA_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(two ? 640 : PF_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (no_pdl || two) ? 0 : 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p, pa, tq, tk, tv, to);
    n++;
    if (le == cudaSuccess && n_seg > 1) {
        fa_combine_pad<PF_D><<<(unsigned)p.total_rows, PF_D, 0, st>>>(part, n_seg, p.total_rows, p.dst, p.dst_type, Dr);
        n++;
        if (cudaGetLastError() != cudaSuccess) return B200FA_ERR_CUDA;
    }
    if (launches) *launches = n;
    return le == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}

}  // namespace b200fa

// prefill_tcgen05.cuh — the tensor-bound prefill path: QK^T and P·V as tcgen05.mma tiles accumulating in
// TMEM, K/V streamed by TMA through an mbarrier ring, online softmax one thread per query row.
//
// Replaces the reference's flash_attn_ext_f16<128,16,128> (flash-llama.h:5-438; WMMA 16x16x16, f16
// accumulators, per-warp K/V fragment loads from global, smem round-trips per S tile).
//
// CTA = TWO 128-row query tiles of one head (256 query rows) sharing every K/V tile: the per-SM L2->smem
// traffic per tensor-core cycle is half that of a one-tile CTA (which would need more than the chip's L2
// bandwidth at full tensor rate), and while one tile's softmax runs the tensor pipe works on the other.
// 384 threads:
//   warps 0-3   softmax / lazy correction / epilogue of tile 0: thread r owns query row r (TMEM lane r), so row
//               max / sum need no shuffles at all;
//   warps 4-7   the same for tile 1;
//   warp  8     TMA producer (one elected lane): Q0, Q1 once, then K_j / V_j tiles into a 2-stage ring;
//   warp  9     MMA issuer (one elected lane): S_t = Q_t K_j^T (SS, both K-major), O_t += P_t V_j (TS: P from
//               TMEM, V MN-major from smem), ordered  PV0(j) QK0(j+1) PV1(j) QK1(j+1)  so each softmax group
//               has a full tile's worth of tensor work to hide behind;
//   warps 10-11 idle (they complete the third warpgroup, which gives its registers to the softmax warpgroups
//               with setmaxnreg).
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512).  P_t (f16) overwrites the first 64
// columns of S_t.  O is rescaled lazily: only when a row max grows by more than 2^8 (P stays within f16 range).
// KV tiles are classified full / mixed / skip per 128x128 tile — from the causal flag arithmetically, or from
// a one-pass scan of the mask tensor (the reference detects all -inf blocks at run time,
// flash-llama.h:276-278) — so the mask is only read on mixed (diagonal) tiles and masked tiles cost nothing.
// Rule kept throughout: every waiter of an mbarrier observes every phase of it, in order.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace b200fa {

constexpr int PF_BM = 128, PF_BN = 128, PF_D = 128;
constexpr int PF_THREADS = 384;
constexpr uint32_t PF_TILE_BYTES = 128 * 128 * 2;  // one 128x128 f16 tile
constexpr uint32_t PF_TMEM_COLS = 512;
constexpr uint32_t PF_TM_S = 0, PF_TM_O = 256;     // tile t: S at PF_TM_S + 128*t, O at PF_TM_O + 128*t
constexpr float PF_RESCALE_THRESHOLD = 8.0f;       // log2 units
#ifndef B200FA_PF_REGS_OTHER
#define B200FA_PF_REGS_OTHER 72
#endif
// 2 x 128 x 216 + 128 x 72 registers = 63 K of the 64 K file (80 would fill it exactly: the launch fails).  At 40 the MMA issuers
// spilled their descriptors inside the issue loop.
constexpr int PF_REGS_SOFTMAX = 216, PF_REGS_OTHER = B200FA_PF_REGS_OTHER;
constexpr int PF_STAGGER_CYCLES = 500;
constexpr int PF_MAX_KV_TILES = 4096;              // schedule capacity: n_kv <= 524288 on this path

struct __align__(1024) PfShared {
    uint8_t q[2][PF_TILE_BYTES];  // per Q tile: [2 k-blocks][128 rows][64 d]  128B-swizzled, K-major
    uint8_t k[2][PF_TILE_BYTES];  // ring stage: same layout, rows = keys
    uint8_t v[2][PF_TILE_BYTES];  // ring stage: [2 d-halves][128 keys][64 d]  128B-swizzled, MN-major B operand
    uint64_t q_full[2], k_full[2], k_empty[2], v_full[2], v_empty[2], pv_done[2];
    uint64_t s_full[2][2], p_full[2][2];  // [query tile][half of the KV tile]
    uint32_t tmem_base;
    int j_lo, j_hi;               // KV tiles outside [j_lo, j_hi) are invisible to both query tiles
    uint8_t cls2[PF_MAX_KV_TILES];  // per KV tile: class for query tile 0 (bits 0-1) and 1 (bits 2-3)
};

struct PfArgs {
    const uint8_t* cls;        // [cls_q_tiles][n_kv_tiles] classes of the 128-position x 128-key mask tiles from the mask scan, or null
    int n_q_tiles, n_kv_tiles; // 128-row / 128-key tiles
    int n_q_pairs;             // CTAs per (head, batch)
    // GQA packing (persistent kernel only; pack_sh = 0 elsewhere): the 2^pack_sh q heads of a KV head share one 128-row tile,
    // row r = position r >> pack_sh, head r & (2^pack_sh - 1); a tile then covers q_rows = 128 >> pack_sh query positions.
    int pack_sh;
    int q_rows;                // query positions per tile (128 unless packed)
    int cls_q_tiles;           // rows of the class table: ceil(n_q / 128)
    float inv_scale;           // 1/scale (mask values are folded into raw scores)
    unsigned long long* dbg;   // timeout codes (mapped host memory), may be null
    float* dump;               // diagnostics, may be null
    int dump_cta;
};

// 0 = every element visible, 1 = mixed (mask / causal edge / ragged tail), 2 = nothing visible.  qt = 128-row tile index.
// slice: index of the (head, batch) mask slice whose classes apply (fa_mask_slice; 0 for the reference's shared mask)
__device__ __forceinline__ int pf_tile_class(const FaParams& p, const PfArgs& a, int qt, int j, bool causal, int slice = 0);
__device__ __forceinline__ int pf_tile_class(const FaParams& p, const PfArgs& a, int qt, int j) { return pf_tile_class(p, a, qt, j, p.causal != 0); }
__device__ __forceinline__ int pf_tile_class(const FaParams& p, const PfArgs& a, int qt, int j, bool causal, int slice) {
    const int kv0 = j * PF_BN;
    if (kv0 >= p.n_kv || qt >= a.n_q_tiles) return 2;
    int c = (kv0 + PF_BN > p.n_kv) ? 1 : 0;
    if (causal) {
        const int64_t q0 = (int64_t)qt * a.q_rows;
        const int64_t first_lim = q0 + p.causal_off;
        const int64_t last_lim = min(q0 + a.q_rows - 1, (int64_t)p.n_q - 1) + p.causal_off;
        if (kv0 > last_lim) return 2;
        if (kv0 + PF_BN - 1 > first_lim) c = 1;
    } else if (a.cls != nullptr) {
        // (a packed tile lies inside one 128-position block of the scan: the block's class is exact for "nothing visible" /
        //  "everything visible" and conservative — mixed — otherwise)
        const int u = a.cls[((int64_t)slice * a.cls_q_tiles + ((int64_t)qt * a.q_rows) / PF_BM) * a.n_kv_tiles + j];
        if (u == 2) return 2;
        if (u == 1) c = 1;
    } else if (p.mask != nullptr) {
        c = 1;
    }
    return c;
}
// next KV tile >= j that tile qt needs, or -1
__device__ __forceinline__ int pf_next_tile(const FaParams& p, const PfArgs& a, int qt, int j) {
    for (; j < a.n_kv_tiles; j++)
        if (pf_tile_class(p, a, qt, j) != 2) return j;
    return -1;
}
// next KV tile >= j that either query tile of the pair needs, or -1
__device__ __forceinline__ int pf_next_union(const FaParams& p, const PfArgs& a, int qt0, int j) {
    for (; j < a.n_kv_tiles; j++)
        if (pf_tile_class(p, a, qt0, j) != 2 || pf_tile_class(p, a, qt0 + 1, j) != 2) return j;
    return -1;
}

// packed fp32 pairs (FFMA2 / FADD2 on sm_100)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// 2^x for a pair, on the FMA/ALU pipes instead of the 16-per-clock MUFU: round-to-nearest split x = n + r with the
// 1.5*2^23 magic add, degree-3 minimax polynomial for 2^r on [-0.5, 0.5] (max relative error 7.5e-5, below the
// half-ulp of the f16 the result is rounded to), n added into the exponent field.  x is clamped at -125.
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& e0, float& e1) {
    float x0, x1;
    unpack2(x2, x0, x1);
    const uint64_t x = pack2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
    const uint64_t xf = add2(x, pack2(12582912.f, 12582912.f));
    const uint64_t t = add2(xf, pack2(-12582912.f, -12582912.f));
    const uint64_t r = fma2(t, pack2(-1.f, -1.f), x);
    uint64_t q = fma2(pack2(0.0551716685f, 0.0551716685f), r, pack2(0.2426111251f, 0.2426111251f));
    q = fma2(q, r, pack2(0.6932609677f, 0.6932609677f));
    q = fma2(q, r, pack2(0.9999280572f, 0.9999280572f));
    float q0, q1, f0, f1;
    unpack2(q, q0, q1);
    unpack2(xf, f0, f1);
    e0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(f0) << 23));
    e1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(f1) << 23));
}

template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// POLY: every POLY-th pair of exponentials of a row is evaluated by exp2_poly2 instead of MUFU.EX2 (0 = none)
template <int POLY>
__global__ void __launch_bounds__(PF_THREADS, 1)
fa_prefill_tcgen05(const __grid_constant__ FaParams p, const __grid_constant__ PfArgs a,
                   const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV) {
    using namespace ptx;
    extern __shared__ uint8_t pf_smem_raw[];
    PfShared& sm = *reinterpret_cast<PfShared*>((reinterpret_cast<uintptr_t>(pf_smem_raw) + 1023) & ~(uintptr_t)1023);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // heavy (late) query tile pairs first: under a causal mask pair qp costs ~ 4*qp + 3 tile products
    const int per_pair = p.n_head * p.n_batch;
    const int qp = a.n_q_pairs - 1 - (int)(blockIdx.x / per_pair);
    const int iq2 = (int)(blockIdx.x % per_pair) % p.n_head;
    const int iq3 = (int)(blockIdx.x % per_pair) / p.n_head;
    const int ik2 = iq2 / p.gqa, ik3 = iq3 / p.rk3;
    const int qt0 = 2 * qp;  // 128-row tiles qt0 and qt0 + 1

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&sm.q_full[s], 1);
            mbar_init(&sm.k_full[s], 1); mbar_init(&sm.k_empty[s], 2);  // one arrival per MMA issuer
            mbar_init(&sm.v_full[s], 1); mbar_init(&sm.v_empty[s], 2);
            for (int h = 0; h < 2; h++) { mbar_init(&sm.s_full[s][h], 1); mbar_init(&sm.p_full[s][h], 4); }  // p_full: one arrival per softmax warp
            mbar_init(&sm.pv_done[s], 1);
        }
        fence_barrier_init();
        sm.j_lo = a.n_kv_tiles; sm.j_hi = 0;
    }
    if (warp == 8) {
        tmem_alloc(&sm.tmem_base, PF_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    // ---- the CTA's schedule, once: class of KV tile j for query tile qt0 (bits 0-1) and qt0+1 (bits 2-3), and the
    //      range [j_lo, j_hi) outside which nothing is visible.  Every role then walks the same list with cheap LDS. ----
    {
        int lo = a.n_kv_tiles, hi = 0;
        for (int j = threadIdx.x; j < a.n_kv_tiles; j += PF_THREADS) {
            const int c = pf_tile_class(p, a, qt0, j) | (pf_tile_class(p, a, qt0 + 1, j) << 2);
            sm.cls2[j] = (uint8_t)c;
            if (c != 0xA) { lo = min(lo, j); hi = max(hi, j + 1); }
        }
        if (hi > 0) { atomicMin(&sm.j_lo, lo); atomicMax(&sm.j_hi, hi); }
    }
    __syncthreads();
    const int j_lo = sm.j_lo, j_hi = sm.j_hi;

    if (warp >= 8) {
        reg_dec<PF_REGS_OTHER>();
        if (warp == 8) {
            // ===================== TMA producer =====================
            if (elect_one()) {
                prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV);
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    if (qt0 + t >= a.n_q_tiles) continue;
                    mbar_arrive_expect_tx(&sm.q_full[t], PF_TILE_BYTES);
                    tma_load_4d(sm.q[t], &tmQ, &sm.q_full[t], 0, (qt0 + t) * PF_BM, iq2, iq3);
                    tma_load_4d(sm.q[t] + PF_TILE_BYTES / 2, &tmQ, &sm.q_full[t], 64, (qt0 + t) * PF_BM, iq2, iq3);
                }
                int u = 0;
                for (int j = j_lo; j < j_hi; j++) {
                    if (sm.cls2[j] == 0xA) continue;
                    const int st = u & 1;
                    const uint32_t ph = (u >> 1) & 1;
                    mbar_wait(&sm.k_empty[st], ph ^ 1, a.dbg, 1);
                    mbar_arrive_expect_tx(&sm.k_full[st], PF_TILE_BYTES);
                    tma_load_4d(sm.k[st], &tmK, &sm.k_full[st], 0, j * PF_BN, ik2, ik3);
                    tma_load_4d(sm.k[st] + PF_TILE_BYTES / 2, &tmK, &sm.k_full[st], 64, j * PF_BN, ik2, ik3);
                    mbar_wait(&sm.v_empty[st], ph ^ 1, a.dbg, 2);
                    mbar_arrive_expect_tx(&sm.v_full[st], PF_TILE_BYTES);
                    tma_load_4d(sm.v[st], &tmV, &sm.v_full[st], 0, j * PF_BN, ik2, ik3);
                    tma_load_4d(sm.v[st] + PF_TILE_BYTES / 2, &tmV, &sm.v_full[st], 64, j * PF_BN, ik2, ik3);
                    u++;
                }
            }
        } else if (warp <= 10) {
            // ===================== MMA issuers: warp 9 drives query tile 0, warp 10 query tile 1 =====================
            // tcgen05.mma issue blocks while the tensor pipe's short queue is full, so any scalar work between two batches
            // of one issuer is tensor idle time — unless the other issuer's batch is queued behind it.  Each issuer is ONE
            // elected thread running a lean loop:  PV_t(previous tile), QK_t(this tile), per KV tile of the schedule.
            const int t = warp - 9;
            if (elect_one()) {
                constexpr uint32_t idesc_pv = make_idesc_f16(PF_BM, PF_D, 0, 1);
                const uint64_t dq = make_smem_desc_sw128(smem_u32(sm.q[t]), 16, 1024);
                const uint64_t dk[2] = {make_smem_desc_sw128(smem_u32(sm.k[0]), 16, 1024), make_smem_desc_sw128(smem_u32(sm.k[1]), 16, 1024)};
                const uint64_t dv[2] = {make_smem_desc_sw128(smem_u32(sm.v[0]), PF_TILE_BYTES / 2, 1024),
                                        make_smem_desc_sw128(smem_u32(sm.v[1]), PF_TILE_BYTES / 2, 1024)};
                const uint32_t tS = tmem + PF_TM_S + 128u * t, tO = tmem + PF_TM_O + 128u * t;
                // A KV tile is processed as two 64-key halves with their own score buffers S^0 / S^1 (64 TMEM columns each,
                // P^h over the first 32 of them): while the softmax group works on one half the tensor pipe does P.V and the
                // next Q.K^T of the other, so neither waits a full tile for the other.
                constexpr uint32_t idesc_qk = make_idesc_f16(PF_BM, 64, 0, 0);
                int n_tiles = 0;   // tiles of this query tile issued so far = phase counter of p_full[t][h]
                int pend = -1;     // stage whose V the pending P.V products need, or -1
                bool have_q = false;
                auto issue_pv = [&](int h) {  // O_t += P^h V[64h .. 64h+63] of the pending tile
                    mbar_wait(&sm.p_full[t][h], (n_tiles - 1) & 1, a.dbg, 5);
                    tc_fence_after();
                    if (n_tiles == 1 && h == 0) {
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) mma_ts(tO, tS + ks * 8, dv[pend] + (uint64_t)(ks * 2048 >> 4), idesc_pv, ks > 0);
                    } else {
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) mma_ts(tO, tS + 64u * h + ks * 8, dv[pend] + (uint64_t)((h * 8192 + ks * 2048) >> 4), idesc_pv, 1u);
                    }
                    tc_commit(&sm.pv_done[t]);
                };
                auto issue_qk = [&](int h, int st) {  // S^h = Q_t K[64h .. 64h+63]^T
#pragma unroll
                    for (int ks = 0; ks < 8; ks++) {
                        const uint64_t off = (uint64_t)(((ks >> 2) * (PF_TILE_BYTES / 2) + (ks & 3) * 32) >> 4);
                        mma_ss(tS + 64u * h, dq + off, dk[st] + off + (uint64_t)(h * 8192 >> 4), idesc_qk, ks > 0);
                    }
                    tc_commit(&sm.s_full[t][h]);
                };
                if (PF_STAGGER_CYCLES > 0 && t == 1 && j_lo < j_hi && (sm.cls2[j_lo] & 3) != 2) {
                    // Start the second tile a little after the first so the two softmax groups do not begin in lock step.
                    // A one-shot wait on phase 0 of the other tile's barrier is safe: that phase cannot be followed by another
                    // complete one before this thread has looked (the next one needs a whole softmax first).
                    mbar_wait(&sm.s_full[0][0], 0, a.dbg, 11);
                    const long long t0 = clock64();
                    while (clock64() - t0 < PF_STAGGER_CYCLES) {}
                }
                int u = 0;
                for (int j = j_lo; j < j_hi; j++) {
                    const int c2 = sm.cls2[j];
                    if (c2 == 0xA) continue;
                    const bool active = ((c2 >> (2 * t)) & 3) != 2;
                    const int st = u & 1;
                    const uint32_t ph = (u >> 1) & 1;
                    // The pending P.V products are always issued before waiting on a later stage (the producer needs v_empty
                    // to move on), interleaved with this tile's Q.K^T halves: PV^0 QK^0 PV^1 QK^1.
                    if (pend >= 0) issue_pv(0);
                    mbar_wait(&sm.k_full[st], ph, a.dbg, 3);
                    if (active) {
                        if (!have_q) { mbar_wait(&sm.q_full[t], 0, a.dbg, 4); have_q = true; }
                        tc_fence_after();
                        issue_qk(0, st);
                    }
                    if (pend >= 0) {
                        issue_pv(1);
                        tc_commit(&sm.v_empty[pend]);
                        pend = -1;
                    }
                    if (active) {
                        issue_qk(1, st);
                        tc_commit(&sm.k_empty[st]);
                        n_tiles++;
                    } else {
                        mbar_arrive(&sm.k_empty[st]);
                    }
                    mbar_wait(&sm.v_full[st], ph, a.dbg, 6);
                    if (active) pend = st;
                    else mbar_arrive(&sm.v_empty[st]);
                    u++;
                }
                if (pend >= 0) {
                    issue_pv(0);
                    issue_pv(1);
                    tc_commit(&sm.v_empty[pend]);
                }
            }
        }
    } else {
        // ===================== softmax / correction / epilogue: thread = query row =====================
        reg_inc<PF_REGS_SOFTMAX>();
        const int t = warp >> 2;                 // query tile of this warpgroup
        const int qt = qt0 + t;
        const int r = threadIdx.x & 127;         // row within the tile = TMEM lane
        const int q0 = qt * PF_BM;
        const int qrow = q0 + r;
        const bool row_valid = qrow < p.n_q;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t tS = trow + PF_TM_S + 128u * t, tO = trow + PF_TM_O + 128u * t;
        const float c = p.scale_log2;
        // rows past n_q read the last real mask row (their results are never stored): every lane of a warp takes the
        // same path, so the .sync.aligned tcgen05 instructions below always see a converged warp
        const char* mrow = (p.mask != nullptr && !p.causal) ? p.mask + (int64_t)min(qrow, p.n_q - 1) * p.nb31 : nullptr;
        const bool mask_vec = (((uintptr_t)p.mask | (uintptr_t)p.nb31) & 15) == 0;
        const int64_t vis = p.causal ? (int64_t)qrow + p.causal_off : (int64_t)p.n_kv;  // last visible key (inclusive)

        float m_ref = -INFINITY, l = 0.f;
        int it = 0;   // tiles of this query tile done
        int g = 0;    // half-tiles done = products the issuer has been handed = phase counter of pv_done[t]
        // diagnostics (b200fa_debug_set): clock64 stamps of the first 64 half-iterations of row 0 of each tile of CTA dump_cta
        long long* tl = (a.dump != nullptr && (int)blockIdx.x == a.dump_cta && r == 0) ? reinterpret_cast<long long*>(a.dump) + t * 64 * 8 : nullptr;
        if (tl) tl[63 * 8 + 6] = clock64();
        for (int j = j_lo; j < j_hi; j++) {
            const int cls = (sm.cls2[j] >> (2 * t)) & 3;
            if (cls == 2) continue;
#pragma unroll 1
            for (int h = 0; h < 2; h++, g++) {
                if (tl && g < 62) tl[g * 8 + 0] = clock64();
                mbar_wait(&sm.s_full[t][h], it & 1, a.dbg, 7);
                if (tl && g < 62) tl[g * 8 + 1] = clock64();
                __syncwarp();
                tc_fence_after();
                const uint32_t tSh = tS + 64u * h;
                if (p.dbg_mode == 2) {  // tuning aid (env B200FA_DBG_MODE=2): no softmax at all -> the loop runs at the pace of the tensor pipe
                    if (g > 0) { mbar_wait(&sm.pv_done[t], (g - 1) & 1, a.dbg, 10); __syncwarp(); }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.p_full[t][h]);
                    if (tl && g < 62) tl[g * 8 + 5] = clock64();
                    continue;
                }
                uint32_t s[2][32];
                tmem_ld32(tSh, s[0]);
                tmem_ld32(tSh + 32u, s[1]);
                tmem_wait_ld();
                if (tl && g < 62) tl[g * 8 + 2] = clock64();
                if (cls == 1) {
                    const int kv0 = j * PF_BN + 64 * h;
                    const int lim = (int)min((int64_t)(p.n_kv - 1), vis) - kv0;  // last visible column of this half for this row
#pragma unroll
                    for (int q2 = 0; q2 < 2; q2++) {
                        if (mrow != nullptr) {
                            if (mask_vec && kv0 + 64 <= p.n_kv) {
#pragma unroll
                                for (int v8 = 0; v8 < 4; v8++) {
                                    const uint4 mv = *reinterpret_cast<const uint4*>(mrow + (int64_t)(kv0 + q2 * 32 + v8 * 8) * 2);
                                    const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                                    for (int e = 0; e < 4; e++) {
                                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                                        s[q2][v8 * 8 + 2 * e] = __float_as_uint(__uint_as_float(s[q2][v8 * 8 + 2 * e]) + f.x * a.inv_scale);
                                        s[q2][v8 * 8 + 2 * e + 1] = __float_as_uint(__uint_as_float(s[q2][v8 * 8 + 2 * e + 1]) + f.y * a.inv_scale);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; i++) {
                                    const int kv = kv0 + q2 * 32 + i;
                                    if (kv < p.n_kv) s[q2][i] = __float_as_uint(__uint_as_float(s[q2][i]) + ld_mask(mrow, kv) * a.inv_scale);
                                }
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 32; i++)
                            if (q2 * 32 + i > lim) s[q2][i] = 0xff800000u;  // -inf: past the sequence end or the causal limit
                    }
                    __syncwarp();
                }
                // ---- row max (raw scores; scale > 0 on this path) ----
                float mx[8];
#pragma unroll
                for (int e = 0; e < 8; e++) mx[e] = -INFINITY;
#pragma unroll
                for (int q2 = 0; q2 < 2; q2++)
#pragma unroll
                    for (int i = 0; i < 32; i++) mx[i & 7] = fmaxf(mx[i & 7], __uint_as_float(s[q2][i]));
                const float m_tile = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7]))) * c;
                const bool need = m_tile > m_ref + PF_RESCALE_THRESHOLD;  // also true for the first finite max
                // pv_done(g-1) is observed in every half-iteration (phase rule): here when O has to be rescaled, otherwise
                // just before this half's P is published
                bool saw_pv = false;
                if (g > 0 && __any_sync(0xffffffffu, need)) {
                    // O currently holds the sum over the halves before g; PV_{g-1} must have landed before we touch it
                    mbar_wait(&sm.pv_done[t], (g - 1) & 1, a.dbg, 8);
                    saw_pv = true;
                    __syncwarp();
                    tc_fence_after();
                    const float alpha = need ? fast_exp2(m_ref - m_tile) : 1.f;
                    l *= alpha;
#pragma unroll
                    for (int q4 = 0; q4 < 4; q4++) {
                        uint32_t o[32];
                        tmem_ld32(tO + 32u * q4, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(tO + 32u * q4, o);
                    }
                    tmem_wait_st();
                }
                if (need) m_ref = m_tile;
                const float m_eff = (m_ref == -INFINITY) ? 0.f : m_ref;
                if (tl && g < 62) tl[g * 8 + 3] = clock64();
                // ---- P = exp2(s*c - m), row sum, pack to f16, store over S^h ----
                // Explicit passes over 32-column blocks keep the independent ex2 of a row in flight ahead of their consumers
                // (sum, pack); FFMA2 / FADD2 halve the fp32 issue slots.
                const uint64_t cc = pack2(c, c), nm = pack2(-m_eff, -m_eff);
                uint64_t ls2[4] = {0ull, 0ull, 0ull, 0ull};
                auto exp_block = [&](int q2) {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        float x0, x1;
                        const uint64_t x2 = fma2(pack2(__uint_as_float(s[q2][2 * i]), __uint_as_float(s[q2][2 * i + 1])), cc, nm);
                        if (POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == 1) {
                            exp2_poly2(x2, x0, x1);
                        } else {
                            unpack2(x2, x0, x1);
                            x0 = fast_exp2(x0); x1 = fast_exp2(x1);
                        }
                        s[q2][2 * i] = __float_as_uint(x0);
                        s[q2][2 * i + 1] = __float_as_uint(x1);
                    }
                };
                auto sum_pack_block = [&](int q2, uint32_t (&pk)[32], int base) {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float p0 = __uint_as_float(s[q2][2 * i]), p1 = __uint_as_float(s[q2][2 * i + 1]);
                        ls2[i & 3] = add2(ls2[i & 3], pack2(p0, p1));
                        pk[base + i] = pack_half2(p0, p1);
                    }
                };
                {
                    uint32_t pk[32];
                    exp_block(0); exp_block(1);
                    sum_pack_block(0, pk, 0);
                    sum_pack_block(1, pk, 16);
                    tmem_st32(tSh, pk);
                }
                {
                    float a0, a1, b0, b1;
                    unpack2(add2(ls2[0], ls2[1]), a0, a1); unpack2(add2(ls2[2], ls2[3]), b0, b1);
                    l += (a0 + a1) + (b0 + b1);
                }
                if (tl && g < 62) tl[g * 8 + 4] = clock64();
                if (g > 0 && !saw_pv) {
                    mbar_wait(&sm.pv_done[t], (g - 1) & 1, a.dbg, 10);
                    __syncwarp();
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.p_full[t][h]);
                if (tl && g < 62) tl[g * 8 + 5] = clock64();
            }
            it++;
        }

        // ---- epilogue: O / l -> dst[(iq3*n_q + q)*n_head + head][D]   (flash-llama.h:434) ----
        if (g > 0) {
            mbar_wait(&sm.pv_done[t], (g - 1) & 1, a.dbg, 9);
            __syncwarp();
            tc_fence_after();
        }
        const float inv_l = l > 0.f ? 1.f / l : 0.f;
        if (tl) tl[63 * 8 + 0] = clock64();
        if (qt < a.n_q_tiles) {
            // Each thread holds one whole output row; storing it directly would scatter 16-byte pieces over 32 lines per
            // instruction.  Instead every warp transposes its 32 rows, 256 bytes of each at a time, through 8 KB of its
            // tile's Q buffer — 16-byte chunk c of row i parked at chunk position c ^ (i & 15), conflict-free both ways —
            // and writes two whole 256-byte row pieces per instruction.  Q_t is idle: pv_done[t](last) fired after every
            // MMA this tile's issuer ever issued.  (K/V buffers may still be feeding the other tile.)
            uint4* stg = reinterpret_cast<uint4*>(sm.q[t] + (warp & 3) * 8192);
            const bool f32out = p.dst_type == B200FA_TYPE_F32;
            const int row0 = q0 + (warp & 3) * 32;                 // first query row of this warp
            const int64_t rstride = (int64_t)p.n_head * PF_D;      // elements between consecutive query rows of dst
            const int64_t obase = (((int64_t)iq3 * p.n_q + row0) * p.n_head + iq2) * PF_D;
            const int ri = lane >> 4, ch = lane & 15;               // store phase: lanes 0-15 one row, 16-31 the next
            const int n_pass = f32out ? 2 : 1;                      // 64 f32 or 128 f16 columns = 256 bytes per pass
            for (int pass = 0; pass < n_pass; pass++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int q4 = f32out ? pass * 2 + h : h * 2;  // f16: blocks (0,1) then (2,3)
                    uint32_t o[32], o2[32];
                    if (it > 0) {
                        tmem_ld32(tO + 32u * q4, o);
                        if (!f32out) tmem_ld32(tO + 32u * (q4 + 1), o2);
                        tmem_wait_ld();
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i++) { o[i] = 0u; o2[i] = 0u; }
                    }
                    if (f32out) {
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int c16 = h * 8 + i;
                            stg[lane * 16 + (c16 ^ (lane & 15))] =
                                make_uint4(__float_as_uint(__uint_as_float(o[4 * i]) * inv_l), __float_as_uint(__uint_as_float(o[4 * i + 1]) * inv_l),
                                           __float_as_uint(__uint_as_float(o[4 * i + 2]) * inv_l), __float_as_uint(__uint_as_float(o[4 * i + 3]) * inv_l));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const uint32_t* src = i < 4 ? o : o2;
                            const int b = (i & 3) * 8;
                            const int c16 = h * 8 + i;
                            stg[lane * 16 + (c16 ^ (lane & 15))] =
                                make_uint4(pack_half2(__uint_as_float(src[b]) * inv_l, __uint_as_float(src[b + 1]) * inv_l),
                                           pack_half2(__uint_as_float(src[b + 2]) * inv_l, __uint_as_float(src[b + 3]) * inv_l),
                                           pack_half2(__uint_as_float(src[b + 4]) * inv_l, __uint_as_float(src[b + 5]) * inv_l),
                                           pack_half2(__uint_as_float(src[b + 6]) * inv_l, __uint_as_float(src[b + 7]) * inv_l));
                        }
                    }
                }
                __syncwarp();
                if (tl) tl[63 * 8 + 1 + 2 * pass] = clock64();
                char* dbase = reinterpret_cast<char*>(p.dst) + (f32out ? (obase * 4 + pass * 256) : obase * 2);
                const int64_t rbytes = rstride * (f32out ? 4 : 2);
#pragma unroll 4
                for (int i = 0; i < 32; i += 2) {
                    const int rr = i + ri;
                    if (row0 + rr < p.n_q) {
                        const uint4 v = stg[rr * 16 + (ch ^ (rr & 15))];
                        *reinterpret_cast<uint4*>(dbase + rr * rbytes + ch * 16) = v;
                    }
                }
                __syncwarp();
                if (tl) tl[63 * 8 + 2 + 2 * pass] = clock64();
            }
        }
    }

    if (a.dump != nullptr && (int)blockIdx.x == a.dump_cta && (threadIdx.x & 127) == 0 && warp < 8)
        reinterpret_cast<long long*>(a.dump)[(warp >> 2) * 64 * 8 + 63 * 8 + 7] = clock64();
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem, PF_TMEM_COLS);
    }
}

// One pass over the mask: class of every 128x128 tile (0 all zero, 2 all -inf, 1 anything else).  Also counts, in
// *not_causal (may be null), the tiles that deviate from the exactly-causal pattern "0 where kv <= q + (n_kv - n_q), -inf
// elsewhere": a caller that passes the usual causal mask tensor without B200FA_FLAG_CAUSAL still gets the synthesised
// mask (no per-element mask reads in the attention kernel).
// blockIdx.z = mask slice (head + m_ne2 * batch): its base is mask + (z % m_ne2) * nb32 + (z / m_ne2) * nb33, its classes follow those of slice z - 1
__device__ __forceinline__ void pf_classify_block(int j, int qt, int z, int n_q_tiles_grid, const char* __restrict__ mask, int64_t nb31, int n_q, int n_kv,
                                                  int n_kv_tiles, uint8_t* __restrict__ cls, unsigned int* not_causal, int m_ne2, int64_t nb32, int64_t nb33) {
    mask += (int64_t)(z % m_ne2) * nb32 + (int64_t)(z / m_ne2) * nb33;
    cls += (int64_t)z * n_q_tiles_grid * n_kv_tiles;
    const int off = n_kv - n_q;
    int has_zero = 0, has_ninf = 0, has_other = 0, deviates = 0;
    const bool vec = ((((uintptr_t)mask | (uintptr_t)nb31) & 15) == 0) && (j + 1) * PF_BN <= n_kv;  // (mask already points at the slice)
    if (vec) {  // aligned, whole tile: 16-byte loads, 8 mask values each — all eight loads of a thread in flight at once (a rolled
                // loop made this kernel eight dependent DRAM round trips long: 8 us for C3's 8 MB mask)
        constexpr int kIters = PF_BM * (PF_BN / 8) / 256;
        uint4 v[kIters];
#pragma unroll
        for (int it = 0; it < kIters; it++) {
            const int idx = threadIdx.x + it * 256;
            const int r = qt * PF_BM + idx / (PF_BN / 8), col = j * PF_BN + (idx % (PF_BN / 8)) * 8;
            v[it] = r < n_q ? ld_nc_v4(mask + (int64_t)r * nb31 + (int64_t)col * 2) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int it = 0; it < kIters; it++) {
            const int idx = threadIdx.x + it * 256;
            const int r = qt * PF_BM + idx / (PF_BN / 8), col = j * PF_BN + (idx % (PF_BN / 8)) * 8;
            if (r >= n_q) continue;
            const uint32_t w[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const uint32_t bits = (w[e >> 1] >> (16 * (e & 1))) & 0xffffu;
                const bool visible = col + e <= r + off;
                if ((bits & 0x7fffu) == 0) { has_zero = 1; if (!visible) deviates = 1; }
                else if (bits == 0xfc00u) { has_ninf = 1; if (visible) deviates = 1; }
                else { has_other = 1; deviates = 1; }
            }
        }
    } else
    for (int idx = threadIdx.x; idx < PF_BM * (PF_BN / 2); idx += blockDim.x) {
        const int r = qt * PF_BM + idx / (PF_BN / 2), col = j * PF_BN + (idx % (PF_BN / 2)) * 2;
        if (r >= n_q) continue;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            if (col + e >= n_kv) continue;
            const uint16_t bits = *reinterpret_cast<const uint16_t*>(mask + (int64_t)r * nb31 + (int64_t)(col + e) * 2);
            const bool visible = col + e <= r + off;
            if ((bits & 0x7fffu) == 0) { has_zero = 1; if (!visible) deviates = 1; }
            else if (bits == 0xfc00u) { has_ninf = 1; if (visible) deviates = 1; }
            else { has_other = 1; deviates = 1; }
        }
    }
    has_zero = __syncthreads_or(has_zero);
    has_ninf = __syncthreads_or(has_ninf);
    has_other = __syncthreads_or(has_other);
    deviates = __syncthreads_or(deviates);
    if (threadIdx.x == 0) {
        cls[(int64_t)qt * n_kv_tiles + j] = (has_other || (has_zero && has_ninf)) ? 1 : (has_ninf ? 2 : 0);
        if (deviates && not_causal != nullptr) atomicAdd(not_causal, 1u);
    }
}
__global__ void __launch_bounds__(256) fa_mask_classify(const char* __restrict__ mask, int64_t nb31, int n_q, int n_kv,
                                                        int n_kv_tiles, uint8_t* __restrict__ cls, unsigned int* not_causal,
                                                        int m_ne2 = 1, int64_t nb32 = 0, int64_t nb33 = 0) {
    pf_classify_block(blockIdx.x, blockIdx.y, blockIdx.z, gridDim.y, mask, nb31, n_q, n_kv, n_kv_tiles, cls, not_causal, m_ne2, nb32, nb33);
}

// f32 Q (any ggml strides) -> dense f16 [batch][head][q][D]; same rounding as the reference (flash-llama.h:80)
__device__ __forceinline__ void pf_q_to_f16_block(int64_t block, const char* __restrict__ q, __half* __restrict__ out, int D, int n_q, int n_head,
                                                  int64_t total_rows, int64_t nb01, int64_t nb02, int64_t nb03) {
    const int chunks = D / 8;
    const int64_t idx = block * blockDim.x + threadIdx.x;
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
__global__ void __launch_bounds__(256) fa_q_to_f16(const char* __restrict__ q, __half* __restrict__ out, int D, int n_q, int n_head,
                                                   int64_t total_rows, int64_t nb01, int64_t nb02, int64_t nb03) {
    pf_q_to_f16_block(blockIdx.x, q, out, D, n_q, n_head, total_rows, nb01, nb02, nb03);
}

// Both helpers of a call in ONE launch (the reference's own call has an f32 Q and a mask tensor: two tiny kernels in front of the
// attention kernel cost two kernel boundaries): blocks [0, q_blocks) convert Q, the rest classify one mask tile each.
struct PfPrepArgs {
    const char* q; __half* q16; int D, n_q, n_head; int64_t total_rows, nb01, nb02, nb03; unsigned q_blocks;
    const char* mask; int64_t nb31; int n_kv, n_kv_tiles, cls_q_tiles; uint8_t* cls; unsigned int* not_causal; int m_ne2; int64_t nb32, nb33;
};
__global__ void __launch_bounds__(256) fa_prefill_prep(const PfPrepArgs a) {
    if (blockIdx.x < a.q_blocks) {
        pf_q_to_f16_block(blockIdx.x, a.q, a.q16, a.D, a.n_q, a.n_head, a.total_rows, a.nb01, a.nb02, a.nb03);
        return;
    }
    const unsigned b = blockIdx.x - a.q_blocks;
    const int j = (int)(b % (unsigned)a.n_kv_tiles), qt = (int)((b / (unsigned)a.n_kv_tiles) % (unsigned)a.cls_q_tiles), z = (int)(b / ((unsigned)a.n_kv_tiles * (unsigned)a.cls_q_tiles));
    pf_classify_block(j, qt, z, a.cls_q_tiles, a.mask, a.nb31, a.n_q, a.n_kv, a.n_kv_tiles, a.cls, a.not_causal, a.m_ne2, a.nb32, a.nb33);
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

// byte tensor [128][lines][heads][batch] (a contiguous byte range per head seen as 128-byte lines); box = 128 x box_lines, no swizzle
inline bool make_line_map(CUtensorMap* m, const void* base, int64_t lines, int64_t heads, int64_t batch, int64_t nb2, int64_t nb3,
                          int box_lines) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[4] = {128, (cuuint64_t)lines, (cuuint64_t)heads, (cuuint64_t)batch};
    cuuint64_t strides[3] = {128, (cuuint64_t)nb2, (cuuint64_t)nb3};
    cuuint32_t box[4] = {128, (cuuint32_t)box_lines, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct PfDebug {
    unsigned long long* dbg = nullptr;
    float* dump = nullptr;
    int dump_cta = 0;
};
// Diagnostics state (b200fa_debug_set): per calling thread, and only in -DB200FA_TUNING builds — the shipped library keeps no
// process-global debug state (concurrent calls from several threads share nothing).
#ifdef B200FA_TUNING
inline PfDebug& pf_debug() { static thread_local PfDebug d; return d; }
#else
inline PfDebug pf_debug() { return PfDebug{}; }
#endif

inline int launch_prefill_tcgen05(const FaParams& p, char* ws, size_t qf16_bytes, size_t cls_bytes, int sm_count,
                                  cudaStream_t st, int* launches) {
    (void)sm_count; (void)cls_bytes;
    if (p.D != PF_D || p.Dr != PF_D || p.kv_type != B200FA_TYPE_F16 || !(p.scale > 0.f) || p.n_kv > PF_MAX_KV_TILES * PF_BN) return B200FA_ERR_UNSUPPORTED;
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
    a.n_q_pairs = (a.n_q_tiles + 1) / 2;
    a.pack_sh = 0; a.q_rows = PF_BM; a.cls_q_tiles = a.n_q_tiles;
    a.inv_scale = 1.0f / p.scale;
    a.dbg = pf_debug().dbg; a.dump = pf_debug().dump; a.dump_cta = pf_debug().dump_cta;
    if (p.mask != nullptr && !p.causal) {
        uint8_t* cls = reinterpret_cast<uint8_t*>(ws + qf16_bytes);
        fa_mask_classify<<<dim3(a.n_kv_tiles, a.n_q_tiles), 256, 0, st>>>(p.mask, p.nb31, p.n_q, p.n_kv, a.n_kv_tiles, cls, nullptr);
        n++;
        a.cls = cls;
    }
    CUtensorMap tq, tk, tv;
    if (!make_tile_map(&tq, qbase, p.n_q, p.n_head, p.n_batch, qnb1, qnb2, qnb3)) return B200FA_ERR_CUDA;
    if (!make_tile_map(&tk, p.k, p.n_kv, p.n_head_kv, p.n_batch_kv, p.nb11, p.nb12, p.nb13)) return B200FA_ERR_CUDA;
    if (!make_tile_map(&tv, p.v, p.n_kv, p.n_head_kv, p.n_batch_kv, p.nb21, p.nb22, p.nb23)) return B200FA_ERR_CUDA;
    constexpr size_t smem_bytes = sizeof(PfShared) + 1024;
    static const int poly = tune_env("B200FA_POLY") ? atoi(tune_env("B200FA_POLY")) : 2;  // default: every 2nd pair on the FMA pipes
#ifdef B200FA_TUNING
    auto kern = poly == 2 ? fa_prefill_tcgen05<2> : (poly == 3 ? fa_prefill_tcgen05<3> : (poly == 4 ? fa_prefill_tcgen05<4> : fa_prefill_tcgen05<0>));
#else
    auto kern = fa_prefill_tcgen05<2>;
#endif
    static thread_local bool attr_set[64][5] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev][poly & 3]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess)
            return B200FA_ERR_CUDA;
        attr_set[dev][poly & 3] = true;
    }
    const unsigned grid = (unsigned)((int64_t)a.n_q_pairs * p.n_head * p.n_batch);
    kern<<<grid, PF_THREADS, smem_bytes, st>>>(p, a, tq, tk, tv);
    n++;
    if (launches) *launches = n;
    return cudaGetLastError() == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}

}  // namespace b200fa

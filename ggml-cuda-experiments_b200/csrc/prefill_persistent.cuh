// prefill_persistent.cuh — the prefill kernel as a persistent grid: one CTA per SM walks a heavy-first list of work
// items (query tile pair x head x batch) handed out by an atomic counter, so that
//   * causal work (items cost 2..2*n_pairs tile products) is balanced dynamically (longest first), and
//   * the fixed cost of an item — tensor-map/Q/K fetch latency at its start, draining the pipe and storing 128 KB of
//     output at its end — overlaps with its neighbours: the producer warp runs one item ahead of the MMA issuers,
//     which run ahead of the softmax groups' epilogue.
// The per-item pipeline is the one of prefill_tcgen05.cuh (two 128-row query tiles sharing a K/V ring, 64-key half
// tiles with double-buffered scores, P in TMEM, lazy rescale); only the hand-offs between items are new:
//   item_full / item_empty [2]   the producer warp publishes (work index, KV range, tile classes) of item k in slot k&1;
//                                the two issuers and the eight softmax warps release the slot when they are done with it
//   q_empty[t]                   issuer t: every Q_t.K^T of the item has completed -> Q_t may be overwritten
//   o_full[t]                    issuer t: every P.V of the item has completed -> O_t may be read out (tcgen05.commit)
//   o_free[t]                    softmax group t: O_t has been read out of TMEM -> the next item's first P.V may overwrite it
// Every barrier keeps running phase counters across items, and every waiter observes every phase in order — except pv_done[t],
// which the softmax groups only wait on before a (rare) lazy rescale (see there for why that is unambiguous).
#pragma once
#include <string.h>

#include "prefill_tcgen05.cuh"

namespace b200fa {

constexpr int PP_MAX_KV_TILES = 1024;  // schedule capacity of the persistent kernel: n_kv <= 131072 (longer: one CTA per item)

struct PpShared {
    uint8_t q[2][PF_TILE_BYTES];
    uint8_t k[2][PF_TILE_BYTES];
    uint8_t v[2][PF_TILE_BYTES];
    uint8_t stage[8][2][2048];        // epilogue transposition: two buffers of 32 rows x 64 bytes per softmax warp
    uint8_t cls2[2][PP_MAX_KV_TILES]; // per item slot
    uint64_t q_full[2], q_empty[2], k_full[2], k_empty[2], v_full[2], v_empty[2], pv_done[2], o_full[2], o_free[2];
    uint64_t s_full[2][2], p_full[2][2];
    uint64_t item_full[2], item_empty[2];
    int it_w[2], it_jlo[2], it_jhi[2];
    uint32_t tmem_base;
};

struct PpArgs {
    PfArgs f;                  // classes, tile counts, inv_scale
    int n_items;               // n_q_pairs * n_head * n_batch
    unsigned int* counters;    // [0] next item to hand out (beyond the first gridDim.x), [1] CTAs finished, [2] mask tiles that deviate
                               // from the exactly-causal pattern (written by fa_mask_classify); all zero between calls
    int detect_causal;         // a mask tensor was scanned: if counters[2] == 0 it is the causal mask and is synthesised instead of read
    // split-KV prefill (few items, long KV — chunked prefill): a work index is item * n_seg + segment; segment s covers KV tiles
    // [s*T/n_seg, (s+1)*T/n_seg) and the epilogue emits unnormalised (O~, m, l) rows [segment][row][D + 4] (tmO describes that
    // buffer) which fa_combine_pad merges.  n_seg == 1: the plain kernel.
    int n_seg;
};

// Role bodies shared by the persistent kernels (one or two softmax threads per query row).
// iq2: the item's q head — or, with GQA packing, its KV head (the tile holds all q heads of the group)
__device__ __forceinline__ void pp_decode_item(const FaParams& p, const PfArgs& a, int w, int& qt0, int& iq2, int& iq3) {
    const int heads = a.pack_sh ? p.n_head_kv : p.n_head;
    const int per_pair = heads * p.n_batch;
    qt0 = 2 * (a.n_q_pairs - 1 - w / per_pair);   // heavy (late) tile pairs first
    const int rem = w % per_pair;
    iq2 = rem % heads; iq3 = rem / heads;
}
// Q tile load (two 64-column boxes).  Unpacked: tmQ is [D][n_q][head][batch], box 64 x 128 rows.  Packed: tmQ is [D][head][n_q][batch],
// box 64 x gqa heads x (128 / gqa) positions, which lands position-major / head-minor in the tile.
__device__ __forceinline__ void pp_load_q(const FaParams& p, const PfArgs& a, void* dst, const CUtensorMap* tmQ, uint64_t* bar, int qt, int iq2, int iq3) {
    using namespace ptx;
#pragma unroll
    for (int c = 0; c < 2; c++) {
        if (a.pack_sh) tma_load_4d(reinterpret_cast<uint8_t*>(dst) + c * (PF_TILE_BYTES / 2), tmQ, bar, 64 * c, iq2 << a.pack_sh, qt * a.q_rows, iq3);
        else tma_load_4d(reinterpret_cast<uint8_t*>(dst) + c * (PF_TILE_BYTES / 2), tmQ, bar, 64 * c, qt * PF_BM, iq2, iq3);
    }
}

// PACK = false: the packing fields are compile-time constants again (pack_sh = 0, 128 positions per tile) — as run-time values they
// lengthened the producer's serial schedule computation, which sits on the critical path between two items (C3 +0.6 us).
template <bool PACK>
__device__ __forceinline__ PfArgs pp_args(const PpArgs& pa) {
    PfArgs a = pa.f;
    if (!PACK) { a.pack_sh = 0; a.q_rows = PF_BM; }
    return a;
}
template <bool PACK>
__device__ __forceinline__ void pp_producer_role(const FaParams& p, const PpArgs& pa, PpShared& sm, int lane, const CUtensorMap& tmQ,
                                                 const CUtensorMap& tmK, const CUtensorMap& tmV) {
    using namespace ptx;
    const PfArgs a = pp_args<PACK>(pa);
    auto decode_item = [&](int w, int& qt0, int& iq2, int& iq3) { pp_decode_item(p, a, w / pa.n_seg, qt0, iq2, iq3); };
    const bool causal = p.causal != 0 || (pa.detect_causal != 0 && __ldcg(pa.counters + 2) == 0u);
    // ===================== producer warp: hands out items, builds their schedule, streams Q / K / V =====================
    if (lane == 0) { prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); }
    int u_tot = 0;           // K/V tiles streamed so far (ring position / phase)
    int nq[2] = {0, 0};      // Q_t loads so far
    for (int k = 0;; k++) {
        const int slot = k & 1;
        if (k >= 2) { mbar_wait(&sm.item_empty[slot], ((k >> 1) - 1) & 1, a.dbg, 20); __syncwarp(); }
        int w = (int)blockIdx.x;
        if (k > 0) {
            if (lane == 0) w = (int)gridDim.x + (int)atomicAdd(pa.counters, 1u);
            w = __shfl_sync(0xffffffffu, w, 0);
        }
        if (w >= pa.n_items) w = -1;
        int qt0 = 0, iq2 = 0, iq3 = 0, lo = a.n_kv_tiles, hi = 0;
        if (w >= 0) {
            decode_item(w, qt0, iq2, iq3);
            for (int j = lane; j < a.n_kv_tiles; j += 32) {
                const int ms = fa_mask_slice(p, iq2, iq3);
                const int c = pf_tile_class(p, a, qt0, j, causal, ms) | (pf_tile_class(p, a, qt0 + 1, j, causal, ms) << 2);
                sm.cls2[slot][j] = (uint8_t)c;
                if (c != 0xA) { lo = min(lo, j); hi = max(hi, j + 1); }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
        }
        if (w >= 0 && pa.n_seg > 1) {  // this work index is one KV segment of its item
            const int seg = w % pa.n_seg;
            lo = max(lo, (int)((long long)seg * a.n_kv_tiles / pa.n_seg));
            hi = min(hi, (int)((long long)(seg + 1) * a.n_kv_tiles / pa.n_seg));
            if (lo >= hi) { lo = a.n_kv_tiles; hi = 0; }
        }
        if (lane == 0) { sm.it_w[slot] = w; sm.it_jlo[slot] = lo; sm.it_jhi[slot] = hi; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.item_full[slot]);
        if (w < 0) break;
        if (lane == 0) {
            const int ik2 = a.pack_sh ? iq2 : iq2 / p.gqa, ik3 = iq3 / p.rk3;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                if (qt0 + t >= a.n_q_tiles) continue;
                if (nq[t] > 0) mbar_wait(&sm.q_empty[t], (nq[t] - 1) & 1, a.dbg, 21);
                mbar_arrive_expect_tx(&sm.q_full[t], PF_TILE_BYTES);
                pp_load_q(p, a, sm.q[t], &tmQ, &sm.q_full[t], qt0 + t, iq2, iq3);
                nq[t]++;
            }
            for (int j = lo; j < hi; j++) {
                if (sm.cls2[slot][j] == 0xA) continue;
                const int st = u_tot & 1;
                const uint32_t ph = (u_tot >> 1) & 1;
                mbar_wait(&sm.k_empty[st], ph ^ 1, a.dbg, 1);
                mbar_arrive_expect_tx(&sm.k_full[st], PF_TILE_BYTES);
                tma_load_4d(sm.k[st], &tmK, &sm.k_full[st], 0, j * PF_BN, ik2, ik3);
                tma_load_4d(sm.k[st] + PF_TILE_BYTES / 2, &tmK, &sm.k_full[st], 64, j * PF_BN, ik2, ik3);
                mbar_wait(&sm.v_empty[st], ph ^ 1, a.dbg, 2);
                mbar_arrive_expect_tx(&sm.v_full[st], PF_TILE_BYTES);
                tma_load_4d(sm.v[st], &tmV, &sm.v_full[st], 0, j * PF_BN, ik2, ik3);
                tma_load_4d(sm.v[st] + PF_TILE_BYTES / 2, &tmV, &sm.v_full[st], 64, j * PF_BN, ik2, ik3);
                u_tot++;
            }
        }
        __syncwarp();
    }
}

template <bool PACK>
__device__ __forceinline__ void pp_issuer_role(const FaParams& p, const PpArgs& pa, PpShared& sm, uint32_t tmem, const int t) {
    using namespace ptx;
    const PfArgs a = pp_args<PACK>(pa);
    auto decode_item = [&](int w, int& qt0, int& iq2, int& iq3) { pp_decode_item(p, a, w / pa.n_seg, qt0, iq2, iq3); };
    // ===================== MMA issuers: warp 9 drives query tile 0, warp 10 query tile 1 =====================
    if (elect_one()) {
        constexpr uint32_t idesc_qk = make_idesc_f16(PF_BM, 64, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_f16(PF_BM, PF_D, 0, 1);
        const uint64_t dq = make_smem_desc_sw128(smem_u32(sm.q[t]), 16, 1024);
        const uint64_t dk[2] = {make_smem_desc_sw128(smem_u32(sm.k[0]), 16, 1024), make_smem_desc_sw128(smem_u32(sm.k[1]), 16, 1024)};
        const uint64_t dv[2] = {make_smem_desc_sw128(smem_u32(sm.v[0]), PF_TILE_BYTES / 2, 1024),
                                make_smem_desc_sw128(smem_u32(sm.v[1]), PF_TILE_BYTES / 2, 1024)};
        const uint32_t tS = tmem + PF_TM_S + 128u * t, tO = tmem + PF_TM_O + 128u * t;
        int u_tot = 0;       // K/V ring position
        int tiles_tot = 0;   // tiles of this query tile issued so far (phase counter of p_full[t][h])
        int nq = 0;          // Q_t tiles consumed so far (phase counter of q_full[t])
        int n_work = 0;      // items in which this query tile had any work (phase counter of o_free[t])
        for (int k = 0;; k++) {
            const int slot = k & 1;
            mbar_wait(&sm.item_full[slot], (k >> 1) & 1, a.dbg, 22);
            const int w = sm.it_w[slot];
            if (w < 0) break;
            const int j_lo = sm.it_jlo[slot], j_hi = sm.it_jhi[slot];
            int qt0, iq2, iq3;
            decode_item(w, qt0, iq2, iq3);
            const bool tile_valid = qt0 + t < a.n_q_tiles;
            int j_last = -1;  // last KV tile this query tile needs: Q_t is released right after its Q.K^T
            for (int j = j_hi - 1; j >= j_lo; j--)
                if (((sm.cls2[slot][j] >> (2 * t)) & 3) != 2) { j_last = j; break; }
            int pend = -1;
            // o_full[t] / o_free[t] are only used by items in which this query tile has work: those keep the softmax group and this
            // issuer in lock step through s_full / p_full.  (An idle softmax group could otherwise run two items — two
            // phases — ahead of its issuer, and a parity wait that is two phases late never returns.)
            bool have_q = false, o_seen = (n_work == 0), first_pv = true;
            auto issue_pv = [&](int h) {  // O_t += P^h V[64h .. 64h+63] of the pending tile
                mbar_wait(&sm.p_full[t][h], (tiles_tot - 1) & 1, a.dbg, 5);
                if (!o_seen) { mbar_wait(&sm.o_free[t], (n_work - 1) & 1, a.dbg, 23); o_seen = true; }  // the previous O_t has been read out
                tc_fence_after();
                if (first_pv) {
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) mma_ts(tO, tS + ks * 8, dv[pend] + (uint64_t)(ks * 2048 >> 4), idesc_pv, ks > 0);
                    first_pv = false;
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
            for (int j = j_lo; j < j_hi; j++) {
                const int c2 = sm.cls2[slot][j];
                if (c2 == 0xA) continue;
                const bool active = ((c2 >> (2 * t)) & 3) != 2;
                const int st = u_tot & 1;
                const uint32_t ph = (u_tot >> 1) & 1;
                if (pend >= 0) issue_pv(0);
                mbar_wait(&sm.k_full[st], ph, a.dbg, 3);
                if (active) {
                    if (!have_q) { mbar_wait(&sm.q_full[t], nq & 1, a.dbg, 4); have_q = true; }
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
                    if (j == j_last) tc_commit(&sm.q_empty[t]);
                    tiles_tot++;
                } else {
                    mbar_arrive(&sm.k_empty[st]);
                }
                mbar_wait(&sm.v_full[st], ph, a.dbg, 6);
                if (active) pend = st;
                else mbar_arrive(&sm.v_empty[st]);
                u_tot++;
            }
            if (pend >= 0) {
                issue_pv(0);
                issue_pv(1);
                tc_commit(&sm.v_empty[pend]);
            }
            if (j_last >= 0) tc_commit(&sm.o_full[t]);  // every product of the item has landed: O_t may be read out
            if (tile_valid) {
                if (!have_q) {  // the tile was loaded but nothing of the KV range is visible to it: hand Q_t straight back
                    mbar_wait(&sm.q_full[t], nq & 1, a.dbg, 4);
                    mbar_arrive(&sm.q_empty[t]);
                }
                nq++;
            }
            if (j_last >= 0) n_work++;
            mbar_arrive(&sm.item_empty[slot]);
        }
    }
}

template <int POLY, bool EXT = false, bool PACK = false>
__global__ void __launch_bounds__(PF_THREADS, 1)
fa_prefill_persistent(const __grid_constant__ FaParams p, const __grid_constant__ PpArgs pa,
                      const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO) {
    using namespace ptx;
    extern __shared__ __align__(1024) uint8_t pp_smem_raw[];  // 128B-swizzled TMA tiles need 1024-byte alignment; no slack to spare
    if ((smem_u32(pp_smem_raw) & 1023u) != 0) __trap();
    PpShared& sm = *reinterpret_cast<PpShared*>(pp_smem_raw);
    const PfArgs a = pp_args<PACK>(pa);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the next kernel may take over SMs as this grid's CTAs retire
    // diagnostics (tuning builds only: a.dump is null otherwise): per CTA %globaltimer at entry / after the dependency wait / at exit and
    // the clock64 cycles between the last two — cycles / ns is the SM clock the kernel really ran at (~1.65 GHz on C3)
#ifdef B200FA_TUNING
    long long* ct = (a.dump != nullptr && threadIdx.x == 0) ? reinterpret_cast<long long*>(a.dump) + 2 * 32 * 8 + blockIdx.x * 4 : nullptr;
#else
    long long* const ct = nullptr;
#endif
    auto gtime = [] { long long x; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(x)); return x; };
    if (ct) ct[0] = gtime();
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&sm.q_full[s], 1); mbar_init(&sm.q_empty[s], 1);
            mbar_init(&sm.k_full[s], 1); mbar_init(&sm.k_empty[s], 2);  // one arrival per MMA issuer
            mbar_init(&sm.v_full[s], 1); mbar_init(&sm.v_empty[s], 2);
            mbar_init(&sm.pv_done[s], 1); mbar_init(&sm.o_full[s], 1); mbar_init(&sm.o_free[s], 4);  // o_free, p_full: one arrival per softmax warp
            for (int h = 0; h < 2; h++) { mbar_init(&sm.s_full[s][h], 1); mbar_init(&sm.p_full[s][h], 4); }
            mbar_init(&sm.item_full[s], 1); mbar_init(&sm.item_empty[s], 10);  // 2 issuers + 8 softmax warps
        }
        fence_barrier_init();
    }
    if (warp == 8) {
        tmem_alloc(&sm.tmem_base, PF_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    // Programmatic dependent launch (see launch_prefill_persistent): everything above ran while the previous kernel of the stream
    // was still draining; nothing below touches global memory before that kernel has completed and flushed.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (ct) { ct[1] = gtime(); ct[3] = clock64(); }

    // work index -> (first 128-row tile, head, batch); heavy (late) tile pairs first
    auto decode_item = [&](int w, int& qt0, int& iq2, int& iq3) { pp_decode_item(p, a, w / pa.n_seg, qt0, iq2, iq3); };

    if (warp >= 8) {
        reg_dec<PF_REGS_OTHER>();
        if (warp == 8) {
            pp_producer_role<PACK>(p, pa, sm, lane, tmQ, tmK, tmV);
        } else if (warp <= 10) {
            pp_issuer_role<PACK>(p, pa, sm, tmem, warp - 9);
        }
    } else {
        // ===================== softmax / correction / epilogue: thread = query row =====================
        reg_inc<PF_REGS_SOFTMAX>();
        const int t = warp >> 2;                 // query tile of this warpgroup
        const int r = threadIdx.x & 127;         // row within the tile = TMEM lane
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t tS = trow + PF_TM_S + 128u * t, tO = trow + PF_TM_O + 128u * t;
        const float c = p.scale_log2;
        const bool mask_vec = (((uintptr_t)p.mask | (uintptr_t)p.nb31 | (uintptr_t)p.nb32 | (uintptr_t)p.nb33) & 15) == 0;
        const bool causal = p.causal != 0 || (pa.detect_causal != 0 && __ldcg(pa.counters + 2) == 0u);
        int it_tot = 0;   // tiles done over all items (phase counter of s_full[t][h])
        int g_tot = 0;    // half tiles done over all items (phase counter of pv_done[t])
        int n_work = 0;   // items in which this query tile had work (phase counter of o_full[t])
        for (int k = 0;; k++) {
            const int slot = k & 1;
            mbar_wait(&sm.item_full[slot], (k >> 1) & 1, a.dbg, 24);
            __syncwarp();
            const int w = sm.it_w[slot];
            if (w < 0) break;
            // diagnostics (b200fa_debug_set): per item of CTA dump_cta: clock64 at item start, after the softmax loop, after O is
            // read out, after the stores; plus the work index and the number of half tiles
            long long* tl = (a.dump != nullptr && (int)blockIdx.x == a.dump_cta && r == 0 && k < 32) ? reinterpret_cast<long long*>(a.dump) + (t * 32 + k) * 8 : nullptr;
            if (tl) { tl[0] = clock64(); tl[4] = w; }
            const int j_lo = sm.it_jlo[slot], j_hi = sm.it_jhi[slot];
            int qt0, iq2, iq3;
            decode_item(w, qt0, iq2, iq3);
            const int qt = qt0 + t;
            const int q0 = qt * a.q_rows;                 // first query position of the tile
            const int qrow = q0 + (r >> a.pack_sh);       // this row's query position (packed: row r = position r >> pack_sh, head r & (gqa - 1))
            // rows past n_q read the last real mask row (their results are never stored): every lane of a warp takes the
            // same path, so the .sync.aligned tcgen05 instructions below always see a converged warp
            const char* mrow = (p.mask != nullptr && !causal) ? p.mask + (int64_t)min(qrow, p.n_q - 1) * p.nb31 + (EXT ? fa_mask_slice_off(p, iq2, iq3) : 0) : nullptr;
            const int64_t vis = causal ? (int64_t)qrow + p.causal_off : (int64_t)p.n_kv;  // last visible key (inclusive)
            float m_ref = -INFINITY, l = 0.f;
            const float mask_mul = EXT ? a.inv_scale * fa_slope(p, iq2) : a.inv_scale;  // mask values are folded into RAW scores: x slope / scale
            int g = 0;  // half tiles of this item done
            for (int j = j_lo; j < j_hi; j++) {
                const int cls = (sm.cls2[slot][j] >> (2 * t)) & 3;
                if (cls == 2) continue;
#pragma unroll 1
                for (int h = 0; h < 2; h++, g++, g_tot++) {
                    mbar_wait(&sm.s_full[t][h], it_tot & 1, a.dbg, 7);
                    __syncwarp();
                    tc_fence_after();
                    const uint32_t tSh = tS + 64u * h;
                    uint32_t s[2][32];
                    tmem_ld32(tSh, s[0]);
                    tmem_ld32(tSh + 32u, s[1]);
                    tmem_wait_ld();
                    if (EXT && p.cap_in != 0.f) {  // logit soft-cap (ext2): s <- cap*tanh(s*scale/cap), kept in raw units (divided by scale)
#pragma unroll
                        for (int q2 = 0; q2 < 2; q2++)
#pragma unroll
                            for (int i = 0; i < 32; i++) s[q2][i] = __float_as_uint(fa_tanh(__uint_as_float(s[q2][i]) * p.cap_in) * p.cap_raw);
                    }
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
                                        const uint32_t wd[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                                        for (int e = 0; e < 4; e++) {
                                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wd[e]));
                                            s[q2][v8 * 8 + 2 * e] = __float_as_uint(__uint_as_float(s[q2][v8 * 8 + 2 * e]) + f.x * mask_mul);
                                            s[q2][v8 * 8 + 2 * e + 1] = __float_as_uint(__uint_as_float(s[q2][v8 * 8 + 2 * e + 1]) + f.y * mask_mul);
                                        }
                                    }
                                } else {
#pragma unroll
                                    for (int i = 0; i < 32; i++) {
                                        const int kv = kv0 + q2 * 32 + i;
                                        if (kv < p.n_kv) s[q2][i] = __float_as_uint(__uint_as_float(s[q2][i]) + ld_mask(mrow, kv) * mask_mul);
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
                    // O_t holds the sum over the halves before this one; the last of those products must have landed before it is
                    // rescaled.  pv_done is NOT observed in every half-iteration (that cost ~100 cycles per half): having seen
                    // s_full of this half, which the issuer committed after P.V(g-2), and with P.V(g) not yet issued, the barrier can only
                    // be in phase g_tot-1 or g_tot here, so the parity wait for P.V(g-1) is unambiguous.
                    if (g > 0 && __any_sync(0xffffffffu, need)) {
                        mbar_wait(&sm.pv_done[t], (g_tot - 1) & 1, a.dbg, 8);
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
                    // ---- P = exp2(s*c - m), row sum, pack to f16, store over S^h ----
                    const uint64_t cc = pack2(c, c), nm = pack2(-m_eff, -m_eff);
                    uint64_t ls2[4] = {0ull, 0ull, 0ull, 0ull};
                    uint32_t pk[32];
#pragma unroll
                    for (int q2 = 0; q2 < 2; q2++) {
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
                    }
#pragma unroll
                    for (int q2 = 0; q2 < 2; q2++) {
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const float p0 = __uint_as_float(s[q2][2 * i]), p1 = __uint_as_float(s[q2][2 * i + 1]);
                            ls2[i & 3] = add2(ls2[i & 3], pack2(p0, p1));
                            pk[q2 * 16 + i] = pack_half2(p0, p1);
                        }
                    }
                    tmem_st32(tSh, pk);
                    {
                        float a0, a1, b0, b1;
                        unpack2(add2(ls2[0], ls2[1]), a0, a1); unpack2(add2(ls2[2], ls2[3]), b0, b1);
                        l += (a0 + a1) + (b0 + b1);
                    }
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.p_full[t][h]);
                }
                it_tot++;
            }

            // ---- epilogue: O / l -> dst[(iq3*n_q + q)*n_head + head][D]   (flash-llama.h:434) ----
            if (tl) { tl[1] = clock64(); tl[5] = g; }
            uint32_t o[4][32];
            if (g > 0) {
                // (pv_done cannot be used here: after the last half its phase may be g_tot-2, g_tot-1 or g_tot.  o_full has one phase per item.)
                mbar_wait(&sm.o_full[t], n_work & 1, a.dbg, 9);
                n_work++;
                __syncwarp();
                tc_fence_after();
#pragma unroll
                for (int q4 = 0; q4 < 4; q4++) tmem_ld32(tO + 32u * q4, o[q4]);
                tmem_wait_ld();
            } else {
#pragma unroll
                for (int q4 = 0; q4 < 4; q4++)
#pragma unroll
                    for (int i = 0; i < 32; i++) o[q4][i] = 0u;
            }
            tc_fence_before();
            __syncwarp();
            if (g > 0 && lane == 0) mbar_arrive(&sm.o_free[t]);  // O_t is in registers: the next item may accumulate into it
            if (tl) tl[2] = clock64();
            const bool partial = pa.n_seg > 1;  // split-KV: unnormalised O~ plus (m, l) go to the partial buffer (tmO describes it)
            const float inv_l = partial ? 1.f : (l > 0.f ? 1.f / l : 0.f);
            if (qt < a.n_q_tiles) {
                // Each thread holds one whole output row.  Per pass the warp parks a 64-byte piece of its 32 rows in 2 KB of
                // shared memory (16-byte chunk c of row i at chunk position c ^ ((i >> 1) & 3) = the TMA 64-byte swizzle) and
                // one lane hands the 32 x 64-byte box to the TMA engine: the store to HBM is asynchronous, rows past n_q are
                // clipped by the tensor map, and the warp only waits until the engine has READ the staging buffer.
                const bool f32out = partial || p.dst_type == B200FA_TYPE_F32;
                const int row0 = q0 + (((warp & 3) * 32) >> a.pack_sh);  // first query position of this warp's 32 rows
                // 16 f32 or 32 f16 columns = 64 bytes per pass; columns past the real head size are never stored.  Partials: a ninth
                // pass carries (m in natural-log units, l) in columns D, D+1 (the box is clipped to the D + 4 columns of a record).
                const int n_pass = partial ? 9 : (p.Dr * (f32out ? 4 : 2) + 63) >> 6;
                const int c3 = partial ? (w % pa.n_seg) * p.n_batch + iq3 : iq3;
#pragma unroll
                for (int pass = 0; pass < 9; pass++) {
                    if (pass >= n_pass) break;
                    uint4 piece[4];
                    if (pass == 8) {
                        piece[0] = make_uint4(__float_as_uint(m_ref * 0.6931471805599453f), __float_as_uint(l), 0u, 0u);
                        piece[1] = piece[2] = piece[3] = make_uint4(0u, 0u, 0u, 0u);
                    } else if (f32out) {
                        const int q4 = pass >> 1, b = (pass & 1) * 16;
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            piece[i] = make_uint4(__float_as_uint(__uint_as_float(o[q4][b + 4 * i]) * inv_l), __float_as_uint(__uint_as_float(o[q4][b + 4 * i + 1]) * inv_l),
                                                  __float_as_uint(__uint_as_float(o[q4][b + 4 * i + 2]) * inv_l), __float_as_uint(__uint_as_float(o[q4][b + 4 * i + 3]) * inv_l));
                    } else {
                        const int q4 = pass;
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            piece[i] = make_uint4(pack_half2(__uint_as_float(o[q4][8 * i]) * inv_l, __uint_as_float(o[q4][8 * i + 1]) * inv_l),
                                                  pack_half2(__uint_as_float(o[q4][8 * i + 2]) * inv_l, __uint_as_float(o[q4][8 * i + 3]) * inv_l),
                                                  pack_half2(__uint_as_float(o[q4][8 * i + 4]) * inv_l, __uint_as_float(o[q4][8 * i + 5]) * inv_l),
                                                  pack_half2(__uint_as_float(o[q4][8 * i + 6]) * inv_l, __uint_as_float(o[q4][8 * i + 7]) * inv_l));
                    }
                    uint4* stg = reinterpret_cast<uint4*>(sm.stage[warp][pass & 1]);
                    if (lane == 0) bulk_wait_read1();   // the box of two passes ago has left this staging buffer
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; i++) stg[lane * 4 + (i ^ ((lane >> 1) & 3))] = piece[i];
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        // box: 64 bytes x 1 head x 32 positions — packed: x gqa heads x 32 / gqa positions, the order of the warp's rows
                        tma_store_4d(&tmO, stg, pass * (f32out ? 16 : 32), iq2 << a.pack_sh, row0, c3);
                        bulk_commit();
                    }
                }
                if (lane == 0) bulk_wait_read0();
                __syncwarp();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.item_empty[slot]);
            if (tl) tl[3] = clock64();
        }
    }

    if (warp < 8 && lane == 0) bulk_wait_all0();  // this lane's output boxes have been written
    tc_fence_before();
    __syncthreads();
    if (ct) { ct[2] = gtime(); ct[3] = clock64() - ct[3]; }
    if (warp == 8) {
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

// GQA packing: log2 of the q heads that share one 128-row tile, or 0.  It pays when the q heads of a group would otherwise each
// occupy their own, mostly empty, tile and stream the group's K/V once per head: 8 query positions x GQA 4 (speculative decoding,
// small prefill chunks) is ONE quarter-filled tile per KV head instead of four 6 %-filled ones.  Power-of-two groups up to 32 heads
// (a warp's 32 rows must be whole positions for the output box); not with the ext2 modifiers (their slopes / mask slices are per item).
inline int pp_pack_shift(int64_t n_q, int64_t n_head, int64_t n_head_kv, bool ext) {
    if (n_head_kv <= 0 || n_head % n_head_kv) return 0;
    const int64_t gqa = n_head / n_head_kv;
    if (ext || gqa < 2 || gqa > 32 || (gqa & (gqa - 1))) return 0;
    const int64_t np = PF_BM / gqa;
    const int64_t pairs_packed = ((n_q + np - 1) / np + 1) / 2, pairs_plain = ((n_q + PF_BM - 1) / PF_BM + 1) / 2;
    if (pairs_packed >= pairs_plain * gqa) return 0;  // no fewer passes over K/V
    int sh = 0;
    while ((1 << sh) < gqa) sh++;
    return sh;
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
    PfPrepArgs prep{};  // the helper launch in front of the attention kernel: Q conversion and / or mask scan
    if (p.q_type == B200FA_TYPE_F32) {
        __half* q16 = reinterpret_cast<__half*>(ws);
        const int64_t work = p.total_rows * (Dr / 8);
        prep.q = p.q; prep.q16 = q16; prep.D = Dr; prep.n_q = p.n_q; prep.n_head = p.n_head; prep.total_rows = p.total_rows;
        prep.nb01 = p.nb01; prep.nb02 = p.nb02; prep.nb03 = p.nb03; prep.q_blocks = (unsigned)((work + 255) / 256);
        qbase = q16;
        qnb1 = Dr * 2; qnb2 = (int64_t)p.n_q * Dr * 2; qnb3 = (int64_t)p.n_head * p.n_q * Dr * 2;
    }
    PpArgs pa{};
    PfArgs& a = pa.f;
    const bool ext = p.cap_in != 0.f || p.alibi_nhl2 != 0 || p.m_ne2 * p.m_ne3 > 1;  // ext2 score modifiers / mask slices: their own instantiation
    a.pack_sh = pp_pack_shift(p.n_q, p.n_head, p.n_head_kv, ext);
    a.q_rows = PF_BM >> a.pack_sh;
    a.cls_q_tiles = (p.n_q + PF_BM - 1) / PF_BM;
    a.n_q_tiles = (p.n_q + a.q_rows - 1) / a.q_rows;
    a.n_kv_tiles = (p.n_kv + PF_BN - 1) / PF_BN;
    a.n_q_pairs = (a.n_q_tiles + 1) / 2;
    a.inv_scale = 1.0f / p.scale;
    a.dbg = pf_debug().dbg; a.dump = pf_debug().dump; a.dump_cta = pf_debug().dump_cta;
    if (p.mask != nullptr && !p.causal) {
        uint8_t* cls = reinterpret_cast<uint8_t*>(ws + qf16_bytes);
        prep.mask = p.mask; prep.nb31 = p.nb31; prep.n_q = p.n_q; prep.n_kv = p.n_kv; prep.n_kv_tiles = a.n_kv_tiles; prep.cls_q_tiles = a.cls_q_tiles;
        prep.cls = cls; prep.not_causal = counters + 2; prep.m_ne2 = p.m_ne2; prep.nb32 = p.nb32; prep.nb33 = p.nb33;
        a.cls = cls;
        pa.detect_causal = 1;
    }
    {
        const unsigned cls_blocks = prep.mask ? (unsigned)(a.n_kv_tiles * a.cls_q_tiles * p.m_ne2 * p.m_ne3) : 0u;
        if (prep.q_blocks + cls_blocks > 0) {
            fa_prefill_prep<<<prep.q_blocks + cls_blocks, 256, 0, st>>>(prep);
            n++;
        }
    }
    if (n_seg < 1 || (n_seg > 1 && (part == nullptr || p.Dr != PF_D))) return B200FA_ERR_INVALID;
    pa.n_seg = n_seg;
    pa.n_items = a.n_q_pairs * (a.pack_sh ? p.n_head_kv : p.n_head) * p.n_batch * n_seg;
    pa.counters = counters;
    CUtensorMap tq, tk, tv;
    if (a.pack_sh) {  // Q as [D][head][n_q][batch]: a box of 64 columns x gqa heads x 128 / gqa positions is one packed tile half
        PFN_encodeTiled enc = get_encode_tiled();
        if (!enc) return B200FA_ERR_CUDA;
        cuuint64_t dims[4] = {(cuuint64_t)Dr, (cuuint64_t)p.n_head, (cuuint64_t)p.n_q, (cuuint64_t)p.n_batch};
        cuuint64_t strides[3] = {(cuuint64_t)qnb2, (cuuint64_t)qnb1, (cuuint64_t)qnb3};
        cuuint32_t box[4] = {64, (cuuint32_t)(1 << a.pack_sh), (cuuint32_t)a.q_rows, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (enc(&tq, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(qbase), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return B200FA_ERR_CUDA;
    } else if (!make_tile_map(&tq, qbase, p.n_q, p.n_head, p.n_batch, qnb1, qnb2, qnb3, 128, Dr)) return B200FA_ERR_CUDA;
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
        cuuint32_t box[4] = {(cuuint32_t)(64 / es), (cuuint32_t)(1 << a.pack_sh), (cuuint32_t)(32 >> a.pack_sh), 1};  // a warp's 32 rows
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (enc(&to, f32o ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, n_seg > 1 ? (void*)part : p.dst, dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return B200FA_ERR_CUDA;
    }
    constexpr size_t smem_bytes = sizeof(PpShared);
    static_assert(smem_bytes <= 227 * 1024, "prefill shared memory budget");
    static const int poly = tune_env("B200FA_POLY") ? atoi(tune_env("B200FA_POLY")) : 2;  // default: every 2nd pair on the FMA pipes
#ifdef B200FA_TUNING
    auto kern = a.pack_sh ? fa_prefill_persistent<2, false, true>
              : ext ? fa_prefill_persistent<2, true>
                    : (poly == 0 ? fa_prefill_persistent<0> : (poly == 3 ? fa_prefill_persistent<3> : (poly == 4 ? fa_prefill_persistent<4> : fa_prefill_persistent<2>)));
    const int ai = a.pack_sh ? 5 : ext ? 4 : (poly == 0 ? 0 : (poly == 3 ? 2 : (poly == 4 ? 3 : 1)));
#else
    auto kern = a.pack_sh ? fa_prefill_persistent<2, false, true> : (ext ? fa_prefill_persistent<2, true> : fa_prefill_persistent<2>);  // (packing excludes the ext2 modifiers)
    const int ai = a.pack_sh ? 5 : (ext ? 4 : 1);
    (void)poly;
#endif
    static thread_local bool attr_set[64][6] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev][ai]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess)
            return B200FA_ERR_CUDA;
        attr_set[dev][ai] = true;
    }
    const unsigned grid = (unsigned)(pa.n_items < sm_count ? pa.n_items : sm_count);
    // programmatic dependent launch: the prologue (barrier init, TMEM allocation) may run while the previous kernel of the stream —
    // the mask classifier, the Q conversion, or the caller's own kernel — is still draining; the kernel waits before its first
    // global access.
    static const bool no_pdl = tune_env("B200FA_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(PF_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
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

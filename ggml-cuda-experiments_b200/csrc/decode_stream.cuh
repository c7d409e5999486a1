// decode_stream.cuh — the bandwidth-bound decode path: a persistent, stream-K split-KV kernel whose K/V
// bytes are moved by the TMA engine through a deep shared-memory ring.
//
// Replaces the reference's flash_attn_row<128,8,2,256> / flash_attn_row_fast (flash_row_float.h:4-413) AND
// its fa_reduce<128,nw> (flash_row_float.h:415-472) for every call with n_q * (n_head / n_head_kv) <= 16 rows
// per KV head (single-token decode of MHA/GQA models, short speculative bursts); GQA bursts of 17..128 rows run as virtual KV
// heads of <= 16 rows (FaParams::kv_div).
//
// Work decomposition ("stream-K"): a *unit* is one (kv head, batch) pair; its keys are cut into 64-key
// chunks; all chunks of all units form one flat list that is divided evenly over the grid (one CTA per
// SM).  A CTA therefore streams one contiguous run of chunks — possibly the tail of one unit, some whole
// units and the head of another — and the HBM stream never stalls at a unit boundary.  Per unit segment the
// CTA emits an (O~, m, l) record; the last CTA to finish a unit (arrival counter) merges the records,
// which is the reference's fa_reduce algebra done in fp32.
//
// CTA = 8 consumer warps + 4 producer warps (one elected lane each; a TMA operation costs its issuing thread ~90 ns, so the
// operations of a chunk are issued side by side).
//   producers: per chunk, wait for a free stage, then issue
//       f16  : 2*D/64 cp.async.bulk.tensor loads (64 keys x 64 dims, 128B swizzle) from the ne/nb-strided tensors, one per warp,
//       q8_0 : one box per tensor of the head's contiguous 34-byte-block rows seen as [lines][128 B] (ragged tails: 1-D bulk copies),
//       plus one 128-byte bulk copy per query row of the mask; all complete on the stage's mbarrier (one arrival per producer).
//   consumers: chunk j belongs to warp group j & 1; warp (w & 3) of the group takes keys 16*(w&3)..+15 of it,
//       reads its mma.sync fragments straight out of the swizzled stage (conflict-light 128-bit LDS), frees
//       the stage, and does QK^T, online softmax and PV in registers — the same permuted-contraction
//       fragment scheme as the rows16 kernel (decode_mma.cuh), so every lane touches 16 contiguous bytes.
// q8_0: K block scales are applied in fp32 to per-block partial dot products (bit-equivalent to dotting
// with f32(d)*q); V is dequantised to f16 = RN(d*q).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "decode_mma.cuh"
#include "sm100_ptx.cuh"

namespace b200fa {

// Diagnostics are compiled in only with -DB200FA_TUNING (profiles/timeline.py builds its own copy of the library): per-CTA and
// per-chunk %globaltimer stamps, and "stream but skip the tile maths" (FaParams::dbg_mode).
#ifdef B200FA_TUNING
#define DK_TL(cond, idx, val) do { if (a.timeline != nullptr && (cond)) a.timeline[idx] = (val); } while (0)
#define DK_DBG_SKIP_MATH (p.dbg_mode == 1)
#else
#define DK_TL(cond, idx, val) do { } while (0)
#define DK_DBG_SKIP_MATH false
#endif

constexpr int DK_CHUNK = 64;   // keys per pipeline stage
constexpr int DK_PWARPS = 4;   // producer warps: K box 0 | K box 1 | V box 0 | V box 1 + mask rows (q8_0: K | - | V | mask)
// consumer warps: groups of four (one 16-key tile of a chunk each).  T8 (the transposed q8_0 tile, below): three groups.
template <bool T8> __host__ __device__ constexpr int dk_cwarps() { return T8 ? 12 : 8; }
template <bool T8> __host__ __device__ constexpr int dk_threads() { return (dk_cwarps<T8>() + DK_PWARPS) * 32; }
constexpr int DK_REC_ROWS = 16;
constexpr int DK_REC_PAD = 4;     // record rows are D + 4 floats (16-byte aligned): O~[D], m, l, 2 unused

template <int D, bool Q8, bool T8 = false>
struct DkGeom {
    static constexpr int kRowBytes = Q8 ? D / 32 * kQ8BlockBytes : D * 2;
    static constexpr int kBoxes = D / 64;                                    // 64-dim TMA boxes per K (or V) chunk
    static constexpr int kKBytes = Q8 ? DK_CHUNK * kRowBytes : kBoxes * 8192;  // bytes of K (== V) per stage
    static constexpr int kVOff = kKBytes;
    static constexpr int kMaskOff = 2 * kKBytes;
    static constexpr int kMaskBytes = (T8 ? 8 : 16) * 128;                   // one 128-byte line per query position
    // q8_0: fragment loads are whole aligned words around a 2-byte-aligned payload, so the last row may be over-read by
    // up to 8 bytes: into V, into the mask lines, and after those into 16 bytes of padding
    static constexpr int kStageBytes = (2 * kKBytes + kMaskBytes + (Q8 ? 16 : 0) + 1023) / 1024 * 1024;
};
template <int D, int RH, bool T8 = false>
struct DkMerge {
    // floats a lane parks per fold: its O fragment + (m, l) per live row (T8: O^T tiles + rows 2t, 2t+1)
    static constexpr int kFloatsPerLane = T8 ? (D / 16) * 4 + 4 : (D / 8) * 2 * RH + 2 * RH;
    static constexpr int kSlotBytes = kFloatsPerLane * 32 * 4;
    static constexpr int kBytes = dk_cwarps<T8>() * kSlotBytes;
};
constexpr int DK_TL_CHUNK0 = 160 * 8;  // diagnostics: per-chunk stamps of CTA 0 start here (4 per chunk, first 512 chunks)
constexpr int DK_TAB = 160;           // contributor table entries (>= SM count)
constexpr int DK_TAIL_BYTES = 2 * 16 * 8 + 64 + DK_TAB * 4;  // barriers (up to 16 full + 16 empty), flag, table
constexpr int DK_SMEM_LIMIT = 227 * 1024;
template <int D, bool Q8, int RH, bool T8 = false>
__host__ __device__ constexpr int dk_stages() {
    // As deep a ring as fits beside the merge slots, at most 8 (T8: 9) — and a MULTIPLE OF THE GROUP COUNT: chunk j is consumed by
    // warp group j % groups, and a stage must always be consumed by the same group, because an mbarrier waiter may never skip a
    // phase (a group that only saw every other phase of a stage could find its parity already satisfied by the phase before).
    constexpr int groups = dk_cwarps<T8>() / 4;
    int n = (DK_SMEM_LIMIT - 1024 - DK_TAIL_BYTES - DkMerge<D, RH, T8>::kBytes) / DkGeom<D, Q8, T8>::kStageBytes;
    n = n > (T8 ? 9 : 8) ? (T8 ? 9 : 8) : n;
    return n / groups * groups;
}
// The DEEP ring: when every CTA's run is a single unit segment (unit-aligned grids: the few-unit shapes, where bytes in flight per SM
// are what bounds the stream), the fold only starts after the CTA's last chunk has been consumed, so the merge slots may alias the
// tail of the ring: the ring takes the whole shared memory.  Multi-segment runs (stream-K over many units) fold in mid-stream and
// keep the ring and the slots apart (dk_stages).  The depth is a kernel argument (DkArgs::ring); slots always start at stage dk_stages().
template <int D, bool Q8, int RH, bool T8 = false>
__host__ __device__ constexpr int dk_stages_deep() {
    constexpr int groups = dk_cwarps<T8>() / 4;
    int n = (DK_SMEM_LIMIT - 1024 - DK_TAIL_BYTES) / DkGeom<D, Q8, T8>::kStageBytes;
    n = n > 12 ? 12 : n;
    n = n / groups * groups;
    return n < dk_stages<D, Q8, RH, T8>() ? dk_stages<D, Q8, RH, T8>() : n;
}
template <int D, bool Q8, int RH, bool T8 = false>
__host__ __device__ constexpr int dk_ring_bytes() {  // ring + merge slots (aliased or not), whichever ends later
    constexpr int safe = dk_stages<D, Q8, RH, T8>() * DkGeom<D, Q8, T8>::kStageBytes + DkMerge<D, RH, T8>::kBytes;
    constexpr int deep = dk_stages_deep<D, Q8, RH, T8>() * DkGeom<D, Q8, T8>::kStageBytes;
    return safe > deep ? safe : deep;
}
template <int D, bool Q8, int RH, bool T8 = false>
__host__ __device__ constexpr int dk_smem_bytes() {
    return dk_ring_bytes<D, Q8, RH, T8>() + DK_TAIL_BYTES + 1024;
}

struct DkArgs {
    int cph;                 // chunks per unit
    int n_units;             // n_head_kv * n_batch
    long long total;         // n_units * cph
    int kv_end;              // keys [0, kv_end) of every unit are streamed (n_kv, clipped by causality)
    int max_slots;           // records a CTA may emit
    float* rec;              // [grid][max_slots][DK_REC_ROWS][D + DK_REC_PAD]
    unsigned int* counters;  // [n_units], zero between calls
    int mask_bulk;           // mask rows can be staged with 128-byte bulk copies (16-byte aligned rows)
    unsigned long long* timeline;  // diagnostics: 8 globaltimer stamps per CTA, or null
    // fused sequence-parallel step (b200fa_flash_attn_seqpar): the unit triples go straight into every rank's exchange buffer
    // over NVLink, the last CTA of this rank waits for all ranks' arrivals and merges into fdst.  null = off.
    char* const* peers;      // device array of `world` exchange-buffer pointers (peers[rank] = own)
    int rank, world;
    void* fdst;              // final merged output [rows][D]
    int fdst_type;
    int q8_lines;            // q8_0: whole 128-byte lines per head covered by the [lines][128 B] tensor maps (0 = 1-D bulk copies only)
    int cluster_k;           // > 1: the grid is launched in clusters of cluster_k CTAs = the CTAs of one unit; their records are
                             // merged through distributed shared memory (no global fence / atomic / L2 round trips)
    int deep_ring;           // every CTA's run is one unit segment: the ring may use up to dk_stages_deep stages
    int ring;                // tuning (B200FA_TUNING builds): ring depth to use, 0 = the default for the mode
};
__device__ __forceinline__ unsigned long long dk_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(ptx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
// 8 bytes at a 2-byte aligned shared address: three aligned words, funnel-shifted
__device__ __forceinline__ uint2 lds_u8x8(uint32_t a) {
    const uint32_t base = a & ~3u, sh = (a & 3u) << 3;
    const uint32_t w0 = lds32(base), w1 = lds32(base + 4), w2 = lds32(base + 8);
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}
template <int CW>
__device__ __forceinline__ void bar_consumers_n() { asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory"); }
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t x) {
    uint32_t r;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}
// D = A * B + C with C in its own registers (the QK accumulator of the transposed q8_0 tile starts at -1152 * sum(Q))
__device__ __forceinline__ void mma_16816_c(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1,
                                            float c0, float c1, float c2, float c3) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }
// address of `local` (a shared-memory address of this CTA) in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cluster_map(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

// owner CTA of flat chunk x when CTA c covers [c*T/G, (c+1)*T/G)
// (32-bit arithmetic throughout: the host only plans this kernel when total * (grid + 1) < 2^31 — DK_MAX_TOTAL_X_GRID — which every
// K/V that fits a 180 GB device satisfies except D = 64 q8_0 caches beyond 125 GB; 64-bit divisions are subroutine calls, and
// five of them in the prologue were most of a microsecond)
__device__ __forceinline__ unsigned dk_owner(unsigned x, unsigned T, unsigned G) { return ((x + 1u) * G - 1u) / T; }
constexpr long long DK_MAX_TOTAL_X_GRID = 0x7fffffffLL;

// EXT: the ext2 score modifiers (ALiBi slope on the mask, tanh soft-cap) are compiled in; the plain entry never pays for them
// T8: the transposed q8_0 tile for units of at most 8 rows (S^T = K Q^T, O^T += V^T P'^T: the quantised rows are the A operands,
// N = 8 covers all live rows, half the MMAs and half the accumulator registers of the row-major tile; see the T8 block below and
// tests/test_q8t_layout.py, which restates its register algebra lane by lane on the CPU).
// One float of a unit's triple (and, with_ml, the row's m and l) to every rank's flag-in-data area: 8-byte stores {value, tag},
// over NVLink for the peers.  Out of line and stateless (step, tag and offsets are re-derived from the exchange header on every
// call): the fused step's bookkeeping must not live in registers across the stream loop of the 128-register q8_0 kernel.
template <int D>
__device__ __noinline__ void xchg_ll_emit(char* const* peers, int rank, int world, int64_t n_floats, int64_t orow, int d, float acc,
                                          bool with_ml, float m_nat, float l_sum) {
    const unsigned int* hdr = reinterpret_cast<const unsigned int*>(peers[rank]);
    const unsigned int step = hdr[32] + 1;   // this rank's step in progress (see decode_mma.cuh)
    const unsigned int tag = xchg_ll_tag(hdr, step);
    const int64_t ll_off = xchg_ll_offset(world, n_floats);
    const int64_t off = ((int64_t)(step & 1u) * world + rank) * n_floats + orow * (D + 2);
    for (int pr = 0; pr < world; pr++) {
        uint2* out = reinterpret_cast<uint2*>(peers[pr] + ll_off) + off;
        st_ll(out + d, acc, tag);
        if (with_ml) { st_ll(out + D, m_nat, tag); st_ll(out + D + 1, l_sum, tag); }
    }
}

// The merge of a share [lo, hi) of the output elements from the flag-in-data area `part` ([rank][row][D + 2] x {value, tag}): for
// eight ranks at a time ALL of an output element's loads (its O~ value, m and l of every rank) are issued before any tag is looked
// at, so a poll costs one L2 round trip, not twenty-four dependent ones (polling element by element: ~20 us at 8 ranks).  Bounded:
// a rank that never shows up raises the exchange's error flag and false is returned (rows already merged stay written).
template <int D>
__device__ __noinline__ bool xchg_ll_merge_share(const uint2* part, int world, int64_t total_rows, int64_t lo, int64_t hi, unsigned int tag,
                                                 void* fdst, int fdst_type, unsigned int* hdr, int n_threads) {
    const unsigned int ms = hdr[kXchgTimeoutWord] ? hdr[kXchgTimeoutWord] : kXchgDefaultTimeoutMs;
    const unsigned long long t0 = xchg_now_ns(), limit = (unsigned long long)ms * 1000000ull;
    bool ok = true;
    for (int64_t idx = lo + threadIdx.x; idx < hi && ok; idx += n_threads) {
        const int64_t row = idx / D;
        const int d = (int)(idx % D);
        float M = -INFINITY, L = 0.f, acc = 0.f;
        for (int s0 = 0; s0 < world && ok; s0 += 8) {
            uint2 vd[8], vm[8], vl[8];
            unsigned int spins = 0;
            for (;;) {
                bool all = true;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (s0 + j < world) {
                        const uint2* rec = part + ((int64_t)(s0 + j) * total_rows + row) * (D + 2);
                        vd[j] = ld_ll(rec + d); vm[j] = ld_ll(rec + D); vl[j] = ld_ll(rec + D + 1);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (s0 + j < world) all = all && vd[j].y == tag && vm[j].y == tag && vl[j].y == tag;
                if (all) break;
                if ((++spins & 15u) == 0 && xchg_now_ns() - t0 > limit) { ok = false; break; }
            }
            if (!ok) break;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (s0 + j < world) {
                    const float mj = __uint_as_float(vm[j].x);
                    const float Mn = fmaxf(M, mj);
                    const float Mu = (Mn == -INFINITY) ? 0.f : Mn;
                    const float w_old = __expf(M - Mu), w_new = __expf(mj - Mu);   // exp(-inf) = 0 for empty partials
                    L = L * w_old + __uint_as_float(vl[j].x) * w_new;
                    acc = acc * w_old + __uint_as_float(vd[j].x) * w_new;
                    M = Mn;
                }
            }
        }
        if (!ok) break;
        const float y = L > 0.f ? acc / L : 0.f;
        if (fdst_type == B200FA_TYPE_F16) reinterpret_cast<__half*>(fdst)[row * D + d] = __float2half_rn(y);
        else reinterpret_cast<float*>(fdst)[row * D + d] = y;
    }
    if (!ok) atomicExch(hdr + kXchgErrWord, 1u);
    return ok;
}

// SP: the fused sequence-parallel step (b200fa_flash_attn_seqpar).  Its own instantiation: even as a cold, out-of-line path the
// exchange code cost the plain kernel 1-4 % (C5 51.8 -> 53.8 us inlined, 52.9 out of line), so the plain entry points carry none of it.
template <int D, int KV_TYPE, int RH, bool EXT = false, bool T8 = false, bool SP = false>
__global__ void __launch_bounds__(dk_threads<T8>(), 1)
fa_decode_stream(const __grid_constant__ FaParams p, const __grid_constant__ DkArgs a,
                 const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV) {
    using namespace ptx;
    static_assert(D == 64 || D == 128, "head size");
    constexpr bool Q8 = (KV_TYPE == B200FA_TYPE_Q8_0);
    static_assert(!T8 || (Q8 && RH == 1), "the transposed tile is the q8_0 kernel for <= 8 rows");
    using Geo = DkGeom<D, Q8, T8>;
    using Mrg = DkMerge<D, RH, T8>;
    constexpr int CW = dk_cwarps<T8>();  // consumer warps
    constexpr int NG = CW / 4;           // consumer groups: chunk j belongs to group j % NG
    constexpr int NS_SAFE = dk_stages<D, Q8, RH, T8>(), NS_DEEP = dk_stages_deep<D, Q8, RH, T8>();
    static_assert(NS_SAFE >= 3 && NS_SAFE % NG == 0 && NS_DEEP % NG == 0 && NS_DEEP <= 16, "ring depth must be a multiple of the group count (see dk_stages)");
#ifdef B200FA_TUNING
    const int NS = a.ring > 0 ? min(a.ring / NG * NG, a.deep_ring ? NS_DEEP : NS_SAFE) : (a.deep_ring ? NS_DEEP : NS_SAFE);
#else
    const int NS = a.deep_ring ? NS_DEEP : NS_SAFE;
#endif
    auto bar_consumers = []() { bar_consumers_n<CW>(); };
    constexpr int NC4 = D / 32;  // 16-byte chunks per lane per K row == q8_0 blocks per row
    constexpr int NCV = D / 64;  // 64-wide halves of a V row
    constexpr int NT = D / 8;    // output n-tiles
    constexpr int RLIVE = 8 * RH; // rows a CTA can hold
    using Tile = KVTile<D, Q8>;

    extern __shared__ uint8_t dk_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dk_smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t stages_u32 = smem_u32(smem);
    float* merge = reinterpret_cast<float*>(smem + NS_SAFE * Geo::kStageBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + dk_ring_bytes<D, Q8, RH, T8>());
    uint64_t* empty = full + 16;
    int* s_flag = reinterpret_cast<int*>(empty + 16);
    int* s_tab = s_flag + 16;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned G = gridDim.x, total = (unsigned)a.total;
    const unsigned start = blockIdx.x * total / G, stop = (blockIdx.x + 1u) * total / G;
    const int my_chunks = (int)(stop - start);
    // Programmatic dependent launch (the host opts in per launch): the NEXT kernel of the stream may be scheduled onto SMs as this
    // grid's CTAs retire, and runs its prologue there; every thread of this kernel in turn waits (griddepcontrol.wait, below) for
    // the previous grid to have completed and flushed before it touches global memory.  Both are no-ops on a plain launch.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    DK_TL(threadIdx.x == 0, blockIdx.x * 8 + 7, dk_now());  // kernel entry
    if (warp == 0 && my_chunks > 0) {
        // Ask L2 for the Q rows of the CTA's first unit before anything else happens.  Once the producers' opening burst (the
        // whole ring: 20-30 MB over the chip) is queued in DRAM, a Q line that misses L2 comes back BEHIND it — measured 3-7 us
        // after kernel entry, with landed stages waiting for their consumers all that time.  (L2 is the coherence point, so
        // a prefetch ahead of griddepcontrol.wait is harmless.)
        const int u0 = (int)(start / (unsigned)a.cph), ik2_0 = u0 % p.n_head_kv, iq3_0 = u0 / p.n_head_kv;
        const int row_bytes = p.Dr * (p.q_type == B200FA_TYPE_F16 ? 2 : 4);
        for (int R = lane >> 2; R < p.n_q * p.gqa; R += 8) {
            const char* qrow = p.q + (int64_t)(R / p.gqa) * p.nb01 + (int64_t)(ik2_0 * p.gqa + R % p.gqa) * p.nb02 + (int64_t)iq3_0 * p.nb03;
            if ((lane & 3) * 128 < row_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(qrow + (lane & 3) * 128));
        }
    }

    if (warp >= CW) {
        // ===================== producers =====================
        // Every TMA / bulk operation costs its issuing thread ~90 ns (measured: +10 us on the C5 shape for one extra 128-byte copy
        // per chunk): a single issuer (5 operations per f16 chunk) was a bottleneck of the stream.  Four warps issue side by side,
        // one operation each: K box 0 | K box 1 | V box 0 | V box 1 and the mask rows (q8_0: K | nothing | V | mask rows); each
        // arms the chunk's barrier with its own byte count.  (Lanes of ONE warp issuing side by side measured slower.)
        // q8_0 has one operation per tensor and chunk, and chunks of half the bytes: the four warps are K even chunks | K odd chunks |
        // V even | V odd (each K warp also copies its chunk's mask rows), so an issuing thread has two chunk periods per iteration.
        const int which = warp - CW;
        const bool isV = which >= 2;   // first two producer warps: K;  last two: V
        const int box = Q8 ? 0 : (which & 1);   // f16: the 64-dim box of the row this warp loads
        constexpr int PSTRIDE = Q8 ? 2 : 1;     // chunks between two issues of one producer
        constexpr int FULL_ARRIVALS = Q8 ? 2 : DK_PWARPS;  // producers that arm a chunk's barrier
        const int pfirst = Q8 ? (which & 1) : 0;
        if (which == 0 && lane == 0) {
            for (int s = 0; s < NS_DEEP; s++) { mbar_init(&full[s], FULL_ARRIVALS); mbar_init(&empty[s], 4); }
            fence_barrier_init();
        }
        if (lane == 0 && (!Q8 || a.q8_lines > 0)) prefetch_tensormap(isV ? &tmV : &tmK);
        asm volatile("bar.sync 2, %0;" ::"n"(DK_PWARPS * 32) : "memory");  // the barriers exist before the other producers touch them
        asm volatile("bar.arrive 3, %0;" ::"n"(dk_threads<T8>()) : "memory");  // ... and tell the consumers so, without waiting for them
        asm volatile("griddepcontrol.wait;" ::: "memory");
        // The per-chunk path of an issuing thread is short on purpose: one thread runs it serially for every chunk of the CTA, and a
        // chunk lasts ~400 ns at the HBM rate (q8_0).  Everything that depends only on the unit (integer divisions by run-time values:
        // ~150 dependent instructions, which used to cost ~400 ns per chunk and bounded the q8_0 stream at 4.5 TB/s) is recomputed
        // only when the run crosses into the next unit.
        int pu = (int)((start + pfirst) / (unsigned)a.cph), pch = (int)(start + pfirst - (unsigned)pu * (unsigned)a.cph);  // unit / chunk of the next issue
        int ik2 = 0, ik3 = 0;  // the REAL kv head / kv batch whose K/V the (possibly virtual) unit pu streams
        int m_head0 = 0;       // EXT, per-head masks: first q head of the unit (its rows' heads are m_head0 + R % gqa)
        int64_t m_unit_off = 0;  // EXT, per-batch masks: byte offset of the unit's batch slice
        auto set_unit = [&](int u) {
            const int iq3 = u / p.n_head_kv;
            ik3 = iq3 / p.rk3;
            ik2 = (u - iq3 * p.n_head_kv) / p.kv_div;
            if constexpr (EXT) {
                m_head0 = (u - iq3 * p.n_head_kv) * p.gqa;
                m_unit_off = p.m_ne3 > 1 ? (int64_t)iq3 * p.nb33 : 0;
            }
        };
        set_unit(pu);
        const int q8_box_chunks = (Q8 && a.q8_lines > 0) ? (int)(((int64_t)a.q8_lines * 128) / Geo::kKBytes) : 0;  // chunks of a head inside the line maps
        const int whole_chunks = p.n_kv / DK_CHUNK;                                        // chunks with all 64 keys
        // one staged 128-byte mask line per query position — or, with per-head masks (EXT, m_ne2 > 1), per row of the unit
        const bool m_per_row = EXT && p.m_ne2 > 1;
        const int mrows_full = ((Q8 ? !isV : which == DK_PWARPS - 1) && a.mask_bulk) ? (m_per_row ? p.n_q * p.gqa : p.n_q) : 0;
        const CUtensorMap* tm = isV ? &tmV : &tmK;
        const bool has_box = box < Geo::kBoxes;                  // (D = 64, f16: one box per tensor)
        const uint32_t half_off = isV ? Geo::kVOff : 0;          // this producer's half of a stage
        int pstage = pfirst % NS, pphase = 1;  // stage of the next issue; parity of its `empty` barrier that means "free" (first lap: free at once)
        auto issue = [&]() {
            const int ch = pch;
            const int key0 = ch * DK_CHUNK;
            const int stage = pstage;
            pstage += PSTRIDE;
            if (pstage >= NS) { pstage -= NS; pphase ^= 1; }
            const uint32_t sb = stages_u32 + stage * Geo::kStageBytes + half_off;
            uint8_t* sp = smem + stage * Geo::kStageBytes + half_off;
            const int mrows = ch < whole_chunks ? mrows_full : 0;
            if (!has_box) {
                mbar_arrive_expect_tx(&full[stage], mrows * 128);
            } else if constexpr (!Q8) {
                mbar_arrive_expect_tx(&full[stage], 8192 + mrows * 128);
                tma_load_4d(sp + box * 8192, tm, &full[stage], 64 * box, key0, ik2, ik3);
            } else if (ch < q8_box_chunks) {
                // q8_0, chunk inside the whole-128-byte-line part of the head: the head's rows are one contiguous byte range, which
                // the tensor maps describe as [lines][128 B] — one box of kKBytes/128 lines per K and per V chunk (a little faster
                // than 1-D bulk copies of the same bytes: 71 vs 74-81 us of pure streaming on the C5 shape).
                mbar_arrive_expect_tx(&full[stage], Geo::kKBytes + mrows * 128);
                tma_load_4d(sp, tm, &full[stage], 0, ch * (Geo::kKBytes / 128), ik2, ik3);
            } else {
                // ragged or unaligned tail of a q8_0 head: a 1-D bulk copy of the 16-byte multiple, the last few words by hand
                const int rows = min(DK_CHUNK, p.n_kv - key0);
                const uint32_t nbytes = (uint32_t)rows * Geo::kRowBytes, nb16 = nbytes & ~15u;
                const char* src = !isV ? p.k + (int64_t)ik2 * p.nb12 + (int64_t)ik3 * p.nb13 + (int64_t)key0 * Geo::kRowBytes
                                       : p.v + (int64_t)ik2 * p.nb22 + (int64_t)ik3 * p.nb23 + (int64_t)key0 * Geo::kRowBytes;
                for (uint32_t o = nb16; o < nbytes; o += 4) *reinterpret_cast<uint32_t*>(sp + o) = __ldg(reinterpret_cast<const uint32_t*>(src + o));
                mbar_arrive_expect_tx(&full[stage], nb16 + mrows * 128);
                if (nb16 > 0) bulk_g2s(sb, src, nb16, &full[stage]);
            }
            if (!m_per_row) {
                for (int r = 0; r < mrows; r++)
                    bulk_g2s(stages_u32 + stage * Geo::kStageBytes + Geo::kMaskOff + r * 128, p.mask + (EXT ? m_unit_off : 0) + (int64_t)r * p.nb31 + (int64_t)key0 * 2, 128, &full[stage]);
            } else {  // line R = row R of the unit: query position R / gqa, head m_head0 + R % gqa
                int iq1 = 0, hq = 0;
                for (int r = 0; r < mrows; r++) {
                    bulk_g2s(stages_u32 + stage * Geo::kStageBytes + Geo::kMaskOff + r * 128,
                             p.mask + m_unit_off + (int64_t)iq1 * p.nb31 + (int64_t)(m_head0 + hq) * p.nb32 + (int64_t)key0 * 2, 128, &full[stage]);
                    if (++hq == p.gqa) { hq = 0; iq1++; }
                }
            }
            pch += PSTRIDE;
            if (pch >= a.cph) {
                do { pch -= a.cph; pu++; } while (pch >= a.cph);
                set_unit(pu);
            }
        };
        if (lane == 0) {
            for (int i = pfirst; i < my_chunks; i += PSTRIDE) {
                if (i >= NS) mbar_wait(&empty[pstage], pphase);  // the first lap needs no wait: the ring starts filling at once
                DK_TL(blockIdx.x == 0 && which == 0 && i < 512, DK_TL_CHUNK0 + i * 4 + 0, dk_now());  // stage free again
                issue();
                DK_TL(blockIdx.x == 0 && which == 0 && i < 512, DK_TL_CHUNK0 + i * 4 + 1, dk_now());  // operations issued
            }
        }
        __syncwarp();
        if (a.cluster_k > 1) { cluster_sync_all(); cluster_sync_all(); }  // every thread of the cluster takes part in both cluster barriers
        DK_TL(threadIdx.x == CW * 32, blockIdx.x * 8 + 6, dk_now());  // producer warp 0 leaves
        return;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ===================== consumers =====================
    // (They meet the producers' "mbarriers are initialised" signal, bar.sync 3, only right before their first wait on a stage: the
    // Q rows of the first unit are requested before that.  Measured: once the producers' opening burst — the whole ring, 20-30 MB
    // over the chip — is queued in DRAM, a Q line that misses L2 comes back BEHIND it, 3-7 us later, and the consumers sat idle
    // on landed stages for that long.)
    const int g = lane >> 2, t = lane & 3;
    const int group = warp >> 2, sub = warp & 3;
    const int rows_total = p.n_q * p.gqa;
    const bool mask_al8 = p.mask != nullptr && ((((uintptr_t)p.mask | (uintptr_t)p.nb31 | (uintptr_t)p.nb32 | (uintptr_t)p.nb33) & 7) == 0);

    // Full accumulator quads also when only fragment rows g are live (RH == 1): rows 8-15 of A are fed zeros, so entries 2-3 stay
    // what they were (zero) and the HMMA accumulates in place — no per-instruction re-zeroing or moves to build its C/D quad.
    // The lane's rows: fragment rows g (+ 8) of the row-major tile; accumulator columns 2t, 2t+1 of the transposed one (T8).
    constexpr int NR = T8 ? 2 : RH;
    constexpr int NB = D / 32;   // q8_0 blocks per row
    constexpr int NMT = D / 16;  // T8: 16-dim m-tiles of O^T
    float o[T8 ? 1 : NT][4];
    float m_run[NR], l_run[NR];
    uint32_t qa[T8 ? 1 : NC4][RH][4];
    // T8 state: O^T tiles (c0/c1 = dim(mt, g) of rows 2t/2t+1, c2/c3 = dim(mt, g + 8)), Q as B fragments (row g, dims 32b + 8t..+7),
    // and -1152 x the block sums of Q rows 2t, 2t+1: the C operand that cancels the bias left in the converted K bytes
    float oT[T8 ? NMT : 1][4];
    uint32_t qb[T8 ? NB : 1][4];
    float qsn[T8 ? NB : 1][2];
    int iq1r[NR], rq[NR];   // query position / q head within the GQA group of the lane's rows
    float mslope[NR];       // log2(e) x ALiBi slope of the row's head (set per unit): the factor on raw mask values
    bool rvalid[NR];
    int lim[NR];
    const char* mrow[NR];
    int mline[NR];          // staged mask line of the row: its query position, or the row itself with per-head masks (EXT)
#pragma unroll
    for (int h = 0; h < NR; h++) {
        const int R = T8 ? 2 * t + h : g + 8 * h;
        rvalid[h] = R < rows_total;
        iq1r[h] = (rvalid[h] ? R : 0) / p.gqa;
        rq[h] = (rvalid[h] ? R : 0) % p.gqa;
        const int64_t vis = p.causal ? (int64_t)iq1r[h] + p.causal_off - p.kv_pos0 + 1 : (int64_t)p.n_kv;
        lim[h] = (int)max((int64_t)0, min((int64_t)a.kv_end, vis));
        mrow[h] = p.mask ? p.mask + (int64_t)iq1r[h] * p.nb31 : nullptr;
        mline[h] = (EXT && p.m_ne2 > 1) ? (rvalid[h] ? R : 0) : iq1r[h];
        mslope[h] = kLog2e;
    }

    // S = Q K^T (fp32) for the lane's 2 x 4 score slots
    auto qk_tile = [&](const Tile& T, float (&s)[2][4]) {
        if constexpr (!T8) {  // the row-major tile (f16 K/V, and q8_0 units of 9-16 rows)
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            if constexpr (Q8) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
            for (int c = 0; c < NC4; c++) {
                uint32_t k0, k1, k2, k3;
                if constexpr (!Q8) {
                    k0 = T.kf[nt][c][0]; k1 = T.kf[nt][c][1]; k2 = T.kf[nt][c][2]; k3 = T.kf[nt][c][3];
                } else {
                    q8x4_to_h2(T.kf[nt][c][0], k0, k1);
                    q8x4_to_h2(T.kf[nt][c][1], k2, k3);
                }
                float acc0[4];
                float(&acc)[4] = Q8 ? acc0 : s[nt];
                const uint32_t q1a = RH == 2 ? qa[c][RH - 1][0] : 0u, q1b = RH == 2 ? qa[c][RH - 1][1] : 0u;
                const uint32_t q1c = RH == 2 ? qa[c][RH - 1][2] : 0u, q1d = RH == 2 ? qa[c][RH - 1][3] : 0u;
                if (Q8 || c == 0) mma_16816_zc(acc, qa[c][0][0], q1a, qa[c][0][1], q1b, k0, k1);  // fresh block sum / first block: C = 0
                else mma_16816(acc, qa[c][0][0], q1a, qa[c][0][1], q1b, k0, k1);
                mma_16816(acc, qa[c][0][2], q1c, qa[c][0][3], q1d, k2, k3);
                if constexpr (Q8) {
                    const float d0 = h_bits_to_f(T.kd[2 * nt][c]), d1 = h_bits_to_f(T.kd[2 * nt + 1][c]);
                    s[nt][0] += acc0[0] * d0; s[nt][1] += acc0[1] * d1;
                    if constexpr (RH == 2) { s[nt][2] += acc0[2] * d0; s[nt][3] += acc0[3] * d1; }
                }
            }
        }
        }
    };
    // scale, mask, online softmax (lane holds keys kv0+4t+j, j=0..3, of its rows), then O += P V
    auto softmax_pv_tile = [&](const Tile& T, const float (&s)[2][4], int kv0) {
        if constexpr (!T8) {  // the row-major tile (f16 K/V, and q8_0 units of 9-16 rows)
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
                    const __half2 d2 = __half2half2(__ushort_as_half(T.vd[i][c]));
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
                mma_16816(o[c * 8 + j], pa0, pa1, pa2, pa3, b0, b1);  // RH == 1: pa1 = pa3 = 0
            }
        }
        }
    };

    // Per-lane byte offsets into a stage, fixed for the whole kernel: every fragment load below is stage base + one of
    // these + an immediate.  f16 stages are TMA boxes [64 keys][64 dims] with the 128-byte swizzle (16-byte chunk c of
    // row r sits at chunk position c ^ (r & 7)); q8_0 stages are the raw 34-byte blocks, rows kRowBytes apart.
    const int r0 = 16 * sub;
    uint32_t kA[2][2], vA[4];   // f16: K rows rho(g) for n-tile nt / chunk parity; V rows 4t+i
    uint32_t kB[2], kdB = 0, vB = 0, vdB = 0, vsh = 0;  // q8_0
    if constexpr (!Q8) {
        const int gb = (g >> 1) & 1, gp = g ^ (4 * (t & 1));
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            const int row = r0 + 4 * (g >> 1) + 2 * nt + (g & 1);
            const int low = t ^ (2 * nt + (g & 1));
#pragma unroll
            for (int par = 0; par < 2; par++) kA[nt][par] = row * 128 + (((par ^ gb) * 4 + low) << 4);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) vA[i] = Geo::kVOff + (r0 + 4 * t + i) * 128 + ((gp ^ i) << 4);
        kB[0] = kB[1] = 0;
    } else {
#pragma unroll
        for (int nt = 0; nt < 2; nt++) kB[nt] = (r0 + 4 * (g >> 1) + 2 * nt + (g & 1)) * Geo::kRowBytes + 8 * t;
        kdB = (r0 + 4 * t) * Geo::kRowBytes;
        const uint32_t vl = (g >> 2) * kQ8BlockBytes + 2 + 8 * (g & 3);
        vB = Geo::kVOff + kdB + (vl & ~3u);
        vsh = (vl & 3u) << 3;
        vdB = Geo::kVOff + kdB + (g >> 2) * kQ8BlockBytes;
        kA[0][0] = kA[0][1] = kA[1][0] = kA[1][1] = 0; vA[0] = vA[1] = vA[2] = vA[3] = 0;
    }

    // fragments of one 16-key sub-tile out of a landed stage: K (+ K scales, mask) first, V once the scores are issued
    auto read_k = [&](Tile& T, uint32_t sb, int kv0, bool staged_mask) {
        if constexpr (!T8) {  // the row-major tile (f16 K/V, and q8_0 units of 9-16 rows)
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
#pragma unroll
            for (int c = 0; c < NC4; c++) {
                if constexpr (!Q8) {
                    const uint4 x = lds128(sb + kA[nt][c & 1] + (c >> 1) * 8192);
                    T.kf[nt][c][0] = x.x; T.kf[nt][c][1] = x.y; T.kf[nt][c][2] = x.z; T.kf[nt][c][3] = x.w;
                } else if ((c & 1) == 0) {  // payload at 34c + 2: two bytes past a word boundary
                    const uint32_t b = sb + kB[nt] + c * kQ8BlockBytes;
                    const uint32_t w0 = lds32(b), w1 = lds32(b + 4), w2 = lds32(b + 8);
                    T.kf[nt][c][0] = prmt(w0, w1, 0x5432); T.kf[nt][c][1] = prmt(w1, w2, 0x5432);
                } else {                    // word aligned
                    const uint32_t b = sb + kB[nt] + c * kQ8BlockBytes + 2;
                    T.kf[nt][c][0] = lds32(b); T.kf[nt][c][1] = lds32(b + 4);
                }
            }
        }
        if constexpr (Q8) {
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int c = 0; c < NC4; c++) T.kd[i][c] = lds_u16(sb + kdB + i * Geo::kRowBytes + c * kQ8BlockBytes);
        }
        if (p.mask != nullptr) {
#pragma unroll
            for (int h = 0; h < RH; h++) {
                if (staged_mask) {
                    T.mk[h] = lds64(sb + Geo::kMaskOff + mline[h] * 128 + (r0 + 4 * t) * 2);
                } else if (mask_al8 && kv0 + 16 <= p.n_kv) {
                    T.mk[h] = __ldg(reinterpret_cast<const uint2*>(mrow[h] + (int64_t)(kv0 + 4 * t) * 2));
                } else {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) w[j] = (kv0 + 4 * t + j < p.n_kv) ? ld_u16(mrow[h] + (int64_t)(kv0 + 4 * t + j) * 2) : 0u;
                    T.mk[h] = make_uint2(w[0] | (w[1] << 16), w[2] | (w[3] << 16));
                }
            }
        }
        }
    };
    auto read_v = [&](Tile& T, uint32_t sb, int kv0) {
        if constexpr (!T8) {  // the row-major tile (f16 K/V, and q8_0 units of 9-16 rows)
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int c = 0; c < NCV; c++) {
                if constexpr (!Q8) {
                    const uint4 x = lds128(sb + vA[i] + c * 8192);
                    T.vf[i][c][0] = x.x; T.vf[i][c][1] = x.y; T.vf[i][c][2] = x.z; T.vf[i][c][3] = x.w;
                } else {
                    const uint32_t b = sb + vB + i * Geo::kRowBytes + c * 2 * kQ8BlockBytes;
                    const uint32_t w0 = lds32(b), w1 = lds32(b + 4), w2 = lds32(b + 8);
                    T.vf[i][c][0] = __funnelshift_r(w0, w1, vsh); T.vf[i][c][1] = __funnelshift_r(w1, w2, vsh);
                    T.vd[i][c] = lds_u16(sb + vdB + i * Geo::kRowBytes + c * 2 * kQ8BlockBytes);
                }
            }
        }
        }
    };

    // Slot layout: [k][lane] floats.  k < KO: the lane's O fragment, k = (n-tile * 2*RH + e) with e = 2h + e1, which is
    // row g + 8h, dim 64c + 16t + j + 8*e1 for n-tile c*8 + j;  k = KO + 2h / + 2h + 1: that row's m / l.
    constexpr int KO = T8 ? NMT * 4 : NT * 2 * RH;
    auto slot_st = [&]() {
        if constexpr (!T8) {  // the row-major tile (f16 K/V, and q8_0 units of 9-16 rows)
        float* sp = merge + warp * (Mrg::kSlotBytes / 4) + lane;
        int k = 0;
#pragma unroll
        for (int n = 0; n < NT; n++)
#pragma unroll
            for (int e = 0; e < 2 * RH; e++) sp[(k++) * 32] = o[n][e];
#pragma unroll
        for (int h = 0; h < RH; h++) { sp[(k++) * 32] = m_run[h]; sp[(k++) * 32] = l_run[h]; }
        }
    };

    // ===================== the transposed q8_0 tile (T8) =====================
    // One warp, 16 keys.  Contraction / accumulator slot s = 8h + 2t' + j of the tile is key 4t' + 2h + j, so that
    //   * the lane's score slots g and g + 8 are keys keyA = 4(g >> 1) + (g & 1) and keyA + 2,
    //   * after movmatrix the lane's P' fragment covers keys 4t .. 4t+3: four consecutive V rows.
    // S^T[key][row] = sum_b dK[key][b] * (sum_{i in b} (1152 + k_i) q_i - 1152 sum_{i in b} q_i): two MMAs per 32-dim block whose
    // A operands are the K bytes converted with ONE byte permute per pair (0x64 high bytes make 1024 + (byte ^ 0x80) = 1152 + k in
    // f16; the subtraction the row-major tile spends an HADD2 per pair on is the accumulator's initial value here), lane t of row g
    // taking payload bytes 8t..8t+7 of the block (the contraction index is permuted the same way in Q's B fragments).
    // O^T[dim][row] += sum_key v[key][dim] * (p[key][row] dV[key][b]): P' = p * dV per block, rounded to f16, transposed by
    // movmatrix into the B fragment; the A operand pairs the bytes of two consecutive keys for one dim (one permute + one logic op
    // + the bias subtraction per pair).  m-tile mt, fragment row m is head dim 32(mt>>1) + 4(m&7) + 2(mt&1) + (m>>3).
    // The running max is lazy: it is raised (warp-wide butterflies, O^T rescaled) only when some score exceeds it by 2^kLazy,
    // so p <= 2^kLazy and the common tile costs one vote instead of six shuffles.
    constexpr float kLazy = 6.f;
    const int keyA = 4 * (g >> 1) + (g & 1);
    const uint32_t t8_ka = (uint32_t)((16 * sub + keyA) * Geo::kRowBytes + 8 * t);   // K row keyA: block 0 scale; payload bytes + 2
    const uint32_t t8_va = (uint32_t)(Geo::kVOff + (16 * sub + 4 * t) * Geo::kRowBytes + 4 * g);  // V rows 4t..: payload word + 2
    const uint32_t t8_sa = (uint32_t)((16 * sub + keyA) * Geo::kRowBytes);           // scales of rows keyA (+ 2 rows: keyB)
    auto t8_tile = [&](uint32_t sb, int key0, bool staged_mask, uint64_t* empty_bar) {
        if constexpr (T8) {
        const int kvA = key0 + 16 * sub + keyA, kvB = kvA + 2;
        const bool vA_ok = kvA < p.n_kv, vB_ok = kvB < p.n_kv;  // rows past a ragged end hold stale bytes: a zero scale keeps them out
        // V operands of the tile: aligned words holding payload bytes 0,1 / 2,3 of the lane's word of keys 4t + i, and the block
        // scales of keys keyA / keyB.  Where they are loaded decides how long the stage is held (B200FA_T8_VLOAD: 0 = per block
        // inside the P.V loop, 1 = after the scores, before the softmax, 2 = before anything else).
        uint32_t vlo[NB][4], vhi[NB][4], vsc[NB][2];
        auto load_v = [&](int b) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t ad = sb + t8_va + i * Geo::kRowBytes + b * kQ8BlockBytes;  // 2 bytes before the payload word
                if ((b & 1) == 0) { vlo[b][i] = lds32(ad); vhi[b][i] = lds32(ad + 4); }
                else vlo[b][i] = vhi[b][i] = lds32(ad + 2);
            }
            vsc[b][0] = vA_ok ? lds_u16(sb + Geo::kVOff + t8_sa + b * kQ8BlockBytes) : 0u;
            vsc[b][1] = vB_ok ? lds_u16(sb + Geo::kVOff + t8_sa + 2 * Geo::kRowBytes + b * kQ8BlockBytes) : 0u;
        };
        auto release = [&]() {  // every byte of the stage has been read: hand it back
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar);
        };
#ifndef B200FA_T8_VLOAD
#define B200FA_T8_VLOAD 0
#endif
        if constexpr (B200FA_T8_VLOAD == 2) {
#pragma unroll
            for (int b = 0; b < NB; b++) load_v(b);
        }
        // ---- S^T = K Q^T ----
        float s[2][2] = {{0.f, 0.f}, {0.f, 0.f}};  // [key slot g / g+8][row 2t / 2t+1]
#pragma unroll
        for (int b = 0; b < NB; b++) {
            uint32_t h[2][4];  // converted pairs of rows keyA / keyB
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const uint32_t ad = sb + t8_ka + r * 2 * Geo::kRowBytes + b * kQ8BlockBytes;  // 2 bytes before the lane's 8 payload bytes
                if ((b & 1) == 0) {  // payload at 2 mod 4: bytes 2,3 of w0, all of w1, bytes 0,1 of w2
                    const uint32_t w0 = lds32(ad) ^ 0x80808080u, w1 = lds32(ad + 4) ^ 0x80808080u, w2 = lds32(ad + 8) ^ 0x80808080u;
                    h[r][0] = prmt(w0, 0x64646464u, 0x4342); h[r][1] = prmt(w1, 0x64646464u, 0x4140);
                    h[r][2] = prmt(w1, 0x64646464u, 0x4342); h[r][3] = prmt(w2, 0x64646464u, 0x4140);
                } else {             // word aligned
                    const uint32_t w0 = lds32(ad + 2) ^ 0x80808080u, w1 = lds32(ad + 6) ^ 0x80808080u;
                    h[r][0] = prmt(w0, 0x64646464u, 0x4140); h[r][1] = prmt(w0, 0x64646464u, 0x4342);
                    h[r][2] = prmt(w1, 0x64646464u, 0x4140); h[r][3] = prmt(w1, 0x64646464u, 0x4342);
                }
            }
            float c[4];
            mma_16816_c(c, h[0][0], h[1][0], h[0][1], h[1][1], qb[b][0], qb[b][1], qsn[b][0], qsn[b][1], qsn[b][0], qsn[b][1]);
            mma_16816(c, h[0][2], h[1][2], h[0][3], h[1][3], qb[b][2], qb[b][3]);
            const float dA = h_bits_to_f(lds_u16(sb + t8_sa + b * kQ8BlockBytes));
            const float dB = h_bits_to_f(lds_u16(sb + t8_sa + 2 * Geo::kRowBytes + b * kQ8BlockBytes));
            s[0][0] = fmaf(c[0], dA, s[0][0]); s[0][1] = fmaf(c[1], dA, s[0][1]);
            s[1][0] = fmaf(c[2], dB, s[1][0]); s[1][1] = fmaf(c[3], dB, s[1][1]);
        }
        if constexpr (B200FA_T8_VLOAD == 1) {
#pragma unroll
            for (int b = 0; b < NB; b++) load_v(b);
        }
        uint32_t mkh[2][2] = {{0u, 0u}, {0u, 0u}};  // mask halves [row][keyA / keyB]
        if (p.mask != nullptr) {
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                if (staged_mask) {
                    const uint32_t ma = sb + Geo::kMaskOff + mline[rr] * 128 + (16 * sub + keyA) * 2;
                    mkh[rr][0] = lds_u16(ma); mkh[rr][1] = lds_u16(ma + 4);
                } else {
                    mkh[rr][0] = kvA < p.n_kv ? ld_u16(mrow[rr] + (int64_t)kvA * 2) : 0u;
                    mkh[rr][1] = kvB < p.n_kv ? ld_u16(mrow[rr] + (int64_t)kvB * 2) : 0u;
                }
            }
        }
        if constexpr (B200FA_T8_VLOAD != 0) release();
        // ---- scale, mask, lazy online softmax ----
        float x[2][2];
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const float ms = EXT ? mslope[rr] : kLog2e;
            const float mA = h_bits_to_f(mkh[rr][0]) * ms, mB = h_bits_to_f(mkh[rr][1]) * ms;
            if (EXT && p.cap_in != 0.f) {
                x[0][rr] = fmaf(fa_tanh(s[0][rr] * p.cap_in), p.cap_out, mA);
                x[1][rr] = fmaf(fa_tanh(s[1][rr] * p.cap_in), p.cap_out, mB);
            } else {
                x[0][rr] = fmaf(s[0][rr], p.scale_log2, mA);
                x[1][rr] = fmaf(s[1][rr], p.scale_log2, mB);
            }
            if (kvA >= lim[rr]) x[0][rr] = -INFINITY;
            if (kvB >= lim[rr]) x[1][rr] = -INFINITY;
        }
        const float ml0 = fmaxf(x[0][0], x[1][0]), ml1 = fmaxf(x[0][1], x[1][1]);
        if (__any_sync(0xffffffffu, ml0 > m_run[0] + kLazy || ml1 > m_run[1] + kLazy)) {
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                float mx = rr == 0 ? ml0 : ml1;
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
                const float m_new = fmaxf(m_run[rr], mx);
                const float alpha = fast_exp2(m_run[rr] - ((m_new == -INFINITY) ? 0.f : m_new));  // m_run = -inf -> 0
                l_run[rr] *= alpha;
#pragma unroll
                for (int mt = 0; mt < NMT; mt++) { oT[mt][rr] *= alpha; oT[mt][2 + rr] *= alpha; }
                m_run[rr] = m_new;
            }
        }
        float pp[2][2];
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const float mu = (m_run[rr] == -INFINITY) ? 0.f : m_run[rr];
            pp[0][rr] = fast_exp2(x[0][rr] - mu);
            pp[1][rr] = fast_exp2(x[1][rr] - mu);
            l_run[rr] += pp[0][rr] + pp[1][rr];
        }
        // ---- O^T += V^T P'^T ----
        const uint32_t ph[2] = {pack_half2(pp[0][0], pp[0][1]), pack_half2(pp[1][0], pp[1][1])};  // (rows 2t, 2t+1) of keys keyA, keyB
#pragma unroll
        for (int b = 0; b < NB; b++) {
            // P' = RN_f16(RN_f16(p) * dV): one packed multiply per key with the scale broadcast to both halves
            if constexpr (B200FA_T8_VLOAD == 0) {
                load_v(b);
                if (b == NB - 1) release();
            }
            const __half dA = __ushort_as_half((unsigned short)vsc[b][0]), dB = __ushort_as_half((unsigned short)vsc[b][1]);
            const __half2 qA = __hmul2(*reinterpret_cast<const __half2*>(&ph[0]), __half2half2(dA));
            const __half2 qB = __hmul2(*reinterpret_cast<const __half2*>(&ph[1]), __half2half2(dB));
            const uint32_t bf0 = movmatrix_trans(*reinterpret_cast<const uint32_t*>(&qA));
            const uint32_t bf1 = movmatrix_trans(*reinterpret_cast<const uint32_t*>(&qB));
            const uint32_t(&lo)[4] = vlo[b];
            const uint32_t(&hi)[4] = vhi[b];
            const uint32_t s0 = (b & 1) == 0 ? 0x6622u : 0x4400u, s1 = (b & 1) == 0 ? 0x7733u : 0x5511u;  // bytes 0,1 of the payload word
            const uint32_t s2 = (b & 1) == 0 ? 0x4400u : 0x6622u, s3 = (b & 1) == 0 ? 0x5511u : 0x7733u;  // bytes 2,3
            auto cv = [](uint32_t z) {  // (byte, byte) of two keys -> (q, q) in f16
                uint32_t y;
                const uint32_t bias = 0x64806480u;
                asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(y) : "r"(z), "r"(0x00FF00FFu), "r"(bias));  // (z & 0x00ff00ff) ^ 0x64806480
                __half2 r = __hsub2(*reinterpret_cast<const __half2*>(&y), *reinterpret_cast<const __half2*>(&bias));
                return *reinterpret_cast<uint32_t*>(&r);
            };
            mma_16816(oT[2 * b], cv(prmt(lo[0], lo[1], s0)), cv(prmt(lo[0], lo[1], s1)), cv(prmt(lo[2], lo[3], s0)), cv(prmt(lo[2], lo[3], s1)), bf0, bf1);
            mma_16816(oT[2 * b + 1], cv(prmt(hi[0], hi[1], s2)), cv(prmt(hi[0], hi[1], s3)), cv(prmt(hi[2], hi[3], s2)), cv(prmt(hi[2], hi[3], s3)), bf0, bf1);
        }
        }
    };
    auto t8_slot_st = [&]() {  // [k][lane]: k < KO: oT[k >> 2][k & 3];  KO + 2rr: m of row 2t + rr;  KO + 2rr + 1: its l
        if constexpr (T8) {
        float* sp = merge + warp * (Mrg::kSlotBytes / 4) + lane;
#pragma unroll
        for (int mt = 0; mt < NMT; mt++)
#pragma unroll
            for (int e = 0; e < 4; e++) sp[(mt * 4 + e) * 32] = oT[mt][e];
#pragma unroll
        for (int rr = 0; rr < 2; rr++) { sp[(KO + 2 * rr) * 32] = m_run[rr]; sp[(KO + 2 * rr + 1) * 32] = l_run[rr]; }
        }
    };
    DK_TL(threadIdx.x == 0, blockIdx.x * 8 + 0, dk_now());

    // ---- fused sequence-parallel step: where the unit triples go, and what the last CTA of the rank does ----
    // Flag-in-data exchange: every float of a unit's triple goes to every rank as one 8-byte store {value, tag of this step} into
    // the rank's flag-in-data area (decode_mma.cuh); nothing is fenced and nothing is counted.  A reader polls the elements it
    // needs until their tags match, which is all the synchronisation there is: one NVLink traversal between the last local
    // chunk and the merge instead of stores + system fence + counter increment + counter poll.
    int units_done = 0;  // units whose final triple this CTA has written
    const int64_t sp_n_floats = p.total_rows * (D + 2);
    auto emit_triple = [&](int64_t orow, int d, float acc, bool with_ml, float m_nat, float l_sum) {
        xchg_ll_emit<D>(a.peers, a.rank, a.world, sp_n_floats, orow, d, acc, with_ml, m_nat, l_sum);
    };
    auto finish_seqpar = [&]() {
        if (!SP || a.peers == nullptr || units_done == 0) return;  // (uniform per CTA) only CTAs that published a unit have anything to do
        unsigned int* hdr = reinterpret_cast<unsigned int*>(a.peers[a.rank]);
        const unsigned int sp_step = hdr[32] + 1;   // this rank's step in progress
        const unsigned int sp_tag = xchg_ll_tag(hdr, sp_step);
        const int64_t sp_ll_off = xchg_ll_offset(a.world, sp_n_floats);
        if (threadIdx.x == 0) s_flag[2] = (int)atomicAdd(hdr + 16, (unsigned int)units_done);   // this CTA's share of the merge ~ units published
        if (threadIdx.x == 0) s_flag[3] = 1;
        bar_consumers();
        const int64_t n_out = p.total_rows * D;
        const int64_t lo = n_out * s_flag[2] / a.n_units, hi = n_out * (s_flag[2] + units_done) / a.n_units;
        const uint2* part = reinterpret_cast<const uint2*>(a.peers[a.rank] + sp_ll_off) + (int64_t)(sp_step & 1u) * a.world * sp_n_floats;
        // ---- merge this CTA's share (fa_reduce algebra), polling as it goes (out of line: the cold path must not cost the stream loop registers) ----
        if (!xchg_ll_merge_share<D>(part, a.world, p.total_rows, lo, hi, sp_tag, a.fdst, a.fdst_type, hdr, CW * 32)) s_flag[3] = 0;
        bar_consumers();
        if (threadIdx.x == 0) {
            const bool good = s_flag[3] != 0;
            const unsigned int done = atomicAdd(hdr + 17, (unsigned int)units_done);
            if (!good) hdr[18] = 1u;   // some CTA of this step timed out: the step must not be counted
            if (done + (unsigned int)units_done == (unsigned int)a.n_units) {
                __threadfence();
                const bool step_ok = good && *reinterpret_cast<volatile unsigned int*>(hdr + 18) == 0u;
                hdr[16] = 0u; hdr[17] = 0u; hdr[18] = 0u;
                if (step_ok) hdr[32] = sp_step;  // the step is complete on this rank
            }
        }
    };

    int i = 0, slot_idx = 0;
    int n_def = 0, def_u[2] = {0, 0}, def_c0[2] = {0, 0}, def_n[2] = {0, 0};  // partial units of this CTA (at most the first and the last segment)
    int cl_u = 0, cl_n = 1;  // cluster mode: the CTA's single unit and its contributor count
    while (i < my_chunks) {
        const unsigned x = start + i;
        const int u = (int)(x / (unsigned)a.cph), ch0 = (int)(x - (unsigned)u * (unsigned)a.cph);
        const int seg_end = min(my_chunks, i + (a.cph - ch0));
        const int ik2 = u % p.n_head_kv, iq3 = u / p.n_head_kv;
        // who else works on this unit (off the critical path: the segment's first stage is still in flight)
        const unsigned c0 = dk_owner((unsigned)u * a.cph, total, G), c1 = dk_owner((unsigned)(u + 1) * a.cph - 1u, total, G);
        const int n_contrib = (int)(c1 - c0 + 1);

        if constexpr (EXT) {  // mask slices: the rows' mask pointers depend on the unit's heads and batch
            if (p.mask != nullptr && (p.m_ne2 > 1 || p.m_ne3 > 1)) {
#pragma unroll
                for (int h = 0; h < NR; h++) mrow[h] = p.mask + (int64_t)iq1r[h] * p.nb31 + fa_mask_slice_off(p, ik2 * p.gqa + rq[h], iq3);
            }
        }
        // ---- Q fragments of this unit's rows (f16; an f32 Q is rounded like the reference does, flash-llama.h:80) ----
        if constexpr (T8) {
            // B fragments: lane (g, t) holds Q[row g][32b + 8t .. + 7]; then -1152 x the block sums of rows 2t, 2t+1
            const bool qv = g < rows_total;
            const int qi1 = (qv ? g : 0) / p.gqa, qh = ik2 * p.gqa + (qv ? g : 0) % p.gqa;
            const char* qrow = p.q + (int64_t)qi1 * p.nb01 + (int64_t)qh * p.nb02 + (int64_t)iq3 * p.nb03;
            // (all loads first: the shuffles below are ordered against memory operations, and one load -> shuffle chain per block
            // made the prologue four L2 round trips long)
            if (!qv) {
#pragma unroll
                for (int b = 0; b < NB; b++) qb[b][0] = qb[b][1] = qb[b][2] = qb[b][3] = 0u;
            } else if (p.q_type == B200FA_TYPE_F16) {
                uint4 v[NB];
#pragma unroll
                for (int b = 0; b < NB; b++) v[b] = *reinterpret_cast<const uint4*>(qrow + (32 * b + 8 * t) * 2);
#pragma unroll
                for (int b = 0; b < NB; b++) { qb[b][0] = v[b].x; qb[b][1] = v[b].y; qb[b][2] = v[b].z; qb[b][3] = v[b].w; }
            } else {
                float4 v[NB], y[NB];
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    v[b] = *reinterpret_cast<const float4*>(qrow + (32 * b + 8 * t) * 4);
                    y[b] = *reinterpret_cast<const float4*>(qrow + (32 * b + 8 * t) * 4 + 16);
                }
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    qb[b][0] = pack_half2(v[b].x, v[b].y); qb[b][1] = pack_half2(v[b].z, v[b].w);
                    qb[b][2] = pack_half2(y[b].x, y[b].y); qb[b][3] = pack_half2(y[b].z, y[b].w);
                }
            }
#pragma unroll
            for (int b = 0; b < NB; b++) {
                float ps = 0.f;
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&qb[b][w]));
                    ps += f.x + f.y;
                }
                ps += __shfl_xor_sync(0xffffffffu, ps, 1);
                ps += __shfl_xor_sync(0xffffffffu, ps, 2);   // block sum of row g, in every lane of the quad
                qsn[b][0] = -1152.f * __shfl_sync(0xffffffffu, ps, (2 * t) * 4);
                qsn[b][1] = -1152.f * __shfl_sync(0xffffffffu, ps, (2 * t + 1) * 4);
            }
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
                if constexpr (EXT) mslope[rr] = kLog2e * fa_slope(p, ik2 * p.gqa + rq[rr]);
#pragma unroll
            for (int mt = 0; mt < NMT; mt++) oT[mt][0] = oT[mt][1] = oT[mt][2] = oT[mt][3] = 0.f;
        }
#pragma unroll
        for (int h = 0; h < (T8 ? 0 : RH); h++) {
            if constexpr (EXT) mslope[h] = kLog2e * fa_slope(p, ik2 * p.gqa + rq[h]);
            const char* qrow = p.q + (int64_t)iq1r[h] * p.nb01 + (int64_t)(ik2 * p.gqa + rq[h]) * p.nb02 + (int64_t)iq3 * p.nb03;
#pragma unroll
            for (int c = 0; c < NC4; c++) {
                const int e0 = 8 * (t + 4 * c);
                if (!rvalid[h] || e0 >= p.Dr) {  // dead row, or a column past the real head size (zero padding)
                    qa[c][h][0] = qa[c][h][1] = qa[c][h][2] = qa[c][h][3] = 0u;
                } else if (p.q_type == B200FA_TYPE_F16) {
                    const uint4 v = *reinterpret_cast<const uint4*>(qrow + e0 * 2);
                    qa[c][h][0] = v.x; qa[c][h][1] = v.y; qa[c][h][2] = v.z; qa[c][h][3] = v.w;
                } else {
                    const float4 v = *reinterpret_cast<const float4*>(qrow + e0 * 4);
                    const float4 y = *reinterpret_cast<const float4*>(qrow + e0 * 4 + 16);
                    qa[c][h][0] = pack_half2(v.x, v.y); qa[c][h][1] = pack_half2(v.z, v.w);
                    qa[c][h][2] = pack_half2(y.x, y.y); qa[c][h][3] = pack_half2(y.z, y.w);
                }
            }
        }
#pragma unroll
        for (int n = 0; n < (T8 ? 0 : NT); n++)
#pragma unroll
            for (int e = 0; e < 4; e++) o[n][e] = 0.f;
#pragma unroll
        for (int h = 0; h < NR; h++) { m_run[h] = -INFINITY; l_run[h] = 0.f; }

        // ---- this warp group's chunks of the segment: chunk j belongs to group j % NG ----
        if (i == 0) asm volatile("bar.sync 3, %0;" ::"n"(dk_threads<T8>()) : "memory");  // the mbarriers are initialised (the producers do not wait here)
        const int j0 = i + (group + NG - i % NG) % NG;
        int stage = j0 % NS, phase = (j0 / NS) & 1;  // advanced by NG per chunk: the ring depth is a multiple of NG
        for (int j = j0; j < seg_end; j += NG, stage += NG) {
            if (stage >= NS) { stage -= NS; phase ^= 1; }
            const int key0 = (ch0 + (j - i)) * DK_CHUNK;
            const int kv0 = key0 + 16 * sub;

            mbar_wait(&full[stage], phase);
            __syncwarp();
            DK_TL(threadIdx.x == 0 && j == 0, blockIdx.x * 8 + 1, dk_now());
            DK_TL(blockIdx.x == 0 && sub == 0 && lane == 0 && j < 512, DK_TL_CHUNK0 + j * 4 + 2, dk_now());  // chunk landed (as seen by its first consumer warp)
            const bool live = kv0 < a.kv_end && !DK_DBG_SKIP_MATH;
            const uint32_t sb = stages_u32 + stage * Geo::kStageBytes;
            if constexpr (T8) {
                if (live) {
                    t8_tile(sb, key0, a.mask_bulk && key0 + DK_CHUNK <= p.n_kv, &empty[stage]);  // hands the stage back after its last load
                } else {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);
                }
                DK_TL(blockIdx.x == 0 && sub == 0 && lane == 0 && j < 512, DK_TL_CHUNK0 + j * 4 + 3, dk_now());  // tile done
                continue;
            }
            Tile T;
            float s[2][4];
            if (live) {
                read_k(T, sb, kv0, a.mask_bulk && key0 + DK_CHUNK <= p.n_kv);
                qk_tile(T, s);
                read_v(T, sb, kv0);
                if constexpr (Q8) {
                    if (key0 + DK_CHUNK > p.n_kv) {  // ragged last chunk: rows past the end hold stale bytes; a zero scale keeps them out of P.V
#pragma unroll
                        for (int i2 = 0; i2 < 4; i2++)
#pragma unroll
                            for (int c = 0; c < NCV; c++)
                                if (kv0 + 4 * t + i2 >= p.n_kv) T.vd[i2][c] = 0u;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);  // fragments are in registers: hand the stage back
            DK_TL(blockIdx.x == 0 && sub == 0 && lane == 0 && j < 512, DK_TL_CHUNK0 + j * 4 + 3, dk_now());  // stage released
            if (live) softmax_pv_tile(T, s, kv0);
        }

        // Cluster mode, first barrier: every CTA of the unit has consumed its last chunk, so the leader's ring (where the records go)
        // is idle.  Arrive now, wait just before the first remote store: the local fold runs in between.
        if (a.cluster_k > 1) cluster_arrive();
        // ---- fold the consumer warps' partial states: every warp parks its (m, l, O) in its slot, then thread (lane, warp)
        //      combines O elements warp, warp + CW, ... of that lane across the slots and stores them ----
        DK_TL(threadIdx.x == 0, blockIdx.x * 8 + 2, dk_now());
#pragma unroll
        for (int h = 0; h < NR; h++) {  // the lane's l covers its own keys: sum over the lanes that share the row
            if constexpr (T8) {
                l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 4);
                l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 8);
                l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 16);
            } else {
                l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
                l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
            }
        }
        bar_consumers();  // everyone is done reading the slots of the previous segment
        if constexpr (T8) t8_slot_st(); else slot_st();
        bar_consumers();
        if (a.cluster_k > 1) cluster_wait();
        {
            float wgt[CW][NR], Ms[NR], Ls[NR];
            const float* sl0 = merge + lane;
#pragma unroll
            for (int h = 0; h < NR; h++) {
                float M = -INFINITY;
#pragma unroll
                for (int w = 0; w < CW; w++) M = fmaxf(M, sl0[w * (Mrg::kSlotBytes / 4) + (KO + 2 * h) * 32]);
                const float Mu = (M == -INFINITY) ? 0.f : M;
                float L = 0.f;
#pragma unroll
                for (int w = 0; w < CW; w++) {
                    wgt[w][h] = fast_exp2(sl0[w * (Mrg::kSlotBytes / 4) + (KO + 2 * h) * 32] - Mu);
                    L += sl0[w * (Mrg::kSlotBytes / 4) + (KO + 2 * h + 1) * 32] * wgt[w][h];
                }
                Ms[h] = M; Ls[h] = L;
            }
#pragma unroll
            for (int k = warp; k < KO; k += CW) {
                // element k of the lane's fragment -> (row R, head dim d); `first`: one thread per row also stores (m, l)
                int h, R, d;
                bool first;
                if constexpr (T8) {
                    const int mt = k >> 2, e = k & 3, mm = g + 8 * (e >> 1);
                    h = e & 1; R = 2 * t + h;
                    d = 32 * (mt >> 1) + 4 * (mm & 7) + 2 * (mt & 1) + (mm >> 3);
                    first = (mt == 0 && (e >> 1) == 0 && g == 0);
                } else {
                    const int n = k / (2 * RH), e = k % (2 * RH), e1 = e & 1;
                    h = e >> 1; R = g + 8 * h;
                    d = 64 * (n >> 3) + 16 * t + (n & 7) + 8 * e1;
                    first = (n == 0 && e1 == 0 && t == 0);
                }
                float acc = 0.f;
#pragma unroll
                // (h is not a compile-time value: selects, not indexed local arrays)
                for (int w = 0; w < CW; w++) acc += sl0[w * (Mrg::kSlotBytes / 4) + k * 32] * ((NR == 1 || h == 0) ? wgt[w][0] : wgt[w][NR - 1]);
                if (R >= rows_total) continue;
                const int row_q1 = (NR == 1 || h == 0) ? iq1r[0] : iq1r[NR - 1], row_h = (NR == 1 || h == 0) ? rq[0] : rq[NR - 1];
                const int64_t orow = ((int64_t)iq3 * p.n_q + row_q1) * p.n_head + ik2 * p.gqa + row_h;  // flash-llama.h:434
                const float M = (NR == 1 || h == 0) ? Ms[0] : Ms[NR - 1], L = (NR == 1 || h == 0) ? Ls[0] : Ls[NR - 1];
                if (n_contrib > 1 && a.cluster_k > 1) {
                    // record [rank][row][D + 2] in the leader CTA's (drained) stage ring; the cluster barrier below orders these
                    // remote stores before the leader's loads.  The leader's own ring is idle: a cluster CTA has one segment,
                    // and its last chunk was consumed before its fold began.
                    const uint32_t rank = blockIdx.x % a.cluster_k;
                    const uint32_t base = cluster_map(stages_u32, 0) + ((rank * RLIVE + R) * (D + 2)) * 4;
                    st_cluster_f32(base + d * 4, acc);
                    if (first) { st_cluster_f32(base + D * 4, M); st_cluster_f32(base + (D + 1) * 4, L); }
                } else if (n_contrib > 1) {
                    float* rec = a.rec + (((int64_t)blockIdx.x * a.max_slots + slot_idx) * DK_REC_ROWS + R) * (D + DK_REC_PAD);
                    rec[d] = acc;
                    if (first) { rec[D] = M; rec[D + 1] = L; }  // log2 units inside the kernel
                } else if (p.dst != nullptr) {
                    const float y = L > 0.f ? acc / L : 0.f;
                    if (d < p.Dr) {
                        if (p.dst_type == B200FA_TYPE_F16) reinterpret_cast<__half*>(p.dst)[orow * p.Dr + d] = __float2half_rn(y);
                        else reinterpret_cast<float*>(p.dst)[orow * p.Dr + d] = y;
                    }
                } else if (SP && a.peers != nullptr) {
                    emit_triple(orow, d, acc, first, M * kLn2, L);
                } else {
                    float* out = p.part_out + orow * (D + 2);
                    out[d] = acc;
                    if (first) { out[D] = M * kLn2; out[D + 1] = L; }
                }
            }
        }
        if (n_contrib == 1) units_done++;
        DK_TL(threadIdx.x == 0, blockIdx.x * 8 + 3, dk_now());
        // Units shared with other CTAs are signalled and merged after the CTA's whole run (below): only its first and its
        // last segment can be partial units, and a fence + atomic round trip in the middle of the stream would stall the
        // consumers for longer than the ring can cover.
        if (n_contrib > 1 && a.cluster_k <= 1) {
            if (n_def < 2) { def_u[n_def] = u; def_c0[n_def] = (int)c0; def_n[n_def] = n_contrib; n_def++; }
        }
        if (a.cluster_k > 1) { cl_u = u; cl_n = n_contrib; }
        slot_idx++;
        i = seg_end;
    }

    if (a.cluster_k > 1) {
        // ---- cluster mode: the unit's CTAs are one thread-block cluster; their records sit in the leader's shared memory ----
        cluster_sync_all();
        DK_TL(threadIdx.x == 0, blockIdx.x * 8 + 5, dk_now());  // past the cluster barrier
        if (blockIdx.x % a.cluster_k != 0 || cl_n <= 1) {
            DK_TL(threadIdx.x == 0, blockIdx.x * 8 + 4, dk_now());
            DK_TL(threadIdx.x == (CW - 1) * 32, blockIdx.x * 8 + 1, dk_now());  // (diagnostic) the last consumer warp leaves
            finish_seqpar();
            return;
        }
        const int u = cl_u, n_contrib = cl_n;
        const int ik2 = u % p.n_head_kv, iq3 = u / p.n_head_kv;
        const float* recs = reinterpret_cast<const float*>(smem);
        for (int idx = threadIdx.x; idx < rows_total * D; idx += CW * 32) {
            const int R = idx / D, d = idx % D;
            float M = -INFINITY;
            for (int cc = 0; cc < n_contrib; cc++) M = fmaxf(M, recs[(cc * RLIVE + R) * (D + 2) + D]);
            const float Mu = (M == -INFINITY) ? 0.f : M;
            float L = 0.f, acc = 0.f;
            for (int cc = 0; cc < n_contrib; cc++) {
                const float* rec = recs + (cc * RLIVE + R) * (D + 2);
                const float wt = fast_exp2(rec[D] - Mu);
                L += rec[D + 1] * wt; acc += rec[d] * wt;
            }
            const int iq1 = R / p.gqa;
            const int64_t orow = ((int64_t)iq3 * p.n_q + iq1) * p.n_head + ik2 * p.gqa + R % p.gqa;
            if (p.dst != nullptr) {
                const float y = L > 0.f ? acc / L : 0.f;
                if (d < p.Dr) {
                    if (p.dst_type == B200FA_TYPE_F16) reinterpret_cast<__half*>(p.dst)[orow * p.Dr + d] = __float2half_rn(y);
                    else reinterpret_cast<float*>(p.dst)[orow * p.Dr + d] = y;
                }
            } else if (SP && a.peers != nullptr) {
                emit_triple(orow, d, acc, d == 0, M * kLn2, L);
            } else {
                float* out = p.part_out + orow * (D + 2);
                out[d] = acc;
                if (d == 0) { out[D] = M * kLn2; out[D + 1] = L; }
            }
        }
        units_done++;
        DK_TL(threadIdx.x == 0, blockIdx.x * 8 + 4, dk_now());
        finish_seqpar();
        return;
    }

    // ---- partial units: publish the records, count arrivals; the last CTA of a unit to arrive merges its records
    //      (fa_reduce, flash_row_float.h:415-472: M = max m_i, L = sum l_i 2^(m_i-M), O = sum O~_i 2^(m_i-M) / L —
    //      here one parallel fp32 pass) ----
    if (n_def == 0) {
        DK_TL(threadIdx.x == 0, blockIdx.x * 8 + 4, dk_now());
        finish_seqpar();
        return;
    }
    __threadfence();
    bar_consumers();
    if (threadIdx.x < n_def) {
        const unsigned int old = atomicInc(a.counters + def_u[threadIdx.x], (unsigned int)def_n[threadIdx.x] - 1);  // wraps to 0: self-resetting
        s_flag[threadIdx.x] = (old == (unsigned int)def_n[threadIdx.x] - 1);
    }
    bar_consumers();
    DK_TL(threadIdx.x == 0, blockIdx.x * 8 + 5, dk_now());  // records published, arrival counted
    for (int k = 0; k < n_def; k++) {
        if (s_flag[k] == 0) continue;
        const int u = def_u[k], n_contrib = def_n[k];
        const int ik2 = u % p.n_head_kv, iq3 = u / p.n_head_kv;
        bar_consumers();  // s_tab is reused
        for (int cc = threadIdx.x; cc < n_contrib; cc += CW * 32) {  // record slot of contributor c0 + cc
            const unsigned st = (unsigned)(def_c0[k] + cc) * total / G;
            s_tab[cc] = (def_c0[k] + cc) * a.max_slots + (u - (int)(st / (unsigned)a.cph));
        }
        bar_consumers();
        __threadfence();
        // The merge is a chain of L2 round trips (~0.6 us each under load) on the critical path of the whole launch: it is laid out so
        // that there are two of them.  (A one-pass version with 72 batched loads per thread spilled in the 128-register transposed
        // q8_0 kernel, and a version that walked the records eight at a time took ~5 us for an 18-CTA unit.)
        // Pass 1, one warp per row: the records' (m, l) pairs -> weights 2^(m_c - M) in shared memory (the merge slots are idle), M, L.
        constexpr int T = CW * 32, D4 = D / 4;
        float* sw = merge;                                   // [rows_total][n_contrib]
        float* sML = sw + rows_total * n_contrib;            // [rows_total][2]
        float4* spart = reinterpret_cast<float4*>(merge + ((rows_total * (n_contrib + 2) + 3) & ~3));  // [groups][rows_total * D/4]
        for (int R = warp; R < rows_total; R += CW) {
            float2 ml[(DK_TAB + 31) / 32];
#pragma unroll
            for (int e = 0; e < (DK_TAB + 31) / 32; e++) {
                const int c = lane + 32 * e;
                ml[e] = c < n_contrib ? __ldcg(reinterpret_cast<const float2*>(a.rec + ((int64_t)s_tab[c] * DK_REC_ROWS + R) * (D + DK_REC_PAD) + D))
                                      : make_float2(-INFINITY, 0.f);
            }
            float M = -INFINITY;
#pragma unroll
            for (int e = 0; e < (DK_TAB + 31) / 32; e++) M = fmaxf(M, ml[e].x);
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, sft));
            const float Mu = (M == -INFINITY) ? 0.f : M;
            float L = 0.f;
#pragma unroll
            for (int e = 0; e < (DK_TAB + 31) / 32; e++) {
                const int c = lane + 32 * e;
                const float wt = fast_exp2(ml[e].x - Mu);
                if (c < n_contrib) sw[R * n_contrib + c] = wt;
                L += ml[e].y * wt;
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) L += __shfl_xor_sync(0xffffffffu, L, sft);
            if (lane == 0) { sML[2 * R] = M; sML[2 * R + 1] = L; }
        }
        bar_consumers();
        // Pass 2: the threads form `groups` teams of E4 (one thread per four output elements); team g sums the records c = g, g + groups,
        // ... with eight 16-byte loads in flight, and the teams' partial sums meet in shared memory.
        const int E4 = rows_total * D4;
        const int groups = E4 <= T ? T / E4 : 1;
        for (int e4 = (int)threadIdx.x % E4, grp = (int)threadIdx.x / E4; grp < groups && e4 < E4; e4 += (groups > 1 ? E4 : T)) {
            const int R = e4 / D4, d4 = e4 % D4;
            const float* wr = sw + R * n_contrib;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = grp; c < n_contrib; c += 8 * groups) {
                float4 v[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int cc = min(c + e * groups, n_contrib - 1);
                    v[e] = __ldcg(reinterpret_cast<const float4*>(a.rec + ((int64_t)s_tab[cc] * DK_REC_ROWS + R) * (D + DK_REC_PAD)) + d4);
                }
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int cc = c + e * groups;
                    const float wt = cc < n_contrib ? wr[cc] : 0.f;
                    acc.x = fmaf(v[e].x, wt, acc.x); acc.y = fmaf(v[e].y, wt, acc.y); acc.z = fmaf(v[e].z, wt, acc.z); acc.w = fmaf(v[e].w, wt, acc.w);
                }
            }
            spart[grp * E4 + e4] = acc;
        }
        bar_consumers();
        for (int e4 = threadIdx.x; e4 < E4; e4 += T) {
            const int R = e4 / D4, d0 = 4 * (e4 % D4);
            float4 s4 = spart[e4];
            for (int g2 = 1; g2 < groups; g2++) {
                const float4 t4 = spart[g2 * E4 + e4];
                s4.x += t4.x; s4.y += t4.y; s4.z += t4.z; s4.w += t4.w;
            }
            const float M = sML[2 * R], L = sML[2 * R + 1];
            const float accs[4] = {s4.x, s4.y, s4.z, s4.w};
            const int iq1 = R / p.gqa;
            const int64_t orow = ((int64_t)iq3 * p.n_q + iq1) * p.n_head + ik2 * p.gqa + R % p.gqa;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int d = d0 + e;
                const float acc = accs[e];
                if (p.dst != nullptr) {
                    const float y = L > 0.f ? acc / L : 0.f;
                    if (d < p.Dr) {
                        if (p.dst_type == B200FA_TYPE_F16) reinterpret_cast<__half*>(p.dst)[orow * p.Dr + d] = __float2half_rn(y);
                        else reinterpret_cast<float*>(p.dst)[orow * p.Dr + d] = y;
                    }
                } else if (SP && a.peers != nullptr) {
                    emit_triple(orow, d, acc, d == 0, M * kLn2, L);
                } else {
                    float* out = p.part_out + orow * (D + 2);
                    out[d] = acc;
                    if (d == 0) { out[D] = M * kLn2; out[D + 1] = L; }
                }
            }
        }
        units_done++;
    }
    finish_seqpar();
    DK_TL(threadIdx.x == 0, blockIdx.x * 8 + 4, dk_now());
}

}  // namespace b200fa

// b200fa_api.cu — the C ABI declared in include/b200fa.h: argument validation, kernel selection,
// split-KV planning and launches.  No allocation, no synchronisation, no CPU compute path.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "decode_mma.cuh"
#include "decode_stream.cuh"
#include "prefill_tcgen05.cuh"
#include "prefill_persistent.cuh"
#include "q8_0.cuh"
#include "tensor_file.cuh"

using namespace b200fa;

namespace {

thread_local const char* g_last_dispatch = "none";
thread_local int g_last_launches = 0;
#ifdef B200FA_TUNING
thread_local unsigned long long* g_timeline = nullptr;  // diagnostics: see b200fa_debug_timeline (tuning builds only, per thread)
#else
constexpr unsigned long long* g_timeline = nullptr;
#endif

struct SeqPar { char* const* peers = nullptr; int rank = 0, world = 1; void* fdst = nullptr; int fdst_type = 0; };
thread_local SeqPar g_seqpar;  // set by b200fa_flash_attn_seqpar around its attn_common call

struct DeviceInfo {
    int sm_count = 0;
    int cc_major = 0;
    bool ok = false;
};

const DeviceInfo& device_info() {
    static thread_local DeviceInfo cache[64];
    static DeviceInfo none;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return none;
    DeviceInfo& d = cache[dev];
    if (!d.ok) {
        if (cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return none;
        if (cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return none;
        d.ok = true;
    }
    return d;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Workspace layout: [arrival counters: fixed kCtrRegion bytes at offset 0][everything else].  The counter region is
// the same for every shape, so B200FA_FLAG_WORKSPACE_ZEROED stays valid when calls of different shapes share a workspace.
constexpr size_t kCtrRegion = 65536 * sizeof(unsigned int) + 256;  // + the persistent prefill kernel's two counters
constexpr size_t kPrefillCtrOff = 65536 * sizeof(unsigned int);

enum PlanKind { kPrefill, kStream, kRows16 };

struct Plan {
    PlanKind kind = kRows16;
    // rows16 split-KV
    int n_splits = 1;
    int split_len = 0;
    int n_groups = 1;
    // stream-K decode
    int cph = 0, n_units = 0, grid = 0, max_slots = 0, kv_end = 0, cluster_k = 0;
    long long total_chunks = 0;
    size_t n_counters = 0;   // arrival counters needed (0 = none)
    size_t part_bytes = 0;   // split-KV partials / stream-K records
    size_t qf16_bytes = 0;   // f16 copy of an f32 Q (tcgen05 path)
    size_t cls_bytes = 0;    // mask tile classes (tcgen05 path)
    size_t kvf16_bytes = 0;  // f16 copies of q8_0 K and V (tcgen05 path on a quantised cache), K then V
    size_t ctr_bytes = 0;    // counter region actually reserved
    size_t total = 0;
};

struct Shape {
    int q_type, kv_type;
    int64_t D, n_q, n_head, n_batch, n_kv, n_head_kv;
    int64_t nb11, nb12, nb13, nb21, nb22, nb23;  // 0 = unknown (workspace sizing): assume the widest plan
    const void* k; const void* v;
    int64_t kv_pos0, n_kv_total;
    int64_t Dr = 0;  // real head size when it differs from the structural D (0 = same)
    int64_t n_batch_kv = 0;  // ne13 (0 = n_batch)
    bool ext = false;  // ALiBi / soft-cap / mask slices requested (ext2 entry): only the persistent prefill kernel implements them
    int64_t mask_slices = 1;  // m_ne2 * m_ne3 (ext2): the prefill kernel keeps one table of mask tile classes per slice
};

// The one place a call's arguments become the planner's view of it (attn_common, b200fa_flash_attn_seqpar and b200fa_plan agree by
// construction): D is the structural head size of the decode kernels (64 or 128), Dr the real one.
Shape make_shape(int q_type, int kv_type, int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03, int64_t ne11, int64_t ne12, int64_t ne13,
                 int64_t nb11, int64_t nb12, int64_t nb13, int64_t nb21, int64_t nb22, int64_t nb23, const void* k, const void* v,
                 int64_t kv_pos0, int64_t n_kv_total) {
    return Shape{q_type, kv_type, ne00 <= 64 ? 64 : (ne00 <= 128 ? 128 : 256), ne01, ne02, ne03, ne11, ne12, nb11, nb12, nb13, nb21, nb22, nb23, k, v, kv_pos0, n_kv_total, ne00, ne13};
}

bool stream_eligible(const Shape& sh, bool sizing) {
    const int64_t rows = sh.n_q * (sh.n_head / sh.n_head_kv);
    if (rows > 16 || (sh.D != 64 && sh.D != 128)) return false;
    if (sh.n_head_kv * sh.n_batch > 65536) return false;
    // the kernel's chunk bookkeeping is 32-bit (DK_MAX_TOTAL_X_GRID): out of reach for any K/V a 180 GB device holds, D = 64 q8_0 aside
    if (sh.n_head_kv * sh.n_batch * ((sh.n_kv + DK_CHUNK - 1) / DK_CHUNK) * (DK_TAB + 1) > DK_MAX_TOTAL_X_GRID) return false;
    static const bool off = tune_env("B200FA_DECODE_IMPL") && !strcmp(tune_env("B200FA_DECODE_IMPL"), "rows16");
    if (off) return false;
    if (sizing || sh.kv_type == B200FA_TYPE_F16) return true;
    // q8_0: the producer copies whole chunks of rows with 16-byte bulk copies -> rows must be contiguous and 16-byte aligned per head
    const int64_t row = sh.D / kQ8BlockElems * kQ8BlockBytes;
    if (sh.nb11 != row || sh.nb21 != row) return false;
    return (((uintptr_t)sh.k | (uintptr_t)sh.v | (uintptr_t)sh.nb12 | (uintptr_t)sh.nb13 | (uintptr_t)sh.nb22 | (uintptr_t)sh.nb23) % 16) == 0;
}

// How many virtual KV heads a real one is split into so that a 17..128-row GQA burst fits the stream kernel's 16 rows (1 = no split)
int virtual_head_split(int64_t n_q, int64_t n_head, int64_t n_head_kv, int64_t n_batch) {
    static const bool no_vh = tune_env("B200FA_NO_VIRTUAL_HEADS") != nullptr;
    const int64_t gqa = n_head / n_head_kv, rows = n_q * gqa;
    if (no_vh || rows <= 16 || rows > 128 || n_q > 16 || n_head_kv * n_batch * 8 > 65536) return 1;
    for (int dv = 2; dv <= gqa; dv++)
        if (gqa % dv == 0 && n_q * (gqa / dv) <= 16) return dv;
    return 1;
}

Plan make_plan(const Shape& sh, uint32_t flags, int sm_count, bool force_partial_out, bool sizing, bool allow_stream = true) {
    Plan pl;
    const int64_t D = sh.D, n_q = sh.n_q, n_head = sh.n_head, n_batch = sh.n_batch, n_kv = sh.n_kv, n_head_kv = sh.n_head_kv;
    const int64_t gqa = n_head / n_head_kv;
    const int64_t rows = n_q * gqa;
    // q8_0 K/V on the prefill path: dequantised once to dense f16 in the workspace (RN_f16(d*q), what the decode kernel feeds its
    // P.V too), then the tensor-core kernel — the 16-row fallback re-reads K/V per row group and measured 20 TFLOP/s on C3's shape.
    const int64_t nbk = sh.n_batch_kv > 0 ? sh.n_batch_kv : n_batch;
    const size_t kv16 = sh.kv_type == B200FA_TYPE_Q8_0 ? align_up((size_t)(n_kv * n_head_kv * nbk * D * 2), 256) : 0;
    static const bool no_q8_prefill = tune_env("B200FA_NO_Q8_PREFILL") != nullptr;
    const bool kv_ok = sh.kv_type == B200FA_TYPE_F16 || (!no_q8_prefill && (D == 64 || D == 128) && kv16 <= ((size_t)512 << 20));
    // More than 16 query positions: the tensor-core kernel, even when a 128-row tile is mostly padding (n_q = 32 against 8 K keys:
    // 122 us on the 16-row-group fallback, which re-reads K/V per group, vs ~40 us here with the KV range split over the SMs).
    // 17..128 rows per KV head from a GQA group with few query positions (bursts): the same kernel with the group's q heads packed
    // into one tile (pp_pack_shift) — one pass over K/V per KV head.
    const int pack_sh = pp_pack_shift(n_q, n_head, n_head_kv, sh.ext);
    static const bool no_pack = tune_env("B200FA_NO_PACK") != nullptr;
    const bool packed_burst = pack_sh > 0 && rows > 16 && !no_pack;
    if (!force_partial_out && !(flags & B200FA_FLAG_NO_TCGEN05) && (n_q > 16 || packed_burst) && D <= 128 && kv_ok &&
        n_kv <= ((sh.Dr == 0 || sh.Dr == 128) && !sh.ext ? (int64_t)PF_MAX_KV_TILES * PF_BN : (int64_t)PP_MAX_KV_TILES * PF_BN)) {
        pl.kind = kPrefill;
        if (sh.q_type == B200FA_TYPE_F32) pl.qf16_bytes = align_up((size_t)(n_q * n_head * n_batch * 128 * 2), 256);
        const int64_t qt = (n_q + 127) / 128, kt = (n_kv + 127) / 128;
        pl.cls_bytes = align_up((size_t)(qt * kt * sh.mask_slices), 256);
        pl.ctr_bytes = kCtrRegion;
        // Split-KV prefill: fewer work items than SMs and a long KV range (chunked prefill, long-context continuation): cut every
        // item's KV tiles into n_splits segments, each its own work item emitting (O~, m, l) rows, merged by a second launch.
        // Pick the segment count with the shortest makespan (waves / segments) among 2..16 with >= 8 KV tiles per segment and at
        // most 128 MB of partial rows.
        const int64_t qt_item = pack_sh ? (n_q + (128 >> pack_sh) - 1) / (128 >> pack_sh) : qt;  // query tiles as the kernel cuts them
        const int64_t n_items = ((qt_item + 1) / 2) * (pack_sh ? n_head_kv : n_head) * n_batch;
        static const bool no_split = tune_env("B200FA_PREFILL_NO_SPLIT") != nullptr;
        if (!no_split && n_items < sm_count && kt >= 16 && kt <= PP_MAX_KV_TILES && (sh.Dr == 0 || sh.Dr == 128)) {
            const size_t row_bytes = (size_t)(n_q * n_head * n_batch) * (128 + 4) * 4;
            int best = 1;
            double best_t = 1.0;
            for (int s = 2; s <= 16 && kt / s >= 8 && row_bytes * s <= ((size_t)128 << 20); s++) {
                const int64_t units = n_items * s, waves = (units + sm_count - 1) / sm_count;
                const double t = (double)waves / s;
                if (t < best_t - 1e-9) { best_t = t; best = s; }
            }
            if (best > 1) {
                pl.n_splits = best;
                pl.part_bytes = align_up(row_bytes * best, 256);
            }
        }
        pl.kvf16_bytes = 2 * kv16;
        pl.total = pl.ctr_bytes + pl.qf16_bytes + pl.cls_bytes + pl.part_bytes + pl.kvf16_bytes;
        return pl;
    }
    if (allow_stream && stream_eligible(sh, sizing)) {
        pl.kind = kStream;
        int64_t kv_end = n_kv;
        if (flags & B200FA_FLAG_CAUSAL) {
            const int64_t lim = n_q + (sh.n_kv_total - n_q) - sh.kv_pos0;  // keys < lim are visible to the last query
            kv_end = lim < 0 ? 0 : (lim < n_kv ? lim : n_kv);
        }
        pl.kv_end = (int)kv_end;
        pl.cph = (int)((kv_end + DK_CHUNK - 1) / DK_CHUNK);
        if (pl.cph < 1) pl.cph = 1;
        if (sizing) pl.cph = (int)((n_kv + DK_CHUNK - 1) / DK_CHUNK);  // the widest case
        pl.n_units = (int)(n_head_kv * n_batch);
        pl.total_chunks = (long long)pl.n_units * pl.cph;
        int grid = sm_count;
        // Few units: give every unit the same whole number of CTAs, so that no CTA's run crosses a unit boundary (a mid-run
        // fold stalls the stream for longer than the ring covers) — as long as that leaves at most a fifth of the SMs idle.
        if (pl.n_units <= sm_count) {
            const int aligned = pl.n_units * (sm_count / pl.n_units);
            if (aligned * 5 >= sm_count * 4) grid = aligned;
            // (Tried: 16 CTAs per unit for 8-9 units, so that a unit is one non-portable 16-CTA cluster: on this B200 fewer than eight
            // such clusters can be resident — cudaOccupancyMaxActiveClusters — and 128 CTAs stream q8_0 slower than 144.)
        }
        if (const char* e = tune_env("B200FA_STREAM_GRID")) grid = atoi(e) > 0 ? atoi(e) : grid;
        // unit-aligned grid with 2..8 CTAs per unit: launch each unit's CTAs as one thread-block cluster (DSMEM merge)
        static const bool no_cluster = tune_env("B200FA_NO_CLUSTER") != nullptr;
        if (!no_cluster && pl.n_units > 0 && grid % pl.n_units == 0 && pl.total_chunks >= grid) {
            const int k = grid / pl.n_units;
            if (k >= 2 && k <= 16) pl.cluster_k = k;  // 9..16: non-portable cluster sizes (launch_stream_t checks that they can be resident)
        }
        if (grid > DK_TAB) grid = DK_TAB;
        pl.grid = (int)(pl.total_chunks < grid ? pl.total_chunks : grid);
        const long long per = (pl.total_chunks + pl.grid - 1) / pl.grid;
        pl.max_slots = (int)((per + pl.cph - 1) / pl.cph) + 1;
        if (sizing) {  // cph (hence the slot count) shrinks under a causal clip: bound both extremes
            const long long per1 = ((long long)pl.n_units + pl.grid - 1) / pl.grid;  // cph = 1
            if (per1 + 1 > pl.max_slots) pl.max_slots = (int)per1 + 1;
        }
        pl.n_counters = (size_t)pl.n_units;
        pl.part_bytes = align_up((size_t)(sizing ? sm_count : pl.grid) * pl.max_slots * DK_REC_ROWS * (size_t)(D + DK_REC_PAD) * 4, 256);
        pl.ctr_bytes = kCtrRegion;
        pl.total = pl.ctr_bytes + pl.part_bytes;
        return pl;
    }
    pl.kind = kRows16;
    pl.n_groups = (int)((rows + kRows - 1) / kRows);
    const int64_t base = (int64_t)pl.n_groups * n_head_kv * n_batch;
    // Split planning: fill whole waves of resident CTAs (2 per SM).  Candidates keep >= 256 keys per split; pick
    // the fewest splits whose wave efficiency is within 3% of the best.
    const int64_t slots = (int64_t)sm_count * 2;
    int64_t max_splits = n_kv / 256;
    if (max_splits < 1) max_splits = 1;
    if (max_splits > 64) max_splits = 64;
    int64_t want = 1;
    if (base * max_splits <= slots) {
        want = max_splits;  // everything fits in one wave: be as parallel as allowed
    } else {
        double best = 0.0;
        for (int64_t n = 1; n <= max_splits; n++) {
            const int64_t ctas = base * n, waves = (ctas + slots - 1) / slots;
            const double eff = (double)ctas / (double)(waves * slots);
            if (eff > best) best = eff;
        }
        for (int64_t n = 1; n <= max_splits; n++) {
            const int64_t ctas = base * n, waves = (ctas + slots - 1) / slots;
            const double eff = (double)ctas / (double)(waves * slots);
            if (eff >= best - 0.03) { want = n; break; }
        }
    }
    if (const char* e = tune_env("B200FA_SPLITS")) want = atoll(e) > 0 ? atoll(e) : want;
    int64_t len = (n_kv + want - 1) / want;
    len = (len + 15) / 16 * 16;
    if (len < 16) len = 16;
    pl.split_len = (int)len;
    pl.n_splits = (int)((n_kv + len - 1) / len);
    if (pl.n_splits < 1) pl.n_splits = 1;
    pl.ctr_bytes = kCtrRegion;
    if (pl.n_splits > 1) {
        pl.part_bytes = align_up((size_t)pl.n_splits * (size_t)(n_batch * n_q * n_head) * (size_t)(D + 2) * 4, 256);
        pl.n_counters = (size_t)base;
        if (pl.n_counters * 4 > kCtrRegion) pl.ctr_bytes = align_up(pl.n_counters * 4, 256);  // oversized: always memset
    }
    pl.total = pl.ctr_bytes + pl.part_bytes;
    return pl;
}

int validate(const void* q, const void* k, const void* v, const void* out, int q_type, int kv_type, int dst_type,
             int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03, int64_t ne10, int64_t ne11, int64_t ne12,
             int64_t ne13, const void* mask, int64_t ne31, int64_t nb31, int64_t nb01, int64_t nb02, int64_t nb03,
             int64_t nb11, int64_t nb12, int64_t nb13, int64_t nb21, int64_t nb22, int64_t nb23, bool out_is_partial = false) {
    if (!q || !k || !v || !out) return B200FA_ERR_INVALID;
    if (ne00 <= 0 || ne01 <= 0 || ne02 <= 0 || ne03 <= 0 || ne11 <= 0 || ne12 <= 0 || ne13 <= 0) return B200FA_ERR_INVALID;
    if (ne00 != ne10) return B200FA_ERR_INVALID;
    if (ne02 % ne12 || ne03 % ne13) return B200FA_ERR_INVALID;
    if (q_type != B200FA_TYPE_F32 && q_type != B200FA_TYPE_F16) return B200FA_ERR_UNSUPPORTED;
    if (kv_type != B200FA_TYPE_F16 && kv_type != B200FA_TYPE_Q8_0) return B200FA_ERR_UNSUPPORTED;
    if (dst_type != B200FA_TYPE_F32 && dst_type != B200FA_TYPE_F16) return B200FA_ERR_UNSUPPORTED;
    // f16 K/V: any head size that is a multiple of 8 up to 256 (80, 96, 112, 160, 192 ... run zero-padded on the 64 / 128 / 256 wide
    // kernels); q8_0 K/V: 64, 128 or 256; the partial (sequence-split) entries: 64 or 128 only
    if (ne00 % 8 || ne00 > 256) return B200FA_ERR_UNSUPPORTED;
    if (kv_type == B200FA_TYPE_Q8_0 && ne00 != 64 && ne00 != 128 && ne00 != 256) return B200FA_ERR_UNSUPPORTED;
    if (out_is_partial && ne00 != 64 && ne00 != 128) return B200FA_ERR_UNSUPPORTED;
    if (ne11 > 0x7fffffff || ne01 > 0x7fffffff || ne02 > 65535 || ne03 * ne12 > 65535) return B200FA_ERR_UNSUPPORTED;
    const int64_t qrow = ne00 * (q_type == B200FA_TYPE_F32 ? 4 : 2);
    if (nb01 < qrow || ((uintptr_t)q | nb01 | nb02 | nb03) % 16) return B200FA_ERR_INVALID;
    if (kv_type == B200FA_TYPE_F16) {
        if (nb11 < ne00 * 2 || nb21 < ne00 * 2) return B200FA_ERR_INVALID;
        if (((uintptr_t)k | (uintptr_t)v | nb11 | nb12 | nb13 | nb21 | nb22 | nb23) % 16) return B200FA_ERR_INVALID;
    } else {
        const int64_t row = ne00 / kQ8BlockElems * kQ8BlockBytes;
        if (nb11 < row || nb21 < row) return B200FA_ERR_INVALID;
        if (((uintptr_t)k | (uintptr_t)v | nb11 | nb12 | nb13 | nb21 | nb22 | nb23) % 2) return B200FA_ERR_INVALID;
    }
    if (mask) {
        if (ne31 < ne01 || nb31 < ne11 * 2 || ((uintptr_t)mask | nb31) % 2) return B200FA_ERR_INVALID;
    }
    if ((uintptr_t)out % (out_is_partial ? 8 : 16)) return B200FA_ERR_INVALID;  // triples are (D + 2) floats: 8-byte aligned rows
    return B200FA_OK;
}

template <int D, int RH, bool EXT>
int launch_rows16(const FaParams& p, int n_groups, cudaStream_t st) {
    dim3 grid(p.n_splits, n_groups, p.n_head_kv * p.n_batch), block(kDecodeWarps * 32);
    constexpr int smem = FifoGeom<D>::kCtaBytes;  // cp.async FIFO (f16) / merge buffers (both)
    static thread_local bool attr_set[64][2] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int ti = p.kv_type == B200FA_TYPE_F16 ? 0 : 1;
    if (dev >= 0 && dev < 64 && !attr_set[dev][ti]) {
        cudaError_t e = ti == 0 ? cudaFuncSetAttribute(fa_rows16_splitkv<D, B200FA_TYPE_F16, RH, true, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                                : cudaFuncSetAttribute(fa_rows16_splitkv<D, B200FA_TYPE_Q8_0, RH, true, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return B200FA_ERR_CUDA;
        attr_set[dev][ti] = true;
    }
#ifdef B200FA_TUNING
    static const int pipe = tune_env("B200FA_DECODE_PIPE") ? atoi(tune_env("B200FA_DECODE_PIPE")) : 1;
    if (ti == 0 && pipe == 0) {  // the register-double-buffered variant without the cp.async FIFO (measured slower; comparison only)
        static thread_local bool a2[64] = {};
        if (!a2[dev]) { cudaFuncSetAttribute(fa_rows16_splitkv<D, B200FA_TYPE_F16, RH, false, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); a2[dev] = true; }
        fa_rows16_splitkv<D, B200FA_TYPE_F16, RH, false, EXT><<<grid, block, smem, st>>>(p);
        return cudaGetLastError() == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
    }
#endif
    if (ti == 0) fa_rows16_splitkv<D, B200FA_TYPE_F16, RH, true, EXT><<<grid, block, smem, st>>>(p);
    else fa_rows16_splitkv<D, B200FA_TYPE_Q8_0, RH, true, EXT><<<grid, block, smem, st>>>(p);
    return cudaGetLastError() == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}

int run_rows16(FaParams& p, const Plan& pl, cudaStream_t st) {
    const bool small = (int64_t)p.n_q * p.gqa <= 8;  // one group of at most 8 rows: only fragment rows g are live
    int rc;
    const bool ext = p.cap_in != 0.f || p.alibi_nhl2 != 0 || p.m_ne2 * p.m_ne3 > 1;
#define B200FA_ROWS16(DD, RR) (ext ? launch_rows16<DD, RR, true>(p, pl.n_groups, st) : launch_rows16<DD, RR, false>(p, pl.n_groups, st))
    if (p.D == 256) rc = small ? B200FA_ROWS16(256, 1) : B200FA_ROWS16(256, 2);  // head sizes 129..256: this kernel only
    else if (p.D == 128) rc = small ? B200FA_ROWS16(128, 1) : B200FA_ROWS16(128, 2);
    else rc = small ? B200FA_ROWS16(64, 1) : B200FA_ROWS16(64, 2);
#undef B200FA_ROWS16
    g_last_launches++;
    return rc;
}

template <int D>
int launch_combine(const float* part, int n_parts, int64_t rows, void* dst, int dst_type, cudaStream_t st) {
    fa_combine<D><<<(unsigned)rows, D, 0, st>>>(part, n_parts, rows, dst, dst_type);
    return cudaGetLastError() == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}

template <int D, int KV, int RH, bool EXT, bool T8 = false, bool SP = false>
int launch_stream_t(const FaParams& p, const DkArgs& a, int grid, const CUtensorMap& tk, const CUtensorMap& tv, cudaStream_t st) {
    constexpr int smem = dk_smem_bytes<D, KV == B200FA_TYPE_Q8_0, RH, T8>();
    static thread_local bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        if (cudaFuncSetAttribute(fa_decode_stream<D, KV, RH, EXT, T8, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return B200FA_ERR_CUDA;
        attr_set[dev] = true;
    }
    // Programmatic dependent launch: this kernel may be scheduled while the previous kernel of the stream drains (it waits in-kernel,
    // griddepcontrol.wait, before its first global access).  Off for the peer-memory variants and with B200FA_NO_PDL.
    static const bool no_pdl = tune_env("B200FA_NO_PDL") != nullptr;
    const bool pdl = !no_pdl && a.peers == nullptr;
    DkArgs args = a;
    if (args.cluster_k > 8) {
        // non-portable cluster size: opt in, and use it only if all clusters of the grid can be resident at once (one 16-CTA cluster per
        // GPC on B200); otherwise the same grid runs without clusters and merges through global records
        static thread_local int max_clusters[64][17] = {};
        int& mc = max_clusters[dev >= 0 && dev < 64 ? dev : 0][args.cluster_k];
        if (mc == 0) {
            mc = -1;
            if (cudaFuncSetAttribute(fa_decode_stream<D, KV, RH, EXT, T8, SP>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
                cudaLaunchConfig_t qc{};
                qc.gridDim = dim3(args.cluster_k); qc.blockDim = dim3(dk_threads<T8>()); qc.dynamicSmemBytes = smem;
                cudaLaunchAttribute qa[1];
                qa[0].id = cudaLaunchAttributeClusterDimension;
                qa[0].val.clusterDim.x = args.cluster_k; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                qc.attrs = qa; qc.numAttrs = 1;
                int n = 0;
                if (cudaOccupancyMaxActiveClusters(&n, fa_decode_stream<D, KV, RH, EXT, T8, SP>, &qc) == cudaSuccess && n > 0) mc = n;
            }
            (void)cudaGetLastError();
        }
        if (mc < grid / args.cluster_k) args.cluster_k = 0;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(dk_threads<T8>()); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (args.cluster_k > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = args.cluster_k; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        na++;
    }
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, fa_decode_stream<D, KV, RH, EXT, T8, SP>, p, args, tk, tv) == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}

int run_stream(const FaParams& p, const Plan& pl, char* ws, cudaStream_t st) {
    DkArgs a{};
    a.cph = pl.cph; a.n_units = pl.n_units; a.total = pl.total_chunks; a.kv_end = pl.kv_end; a.max_slots = pl.max_slots;
    a.counters = reinterpret_cast<unsigned int*>(ws);
    a.rec = reinterpret_cast<float*>(ws + pl.ctr_bytes);
    a.timeline = g_timeline;
    a.cluster_k = pl.cluster_k;
    // every CTA's run lies inside one unit (unit-aligned grid, or one chunk per CTA): the fold follows the CTA's last chunk, the deep ring applies
    a.deep_ring = (pl.grid > 0 && pl.n_units > 0 && pl.grid % pl.n_units == 0 && pl.total_chunks >= pl.grid) ? 1 : 0;
    { static const bool nd = tune_env("B200FA_NO_DEEP_RING") != nullptr; if (nd) a.deep_ring = 0; }
    { static const int rg = tune_env("B200FA_RING") ? atoi(tune_env("B200FA_RING")) : 0; a.ring = rg; }
    a.peers = g_seqpar.peers; a.rank = g_seqpar.rank; a.world = g_seqpar.world; a.fdst = g_seqpar.fdst; a.fdst_type = g_seqpar.fdst_type;
    a.mask_bulk = (p.mask != nullptr && ((((uintptr_t)p.mask) | (uintptr_t)p.nb31 | (uintptr_t)p.nb32 | (uintptr_t)p.nb33) % 16) == 0) ? 1 : 0;
    { static const bool nb = tune_env("B200FA_NO_MASK_BULK") != nullptr; if (nb) a.mask_bulk = 0; }
    CUtensorMap tk{}, tv{};
    if (p.kv_type == B200FA_TYPE_F16) {
        if (!make_tile_map(&tk, p.k, p.n_kv, p.n_head_kv / p.kv_div, p.n_batch_kv, p.nb11, p.nb12, p.nb13, DK_CHUNK, p.Dr)) return B200FA_ERR_CUDA;
        if (!make_tile_map(&tv, p.v, p.n_kv, p.n_head_kv / p.kv_div, p.n_batch_kv, p.nb21, p.nb22, p.nb23, DK_CHUNK, p.Dr)) return B200FA_ERR_CUDA;
    }
    const bool q8 = p.kv_type == B200FA_TYPE_Q8_0;
    if (q8) {
        // the contiguous rows of a head as a [lines][128 B] byte tensor (whole lines only: nothing past the head is ever read)
        static const bool no_lines = tune_env("B200FA_Q8_BULK1D") != nullptr;
        const int64_t lines = (int64_t)p.n_kv * (p.D / kQ8BlockElems * kQ8BlockBytes) / 128;
        if (!no_lines && lines > 0) {
            const int box_lines = DK_CHUNK * (p.D / kQ8BlockElems * kQ8BlockBytes) / 128;
            if (!make_line_map(&tk, p.k, lines, p.n_head_kv / p.kv_div, p.n_batch_kv, p.nb12, p.nb13, box_lines)) return B200FA_ERR_CUDA;
            if (!make_line_map(&tv, p.v, lines, p.n_head_kv / p.kv_div, p.n_batch_kv, p.nb22, p.nb23, box_lines)) return B200FA_ERR_CUDA;
            a.q8_lines = (int)(lines > 0x7fffffff ? 0x7fffffff : lines);
        }
    }
    static const int force_rh = tune_env("B200FA_STREAM_RH") ? atoi(tune_env("B200FA_STREAM_RH")) : 0;  // tuning: 2 = always the 16-row variant
    const bool small = (int64_t)p.n_q * p.gqa <= 8 && force_rh != 2;
    g_last_launches++;
    const bool ext = p.cap_in != 0.f || p.alibi_nhl2 != 0 || p.m_ne2 * p.m_ne3 > 1;
    // the fused sequence-parallel step has its own instantiations (SP; the entry takes no score modifiers)
    const bool sp = a.peers != nullptr;
    if (sp && ext) return B200FA_ERR_UNSUPPORTED;
#define B200FA_STREAM_E(DD, KK, EE, SS) (small ? launch_stream_t<DD, KK, 1, EE, false, SS>(p, a, pl.grid, tk, tv, st) : launch_stream_t<DD, KK, 2, EE, false, SS>(p, a, pl.grid, tk, tv, st))
#define B200FA_STREAM(DD, KK) (sp ? B200FA_STREAM_E(DD, KK, false, true) : (ext ? B200FA_STREAM_E(DD, KK, true, false) : B200FA_STREAM_E(DD, KK, false, false)))
#define B200FA_STREAM_T8(DD) (sp ? launch_stream_t<DD, B200FA_TYPE_Q8_0, 1, false, true, true>(p, a, pl.grid, tk, tv, st) \
                                 : (ext ? launch_stream_t<DD, B200FA_TYPE_Q8_0, 1, true, true>(p, a, pl.grid, tk, tv, st) : launch_stream_t<DD, B200FA_TYPE_Q8_0, 1, false, true>(p, a, pl.grid, tk, tv, st)))
    // q8_0 units of at most 8 rows: the transposed tile (decode_stream.cuh, T8)
    static const bool q8_rowmajor = tune_env("B200FA_Q8_ROWMAJOR") != nullptr;
    if (q8 && small && !q8_rowmajor) return p.D == 128 ? B200FA_STREAM_T8(128) : B200FA_STREAM_T8(64);
    if (p.D == 128) return q8 ? B200FA_STREAM(128, B200FA_TYPE_Q8_0) : B200FA_STREAM(128, B200FA_TYPE_F16);
    return q8 ? B200FA_STREAM(64, B200FA_TYPE_Q8_0) : B200FA_STREAM(64, B200FA_TYPE_F16);
#undef B200FA_STREAM_T8
#undef B200FA_STREAM
#undef B200FA_STREAM_E
}

}  // namespace

extern "C" {

const char* b200fa_status_string(int s) {
    switch (s) {
        case B200FA_OK: return "ok";
        case B200FA_ERR_INVALID: return "invalid argument";
        case B200FA_ERR_UNSUPPORTED: return "unsupported shape or type";
        case B200FA_ERR_WORKSPACE: return "workspace missing or too small";
        case B200FA_ERR_CUDA: return "no sm_100 device or launch failed";
        case B200FA_ERR_IO: return "tensor file missing, truncated or not writable";
        default: return "unknown status";
    }
}

int b200fa_version(void) { return 100; }
int b200fa_tensor_file_info(const char* path, b200fa_tensor_info* info) { return tf_info(path, info); }
int b200fa_tensor_file_read(const char* path, void* dst, size_t dst_bytes) { return tf_read(path, dst, dst_bytes); }
int b200fa_tensor_file_write(const char* path, const char* name, int type, int n_dims, const int64_t* ne, const void* data) {
    return tf_write(path, name, type, n_dims, ne, data);
}
void b200fa_debug_set(void* timeout_word, float* dump, int dump_cta) {
#ifdef B200FA_TUNING
    pf_debug().dbg = (unsigned long long*)timeout_word;
    pf_debug().dump = dump;
    pf_debug().dump_cta = dump_cta;
#else
    (void)timeout_word; (void)dump; (void)dump_cta;  // inert in the shipped library
#endif
}
const char* b200fa_last_dispatch(void) { return g_last_dispatch; }
void b200fa_debug_timeline(void* stamps) {
#ifdef B200FA_TUNING
    g_timeline = (unsigned long long*)stamps;
#else
    (void)stamps;  // inert in the shipped library
#endif
}
int b200fa_last_launch_count(void) { return g_last_launches; }

size_t b200fa_workspace_size(int q_type, int kv_type, int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
                             int64_t ne11, int64_t ne12, int64_t ne13, uint32_t flags) {
    return b200fa_workspace_size_ext2(q_type, kv_type, ne00, ne01, ne02, ne03, ne11, ne12, ne13, nullptr, flags);
}

size_t b200fa_workspace_size_ext2(int q_type, int kv_type, int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
                                  int64_t ne11, int64_t ne12, int64_t ne13, const b200fa_ext_params* ext, uint32_t flags) {
    const DeviceInfo& di = device_info();
    const int sms = di.ok ? di.sm_count : 148;
    if (ne12 <= 0 || ne02 % ne12 || ne00 <= 0 || ne01 <= 0 || ne03 <= 0 || ne11 <= 0) return 0;
    if (ne00 > 256) return 0;
    Shape sh = make_shape(q_type, kv_type, ne00, ne01, ne02, ne03, ne11, ne12, ne13 > 0 ? ne13 : ne03, 0, 0, 0, 0, 0, 0, nullptr, nullptr, 0, ne11);
    if (ext) {  // mask slices: one table of mask tile classes per slice on the prefill path
        sh.mask_slices = (ext->mask_ne2 > 0 ? ext->mask_ne2 : 1) * (ext->mask_ne3 > 0 ? ext->mask_ne3 : 1);
        sh.ext = ext->max_bias > 0.f || ext->logit_softcap != 0.f || sh.mask_slices > 1;
    }
    size_t m = 0;
    for (int variant = 0; variant < 5; variant++) {  // every path the two entry points can take for this shape
        const bool partial = variant == 1 || variant == 3;
        const bool stream = variant <= 1;
        const uint32_t f = variant == 4 ? (flags | B200FA_FLAG_NO_TCGEN05) : flags;
        const Plan pl = make_plan(sh, f, sms, partial, true, stream);
        if (pl.total > m) m = pl.total;
    }
    if (const int dv = virtual_head_split(ne01, ne02, ne12, ne03); dv > 1) {  // the virtual-head stream plan has more units and slots
        Shape sv = sh;
        sv.n_head_kv = ne12 * dv;
        for (int partial = 0; partial < 2; partial++) {
            const Plan pl = make_plan(sv, flags, sms, partial != 0, true, true);
            if (pl.total > m) m = pl.total;
        }
    }
    return m + 256;
}

static int attn_common(const void* q, const void* k, const void* v, const void* mask, void* dst, float* partial_out,
                       float scale, int q_type, int kv_type, int dst_type,
                       int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
                       int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
                       int64_t ne31, int64_t nb31, int64_t nb01, int64_t nb02, int64_t nb03,
                       int64_t nb11, int64_t nb12, int64_t nb13, int64_t nb21, int64_t nb22, int64_t nb23,
                       int64_t kv_pos0, int64_t n_kv_total, uint32_t flags, void* workspace, size_t workspace_bytes,
                       cudaStream_t st, const b200fa_ext_params* ext = nullptr) {
    g_last_dispatch = "none";
    g_last_launches = 0;
    const bool want_partial = partial_out != nullptr;
    int rc = validate(q, k, v, want_partial ? (void*)partial_out : dst, q_type, kv_type, dst_type, ne00, ne01, ne02,
                      ne03, ne10, ne11, ne12, ne13, mask, ne31, nb31, nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23, want_partial);
    if (rc != B200FA_OK) return rc;
    const DeviceInfo& di = device_info();
    if (!di.ok || di.cc_major != 10) return B200FA_ERR_CUDA;  // sm_100a only, no fallback

    if (!(scale > 0.f)) flags |= B200FA_FLAG_NO_TCGEN05;  // the tile kernel takes row maxima of raw scores
    const int64_t Dp = ne00 <= 64 ? 64 : (ne00 <= 128 ? 128 : 256);  // structural head size of the decode kernels; the prefill kernel is always 128 wide
    const float max_bias = ext ? ext->max_bias : 0.f, softcap = ext ? ext->logit_softcap : 0.f;
    if (!(max_bias >= 0.f) || !(softcap == softcap) || isinf(softcap) || isinf(max_bias)) return B200FA_ERR_INVALID;
    // mask slices (upstream ggml's ne32 / ne33 broadcast): one mask per head and / or per batch entry instead of the reference's shared one
    const int64_t m_ne2 = ext && ext->mask_ne2 > 0 ? ext->mask_ne2 : 1, m_ne3 = ext && ext->mask_ne3 > 0 ? ext->mask_ne3 : 1;
    if (m_ne2 * m_ne3 > 1) {
        if (!mask || (m_ne2 != 1 && m_ne2 != ne02) || (m_ne3 != 1 && m_ne3 != ne03)) return B200FA_ERR_INVALID;
        if ((m_ne2 > 1 && (ext->mask_nb2 < ne31 * nb31 || ext->mask_nb2 % 2)) || (m_ne3 > 1 && (ext->mask_nb3 <= 0 || ext->mask_nb3 % 2))) return B200FA_ERR_INVALID;
        if (want_partial) return B200FA_ERR_UNSUPPORTED;  // the sequence-split entries take the shared mask only
    }
    Shape sh = make_shape(q_type, kv_type, ne00, ne01, ne02, ne03, ne11, ne12, ne13, nb11, nb12, nb13, nb21, nb22, nb23, k, v, kv_pos0, n_kv_total);
    sh.ext = max_bias > 0.f || softcap != 0.f || m_ne2 * m_ne3 > 1;
    sh.mask_slices = m_ne2 * m_ne3;
    // 17..128 rows per KV head from a GQA group (a burst of up to 16 query positions, e.g. speculative decoding): split every real
    // KV head into kv_div VIRTUAL heads of gqa / kv_div q heads each, so that a unit has <= 16 rows and the stream kernel applies.
    // K/V are streamed kv_div times, but by CTAs working side by side: the repeats are L2 hits (8 positions x GQA 4, batch 8,
    // KV 8192: 190 us on the 16-row-group fallback -> 134 us).
    // (Not when the tensor-core kernel takes the burst with the group packed into one tile: one pass over K/V instead of kv_div.)
    int kv_div = make_plan(sh, flags, di.sm_count, want_partial, false).kind == kPrefill ? 1 : virtual_head_split(ne01, ne02, ne12, ne03);
    if (kv_div > 1) {
        Shape sv = sh;
        sv.n_head_kv = ne12 * kv_div;
        if (stream_eligible(sv, false)) sh = sv; else kv_div = 1;
    }
    Plan pl = make_plan(sh, flags, di.sm_count, want_partial, false);
    if (!workspace || workspace_bytes < pl.total || ((uintptr_t)workspace % 256)) return B200FA_ERR_WORKSPACE;
    char* ws = (char*)workspace;

    FaParams p{};
    p.q = (const char*)q; p.k = (const char*)k; p.v = (const char*)v; p.mask = (const char*)mask;
    p.dst = dst; p.part = nullptr;
    p.scale = scale; p.scale_log2 = scale * kLog2e;
    p.q_type = q_type; p.kv_type = kv_type; p.dst_type = dst_type;
    p.D = (int)Dp; p.Dr = (int)ne00; p.n_q = (int)ne01; p.n_head = (int)ne02; p.n_batch = (int)ne03;
    p.n_kv = (int)ne11; p.n_head_kv = (int)ne12; p.n_batch_kv = (int)ne13;
    p.gqa = (int)(ne02 / ne12); p.rk3 = (int)(ne03 / ne13);
    p.kv_div = 1;
    p.nb01 = nb01; p.nb02 = nb02; p.nb03 = nb03;
    p.nb11 = nb11; p.nb12 = nb12; p.nb13 = nb13;
    p.nb21 = nb21; p.nb22 = nb22; p.nb23 = nb23;
    p.nb31 = nb31;
    p.m_ne2 = (int)m_ne2; p.m_ne3 = (int)m_ne3;
    p.nb32 = m_ne2 > 1 ? ext->mask_nb2 : 0; p.nb33 = m_ne3 > 1 ? ext->mask_nb3 : 0;
    p.causal = (flags & B200FA_FLAG_CAUSAL) ? 1 : 0;
    p.kv_pos0 = kv_pos0;
    p.causal_off = n_kv_total - ne01;
    p.total_rows = ne03 * ne01 * ne02;
    { static const int dm = tune_env("B200FA_DBG_MODE") ? atoi(tune_env("B200FA_DBG_MODE")) : 0; p.dbg_mode = dm; }
    if (max_bias > 0.f) {  // ALiBi slopes (upstream ggml: m0 = 2^(-max_bias/n_head_log2), m1 = 2^(-(max_bias/2)/n_head_log2))
        int nhl2 = 1;
        while (nhl2 * 2 <= (int)ne02) nhl2 *= 2;
        p.alibi_nhl2 = nhl2;
        p.alibi_m0l = -max_bias / (float)nhl2;
        p.alibi_m1l = -(max_bias * 0.5f) / (float)nhl2;
    }
    if (softcap != 0.f) {  // s = cap * tanh(q.k * scale / cap)
        const float cap = fabsf(softcap);  // the expression is even in cap
        p.cap_in = scale / cap;
        p.cap_out = cap * kLog2e;
        p.cap_raw = cap / scale;
    }

    if (pl.kind == kPrefill) {
        g_last_dispatch = "prefill_tcgen05";
        p.D = PF_D;
        int launches = 0;
        if (kv_type == B200FA_TYPE_Q8_0) {  // f16 copies of K and V, dense [batch][head][row][D]
            __half* k16 = reinterpret_cast<__half*>(ws + pl.ctr_bytes + pl.qf16_bytes + pl.cls_bytes + pl.part_bytes);
            __half* v16 = reinterpret_cast<__half*>(reinterpret_cast<char*>(k16) + pl.kvf16_bytes / 2);
            const int64_t n_oct = ne11 * ne12 * ne13 * (ne00 / 8);
            const unsigned gb = (unsigned)((n_oct + 255) / 256);
            q8_rows_to_f16_kernel<<<gb, 256, 0, st>>>(p.k, k16, (int)ne00, (int)ne11, (int)ne12, n_oct, nb11, nb12, nb13);
            q8_rows_to_f16_kernel<<<gb, 256, 0, st>>>(p.v, v16, (int)ne00, (int)ne11, (int)ne12, n_oct, nb21, nb22, nb23);
            if (cudaGetLastError() != cudaSuccess) return B200FA_ERR_CUDA;
            p.k = reinterpret_cast<const char*>(k16); p.v = reinterpret_cast<const char*>(v16);
            p.kv_type = B200FA_TYPE_F16;
            p.nb11 = p.nb21 = ne00 * 2; p.nb12 = p.nb22 = ne11 * ne00 * 2; p.nb13 = p.nb23 = ne12 * ne11 * ne00 * 2;
        }
        static const bool per_cta = tune_env("B200FA_PREFILL") && !strcmp(tune_env("B200FA_PREFILL"), "cta");
        if ((per_cta && p.Dr == PF_D && !sh.ext) || ne11 > (int64_t)PP_MAX_KV_TILES * PF_BN) {
            rc = launch_prefill_tcgen05(p, ws + pl.ctr_bytes, pl.qf16_bytes, pl.cls_bytes, di.sm_count, st, &launches);
        } else {
            if (!(flags & B200FA_FLAG_WORKSPACE_ZEROED) && cudaMemsetAsync(ws + kPrefillCtrOff, 0, 256, st) != cudaSuccess) return B200FA_ERR_CUDA;
            const bool split = pl.n_splits > 1 && !per_cta;
            rc = launch_prefill_persistent(p, ws + pl.ctr_bytes, pl.qf16_bytes, reinterpret_cast<unsigned int*>(ws + kPrefillCtrOff), di.sm_count, st, &launches,
                                           split ? pl.n_splits : 1, split ? reinterpret_cast<float*>(ws + pl.ctr_bytes + pl.qf16_bytes + pl.cls_bytes) : nullptr);
        }
        g_last_launches = launches + (kv_type == B200FA_TYPE_Q8_0 ? 2 : 0);
        return rc;
    }

    p.part_out = partial_out;
    if (pl.n_counters > 0 && (!(flags & B200FA_FLAG_WORKSPACE_ZEROED) || pl.ctr_bytes != kCtrRegion)) {
        if (cudaMemsetAsync(ws, 0, align_up(pl.n_counters * 4, 256), st) != cudaSuccess) return B200FA_ERR_CUDA;
    }
    if (g_seqpar.peers != nullptr && pl.kind != kStream) return B200FA_ERR_UNSUPPORTED;  // the fused step exists in the stream kernel only
    if (pl.kind == kStream) {
        if (kv_div > 1) { p.kv_div = kv_div; p.n_head_kv = (int)ne12 * kv_div; p.gqa = (int)(ne02 / ne12) / kv_div; }
        g_last_dispatch = want_partial ? "decode_stream_partial" : "decode_stream";
        return run_stream(p, pl, ws, st);
    }
    p.n_splits = pl.n_splits; p.split_len = pl.split_len;
    g_last_dispatch = want_partial ? "decode_splitkv_partial" : ((ne01 * p.gqa <= 64) ? "decode_splitkv" : "rows16_mma");
    if (pl.n_splits > 1) {
        p.counters = (unsigned int*)ws;
        p.part = (float*)(ws + pl.ctr_bytes);
    }
    return run_rows16(p, pl, st);
}

int b200fa_flash_attn_ext(const void* q, const void* k, const void* v, const void* mask, void* dst, float scale,
                          int q_type, int kv_type, int dst_type,
                          int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
                          int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
                          int64_t ne31, int64_t nb31, int64_t nb01, int64_t nb02, int64_t nb03,
                          int64_t nb11, int64_t nb12, int64_t nb13, int64_t nb21, int64_t nb22, int64_t nb23,
                          int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3,
                          uint32_t flags, void* workspace, size_t workspace_bytes, b200fa_stream_t stream) {
    if (ne0 != ne00 || ne1 != ne02 || ne2 != ne01 || ne3 != ne03) return B200FA_ERR_INVALID;  // flash-llama.h:434
    return attn_common(q, k, v, mask, dst, nullptr, scale, q_type, kv_type, dst_type, ne00, ne01, ne02, ne03, ne10, ne11,
                       ne12, ne13, ne31, nb31, nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23, 0, ne11, flags,
                       workspace, workspace_bytes, (cudaStream_t)stream);
}

// The host-side decision for a shape, without touching a device: which kernel family, how the work is cut, how much workspace.
// Mirrors attn_common's planning (no pointers: q8_0 rows are assumed contiguous and aligned, the modifiers off).
int b200fa_plan(int q_type, int kv_type, int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03, int64_t ne11, int64_t ne12,
                int64_t ne13, uint32_t flags, int sm_count, b200fa_plan_info* out) {
    if (!out || sm_count < 1) return B200FA_ERR_INVALID;
    if (ne00 <= 0 || ne01 <= 0 || ne02 <= 0 || ne03 <= 0 || ne11 <= 0 || ne12 <= 0 || ne13 <= 0 || ne02 % ne12 || ne03 % ne13) return B200FA_ERR_INVALID;
    if (ne00 % 8 || ne00 > 256) return B200FA_ERR_UNSUPPORTED;
    if (kv_type == B200FA_TYPE_Q8_0 && ne00 != 64 && ne00 != 128 && ne00 != 256) return B200FA_ERR_UNSUPPORTED;
    if (kv_type != B200FA_TYPE_F16 && kv_type != B200FA_TYPE_Q8_0) return B200FA_ERR_UNSUPPORTED;
    const int64_t Dp = ne00 <= 64 ? 64 : (ne00 <= 128 ? 128 : 256);
    const int64_t row = kv_type == B200FA_TYPE_Q8_0 ? Dp / kQ8BlockElems * kQ8BlockBytes : ne00 * 2;
    Shape sh{q_type, kv_type, Dp, ne01, ne02, ne03, ne11, ne12, row, row * ne11, row * ne11 * ne12, row, row * ne11, row * ne11 * ne12,
             (const void*)256, (const void*)256, 0, ne11, ne00, ne13};
    int kv_div = make_plan(sh, flags, sm_count, false, false).kind == kPrefill ? 1 : virtual_head_split(ne01, ne02, ne12, ne03);
    if (kv_div > 1) {
        Shape sv = sh;
        sv.n_head_kv = ne12 * kv_div;
        if (stream_eligible(sv, false)) sh = sv; else kv_div = 1;
    }
    const Plan pl = make_plan(sh, flags, sm_count, false, false);
    memset(out, 0, sizeof(*out));
    out->kind = pl.kind == kPrefill ? B200FA_PLAN_PREFILL : (pl.kind == kStream ? B200FA_PLAN_STREAM : B200FA_PLAN_ROWS16);
    out->kv_div = pl.kind == kStream ? kv_div : 1;
    out->n_splits = pl.kind == kStream ? 1 : pl.n_splits;
    out->grid = pl.kind == kStream ? pl.grid : 0;
    out->cluster_k = pl.kind == kStream ? pl.cluster_k : 0;
    out->kv_f16_copy_bytes = (int64_t)pl.kvf16_bytes;
    out->workspace_bytes = (int64_t)pl.total;
    return B200FA_OK;
}

int b200fa_flash_attn_ext2(const void* q, const void* k, const void* v, const void* mask, void* dst, float scale,
                           int q_type, int kv_type, int dst_type,
                           int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
                           int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
                           int64_t ne31, int64_t nb31, int64_t nb01, int64_t nb02, int64_t nb03,
                           int64_t nb11, int64_t nb12, int64_t nb13, int64_t nb21, int64_t nb22, int64_t nb23,
                           int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3, const b200fa_ext_params* ext,
                           uint32_t flags, void* workspace, size_t workspace_bytes, b200fa_stream_t stream) {
    if (ne0 != ne00 || ne1 != ne02 || ne2 != ne01 || ne3 != ne03) return B200FA_ERR_INVALID;
    return attn_common(q, k, v, mask, dst, nullptr, scale, q_type, kv_type, dst_type, ne00, ne01, ne02, ne03, ne10, ne11,
                       ne12, ne13, ne31, nb31, nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23, 0, ne11, flags,
                       workspace, workspace_bytes, (cudaStream_t)stream, ext);
}

int b200fa_flash_attn_partial(const void* q, const void* k, const void* v, const void* mask, float* partial, float scale,
                              int q_type, int kv_type,
                              int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
                              int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
                              int64_t ne31, int64_t nb31, int64_t nb01, int64_t nb02, int64_t nb03,
                              int64_t nb11, int64_t nb12, int64_t nb13, int64_t nb21, int64_t nb22, int64_t nb23,
                              int64_t kv_pos0, int64_t n_kv_total,
                              uint32_t flags, void* workspace, size_t workspace_bytes, b200fa_stream_t stream) {
    return b200fa_flash_attn_partial2(q, k, v, mask, partial, scale, q_type, kv_type, ne00, ne01, ne02, ne03, ne10, ne11, ne12, ne13, ne31, nb31,
                                      nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23, kv_pos0, n_kv_total, nullptr, flags, workspace,
                                      workspace_bytes, stream);
}

int b200fa_flash_attn_partial2(const void* q, const void* k, const void* v, const void* mask, float* partial, float scale,
                               int q_type, int kv_type,
                               int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
                               int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
                               int64_t ne31, int64_t nb31, int64_t nb01, int64_t nb02, int64_t nb03,
                               int64_t nb11, int64_t nb12, int64_t nb13, int64_t nb21, int64_t nb22, int64_t nb23,
                               int64_t kv_pos0, int64_t n_kv_total, const b200fa_ext_params* ext,
                               uint32_t flags, void* workspace, size_t workspace_bytes, b200fa_stream_t stream) {
    if (!partial || kv_pos0 < 0 || n_kv_total < kv_pos0 + ne11) return B200FA_ERR_INVALID;
    return attn_common(q, k, v, mask, nullptr, partial, scale, q_type, kv_type, B200FA_TYPE_F32, ne00, ne01, ne02, ne03,
                       ne10, ne11, ne12, ne13, ne31, nb31, nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23, kv_pos0,
                       n_kv_total, flags, workspace, workspace_bytes, (cudaStream_t)stream, ext);
}

size_t b200fa_xchg_bytes(int world, int64_t n_rows, int64_t D) {
    if (world < 1 || n_rows < 1 || D < 1) return 0;
    const size_t n_floats = (size_t)n_rows * (size_t)(D + 2);
    return (size_t)xchg_ll_offset(world, (int64_t)n_floats) + 2 * (size_t)world * n_floats * 8;  // header, staging, gathered x 2, flag-in-data x 2
}

int b200fa_flash_attn_partial_scatter(const void* q, const void* k, const void* v, const void* mask, float scale,
                                      int q_type, int kv_type,
                                      int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
                                      int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
                                      int64_t ne31, int64_t nb31, int64_t nb01, int64_t nb02, int64_t nb03,
                                      int64_t nb11, int64_t nb12, int64_t nb13, int64_t nb21, int64_t nb22, int64_t nb23,
                                      int64_t kv_pos0, int64_t n_kv_total, void* xchg, void* const* peers, int rank, int world,
                                      uint32_t flags, void* workspace, size_t workspace_bytes, b200fa_stream_t stream) {
    if (!xchg || !peers || world < 1 || rank < 0 || rank >= world || ((uintptr_t)xchg % 256)) return B200FA_ERR_INVALID;
    if (kv_pos0 < 0 || n_kv_total < kv_pos0 + ne11) return B200FA_ERR_INVALID;
    const int64_t n_rows = ne03 * ne01 * ne02, n_floats = n_rows * (ne00 + 2);
    float* staging = reinterpret_cast<float*>((char*)xchg + kXchgHeader);
    int rc = attn_common(q, k, v, mask, nullptr, staging, scale, q_type, kv_type, B200FA_TYPE_F32, ne00, ne01, ne02, ne03,
                         ne10, ne11, ne12, ne13, ne31, nb31, nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23, kv_pos0,
                         n_kv_total, flags, workspace, workspace_bytes, (cudaStream_t)stream);
    if (rc != B200FA_OK) return rc;
    unsigned blocks = (unsigned)((n_floats + 1023) / 1024);
    if (blocks > 32) blocks = 32;
    if (blocks < 1) blocks = 1;
    fa_scatter_signal<<<blocks, 256, 0, (cudaStream_t)stream>>>((char* const*)peers, rank, world, n_floats);
    g_last_launches++;
    return cudaGetLastError() == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}

int b200fa_flash_attn_seqpar(const void* q, const void* k, const void* v, const void* mask, void* dst, float scale,
                             int q_type, int kv_type, int dst_type,
                             int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
                             int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
                             int64_t ne31, int64_t nb31, int64_t nb01, int64_t nb02, int64_t nb03,
                             int64_t nb11, int64_t nb12, int64_t nb13, int64_t nb21, int64_t nb22, int64_t nb23,
                             int64_t kv_pos0, int64_t n_kv_total, void* xchg, void* const* peers, int rank, int world,
                             uint32_t flags, void* workspace, size_t workspace_bytes, b200fa_stream_t stream) {
    if (!dst || !xchg || !peers || world < 1 || rank < 0 || rank >= world || ((uintptr_t)xchg % 256) || ((uintptr_t)dst % 16)) return B200FA_ERR_INVALID;
    if (dst_type != B200FA_TYPE_F32 && dst_type != B200FA_TYPE_F16) return B200FA_ERR_UNSUPPORTED;
    if (kv_pos0 < 0 || n_kv_total < kv_pos0 + ne11) return B200FA_ERR_INVALID;
    // one launch when the stream decode kernel takes the shape; otherwise the three-launch sequence
    const DeviceInfo& di = device_info();
    if (!di.ok || di.cc_major != 10) return B200FA_ERR_CUDA;
    {   // the same argument checks attn_common applies, BEFORE any planning arithmetic (ne12 = 0 must be an error, not a division)
        const int rcv = validate(q, k, v, (char*)xchg + kXchgHeader, q_type, kv_type, B200FA_TYPE_F32, ne00, ne01, ne02, ne03, ne10, ne11, ne12, ne13,
                                 mask, ne31, nb31, nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23, true);
        if (rcv != B200FA_OK) return rcv;
    }
    const Shape sh = make_shape(q_type, kv_type, ne00, ne01, ne02, ne03, ne11, ne12, ne13, nb11, nb12, nb13, nb21, nb22, nb23, k, v, kv_pos0, n_kv_total);
    const Plan pl = make_plan(sh, flags, di.sm_count, true, false);
    if (pl.kind == kStream) {
        float* staging = reinterpret_cast<float*>((char*)xchg + kXchgHeader);  // never written in this mode; only a non-null partial_out
        g_seqpar.peers = (char* const*)peers; g_seqpar.rank = rank; g_seqpar.world = world; g_seqpar.fdst = dst; g_seqpar.fdst_type = dst_type;
        const int rc = attn_common(q, k, v, mask, nullptr, staging, scale, q_type, kv_type, B200FA_TYPE_F32, ne00, ne01, ne02, ne03,
                                   ne10, ne11, ne12, ne13, ne31, nb31, nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23, kv_pos0,
                                   n_kv_total, flags, workspace, workspace_bytes, (cudaStream_t)stream);
        g_seqpar = SeqPar{};
        if (rc == B200FA_OK) g_last_dispatch = "decode_stream_seqpar";
        return rc;
    }
    int rc = b200fa_flash_attn_partial_scatter(q, k, v, mask, scale, q_type, kv_type, ne00, ne01, ne02, ne03, ne10, ne11, ne12, ne13, ne31,
                                               nb31, nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23, kv_pos0, n_kv_total, xchg, peers, rank,
                                               world, flags, workspace, workspace_bytes, stream);
    if (rc != B200FA_OK) return rc;
    return b200fa_merge_partials_wait(xchg, world, ne03 * ne01 * ne02, ne00, dst, dst_type, stream);
}

int b200fa_merge_partials_wait(void* xchg, int world, int64_t n_rows, int64_t D, void* dst, int dst_type, b200fa_stream_t stream) {
    if (!xchg || !dst || world < 1 || n_rows < 1) return B200FA_ERR_INVALID;
    if (dst_type != B200FA_TYPE_F32 && dst_type != B200FA_TYPE_F16) return B200FA_ERR_UNSUPPORTED;
    const DeviceInfo& di = device_info();
    if (!di.ok || di.cc_major != 10) return B200FA_ERR_CUDA;
    if (D == 128) fa_combine_wait<128><<<(unsigned)n_rows, 128, 0, (cudaStream_t)stream>>>((char*)xchg, world, n_rows, dst, dst_type);
    else if (D == 64) fa_combine_wait<64><<<(unsigned)n_rows, 64, 0, (cudaStream_t)stream>>>((char*)xchg, world, n_rows, dst, dst_type);
    else return B200FA_ERR_UNSUPPORTED;
    return cudaGetLastError() == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}

int b200fa_peer_alloc(size_t bytes, void** ptr, unsigned char handle[64]) {
    if (!ptr || !handle || bytes == 0) return B200FA_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return B200FA_ERR_CUDA;
    if (cudaMemset(p, 0, bytes) != cudaSuccess) { cudaFree(p); return B200FA_ERR_CUDA; }
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); return B200FA_ERR_CUDA; }
    memcpy(handle, &h, 64);
    *ptr = p;
    return B200FA_OK;
}
int b200fa_peer_open(const unsigned char handle[64], void** ptr) {
    if (!ptr || !handle) return B200FA_ERR_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    return cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}
int b200fa_peer_set_timeout(void* xchg, int timeout_ms, b200fa_stream_t stream) {
    if (!xchg || timeout_ms < 0) return B200FA_ERR_INVALID;
    fa_xchg_set_word<<<1, 1, 0, (cudaStream_t)stream>>>((char*)xchg, kXchgTimeoutWord, (unsigned int)timeout_ms);
    return cudaGetLastError() == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}
int b200fa_peer_status(const void* xchg, int* timed_out, b200fa_stream_t stream) {
    if (!xchg || !timed_out) return B200FA_ERR_INVALID;
    unsigned int w = 0;
    if (cudaMemcpyAsync(&w, (const char*)xchg + kXchgErrWord * 4, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return B200FA_ERR_CUDA;
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return B200FA_ERR_CUDA;
    *timed_out = w != 0u;
    return B200FA_OK;
}
int b200fa_peer_reset(void* xchg, b200fa_stream_t stream) {
    if (!xchg) return B200FA_ERR_INVALID;
    fa_xchg_reset<<<1, kXchgHeader / 4, 0, (cudaStream_t)stream>>>((char*)xchg);
    return cudaGetLastError() == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}
int b200fa_peer_close(void* ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA; }
int b200fa_peer_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA; }

int b200fa_workspace_init(void* workspace, size_t workspace_bytes, b200fa_stream_t stream) {
    if (!workspace || !workspace_bytes) return B200FA_ERR_INVALID;
    return cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream) == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}

int b200fa_merge_partials(const float* partials, int n_parts, int64_t n_rows, int64_t D, void* dst, int dst_type,
                          b200fa_stream_t stream) {
    if (!partials || !dst || n_parts < 1 || n_rows < 1) return B200FA_ERR_INVALID;
    if (dst_type != B200FA_TYPE_F32 && dst_type != B200FA_TYPE_F16) return B200FA_ERR_UNSUPPORTED;
    const DeviceInfo& di = device_info();
    if (!di.ok || di.cc_major != 10) return B200FA_ERR_CUDA;
    if (D == 128) return launch_combine<128>(partials, n_parts, n_rows, dst, dst_type, (cudaStream_t)stream);
    if (D == 64) return launch_combine<64>(partials, n_parts, n_rows, dst, dst_type, (cudaStream_t)stream);
    return B200FA_ERR_UNSUPPORTED;
}

int b200fa_quantize_q8_0(const void* src, int src_type, void* dst, int64_t n, b200fa_stream_t stream) {
    if (!src || !dst || n <= 0 || n % kQ8BlockElems) return B200FA_ERR_INVALID;
    const DeviceInfo& di = device_info();
    if (!di.ok || di.cc_major != 10) return B200FA_ERR_CUDA;
    const int64_t nb = n / kQ8BlockElems;
    const unsigned grid = (unsigned)((nb + 7) / 8);
    if (src_type == B200FA_TYPE_F32) q8_0_quantize_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, (uint8_t*)dst, nb);
    else if (src_type == B200FA_TYPE_F16) q8_0_quantize_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)src, (uint8_t*)dst, nb);
    else return B200FA_ERR_UNSUPPORTED;
    return cudaGetLastError() == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}

int b200fa_dequantize_q8_0(const void* src, float* dst, int64_t n, b200fa_stream_t stream) {
    if (!src || !dst || n <= 0 || n % kQ8BlockElems) return B200FA_ERR_INVALID;
    const DeviceInfo& di = device_info();
    if (!di.ok || di.cc_major != 10) return B200FA_ERR_CUDA;
    const int64_t nb = n / kQ8BlockElems;
    q8_0_dequantize_kernel<<<(unsigned)((nb + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)src, dst, nb);
    return cudaGetLastError() == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}

int b200fa_kv_cache_append(const void* src, int src_type, void* cache, int cache_type, int64_t D, int64_t n_tokens, int64_t n_head_kv,
                           int64_t n_batch, int64_t src_nb1, int64_t src_nb2, int64_t src_nb3, int64_t cache_nb1, int64_t cache_nb2,
                           int64_t cache_nb3, int64_t n_past, int64_t n_kv_max, b200fa_stream_t stream) {
    if (!src || !cache || D <= 0 || n_tokens <= 0 || n_head_kv <= 0 || n_batch <= 0 || n_past < 0) return B200FA_ERR_INVALID;
    if (D % (cache_type == B200FA_TYPE_Q8_0 ? 32 : 8)) return B200FA_ERR_INVALID;  // the head sizes the attention entries take for that cache type
    if (n_kv_max <= 0 || n_past + n_tokens > n_kv_max) return B200FA_ERR_INVALID;   // the rows would land past the end of the cache
    if (src_type != B200FA_TYPE_F32 && src_type != B200FA_TYPE_F16) return B200FA_ERR_UNSUPPORTED;
    if (cache_type != B200FA_TYPE_F16 && cache_type != B200FA_TYPE_Q8_0) return B200FA_ERR_UNSUPPORTED;
    const int64_t es = src_type == B200FA_TYPE_F32 ? 4 : 2;
    if (((uintptr_t)src | src_nb1 | src_nb2 | src_nb3) % es || ((uintptr_t)cache | cache_nb1 | cache_nb2 | cache_nb3) % 2) return B200FA_ERR_INVALID;
    const DeviceInfo& di = device_info();
    if (!di.ok || di.cc_major != 10) return B200FA_ERR_CUDA;
    const int64_t rows = n_tokens * n_head_kv * n_batch;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (src_type == B200FA_TYPE_F32)
        kv_append_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const char*)src, (char*)cache, cache_type, (int)D, (int)n_tokens, (int)n_head_kv,
                                                                      rows, src_nb1, src_nb2, src_nb3, cache_nb1, cache_nb2, cache_nb3, n_past);
    else
        kv_append_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>((const char*)src, (char*)cache, cache_type, (int)D, (int)n_tokens, (int)n_head_kv,
                                                                       rows, src_nb1, src_nb2, src_nb3, cache_nb1, cache_nb2, cache_nb3, n_past);
    return cudaGetLastError() == cudaSuccess ? B200FA_OK : B200FA_ERR_CUDA;
}

}  // extern "C"

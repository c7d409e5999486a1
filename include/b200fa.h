/*
 * b200fa.h — C ABI of the B200-native flash-attention path (sm_100a).
 *
 * Drop-in boundary for the launcher interface prototyped in FSSRepo/ggml-cuda-experiments.
 * The reference has no plugin/FFI layer: its boundary is the kernel signature plus the launch
 * geometry at the call sites.  Each entry point below names the reference interface it replaces.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes; every pointer is caller-owned DEVICE memory unless stated;
 *   - work is enqueued on `stream` of the CURRENT device; nothing synchronises, nothing allocates;
 *   - returns B200FA_OK (0) or a negative B200FA_ERR_* code; never aborts.  Asynchronous faults
 *     surface at the caller's next cudaGetLastError/cudaStreamSynchronize, as with the reference's
 *     raw <<<>>> launches (flash-matrix.cu:198-206, which check nothing);
 *   - thread-safe for concurrent calls on distinct streams with distinct workspaces.
 *   - there is NO CPU fallback: on a machine without an sm_100 device every launch entry returns
 *     B200FA_ERR_CUDA.
 *
 * Tensor description is ggml's: ne = element counts, nb = byte strides, fastest dimension first.
 *   q    : [ne00=D, ne01=n_q,  ne02=n_head,    ne03=n_batch]  f32 (reference) or f16
 *   k, v : [ne10=D, ne11=n_kv, ne12=n_head_kv, ne13=n_batch_kv] f16, or q8_0 (34-byte blocks of 32)
 *   mask : f16 [n_kv, ne31 >= n_q] rows = queries, row stride nb31 bytes, shared by heads/batches (per-head / per-sequence
 *          slices: b200fa_flash_attn_ext2); may be NULL
 *   dst  : [D, n_head, n_q, n_batch] contiguous, f32 (reference) or f16
 *   GQA  : kv head = q head / (ne02/ne12); batch broadcast likewise (flash-llama.h:128-140).
 */
#ifndef B200FA_H
#define B200FA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ggml type tags (values are ggml's GGML_TYPE_*) */
#define B200FA_TYPE_F32 0
#define B200FA_TYPE_F16 1
#define B200FA_TYPE_Q8_0 8

/* flags */
#define B200FA_FLAG_CAUSAL 1u      /* mask (if given) is exactly the causal one: 0 where kv <= q + (n_kv - n_q),
                                      -inf elsewhere.  Lets the kernels synthesise it and skip masked tiles. */
#define B200FA_FLAG_NO_TCGEN05 2u  /* diagnostics: force the register-streaming kernel even for prefill shapes */
#define B200FA_FLAG_WORKSPACE_ZEROED 4u /* the caller zero-filled the workspace once (b200fa_workspace_init or cudaMemset) and
                                          only b200fa calls have touched it since: skips the per-call cudaMemsetAsync of the
                                          split-KV arrival counters (every call leaves them zero again) */

/* status codes */
#define B200FA_OK 0
#define B200FA_ERR_INVALID (-1)      /* NULL pointer, inconsistent ne/nb, misaligned rows */
#define B200FA_ERR_UNSUPPORTED (-2)  /* valid but not built: head size, type combination */
#define B200FA_ERR_WORKSPACE (-3)    /* workspace NULL or smaller than b200fa_workspace_size() */
#define B200FA_ERR_CUDA (-4)         /* no sm_100 device, or the launch itself failed */
#define B200FA_ERR_IO (-5)           /* tensor-dump file missing, unreadable, truncated or not writable */

typedef void* b200fa_stream_t; /* a cudaStream_t */

const char* b200fa_status_string(int status);
int b200fa_version(void); /* major*10000 + minor*100 + patch */

/*
 * dst = softmax(scale * Q K^T + mask) V
 *
 * Replaces:  flash_attn_ext_f16<D,Q,C><<<grid,block,smem,stream>>>(q,k,v,mask,dst,scale, ne.., nb..)
 *            flash-llama.h:5-32 (signature), launched at flash-matrix.cu:198-206 and kernel_test.h:191-198;
 *            and, for batch-1 decode, the pair
 *            flash_attn_row<128,8,2,256> + fa_reduce<128,8>   flash_row_float.h:4-6, :415-416,
 *            launched at flash-matrix.cu:226-227 and kernel_test.h:161-162
 *            (whose V^T / dense-K repacking and cudaMalloc'd scratch the caller no longer needs).
 * The argument list is the reference's, widened to int64 and extended by: type tags (the reference
 * fixes f32 Q / f16 KV / f32 dst), separate V strides (the reference aliases nb2x = nb1x,
 * flash-llama.h:123-125), flags, a caller-owned workspace (the reference cudaMallocs its split-KV
 * scratch at flash-matrix.cu:223-224) and the stream.
 *
 * Dispatch (all on the GPU):
 *   n_q <= 16, rows per KV head (n_q * n_head/n_head_kv) <= 16
 *                      -> stream-K decode kernel: TMA-fed ring, mma.sync fragments, split-KV merged in the same launch (HBM-bound)
 *   n_q <= 16, more than 16 rows under GQA with a power-of-two group of 2..32 q heads
 *                      -> the tile kernel below with the group's q heads PACKED into one 128-row tile (row = position x head):
 *                         one pass over K/V per KV head
 *   n_q <= 16, 17..128 rows under GQA with any other group size (or with the ext2 modifiers)
 *                      -> the stream kernel over virtual KV heads of <= 16 rows each
 *   n_q > 16           -> tcgen05/TMEM/TMA tile kernel (prefill; tensor-bound); q8_0 K/V are dequantised once to f16 workspace copies;
 *                         with fewer work items than SMs and a long KV range the KV tiles are split into segments merged by a
 *                         second launch; under GQA (power-of-two groups) tiles are packed whenever that saves passes over K/V
 *   anything else (more than 128 rows from <= 16 positions, scale <= 0, odd q8_0 alignments, sequences beyond the tile schedule)
 *                      -> the register-streaming kernel over 16-row groups
 * Requirements: head size D = ne00: any multiple of 8 up to 128 with f16 K/V (64 and 128 run natively; the others run
 * zero-padded on the fly on the 64/128-wide kernels, nothing is copied), 64 or 128 with q8_0 K/V and in the partial /
 * sequence-parallel entries; rows of q/k/v 16-byte aligned for f16/f32 (nb % 16 == 0), 2-byte for q8_0.
 */
int b200fa_flash_attn_ext(
    const void* q, const void* k, const void* v, const void* mask, void* dst, float scale,
    int q_type, int kv_type, int dst_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
    int64_t ne31, int64_t nb31,
    int64_t nb01, int64_t nb02, int64_t nb03,
    int64_t nb11, int64_t nb12, int64_t nb13,
    int64_t nb21, int64_t nb22, int64_t nb23,
    int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3,
    uint32_t flags, void* workspace, size_t workspace_bytes, b200fa_stream_t stream);

/*
 * The same call with the two score modifiers of upstream ggml's flash_attn_ext (NOT in the reference,
 * whose kernel has neither — semantics restated from ggml, parity unpinned by the reference):
 *     s = q.k * scale;   if (logit_softcap != 0) s = logit_softcap * tanh(s / logit_softcap);   s += slope(head) * mask
 *     slope(h) = 1 when max_bias == 0, else ALiBi: n2 = 2^floor(log2(n_head)), m0 = 2^(-max_bias/n2), m1 = 2^(-max_bias/2/n2),
 *                h < n2 ? m0^(h+1) : m1^(2(h-n2)+1)
 * ext == NULL or {0, 0} is exactly b200fa_flash_attn_ext.  B200FA_FLAG_CAUSAL still means "the mask is exactly 0 / -inf causal".
 */
typedef struct b200fa_ext_params {
    float max_bias;       /* >= 0; 0 = no ALiBi */
    float logit_softcap;  /* 0 = off */
    /* Mask slices — upstream ggml's mask broadcast over heads (ne32) and batch entries (ne33); the reference shares ONE mask between
     * all heads and batches (flash-llama.h:151,194), which is mask_ne2 = mask_ne3 = 0 or 1 here.  mask_ne2 in {1, ne02}: one mask per
     * head; mask_ne3 in {1, ne03}: one per batch entry (per-sequence masks of a batched decode).  The row of (query iq1, head iq2,
     * batch iq3) starts at mask + iq1*nb31 + (iq2 % mask_ne2)*mask_nb2 + (iq3 % mask_ne3)*mask_nb3; every slice has ne31 rows.
     * Not accepted by the sequence-split entries.  Workspace: b200fa_workspace_size_ext2. */
    int64_t mask_ne2, mask_ne3, mask_nb2, mask_nb3;
} b200fa_ext_params;
int b200fa_flash_attn_ext2(
    const void* q, const void* k, const void* v, const void* mask, void* dst, float scale,
    int q_type, int kv_type, int dst_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
    int64_t ne31, int64_t nb31,
    int64_t nb01, int64_t nb02, int64_t nb03,
    int64_t nb11, int64_t nb12, int64_t nb13,
    int64_t nb21, int64_t nb22, int64_t nb23,
    int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3,
    const b200fa_ext_params* ext,
    uint32_t flags, void* workspace, size_t workspace_bytes, b200fa_stream_t stream);

/* Bytes of workspace b200fa_flash_attn_ext needs for this shape on the current device
 * (split-KV partials, an f16 copy of an f32 Q for the tcgen05 path, mask tile classes).
 * Replaces the inline cudaMalloc of flash-matrix.cu:223-224 / kernel_test.h:153-155. */
size_t b200fa_workspace_size(
    int q_type, int kv_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne11, int64_t ne12, int64_t ne13, uint32_t flags);

/* The same for b200fa_flash_attn_ext2 with these extensions (mask slices add one table of mask tile classes per slice on the
 * prefill path); ext == NULL is b200fa_workspace_size. */
size_t b200fa_workspace_size_ext2(
    int q_type, int kv_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne11, int64_t ne12, int64_t ne13, const struct b200fa_ext_params* ext, uint32_t flags);

/* cudaMemsetAsync(workspace, 0, bytes) on `stream`; see B200FA_FLAG_WORKSPACE_ZEROED. */
int b200fa_workspace_init(void* workspace, size_t workspace_bytes, b200fa_stream_t stream);

/*
 * Sequence-split building blocks (cross-GPU split-KV).  Same maths as the reference's intra-GPU pair
 * flash_attn_row (per-block partial O, m, l; flash_row_float.h:164-197) and fa_reduce (:415-472),
 * with the state kept in f32 instead of f16.
 *
 * b200fa_flash_attn_partial: like b200fa_flash_attn_ext over THIS device's slice of the KV sequence, but
 * instead of dst it writes, per output row r = (batch*n_q + q)*n_head + head, the unnormalised triple
 *     partial[r*(D+2) + 0..D-1] = sum_kv exp(s - m) * V      partial[r*(D+2) + D] = m      [.. + D+1] = l
 * with s = scale*q.k + mask in natural-log units; m = -inf and l = 0 for a row that saw no visible key.
 * `kv_pos0` is the global position of this slice's first key (used only with B200FA_FLAG_CAUSAL).
 */
int b200fa_flash_attn_partial(
    const void* q, const void* k, const void* v, const void* mask, float* partial, float scale,
    int q_type, int kv_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
    int64_t ne31, int64_t nb31,
    int64_t nb01, int64_t nb02, int64_t nb03,
    int64_t nb11, int64_t nb12, int64_t nb13,
    int64_t nb21, int64_t nb22, int64_t nb23,
    int64_t kv_pos0, int64_t n_kv_total,
    uint32_t flags, void* workspace, size_t workspace_bytes, b200fa_stream_t stream);

/* The same with the score modifiers of b200fa_flash_attn_ext2 (ALiBi slopes on this slice's mask columns, logit soft-cap); mask
 * slices are not accepted here.  ext == NULL is b200fa_flash_attn_partial. */
int b200fa_flash_attn_partial2(
    const void* q, const void* k, const void* v, const void* mask, float* partial, float scale,
    int q_type, int kv_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
    int64_t ne31, int64_t nb31,
    int64_t nb01, int64_t nb02, int64_t nb03,
    int64_t nb11, int64_t nb12, int64_t nb13,
    int64_t nb21, int64_t nb22, int64_t nb23,
    int64_t kv_pos0, int64_t n_kv_total, const struct b200fa_ext_params* ext,
    uint32_t flags, void* workspace, size_t workspace_bytes, b200fa_stream_t stream);

/* Merge n_parts partial triples per row (partials[p][n_rows][D+2], e.g. the all-gathered per-GPU
 * results) into dst[n_rows][D].  Replaces fa_reduce<128,nw> (flash_row_float.h:415-472).  (Inside one GPU the
 * split-KV kernel merges its own splits: the last CTA of a row group to finish does it.) */
int b200fa_merge_partials(const float* partials, int n_parts, int64_t n_rows, int64_t D,
                          void* dst, int dst_type, b200fa_stream_t stream);

/*
 * Sequence-split combine over peer-mapped memory (NVLink 5 / NVSwitch), without NCCL on the path.
 * Every rank owns an *exchange buffer* of b200fa_xchg_bytes(world, n_rows, D) bytes, zero-filled once, mapped into every
* peer (cudaIpc: b200fa_peer_alloc / b200fa_peer_open below, or any other peer mapping).  `xchg` is this rank's own buffer,
 * `peers` a DEVICE array of `world` pointers to all ranks' buffers as seen from this rank (peers[rank] == xchg).  Per step
 * (the step number is kept on the device, so the pair of calls is a fixed launch sequence that a CUDA graph can replay):
 *   b200fa_flash_attn_partial_scatter : like b200fa_flash_attn_partial, but the triples go into this rank's slot of its
 *        own exchange buffer and are then stored into the same slot of every peer's buffer (plain NVLink stores), followed
 *        by a system-scope fence and one atomic increment of each peer's arrival counter;
*   b200fa_merge_partials_wait        : waits (on the device) until all `world` ranks have published this step, then merges
 *        (fa_reduce algebra) into dst.  The wait is BOUNDED and never traps: after the exchange's timeout (b200fa_peer_set_timeout,
 *        default 4 s) the kernel raises the buffer's error flag, skips the merge (dst is left untouched, the step is not counted)
 *        and ends normally — poll b200fa_peer_status, and b200fa_peer_reset on every rank before using the exchange again.
 * The ranks' kernels must be able to run AT THE SAME TIME: one rank per device.  Several ranks on one device cannot
 * (a kernel that waits for a peer launched behind it never sees it arrive; B200_PROFILING.md) — emulate them with the plain
 * b200fa_flash_attn_partial + b200fa_merge_partials pair instead.
 * Two generations of the gathered area (step parity) make it safe for a fast rank to start the next step early.
 * b200fa_flash_attn_seqpar (the fused one-kernel step, below) uses a second, flag-in-data gathered area of the same buffer: every
 * float travels as one 8-byte store {value, tag of the step} and readers poll the elements they need — no fence, no counter.
 * (On a timeout of the fused step the rows merged before the missing rank was given up on stay written; the step is not counted.)
 */
size_t b200fa_xchg_bytes(int world, int64_t n_rows, int64_t D);
int b200fa_flash_attn_partial_scatter(
    const void* q, const void* k, const void* v, const void* mask, float scale,
    int q_type, int kv_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
    int64_t ne31, int64_t nb31,
    int64_t nb01, int64_t nb02, int64_t nb03,
    int64_t nb11, int64_t nb12, int64_t nb13,
    int64_t nb21, int64_t nb22, int64_t nb23,
    int64_t kv_pos0, int64_t n_kv_total,
    void* xchg, void* const* peers, int rank, int world,
    uint32_t flags, void* workspace, size_t workspace_bytes, b200fa_stream_t stream);
int b200fa_merge_partials_wait(void* xchg, int world, int64_t n_rows, int64_t D, void* dst, int dst_type, b200fa_stream_t stream);
/* The whole sequence-parallel step in one call — and, for decode shapes (<= 16 rows per KV head), in ONE kernel: the stream
 * decode kernel stores each unit's triple straight into every rank's exchange buffer over NVLink as {value, step tag} pairs, and
 * the CTAs that published a unit poll the other ranks' elements of their share of the output until the tags match, then merge into
 * dst [rows][D] (no fence, no counter, no second NVLink round trip).  dst holds the same result on every rank. */
int b200fa_flash_attn_seqpar(
    const void* q, const void* k, const void* v, const void* mask, void* dst, float scale,
    int q_type, int kv_type, int dst_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
    int64_t ne31, int64_t nb31,
    int64_t nb01, int64_t nb02, int64_t nb03,
    int64_t nb11, int64_t nb12, int64_t nb13,
    int64_t nb21, int64_t nb22, int64_t nb23,
    int64_t kv_pos0, int64_t n_kv_total,
    void* xchg, void* const* peers, int rank, int world,
    uint32_t flags, void* workspace, size_t workspace_bytes, b200fa_stream_t stream);
/* cudaMalloc'd, zero-filled buffer + its cudaIpc handle (64 bytes) / mapping of a peer's handle / release. */
int b200fa_peer_alloc(size_t bytes, void** ptr, unsigned char handle[64]);
int b200fa_peer_open(const unsigned char handle[64], void** ptr);
int b200fa_peer_close(void* ptr);
int b200fa_peer_free(void* ptr);
/* Timeout (ms, 0 = the 4 s default) of the device-side waits on this rank's exchange buffer; error flag raised by a timed-out wait
 * (this call synchronises `stream`); header back to step 0 (arrivals, counters, error flag — the timeout is kept): to be called by
 * EVERY rank at the same point of the protocol, after a step that failed or timed out on any of them. */
int b200fa_peer_set_timeout(void* xchg, int timeout_ms, b200fa_stream_t stream);
int b200fa_peer_status(const void* xchg, int* timed_out, b200fa_stream_t stream);
int b200fa_peer_reset(void* xchg, b200fa_stream_t stream);

/*
 * ggml q8_0 rows (block {f16 d; int8 qs[32]}, 34 bytes).  Not in the reference;
 * format and rounding are ggml's published ones.  n_elements % 32 == 0.
 *   quantize : d = amax/127 (stored f16), q = roundf(x * (1/d))       src f32 or f16
 *   dequantize: y = f32(d) * q                                         bit-exact, dst f32
 */
int b200fa_quantize_q8_0(const void* src, int src_type, void* dst, int64_t n_elements, b200fa_stream_t stream);
int b200fa_dequantize_q8_0(const void* src, float* dst, int64_t n_elements, b200fa_stream_t stream);

/*
 * KV-cache append: writes `n_tokens` new K (or V) rows per (kv head, batch) into the cache tensor at row `n_past`,
 * converting f32/f16 -> f16 or -> q8_0 blocks on the way.  Replaces the host-side repacking loops of the reference's driver
 * (flash-matrix.cu:130-165).  src element (d, tok, head, b) at src + tok*src_nb1 + head*src_nb2 + b*src_nb3 (+ d*elem);
 * cache row (n_past + tok, head, b) at cache + (n_past+tok)*cache_nb1 + head*cache_nb2 + b*cache_nb3 — the same nb the
 * attention entry takes for that tensor.  D: a multiple of 8 for an f16 cache, of 32 for q8_0 (the head sizes the attention entries
 * accept for that cache type).  n_kv_max: rows the cache holds per (head, batch); n_past + n_tokens > n_kv_max is B200FA_ERR_INVALID.
 */
int b200fa_kv_cache_append(const void* src, int src_type, void* cache, int cache_type, int64_t D, int64_t n_tokens,
                           int64_t n_head_kv, int64_t n_batch, int64_t src_nb1, int64_t src_nb2, int64_t src_nb3,
                           int64_t cache_nb1, int64_t cache_nb2, int64_t cache_nb3, int64_t n_past, int64_t n_kv_max,
                           b200fa_stream_t stream);

/*
 * ggml tensor-dump files (host side, no GPU involved): the fixture format the reference replays
 * (loader: utils.h:110-150; use: flash-matrix.cu:69-73, files fa-cuda-{q,k,v,mask,qkv}-256.tensor).
 *   i32 n_dims | i32 type (0 f32, 1 f16) | i32 ne[n_dims] | i32 name_len | name | raw data (ne[0] fastest)
 * _info parses the header and checks that the whole payload is present; _read copies the payload into `dst`
 * (host memory, >= data_bytes); _write produces a file the reference's loader accepts (name <= 19 chars, its
 * field is char[20]).
 */
typedef struct b200fa_tensor_info {
    int32_t n_dims;
    int32_t type;         /* B200FA_TYPE_F32 | B200FA_TYPE_F16 */
    int64_t ne[4];        /* unused trailing dimensions are 1 */
    char name[64];
    int64_t data_offset;  /* byte offset of the payload in the file */
    int64_t data_bytes;
} b200fa_tensor_info;
int b200fa_tensor_file_info(const char* path, b200fa_tensor_info* info);
int b200fa_tensor_file_read(const char* path, void* dst, size_t dst_bytes);
int b200fa_tensor_file_write(const char* path, const char* name, int type, int n_dims, const int64_t* ne, const void* data);

/*
 * Planning query (host only, no device needed): what b200fa_flash_attn_ext would do for a shape on a GPU with `sm_count` SMs —
 * kernel family, virtual KV heads per real one (GQA bursts), KV splits (16-row kernel: splits across CTAs; tile kernel: split-KV
 * prefill segments), stream-K grid and cluster size, bytes of f16 copies made of a q8_0 cache, workspace bytes of this plan.
 */
#define B200FA_PLAN_PREFILL 0 /* tcgen05 tile kernel */
#define B200FA_PLAN_STREAM 1  /* stream-K decode kernel */
#define B200FA_PLAN_ROWS16 2  /* register-streaming kernel over 16-row groups */
typedef struct b200fa_plan_info {
    int32_t kind, kv_div, n_splits, grid, cluster_k, reserved;
    int64_t kv_f16_copy_bytes, workspace_bytes;
} b200fa_plan_info;
int b200fa_plan(int q_type, int kv_type, int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03, int64_t ne11, int64_t ne12,
                int64_t ne13, uint32_t flags, int sm_count, b200fa_plan_info* out);

/* Diagnostics: name of the kernel family the last b200fa_flash_attn_ext call on this thread dispatched to
 * ("decode_splitkv", "prefill_tcgen05", "rows16_mma"), and how many kernels it launched. */
const char* b200fa_last_dispatch(void);
/* The two hooks below are INERT in the shipped library: they (and every environment knob) exist only in builds with
 * -DB200FA_TUNING (ggml-cuda-experiments_b200/build.py build(tuning=True)), where their state is per calling thread.
 * Diagnostics for the tcgen05 kernel: `timeout_word` (8 bytes, device-visible, e.g. mapped host memory) receives a
 * code if an mbarrier wait times out (the kernel then traps instead of hanging); `dump` (device, (2*128*128+256) f32)
 * receives the first raw score tile, the unnormalised output tile and (l, m) of CTA `dump_cta`.  NULL disables. */
void b200fa_debug_set(void* timeout_word, float* dump, int dump_cta);
int b200fa_last_launch_count(void);
/* Diagnostics for the stream decode kernel: `stamps` (device, 8 x uint64 per CTA) receives %globaltimer values at
 * kernel start, first landed stage, end of streaming, after the in-CTA fold and at the end.  NULL disables. */
void b200fa_debug_timeline(void* stamps);

#ifdef __cplusplus
}
#endif
#endif /* B200FA_H */

/*
 * attn_oracle.c — CPU oracle for the flash-attention hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (ggml-cuda-experiments_b200/) never links, imports or calls it.
 *
 * It is a plain-C restatement of the reference's host attention, generalised from the dense
 * per-head buffers the reference's drivers use to the ggml ne/nb strides the kernel boundary
 * takes (flash-llama.h:6-32):
 *
 *   scores  = scale * Q K^T + mask      reference: mulmat_cpu(float*, half*, half*)  utils.h:18-28
 *   P       = softmax(scores)           reference: softmax()                          utils.h:30-49
 *   O       = P V                       reference: mulmat_cpu(..., B_transposed)      utils.h:18-28,
 *                                       sequenced as test_llama does                  flash-matrix.cu:88-102
 *   dst     = [batch][q][head][d]       reference: permute                            flash-matrix.cu:105-111,
 *                                       and the kernel's own store                    flash-llama.h:434
 *   GQA     : kv head = q head / (ne02/ne12), batch broadcast likewise               flash-llama.h:128-140
 *
 * Arithmetic follows the reference exactly: fp32 accumulation in k order, the online (M,S)
 * pass followed by expf(s-M)/S, fp32 accumulation of P·V in kv order.  The Q operand is used
 * as given when it is f32 (test_llama path) or widened from f16; `round_q_f16` reproduces the
 * kernel_test path (utils.h:5-16) that rounds Q through f16 first.
 *
 * One deliberate, switchable deviation: the reference softmax returns NaN when the FIRST score
 * of a row is -inf (utils.h:37-41: expf(-inf - -inf)).  `strict_ref != 0` keeps that behaviour;
 * `strict_ref == 0` treats exp(-inf - M) as 0 so rows with a masked prefix stay finite (what the
 * reference's CUDA kernels do, flash-llama.h:240-243).  A row that is masked everywhere yields
 * NaN in strict mode and zeros otherwise.
 *
 * q8_0 is NOT in the reference.  The block format and rounding restated here are
 * ggml's published ones (block_q8_0 {f16 d; int8 qs[32]}, d = amax/127, q = roundf(x * (1/d)),
 * y = f32(d) * q) and are pinned against gguf==0.19.0 `gguf/quants.py:378-401` by
 * tests/golden/make_golden.py — "parity unpinned by the reference" for that format.
 *
 * Build: see oracle/Makefile (gcc -O2 -fPIC -shared -pthread).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_TYPE_F32 0
#define ORACLE_TYPE_F16 1
#define ORACLE_TYPE_Q8_0 8
#define QK8_0 32

/* ---- IEEE binary16 <-> binary32, bit exact (round to nearest even on narrowing) ---- */
static inline float h2f(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f; memcpy(&f, &bits, 4); return f;
}

static inline uint16_t f2h(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) { /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u | ((ax >> 13) & 0x3ffu) : 0));
    }
    if (ax >= 0x477ff000u) { /* rounds to >= 65520 -> inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x38800000u) { /* subnormal half or zero */
        if (ax < 0x33000000u) return (uint16_t)sign; /* < 2^-25 -> 0 (ties to even -> 0) */
        int e = (int)(ax >> 23);
        uint32_t man = (ax & 0x7fffffu) | 0x800000u;
        int shift = 126 - e; /* 14..24 */
        uint32_t half_man = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half_man & 1))) half_man++;
        return (uint16_t)(sign | half_man);
    }
    uint32_t e = (ax >> 23) - 127 + 15;
    uint32_t man = ax & 0x7fffffu;
    uint32_t h = (e << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;
    return (uint16_t)(sign | h);
}

void oracle_f16_to_f32(const uint16_t* src, float* dst, int64_t n) {
    for (int64_t i = 0; i < n; i++) dst[i] = h2f(src[i]);
}
void oracle_f32_to_f16(const float* src, uint16_t* dst, int64_t n) {
    for (int64_t i = 0; i < n; i++) dst[i] = f2h(src[i]);
}

/* ---- ggml q8_0 (published format; see header) ---- */
void oracle_quantize_q8_0(const float* x, uint8_t* y, int64_t n) {
    int64_t nb = n / QK8_0;
    for (int64_t b = 0; b < nb; b++) {
        float amax = 0.0f;
        for (int j = 0; j < QK8_0; j++) {
            float v = fabsf(x[b * QK8_0 + j]);
            if (v > amax) amax = v;
        }
        const float d = amax / 127.0f;
        const float id = d ? 1.0f / d : 0.0f;
        uint16_t dh = f2h(d);
        memcpy(y + b * 34, &dh, 2);
        for (int j = 0; j < QK8_0; j++) {
            float v = x[b * QK8_0 + j] * id;
            ((int8_t*)(y + b * 34 + 2))[j] = (int8_t)roundf(v);
        }
    }
}

void oracle_dequantize_q8_0(const uint8_t* x, float* y, int64_t n) {
    int64_t nb = n / QK8_0;
    for (int64_t b = 0; b < nb; b++) {
        uint16_t dh; memcpy(&dh, x + b * 34, 2);
        const float d = h2f(dh);
        const int8_t* qs = (const int8_t*)(x + b * 34 + 2);
        for (int j = 0; j < QK8_0; j++) y[b * QK8_0 + j] = d * (float)qs[j];
    }
}

/* ---- the attention restatement ---- */
typedef struct {
    const char* q; const char* k; const char* v; const char* mask; char* dst;
    float scale;
    int q_type, kv_type, dst_type;
    int64_t ne00, ne01, ne02, ne03, ne10, ne11, ne12, ne13, ne31, nb31;
    int64_t nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23;
    int64_t ne0, ne1, ne2, ne3;
    int round_q_f16, strict_ref;
    float max_bias, logit_softcap; /* upstream ggml_flash_attn_ext extensions (0 = off); see oracle_flash_attn_ext2 */
    int64_t ne32, ne33, nb32, nb33; /* mask slices per head / batch (upstream ggml broadcast; 1, 1 = the reference's shared mask) */
} oracle_args;

/* widen one K/V row (ne10 elements) to f32 */
static void load_kv_row(const char* row, int kv_type, int64_t D, float* out) {
    if (kv_type == ORACLE_TYPE_F16) {
        const uint16_t* r = (const uint16_t*)row;
        for (int64_t i = 0; i < D; i++) out[i] = h2f(r[i]);
    } else if (kv_type == ORACLE_TYPE_Q8_0) {
        oracle_dequantize_q8_0((const uint8_t*)row, out, D);
    } else {
        memcpy(out, row, (size_t)D * 4);
    }
}

/* one (batch, head): all ne01 query rows.  scratch: scores[ne11], kvrow[D], qrow[D], acc[D] */
static void attend_head(const oracle_args* a, int64_t iq3, int64_t iq2, float* scratch) {
    const int64_t D = a->ne00, n_q = a->ne01, n_kv = a->ne11;
    const int64_t rk2 = a->ne02 / a->ne12, rk3 = a->ne03 / a->ne13; /* flash-llama.h:128-140 */
    const int64_t ik2 = iq2 / rk2, ik3 = iq3 / rk3;
    float* scores = scratch;
    float* kvrow = scores + n_kv;
    float* qrow = kvrow + D;
    float* acc = qrow + D;

    for (int64_t iq1 = 0; iq1 < n_q; iq1++) {
        const char* qp = a->q + iq1 * a->nb01 + iq2 * a->nb02 + iq3 * a->nb03; /* flash-llama.h:71 */
        for (int64_t i = 0; i < D; i++) {
            float x = (a->q_type == ORACLE_TYPE_F16) ? h2f(((const uint16_t*)qp)[i]) : ((const float*)qp)[i];
            if (a->round_q_f16) x = h2f(f2h(x)); /* utils.h:10 */
            qrow[i] = x;
        }
        const uint16_t* mrow = a->mask ? (const uint16_t*)(a->mask + iq1 * a->nb31 + (iq2 % a->ne32) * a->nb32 + (iq3 % a->ne33) * a->nb33)
                                       : NULL; /* flash-llama.h:151 (shared mask: ne32 = ne33 = 1) */
        /* ALiBi slope of this head and logit soft-cap — upstream ggml semantics, NOT in the reference (see oracle_flash_attn_ext2) */
        float slope = 1.0f;
        if (a->max_bias > 0.0f) {
            const uint32_t n_head_log2 = 1u << (uint32_t)floor(log2((double)a->ne02));
            const float m0 = powf(2.0f, -(a->max_bias) / n_head_log2);
            const float m1 = powf(2.0f, -(a->max_bias / 2.0f) / n_head_log2);
            const uint32_t h = (uint32_t)iq2;
            slope = h < n_head_log2 ? powf(m0, (float)(h + 1)) : powf(m1, (float)(2 * (h - n_head_log2) + 1));
        }
        const float cap = a->logit_softcap;
        const float qk_scale = cap != 0.0f ? a->scale / cap : a->scale;

        /* scores = scale * q·k + mask   (utils.h:18-28: acc over k, then acc*scale + mask) */
        for (int64_t ic = 0; ic < n_kv; ic++) {
            load_kv_row(a->k + ic * a->nb11 + ik2 * a->nb12 + ik3 * a->nb13, a->kv_type, D, kvrow);
            float s = 0.0f;
            for (int64_t i = 0; i < D; i++) s += qrow[i] * kvrow[i];
            s *= qk_scale;
            if (cap != 0.0f) s = cap * tanhf(s);
            scores[ic] = s + (mrow ? slope * h2f(mrow[ic]) : 0.0f);
        }

        /* softmax (utils.h:30-49): online (M,S), then expf(s-M)/S */
        float M = -INFINITY, S = 0.0f;
        for (int64_t ic = 0; ic < n_kv; ic++) {
            float s = scores[ic];
            if (s > M) {
                S = 1.0f + S * expf(M - s);
                M = s;
            } else if (a->strict_ref || s != -INFINITY) {
                S += expf(s - M);
            }
        }
        for (int64_t ic = 0; ic < n_kv; ic++) {
            if (!a->strict_ref && (scores[ic] == -INFINITY)) scores[ic] = 0.0f;
            else scores[ic] = expf(scores[ic] - M) / S;
        }
        if (!a->strict_ref && M == -INFINITY) {
            for (int64_t ic = 0; ic < n_kv; ic++) scores[ic] = 0.0f;
        }

        /* O = P V, accumulate over kv in order (utils.h:18-28 with K=n_kv) */
        for (int64_t i = 0; i < D; i++) acc[i] = 0.0f;
        for (int64_t ic = 0; ic < n_kv; ic++) {
            load_kv_row(a->v + ic * a->nb21 + ik2 * a->nb22 + ik3 * a->nb23, a->kv_type, D, kvrow);
            const float p = scores[ic];
            for (int64_t i = 0; i < D; i++) acc[i] += p * kvrow[i];
        }

        /* dst[(iq3*ne2*ne1 + iq2 + iq1*ne1)*D + i]   (flash-llama.h:434) */
        const int64_t o = (iq3 * a->ne2 * a->ne1 + iq2 + iq1 * a->ne1) * D;
        if (a->dst_type == ORACLE_TYPE_F16) {
            uint16_t* d = (uint16_t*)a->dst + o;
            for (int64_t i = 0; i < D; i++) d[i] = f2h(acc[i]);
        } else {
            float* d = (float*)a->dst + o;
            for (int64_t i = 0; i < D; i++) d[i] = acc[i];
        }
    }
}

typedef struct { const oracle_args* a; int tid, nthreads; } worker_arg;

static void* worker(void* p) {
    const worker_arg* w = (const worker_arg*)p;
    const oracle_args* a = w->a;
    float* scratch = (float*)malloc(sizeof(float) * (size_t)(a->ne11 + 3 * a->ne00));
    const int64_t units = a->ne02 * a->ne03;
    for (int64_t u = w->tid; u < units; u += w->nthreads) attend_head(a, u / a->ne02, u % a->ne02, scratch);
    free(scratch);
    return NULL;
}

/* Same argument list as the C-ABI entry (include/b200fa.h), minus flags/workspace/stream, plus the
 * two oracle switches and a thread count.  Returns 0, or -1 on an argument it cannot interpret.
 *
 * max_bias / logit_softcap (b200fa_flash_attn_ext2) are NOT in the reference: PARITY UNPINNED for them.
 * They restate the published semantics of upstream ggml's ggml_flash_attn_ext (ggml.c, flash_attn_ext_f16 forward, not vendored):
 *     n_head_log2 = 2^floor(log2(n_head));  m0 = 2^(-max_bias/n_head_log2);  m1 = 2^(-(max_bias/2)/n_head_log2)
 *     slope(h)    = max_bias > 0 ? (h < n_head_log2 ? m0^(h+1) : m1^(2(h-n_head_log2)+1)) : 1
 *     s           = q.k * scale            (scale /= softcap first when softcap != 0)
 *     s           = softcap * tanhf(s)     (when softcap != 0)
 *     s          += slope(h) * mask[q][k]
 */
int oracle_flash_attn_ext2(
    const void* q, const void* k, const void* v, const void* mask, void* dst, float scale,
    int q_type, int kv_type, int dst_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
    int64_t ne31, int64_t nb31,
    int64_t nb01, int64_t nb02, int64_t nb03,
    int64_t nb11, int64_t nb12, int64_t nb13,
    int64_t nb21, int64_t nb22, int64_t nb23,
    int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3,
    int round_q_f16, int strict_ref, int nthreads, float max_bias, float logit_softcap);

int oracle_flash_attn_ext3(
    const void* q, const void* k, const void* v, const void* mask, void* dst, float scale,
    int q_type, int kv_type, int dst_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
    int64_t ne31, int64_t nb31,
    int64_t nb01, int64_t nb02, int64_t nb03,
    int64_t nb11, int64_t nb12, int64_t nb13,
    int64_t nb21, int64_t nb22, int64_t nb23,
    int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3,
    int round_q_f16, int strict_ref, int nthreads, float max_bias, float logit_softcap,
    int64_t ne32, int64_t ne33, int64_t nb32, int64_t nb33);

int oracle_flash_attn_ext(
    const void* q, const void* k, const void* v, const void* mask, void* dst, float scale,
    int q_type, int kv_type, int dst_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
    int64_t ne31, int64_t nb31,
    int64_t nb01, int64_t nb02, int64_t nb03,
    int64_t nb11, int64_t nb12, int64_t nb13,
    int64_t nb21, int64_t nb22, int64_t nb23,
    int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3,
    int round_q_f16, int strict_ref, int nthreads)
{
    return oracle_flash_attn_ext2(q, k, v, mask, dst, scale, q_type, kv_type, dst_type, ne00, ne01, ne02, ne03, ne10, ne11, ne12, ne13,
                                  ne31, nb31, nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23, ne0, ne1, ne2, ne3,
                                  round_q_f16, strict_ref, nthreads, 0.0f, 0.0f);
}

int oracle_flash_attn_ext2(
    const void* q, const void* k, const void* v, const void* mask, void* dst, float scale,
    int q_type, int kv_type, int dst_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
    int64_t ne31, int64_t nb31,
    int64_t nb01, int64_t nb02, int64_t nb03,
    int64_t nb11, int64_t nb12, int64_t nb13,
    int64_t nb21, int64_t nb22, int64_t nb23,
    int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3,
    int round_q_f16, int strict_ref, int nthreads, float max_bias, float logit_softcap)
{
    return oracle_flash_attn_ext3(q, k, v, mask, dst, scale, q_type, kv_type, dst_type, ne00, ne01, ne02, ne03, ne10, ne11, ne12, ne13,
                                  ne31, nb31, nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23, ne0, ne1, ne2, ne3,
                                  round_q_f16, strict_ref, nthreads, max_bias, logit_softcap, 1, 1, 0, 0);
}

/* + mask slices: the mask row of (iq1, iq2, iq3) is mask + iq1*nb31 + (iq2 % ne32)*nb32 + (iq3 % ne33)*nb33 (upstream ggml's
 * broadcast of the mask over heads and batch entries; not in the reference, whose mask is shared: flash-llama.h:151,194) */
int oracle_flash_attn_ext3(
    const void* q, const void* k, const void* v, const void* mask, void* dst, float scale,
    int q_type, int kv_type, int dst_type,
    int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
    int64_t ne10, int64_t ne11, int64_t ne12, int64_t ne13,
    int64_t ne31, int64_t nb31,
    int64_t nb01, int64_t nb02, int64_t nb03,
    int64_t nb11, int64_t nb12, int64_t nb13,
    int64_t nb21, int64_t nb22, int64_t nb23,
    int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3,
    int round_q_f16, int strict_ref, int nthreads, float max_bias, float logit_softcap,
    int64_t ne32, int64_t ne33, int64_t nb32, int64_t nb33)
{
    if (!q || !k || !v || !dst) return -1;
    if (ne32 < 1 || ne33 < 1) return -1;
    if (ne00 != ne10 || ne00 <= 0 || ne12 <= 0 || ne13 <= 0) return -1;
    if (ne02 % ne12 || ne03 % ne13) return -1;
    if (kv_type == ORACLE_TYPE_Q8_0 && (ne00 % QK8_0)) return -1;
    if (q_type != ORACLE_TYPE_F32 && q_type != ORACLE_TYPE_F16) return -1;
    oracle_args a = {
        (const char*)q, (const char*)k, (const char*)v, (const char*)mask, (char*)dst, scale,
        q_type, kv_type, dst_type,
        ne00, ne01, ne02, ne03, ne10, ne11, ne12, ne13, ne31, nb31,
        nb01, nb02, nb03, nb11, nb12, nb13, nb21, nb22, nb23,
        ne0, ne1, ne2, ne3, round_q_f16, strict_ref, max_bias, logit_softcap, ne32, ne33, nb32, nb33 };
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads == 1) {
        worker_arg w = { &a, 0, 1 };
        worker(&w);
        return 0;
    }
    pthread_t th[256];
    worker_arg wa[256];
    for (int t = 0; t < nthreads; t++) {
        wa[t].a = &a; wa[t].tid = t; wa[t].nthreads = nthreads;
        pthread_create(&th[t], NULL, worker, &wa[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}

/* Split-KV merge algebra restated from fa_reduce (flash_row_float.h:429-451, 461-471):
 * running pairwise merge M=max(M0,M1); S=S0*e^(M0-M)+S1*e^(M1-M); O likewise; final O/S.
 * Here in fp32 (the reference keeps the state in f16).  partials: n_parts x (m, l, O[D]). */
void oracle_merge_partials(const float* m, const float* l, const float* O, int64_t n_parts, int64_t D, float* out) {
    float M0 = m[0], S0 = l[0];
    for (int64_t i = 0; i < D; i++) out[i] = O[i];
    for (int64_t p = 1; p < n_parts; p++) {
        float M1 = m[p], S1 = l[p];
        float M = M0 > M1 ? M0 : M1;
        float ms0 = (M0 == -INFINITY) ? 0.0f : expf(M0 - M);
        float ms1 = (M1 == -INFINITY) ? 0.0f : expf(M1 - M);
        S0 = S0 * ms0 + S1 * ms1;
        for (int64_t i = 0; i < D; i++) out[i] = out[i] * ms0 + O[p * D + i] * ms1;
        M0 = M;
    }
    for (int64_t i = 0; i < D; i++) out[i] = S0 > 0.0f ? out[i] / S0 : 0.0f;
}

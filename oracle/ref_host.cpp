// ref_host.cpp — thin extern "C" shim around the reference's OWN host attention.
// TEST INFRASTRUCTURE ONLY (see oracle/attn_oracle.c header for who may load it).
//
// Nothing from the reference is copied here: utils.h is #included from /root/reference/src at
// build time (oracle/Makefile passes -I$(REF)/src) and the functions below only sequence its
// mulmat_cpu / softmax calls the way the reference's drivers do:
//   ref_host_attention_llama  -> test_llama   (flash-matrix.cu:88-111)  f32 Q, f16 K, f16 V^T, 2-D f16 mask
//   ref_host_attention_ktest  -> kernel_test  (kernel_test.h:50-61)     f32 Q/K/V rounded via f16, 1-D f32 mask
// Output goes to oracle/_ref/libref_host.so (git-ignored, travels to the GPU box).
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
#include "utils.h"

extern "C" {

// Q  f32 [n_head][n_q][D]; K f16 [n_head_kv][n_kv][D]; VT f16 [n_head_kv][D][n_kv];
// mask f16 [n_q][n_kv] or NULL; out f32 [n_q][n_head][D] (the permuted layout, flash-matrix.cu:105-111).
// scores: scratch f32 [n_head][n_q][n_kv].  Heads are spread over `nthreads` std::threads.
int ref_host_attention_llama(const float* Q, const uint16_t* K, const uint16_t* VT, const uint16_t* mask,
                             float* out, float* scores, int D, int n_q, int n_kv, int n_head, int n_head_kv,
                             float scale, int nthreads) {
    if (n_head % n_head_kv) return -1;
    const int r = n_head / n_head_kv;
    std::vector<float> tmp((size_t)n_head * n_q * D);
    auto run = [&](int t) {
        for (int h = t; h < n_head; h += nthreads) {
            float* sc = scores + (size_t)h * n_kv * n_q;
            mulmat_cpu(Q + (size_t)h * D * n_q, (const half*)K + (size_t)(h / r) * D * n_kv, (const half*)mask, sc,
                       n_q, n_kv, D, scale, true);
            softmax(sc, n_kv, n_q, h);
            mulmat_cpu(sc, (const half*)VT + (size_t)(h / r) * D * n_kv, nullptr, tmp.data() + (size_t)h * D * n_q,
                       n_q, D, n_kv, 1.0f, true);
        }
    };
    if (nthreads <= 1) { nthreads = 1; run(0); }
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(run, t);
        for (auto& x : th) x.join();
    }
    for (int h = 0; h < n_head; h++)
        for (int b = 0; b < n_q; b++)
            for (int i = 0; i < D; i++)
                out[(size_t)b * n_head * D + (size_t)h * D + i] = tmp[(size_t)h * n_q * D + (size_t)b * D + i];
    return 0;
}

// kernel_test.h:50-61 — batch 1, f32 buffers rounded through f16 inside mulmat_cpu, 1-D mask, V [kv][D].
int ref_host_attention_ktest(const float* query, const float* key, const float* value, const float* mask,
                             float* qkv, float* scores, int D, int n_kv, int n_head, int n_head_kv, float scale) {
    const int r = n_head / n_head_kv;
    for (int h = 0; h < n_head; h++) {
        mulmat_cpu(query + h * D, key + (size_t)(h / r) * D * n_kv, mask, scores + (size_t)h * n_kv, 1, n_kv, D, scale, true);
        softmax(scores + (size_t)h * n_kv, n_kv, 1, h);
    }
    for (int h = 0; h < n_head; h++)
        mulmat_cpu(scores + (size_t)h * n_kv, value + (size_t)(h / r) * D * n_kv, nullptr, qkv + h * D, 1, D, n_kv, 1.0f);
    return 0;
}

// utils.h:110-150 — the reference's own tensor-dump loader, used to check that files written by the product's writer
// load there.  The loader does not return ne, so the caller says how many payload bytes to copy out.
int ref_host_load_tensor(const char* path, int* type, char* name20, void* dst, int64_t bytes) {
    tensor* t = load_tensor_from_file(path);
    if (!t) return -1;
    *type = t->type;
    memcpy(name20, t->name, 20);
    memcpy(dst, t->data, (size_t)bytes);
    free(t->data);
    delete t;
    return 0;
}

int ref_host_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"

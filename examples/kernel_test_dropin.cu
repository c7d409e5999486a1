// kernel_test_dropin.cu — the reference's `kernel_test` flow with its kernel launches replaced by ONE call of the C ABI.
//
// The reference driver (src/kernel_test.h:25-61, 125-234) fills query / key / value / mask with 1 - 2 rand()/RAND_MAX, computes
// softmax(scale Q K^T + mask) V on the host, repacks K and V^T by hand, cudaMallocs the split-KV scratch and launches
//   flash_attn_row<128,8,2,256><<<(n_kv/256, n_head), (32,8)>>> + fa_reduce<128,8>        (kernel_test.h:161-162)   and
//   flash_attn_ext_f16<128,16,128><<<(1, n_head, 1), (32,2)>>>                              (kernel_test.h:191-198),
// then prints the result beside the host values.  This file is that caller, written against include/b200fa.h: CUDA C++ host code,
// plain cudaMalloc'd buffers, the reference's argument order and the SAME tensor layouts (f32 Q [head][D], f16 K/V [kv head][kv][D]
// — V is NOT transposed —, f16 mask [32 padded rows][kv]), no repack, no second launch.  Build (also done by __graft_entry__.build()):
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iinclude examples/kernel_test_dropin.cu \
//        -Lggml-cuda-experiments_b200/_build -lb200fa -Xlinker -rpath -Xlinker '$ORIGIN/../../ggml-cuda-experiments_b200/_build' \
//        -o examples/_build/kernel_test_dropin
// Prints "max_abs_diff <x>" and exits 0 when every output is inside |x - ref| <= 2e-3 + 1e-2 |ref|.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "b200fa.h"

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

static float rnd() { return 1.0f - 2.0f * ((float)rand() / (float)RAND_MAX); }  // the reference's recipe (utils.h:57-61), seeded

int main(int argc, char** argv) {
    const int head_dim = 128, n_head = 32, r_kv_heads = 4, n_kv_head = n_head / r_kv_heads;  // kernel_test.h:25-33
    const int kv_size = argc > 1 ? atoi(argv[1]) : 4096;
    const float scale = 1.0f / sqrtf((float)head_dim);
    srand(1234);
    std::vector<float> query((size_t)n_head * head_dim), ref((size_t)n_head * head_dim);
    std::vector<__half> key((size_t)n_kv_head * kv_size * head_dim), value(key.size()), mask((size_t)32 * kv_size);
    for (auto& x : query) x = rnd();
    for (auto& x : key) x = __float2half(rnd());
    for (auto& x : value) x = __float2half(rnd());
    for (int i = 0; i < kv_size; i++) mask[i] = __float2half(i % 7 == 3 ? -INFINITY : 0.25f * rnd());  // row 0 is the live row
    for (size_t i = kv_size; i < mask.size(); i++) mask[i] = __float2half(0.f);                          // 31 rows of padding

    // host attention, fp32, as the reference's driver computes it (kernel_test.h:50-61: Q and K rounded through f16)
    std::vector<float> s(kv_size);
    for (int h = 0; h < n_head; h++) {
        const __half* k = key.data() + (size_t)(h / r_kv_heads) * kv_size * head_dim;
        const __half* v = value.data() + (size_t)(h / r_kv_heads) * kv_size * head_dim;
        float m = -INFINITY;
        for (int j = 0; j < kv_size; j++) {
            float acc = 0.f;
            for (int d = 0; d < head_dim; d++) acc += __half2float(__float2half(query[(size_t)h * head_dim + d])) * __half2float(k[(size_t)j * head_dim + d]);
            s[j] = acc * scale + __half2float(mask[j]);
            m = fmaxf(m, s[j]);
        }
        double sum = 0.0;
        for (int j = 0; j < kv_size; j++) { s[j] = expf(s[j] - m); sum += s[j]; }
        for (int d = 0; d < head_dim; d++) {
            double acc = 0.0;
            for (int j = 0; j < kv_size; j++) acc += (double)s[j] * __half2float(v[(size_t)j * head_dim + d]);
            ref[(size_t)h * head_dim + d] = (float)(acc / sum);
        }
    }

    float *d_q = nullptr, *d_dst = nullptr;
    __half *d_k = nullptr, *d_v = nullptr, *d_mask = nullptr;
    void* d_ws = nullptr;
    CHECK(cudaMalloc(&d_q, query.size() * 4)); CHECK(cudaMalloc(&d_dst, ref.size() * 4));
    CHECK(cudaMalloc(&d_k, key.size() * 2)); CHECK(cudaMalloc(&d_v, value.size() * 2)); CHECK(cudaMalloc(&d_mask, mask.size() * 2));
    CHECK(cudaMemcpy(d_q, query.data(), query.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_k, key.data(), key.size() * 2, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_v, value.data(), value.size() * 2, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_mask, mask.data(), mask.size() * 2, cudaMemcpyHostToDevice));
    // the caller-owned scratch that replaces the reference's cudaMalloc of (n_blocks*128 + 2*n_blocks)*n_head halves (flash-matrix.cu:223-224)
    const size_t ws_bytes = b200fa_workspace_size(B200FA_TYPE_F32, B200FA_TYPE_F16, head_dim, 1, n_head, 1, kv_size, n_kv_head, 1, 0);
    CHECK(cudaMalloc(&d_ws, ws_bytes));
    cudaStream_t stream;
    CHECK(cudaStreamCreate(&stream));

    // the reference's call (kernel_test.h:191-198), argument for argument: ne0x = Q dims, ne1x = K dims, ne31/nb31 = mask rows / row bytes,
    // nb0x = Q strides (f32, dense per head), nb1x = K strides (f16, [kv head][kv][D]), ne0..3 = dst dims
    const int rc = b200fa_flash_attn_ext(
        d_q, d_k, d_v, d_mask, d_dst, scale, B200FA_TYPE_F32, B200FA_TYPE_F16, B200FA_TYPE_F32,
        head_dim, 1, n_head, 1, head_dim, kv_size, n_kv_head, 1, 32, (int64_t)kv_size * 2,
        head_dim * 4, head_dim * 4, (int64_t)head_dim * n_head * 4,
        head_dim * 2, (int64_t)head_dim * kv_size * 2, (int64_t)head_dim * kv_size * n_kv_head * 2,
        head_dim * 2, (int64_t)head_dim * kv_size * 2, (int64_t)head_dim * kv_size * n_kv_head * 2,
        head_dim, n_head, 1, 1, 0, d_ws, ws_bytes, stream);
    if (rc != B200FA_OK) { fprintf(stderr, "b200fa_flash_attn_ext: %s\n", b200fa_status_string(rc)); return 3; }
    CHECK(cudaStreamSynchronize(stream));
    std::vector<float> got(ref.size());
    CHECK(cudaMemcpy(got.data(), d_dst, got.size() * 4, cudaMemcpyDeviceToHost));

    double max_abs = 0.0; int bad = 0;
    for (size_t i = 0; i < got.size(); i++) {
        const double e = fabs((double)got[i] - ref[i]);
        if (e > max_abs) max_abs = e;
        if (!(e <= 2e-3 + 1e-2 * fabs(ref[i]))) bad++;
    }
    printf("kernel_test drop-in: %d heads (%d kv heads), head_dim %d, kv %d: dispatch %s, launches %d\n", n_head, n_kv_head, head_dim, kv_size,
           b200fa_last_dispatch(), b200fa_last_launch_count());
    printf("R (-0.314) CUDA: %.4f  host: %.4f\n", got[0], ref[0]);  // the reference prints its first values the same way (kernel_test.h:226-233)
    printf("max_abs_diff %.3e\n", max_abs);
    cudaFree(d_q); cudaFree(d_dst); cudaFree(d_k); cudaFree(d_v); cudaFree(d_mask); cudaFree(d_ws); cudaStreamDestroy(stream);
    return bad == 0 ? 0 : 1;
}

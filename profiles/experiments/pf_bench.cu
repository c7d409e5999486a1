// pf_bench.cu — tuning harness for the prefill kernel alone: compiles csrc/prefill_persistent.cuh into a small binary (seconds,
// not the minutes of the whole library), times it with CUDA events on rotating K/V sets and compares the result with the
// shipped library (libb200fa.so, parity-green against the oracle) on the same inputs.  Not part of the product or the tests.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I ggml-cuda-experiments_b200/csrc \
//        profiles/experiments/pf_bench.cu -L ggml-cuda-experiments_b200/_build -lb200fa -o pf_bench
//   ./pf_bench [n_q] [n_kv] [heads] [kv_heads] [causal: 0 none | 1 flag | 2 mask tensor] [iters] [f16out] [K/V sets to rotate over, <= 8]
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "common.cuh"
#include "decode_mma.cuh"
#include "prefill_tcgen05.cuh"
#include "prefill_persistent.cuh"

using namespace b200fa;

__global__ void fill_half(__half* x, size_t n, uint32_t seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u ^ seed;
        h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
        x[i] = __float2half((float)(h & 0xffffff) / 8388608.f - 1.f);
    }
}
__global__ void fill_causal_mask(__half* m, int n_q, int n_kv) {
    const int off = n_kv - n_q;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)n_q * n_kv; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / n_kv), c = (int)(i % n_kv);
        m[i] = c <= r + off ? __float2half(0.f) : __ushort_as_half(0xfc00);
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
    const int n_q = argc > 1 ? atoi(argv[1]) : 2048, n_kv = argc > 2 ? atoi(argv[2]) : 2048;
    const int H = argc > 3 ? atoi(argv[3]) : 32, Hk = argc > 4 ? atoi(argv[4]) : H;
    const int causal = argc > 5 ? atoi(argv[5]) : 1, iters = argc > 6 ? atoi(argv[6]) : 200;
    const int f16out = argc > 7 ? atoi(argv[7]) : 0;
    const int D = 128, kMaxSets = 8;
    const int nsets = argc > 8 ? atoi(argv[8]) : 4;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const size_t nq_el = (size_t)H * n_q * D, nk_el = (size_t)Hk * n_kv * D;
    __half *q, *k[kMaxSets], *v[kMaxSets], *mask = nullptr;
    CK(cudaMalloc(&q, nq_el * 2)); fill_half<<<1024, 256>>>(q, nq_el, 1);
    for (int s = 0; s < nsets; s++) {
        CK(cudaMalloc(&k[s], nk_el * 2)); CK(cudaMalloc(&v[s], nk_el * 2));
        fill_half<<<1024, 256>>>(k[s], nk_el, 10 + s); fill_half<<<1024, 256>>>(v[s], nk_el, 20 + s);
    }
    if (causal == 2) { CK(cudaMalloc(&mask, (size_t)n_q * n_kv * 2)); fill_causal_mask<<<1024, 256>>>(mask, n_q, n_kv); }
    const size_t out_bytes = nq_el * (f16out ? 2 : 4);
    void *dst, *dst_ref; CK(cudaMalloc(&dst, out_bytes)); CK(cudaMalloc(&dst_ref, out_bytes));
    const int64_t wsz = b200fa_workspace_size(B200FA_TYPE_F16, B200FA_TYPE_F16, D, n_q, H, 1, n_kv, Hk, 1, 0);
    char *ws, *ws_ref; CK(cudaMalloc(&ws, wsz)); CK(cudaMalloc(&ws_ref, wsz));
    CK(cudaMemset(ws, 0, wsz)); CK(cudaMemset(ws_ref, 0, wsz));
    const float scale = 1.f / sqrtf((float)D);
    cudaStream_t st; CK(cudaStreamCreate(&st));

    auto params = [&](int s) {
        FaParams p{};
        p.q = (const char*)q; p.k = (const char*)k[s]; p.v = (const char*)v[s]; p.mask = (const char*)mask; p.dst = dst;
        p.scale = scale; p.scale_log2 = scale * kLog2e;
        p.q_type = B200FA_TYPE_F16; p.kv_type = B200FA_TYPE_F16; p.dst_type = f16out ? B200FA_TYPE_F16 : B200FA_TYPE_F32;
        p.D = D; p.Dr = D; p.n_q = n_q; p.n_head = H; p.n_batch = 1; p.n_kv = n_kv; p.n_head_kv = Hk; p.n_batch_kv = 1;
        p.gqa = H / Hk; p.rk3 = 1; p.kv_div = 1;
        p.nb01 = D * 2; p.nb02 = (int64_t)n_q * D * 2; p.nb03 = (int64_t)H * n_q * D * 2;
        p.nb11 = p.nb21 = D * 2; p.nb12 = p.nb22 = (int64_t)n_kv * D * 2; p.nb13 = p.nb23 = (int64_t)Hk * n_kv * D * 2;
        p.nb31 = (int64_t)n_kv * 2; p.m_ne2 = p.m_ne3 = 1;
        p.causal = causal == 1; p.causal_off = n_kv - n_q; p.total_rows = (int64_t)n_q * H;
        return p;
    };
    // workspace: [256 KiB counters][...]; the persistent kernel's counters sit at the end of the counter region (b200fa_api.cu)
    const size_t ctr_region = 65536 * sizeof(unsigned int) + 256;
    auto run = [&](int s) {
        FaParams p = params(s);
        int launches = 0;
        return launch_prefill_persistent(p, ws + ctr_region, 0, reinterpret_cast<unsigned int*>(ws + 65536 * sizeof(unsigned int)), prop.multiProcessorCount, st, &launches);
    };
    int rc = run(0);
    CK(cudaStreamSynchronize(st));
    if (rc != 0) { printf("launch rc=%d\n", rc); return 1; }
    // reference: the shipped library on set 0
    rc = b200fa_flash_attn_ext(q, k[0], v[0], mask, dst_ref, scale, B200FA_TYPE_F16, B200FA_TYPE_F16, f16out ? B200FA_TYPE_F16 : B200FA_TYPE_F32,
                               D, n_q, H, 1, D, n_kv, Hk, 1, mask ? n_q : 0, (int64_t)n_kv * 2, D * 2, (int64_t)n_q * D * 2, (int64_t)H * n_q * D * 2,
                               D * 2, (int64_t)n_kv * D * 2, (int64_t)Hk * n_kv * D * 2, D * 2, (int64_t)n_kv * D * 2, (int64_t)Hk * n_kv * D * 2,
                               D, H, n_q, 1, causal == 1 ? B200FA_FLAG_CAUSAL : 0, ws_ref, wsz, st);
    CK(cudaStreamSynchronize(st));
    if (rc != 0) { printf("reference rc=%d\n", rc); return 1; }
    {
        std::vector<char> a(out_bytes), b(out_bytes);
        CK(cudaMemcpy(a.data(), dst, out_bytes, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), dst_ref, out_bytes, cudaMemcpyDeviceToHost));
        double mx = 0, mref = 0; size_t bad = 0;
        for (size_t i = 0; i < nq_el; i++) {
            const float x = f16out ? __half2float(((__half*)a.data())[i]) : ((float*)a.data())[i];
            const float y = f16out ? __half2float(((__half*)b.data())[i]) : ((float*)b.data())[i];
            const double d = fabs((double)x - y);
            if (!(d <= 2e-3 + 1e-2 * fabs(y))) bad++;
            if (d > mx || d != d) mx = d;
            if (fabs(y) > mref) mref = fabs(y);
        }
        printf("parity vs shipped library: max_abs=%.3e (max |ref| %.3f) out-of-tolerance=%zu of %zu\n", mx, mref, bad, nq_el);
    }
#ifdef B200FA_TUNING
    if (getenv("PF_DUMP_CTA")) {  // clock64 / %globaltimer stamps (tuning build: -DB200FA_TUNING): per-item timeline of one CTA, per-CTA entry / exit
        const size_t dump_bytes = 2 * 32 * 8 * 8 + 160 * 4 * 8;
        float* dump; CK(cudaMalloc(&dump, dump_bytes));
        pf_debug().dump = dump; pf_debug().dump_cta = atoi(getenv("PF_DUMP_CTA"));
        CK(cudaMemset(dump, 0, dump_bytes));
        run(0); CK(cudaStreamSynchronize(st)); run(0); CK(cudaStreamSynchronize(st));   // the second (warm) run's stamps stay
        long long h[2][32][8]; CK(cudaMemcpy(h, dump, sizeof(h), cudaMemcpyDeviceToHost));
        const long long t0 = h[0][0][0] && h[0][0][0] < h[1][0][0] ? h[0][0][0] : h[1][0][0];
        for (int t = 0; t < 2; t++) {
            printf("tile %d: item | work | halves | start | loop end | O read out | stored   (cycles since the CTA's first item)\n", t);
            for (int k = 0; k < 32 && h[t][k][0]; k++)
                printf("   %2d | %5lld | %3lld | %7lld | %7lld | %7lld | %7lld | %5.0f per half (incl. the wait for the item's first scores)\n", k, h[t][k][4], h[t][k][5], h[t][k][0] - t0, h[t][k][1] - t0,
                       h[t][k][2] - t0, h[t][k][3] - t0, h[t][k][5] ? (double)(h[t][k][1] - h[t][k][0]) / h[t][k][5] : 0.0);
        }
        // back-to-back launches as in the timed loop: CTA entry / dependency wait / exit of the LAST launch (%globaltimer, ns)
        for (int i = 0; i < 6; i++) run(i % nsets);
        CK(cudaStreamSynchronize(st));
        long long c[160][4]; CK(cudaMemcpy(c, (char*)dump + 2 * 32 * 8 * 8, sizeof(c), cudaMemcpyDeviceToHost));
        const int G = prop.multiProcessorCount < 160 ? prop.multiProcessorCount : 160;
        long long e0 = c[0][0], w0 = 0, x0 = c[0][2], x1 = 0, e1 = 0;
        for (int i = 0; i < G; i++) { if (c[i][0] < e0) e0 = c[i][0]; if (c[i][0] > e1) e1 = c[i][0]; if (c[i][1] > w0) w0 = c[i][1]; if (c[i][2] < x0) x0 = c[i][2]; if (c[i][2] > x1) x1 = c[i][2]; }
        printf("last of 6 back-to-back launches (ns after the first CTA entry): last entry %lld | dependency wait over %lld | first exit %lld | last exit %lld\n", e1 - e0, w0 - e0, x0 - e0, x1 - e0);
        printf("  CTA 0: %lld cycles in %lld ns between the dependency wait and the exit -> %.0f MHz\n", c[0][3], c[0][2] - c[0][1], 1e3 * c[0][3] / (double)(c[0][2] - c[0][1]));
        printf("  exits by CTA (us after the dependency wait):"); for (int i = 0; i < G; i++) printf("%s%.1f", i % 16 ? " " : "\n    ", (c[i][2] - w0) * 1e-3); printf("\n");
        pf_debug().dump = nullptr;
    }
#endif
    for (int i = 0; i < 10; i++) run(i % nsets);
    CK(cudaStreamSynchronize(st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, tot = 0;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0, st));
        for (int i = 0; i < iters; i++) run(i % nsets);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = fminf(best, ms); tot += ms;
    }
    const double flops = 4.0 * H * (double)n_q * n_kv * D * (causal && n_q == n_kv ? 0.5 : 1.0);
    const double us_best = best * 1e3 / iters, us_avg = tot * 1e3 / iters / 5;
    printf("n_q=%d n_kv=%d H=%d/%d causal=%d: %.2f us best, %.2f us avg -> %.1f TFLOP/s (best) %.1f (avg)\n", n_q, n_kv, H, Hk, causal, us_best, us_avg,
           flops / us_best / 1e6, flops / us_avg / 1e6);
    CK(cudaGetLastError());
    return 0;
}

// ldpattern.cu — isolates ONE question for the decode kernel design: with a fixed, small number of resident warps
// (8 per SM, as a 200-register kernel gets), how does achieved HBM read bandwidth depend on the SHAPE of each
// warp-level load request?  Every variant streams the same bytes (rows of 256 B) and keeps 16 x 16 B or 8 x 32 B
// per lane in flight before consuming them.
//   mode 0: LDG.128, quad covers 64 B of a row  -> 8 rows x 64 B per instruction  (half-used 128 B lines)
//   mode 1: LDG.128, 8 lanes cover 128 B of a row -> 4 rows x 128 B per instruction (full lines)
//   mode 2: LDG.256, quad covers 128 B of a row  -> 8 rows x 128 B per instruction (full lines)
//   mode 3: LDG.256, 8 lanes cover a 256 B row   -> 4 rows x 256 B per instruction
//   mode 4: LDG.128 fully linear (lane i -> 16 B chunk i): 512 contiguous bytes per instruction
// build + run: python profiles/experiments/ldpattern.py  (shared library driven through ctypes; torch owns the buffer)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint4 ld128(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void ld256(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}

// each warp owns a contiguous run of 16-row tiles (4 KB each); `tiles_per_warp` tiles, stride between warps' tiles
template <int MODE>
__global__ void __launch_bounds__(128) stream(const char* __restrict__ base, size_t bytes_per_cta, int tiles_per_warp, unsigned* sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const char* cta = base + (size_t)blockIdx.x * bytes_per_cta;
    unsigned acc = 0;
    for (int it = 0; it < tiles_per_warp; it++) {
        const char* tile = cta + ((size_t)it * 4 + warp) * 8192;  // two 4 KB half-tiles ("K" and "V") per warp iteration
        uint4 r[16];
        if (MODE == 0) {
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int nt = 0; nt < 2; nt++)
#pragma unroll
                    for (int c = 0; c < 4; c++) r[h * 8 + nt * 4 + c] = ld128(tile + h * 4096 + (nt * 8 + g) * 256 + (t + 4 * c) * 16);
        } else if (MODE == 1) {
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int c = 0; c < 2; c++) r[h * 8 + i * 2 + c] = ld128(tile + h * 4096 + (4 * t + i) * 256 + (8 * c + g) * 16);
        } else if (MODE == 2) {
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int nt = 0; nt < 2; nt++)
#pragma unroll
                    for (int c = 0; c < 2; c++) ld256(tile + h * 4096 + (nt * 8 + g) * 256 + (t + 4 * c) * 32, r[h * 8 + nt * 4 + 2 * c], r[h * 8 + nt * 4 + 2 * c + 1]);
        } else if (MODE == 3) {
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int i = 0; i < 4; i++) ld256(tile + h * 4096 + (4 * t + i) * 256 + g * 32, r[h * 8 + i * 2], r[h * 8 + i * 2 + 1]);
        } else if (MODE == 5 || MODE == 6) {
            // the decode kernel's real shape: "K" tile from the first half of the buffer (8 rows x 64 B per instruction),
            // "V" tile from the second half, 1 GiB away (4 rows x 128 B per instruction); MODE 6: CTA regions 512 KB apart
            // at power-of-two bases like (batch, head) slabs
            const size_t half = (size_t)1 << 30;
            const char* kt = (MODE == 6 ? base + (size_t)blockIdx.x * (512 << 10) : cta) + ((size_t)it * 4 + warp) * 4096;
            if (MODE == 5) kt = base + (size_t)blockIdx.x * (bytes_per_cta / 2) + ((size_t)it * 4 + warp) * 4096;
            const char* vt = kt + half;
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int c = 0; c < 4; c++) r[nt * 4 + c] = ld128(kt + (nt * 8 + g) * 256 + (t + 4 * c) * 16);
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int c = 0; c < 2; c++) r[8 + i * 2 + c] = ld128(vt + (4 * t + i) * 256 + (8 * c + g) * 16);
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) r[j] = ld128(tile + j * 512 + lane * 16);
        }
#pragma unroll
        for (int j = 0; j < 16; j++) acc += r[j].x ^ r[j].y ^ r[j].z ^ r[j].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); fflush(stdout); return -1.f; } } while (0)

template <int MODE>
float run(const char* buf, size_t total, int ctas, unsigned* sink) {
    const int tiles = (int)((total / ctas) / (4 * 8192));
    const size_t per_cta = (size_t)tiles * 4 * 8192;  // whole tiles only: keeps every CTA base 32 KB aligned
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) stream<MODE><<<ctas, 128>>>(buf, per_cta, tiles, sink);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 10; i++) stream<MODE><<<ctas, 128>>>(buf, per_cta, tiles, sink);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = 0.f; CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return (float)((double)ctas * tiles * 4 * 8192 * 10 / (ms * 1e-3) / 1e9);
}

extern "C" float ldpattern_run(int mode, const char* buf, size_t total, int ctas, unsigned* sink) {
    switch (mode) {
        case 0: return run<0>(buf, total, ctas, sink);
        case 1: return run<1>(buf, total, ctas, sink);
        case 2: return run<2>(buf, total, ctas, sink);
        case 3: return run<3>(buf, total, ctas, sink);
        case 5: return run<5>(buf, total, ctas, sink);
        case 6: return run<6>(buf, total, ctas, sink);
        default: return run<4>(buf, total, ctas, sink);
    }
}

// ref_gpu.cu — extern "C" launchers for the reference's OWN CUDA kernels, compiled for sm_100a.
// TEST INFRASTRUCTURE ONLY.  The kernels are #included from /root/reference/src at build time
// (never copied); this file reproduces only the launch geometry of the reference call sites:
//   flash_attn_ext_f16<128,16,128>  grid (ceil(ne01/16), ne02, ne03), block (32,2), smem 16*(128+2*(128+16))*2
//                                   flash-matrix.cu:180-206, kernel_test.h:182-198
//   flash_attn_row<128,8,2,256> + fa_reduce<128,8>
//                                   grid (n_kv/256, n_head), block (32,8)   flash-matrix.cu:210-227
// -DINFINITE=INFINITY works around flash_row_float.h:265 (a Windows macro); see SURVEY.md.
// Output: oracle/_ref/libref_gpu.so (git-ignored, travels to the GPU box).  On B200 these run
// as legacy HMMA kernels (no tcgen05/TMA) and serve as the "reference CUDA kernel" parity leg.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <mma.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <algorithm>
#include <cooperative_groups.h>
#include <cooperative_groups/memcpy_async.h>
#include "cuda_info.h"
#include "tensor-mma.h"
#include "flash-llama.h"
#include "flash_row_float.h"

extern "C" {

int ref_gpu_flash_attn_ext_f16(const void* q, const void* k, const void* v, const void* mask, float* dst, float scale,
                               int ne00, int ne01, int ne02, int ne03, int ne10, int ne11, int ne12, int ne13,
                               int ne31, int nb31, int nb01, int nb02, int nb03, int nb11, int nb12, int nb13,
                               int ne0, int ne1, int ne2, int ne3, cudaStream_t stream) {
    if (ne00 != 128) return -1;
    constexpr int nqpb = 16, ncpw = 128, nwarps = 2;
    dim3 grid((ne01 + nqpb - 1) / nqpb, ne02, ne03), block(32, nwarps, 1);
    const size_t shmem = nqpb * (128 + nwarps * (ncpw + nqpb)) * (sizeof(float) / 2);
    flash_attn_ext_f16<128, nqpb, ncpw><<<grid, block, shmem, stream>>>(
        (const char*)q, (const char*)k, (const char*)v, (const char*)mask, dst, scale,
        ne00, ne01, ne02, ne03, ne10, ne11, ne12, ne13, ne31, nb31, nb01, nb02, nb03, nb11, nb12, nb13,
        ne0, ne1, ne2, ne3);
    return (int)cudaGetLastError();
}

// query f32 [n_head][128]; key f16 [n_head_kv][n_kv][128]; valueT f16 [n_head_kv][128][n_kv];
// mask f16 [n_kv]; tmp f16 [(n_blocks*128 + 2*n_blocks) * n_head]; dst f32 [n_head][128].
int ref_gpu_flash_attn_row(const float* query, void* key, const void* valueT, const void* mask, void* tmp, float* dst,
                           int n_kv, float scale, int n_head, int r_kv_heads, cudaStream_t stream) {
    if (n_kv % 256) return -1;
    constexpr int num_warps = 8, kv_per_block = 256;
    dim3 grid(n_kv / kv_per_block, n_head, 1), block(32, num_warps, 1);
    const int shmem = 128 * 2 * sizeof(half) + 2 * sizeof(half) + num_warps * (256 + 2) * sizeof(half);
    flash_attn_row<128, num_warps, 2, kv_per_block><<<grid, block, shmem, stream>>>(
        query, (half*)key, (const half*)valueT, (const half*)mask, (half*)tmp, n_kv, scale, 128 * n_kv, r_kv_heads);
    fa_reduce<128, num_warps><<<n_head, block, shmem + n_kv / kv_per_block * 4, stream>>>(
        (const half*)tmp, dst, n_kv, n_kv / kv_per_block, r_kv_heads);
    return (int)cudaGetLastError();
}

size_t ref_gpu_row_tmp_halves(int n_kv, int n_head) {
    const int nb = n_kv / 256;
    return (size_t)(nb * 128 + nb * 2) * n_head;
}

}  // extern "C"

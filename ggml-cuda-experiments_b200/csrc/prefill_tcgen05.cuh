#pragma once
#include "common.cuh"
namespace b200fa {
inline int launch_prefill_tcgen05(const FaParams&, char*, size_t, size_t, int, cudaStream_t, int*) { return B200FA_ERR_UNSUPPORTED; }
}

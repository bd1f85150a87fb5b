// placeholder until the tcgen05 kernels land
#pragma once
#include "gin_common.cuh"
namespace gin {
inline bool tc_supported(int, int) { return false; }
inline bool tc_wgrad_supported(int, int) { return false; }
inline int launch_gather_gemm_tc(const int32_t*, const GinSide&, int, const float*, const void*, const float*, float*, int, int, int, int, cudaStream_t) { return -4; }
inline int launch_wgrad_tc(const int32_t*, const GinSide&, int, const float*, const float*, float*, int, int, int, int, cudaStream_t) { return -4; }
}

// Shared device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gin_plan.h"

#define GIN_DEVINL __device__ __forceinline__

struct GinSrcView {      // a gathered fp32 tensor with arbitrary element strides
  const float* p;
  long long sb, sp, sc;  // element (b, pixel, channel) -> p[b*sb + pixel*sp + channel*sc]
  int P;                 // pixels per sample
};

__host__ GIN_DEVINL const int32_t* plan_words(const void* plan) { return reinterpret_cast<const int32_t*>(plan); }

// Packed-weight blob layout (gin_hexconv_pack_weights), offsets in bytes from the start:
//   [0]                      float wf[7][Cin][Cout]   forward   B operand (SIMT)
//   [28*Cin*Cout]            float wd[7][Cout][Cin]   dgrad     B operand (SIMT)
//   [56*Cin*Cout]            bf16  bf[7][Cin/64][Cout][64]  forward B tiles (tcgen05), pre-swizzled smem image
//   [70*Cin*Cout]            bf16  bd[7][Cout/64][Cin][64]  dgrad   B tiles (tcgen05), pre-swizzled smem image
GIN_DEVINL __host__ size_t packed_off_wf(int, int) { return 0; }
GIN_DEVINL __host__ size_t packed_off_wd(int Cin, int Cout) { return (size_t)28 * Cin * Cout; }
GIN_DEVINL __host__ size_t packed_off_bf(int Cin, int Cout) { return (size_t)56 * Cin * Cout; }
GIN_DEVINL __host__ size_t packed_off_bd(int Cin, int Cout) { return (size_t)70 * Cin * Cout; }
GIN_DEVINL __host__ size_t packed_total(int Cin, int Cout) { return (size_t)84 * Cin * Cout; }

GIN_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Shared device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "gin_plan.h"

#define GIN_DEVINL __device__ __forceinline__

// Programmatic dependent launch (GIN_PDL=1): a kernel launched through launch_pdl may start while its predecessor in the stream is
// still draining; its CTAs set up (barriers, TMEM, constant tables) and then block in GIN_PDL_SYNC() until the predecessor grid has
// completed and its memory is visible.  EVERY kernel launched through launch_pdl must execute GIN_PDL_SYNC() in all threads before
// its first access to memory another kernel produces or consumes (completion of a grid then implies completion of everything
// before it in the stream).  Without the launch attribute the instructions are no-ops.
#define GIN_PDL_SYNC() asm volatile("griddepcontrol.wait;\n\tgriddepcontrol.launch_dependents;" ::: "memory")

struct GinSrcView {      // a gathered fp32 tensor with arbitrary element strides
  const float* p;
  long long sb, sp, sc;  // element (b, pixel, channel) -> p[b*sb + pixel*sp + channel*sc]
  int P;                 // pixels per sample
};

namespace gin {
// Per-device host state: function attributes (dynamic shared-memory size) are per context, so "configured once" must be
// tracked per device; persistent grids size themselves from the device's SM count (148 on B200; capped there because the
// per-CTA workspaces are sized for 148).
constexpr int kMaxDevices = 64, kMaxSMs = 148;
inline int current_device() { int d = 0; cudaGetDevice(&d); return d & (kMaxDevices - 1); }
inline int num_sms() {
  static int n[kMaxDevices] = {};
  const int d = current_device();
  if (!n[d]) {
    int v = kMaxSMs;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = kMaxSMs;
    n[d] = v < kMaxSMs ? v : kMaxSMs;
  }
  return n[d];
}
struct PerDeviceFlag {
  bool done[kMaxDevices] = {};
  bool& here() { return done[current_device()]; }
};

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GIN_PDL"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
}  // namespace gin

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

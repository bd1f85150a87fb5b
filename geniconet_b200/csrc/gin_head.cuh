// The 1x1 head of the decoder, `Conv2d(64, 3, 1) -> Tanh` (models.py:151-154), as two bandwidth-bound kernels over the
// pixel-major activation the last Up block leaves behind (x [B*P][64] fp32): forward reads 256 B and writes 12 B per pixel,
// backward reads x once more, writes dx, and leaves per-CTA partial sums of dW / db for a second, deterministic pass.
//
// Mapping (both kernels): a half-warp owns one pixel, each lane one float4 of its 64 channels, so every x / dx access is a
// full 256-byte row per half-warp.  A warp walks 32 pixels in 16 such steps; the three outputs of pixel 2*it+h end up in lane
// h*16+it, so the planar y / dy / dz accesses of a 32-pixel block are one 128-byte line per channel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gin_common.cuh"

namespace gin {
namespace head {

constexpr int CIN = 64, COUT = 3, kWarps = 8, kThreads = kWarps * 32;
constexpr int PART = COUT * CIN + 4;                 // floats per CTA partial: dW[3][64], db[3], pad
constexpr int MAX_CTAS = 148 * 4;

__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// lane j of a 32-pixel block <-> pixel block + 2*(j & 15) + (j >> 4)
__device__ __forceinline__ long long lane_pixel(long long blk, int lane) { return blk + 2 * (lane & 15) + (lane >> 4); }

__global__ void __launch_bounds__(kThreads) fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ y, long long rows, long long P) {
  GIN_PDL_SYNC();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l16 = lane & 15, h = lane >> 4;
  float4 wr[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) wr[o] = __ldg(reinterpret_cast<const float4*>(w + o * CIN) + l16);
  const float b0 = __ldg(bias), b1 = __ldg(bias + 1), b2 = __ldg(bias + 2);
  for (long long blk = ((long long)blockIdx.x * kWarps + warp) * 32; blk < rows; blk += (long long)gridDim.x * kWarps * 32) {
    float k0 = 0.f, k1 = 0.f, k2 = 0.f;
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
      const long long r = blk + 2 * it + h;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) v = __ldcs(reinterpret_cast<const float4*>(x + r * CIN) + l16);
      float s0 = dot4(v, wr[0]), s1 = dot4(v, wr[1]), s2 = dot4(v, wr[2]);
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, m); s1 += __shfl_xor_sync(0xffffffffu, s1, m); s2 += __shfl_xor_sync(0xffffffffu, s2, m);
      }
      if (l16 == it) { k0 = s0; k1 = s1; k2 = s2; }
    }
    const long long r = lane_pixel(blk, lane);
    if (r < rows) {
      const long long b = r / P, p = r - b * P;
      float* yp = y + (b * COUT) * P + p;
      yp[0] = tanhf(k0 + b0); yp[P] = tanhf(k1 + b1); yp[2 * P] = tanhf(k2 + b2);
    }
  }
}

// dz = dy * (1 - y^2);  dx[r][c] = sum_o dz[o] * w[o][c];  partial[cta] = sum over the CTA's pixels of dz[o] * x[r][c] and of dz[o]
__global__ void __launch_bounds__(kThreads) bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ y,
                                                       const float* __restrict__ dy, float* __restrict__ dx, float* __restrict__ partial,
                                                       long long rows, long long P) {
  GIN_PDL_SYNC();
  __shared__ float red[kWarps][PART];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l16 = lane & 15, h = lane >> 4;
  float4 wr[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) wr[o] = __ldg(reinterpret_cast<const float4*>(w + o * CIN) + l16);
  float4 acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
  float db0 = 0.f, db1 = 0.f, db2 = 0.f;
  for (long long blk = ((long long)blockIdx.x * kWarps + warp) * 32; blk < rows; blk += (long long)gridDim.x * kWarps * 32) {
    float z0 = 0.f, z1 = 0.f, z2 = 0.f;
    {
      const long long r = lane_pixel(blk, lane);
      if (r < rows) {
        const long long b = r / P, p = r - b * P, o0 = (b * COUT) * P + p;
        const float y0 = __ldg(y + o0), y1 = __ldg(y + o0 + P), y2 = __ldg(y + o0 + 2 * P);
        z0 = __ldg(dy + o0) * (1.f - y0 * y0); z1 = __ldg(dy + o0 + P) * (1.f - y1 * y1); z2 = __ldg(dy + o0 + 2 * P) * (1.f - y2 * y2);
      }
      db0 += z0; db1 += z1; db2 += z2;
    }
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
      const long long r = blk + 2 * it + h;
      const int src = (lane & 16) + it;
      const float a0 = __shfl_sync(0xffffffffu, z0, src), a1 = __shfl_sync(0xffffffffu, z1, src), a2 = __shfl_sync(0xffffffffu, z2, src);
      if (r < rows) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(x + r * CIN) + l16);
        float4 g;
        g.x = a0 * wr[0].x + a1 * wr[1].x + a2 * wr[2].x; g.y = a0 * wr[0].y + a1 * wr[1].y + a2 * wr[2].y;
        g.z = a0 * wr[0].z + a1 * wr[1].z + a2 * wr[2].z; g.w = a0 * wr[0].w + a1 * wr[1].w + a2 * wr[2].w;
        *(reinterpret_cast<float4*>(dx + r * CIN) + l16) = g;
        acc[0].x += a0 * v.x; acc[0].y += a0 * v.y; acc[0].z += a0 * v.z; acc[0].w += a0 * v.w;
        acc[1].x += a1 * v.x; acc[1].y += a1 * v.y; acc[1].z += a1 * v.z; acc[1].w += a1 * v.w;
        acc[2].x += a2 * v.x; acc[2].y += a2 * v.y; acc[2].z += a2 * v.z; acc[2].w += a2 * v.w;
      }
    }
  }
  // warp: fold the two half-warps (same channels, different pixels), then the eight warps through shared memory
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    acc[o].x += __shfl_xor_sync(0xffffffffu, acc[o].x, 16); acc[o].y += __shfl_xor_sync(0xffffffffu, acc[o].y, 16);
    acc[o].z += __shfl_xor_sync(0xffffffffu, acc[o].z, 16); acc[o].w += __shfl_xor_sync(0xffffffffu, acc[o].w, 16);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    db0 += __shfl_xor_sync(0xffffffffu, db0, m); db1 += __shfl_xor_sync(0xffffffffu, db1, m); db2 += __shfl_xor_sync(0xffffffffu, db2, m);
  }
  if (h == 0) {
#pragma unroll
    for (int o = 0; o < COUT; ++o) *reinterpret_cast<float4*>(&red[warp][o * CIN + 4 * l16]) = acc[o];
  }
  if (lane == 0) { red[warp][COUT * CIN] = db0; red[warp][COUT * CIN + 1] = db1; red[warp][COUT * CIN + 2] = db2; red[warp][COUT * CIN + 3] = 0.f; }
  __syncthreads();
  for (int i = threadIdx.x; i < PART; i += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kWarps; ++k) s += red[k][i];
    partial[(size_t)blockIdx.x * PART + i] = s;
  }
}

// one WARP per output: lane l adds partials l, l+32, ... in order, then a fixed shuffle tree -> deterministic, and 32x the
// memory parallelism of a serial loop (the serial version took 25 us for 592 partials)
__global__ void __launch_bounds__(256) bwd_final_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ dw, float* __restrict__ db) {
  GIN_PDL_SYNC();
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= COUT * CIN + COUT) return;
  float s = 0.f;
  for (int k = lane; k < nparts; k += 32) s += partial[(size_t)k * PART + i];
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (lane == 0) { if (i < COUT * CIN) dw[i] = s; else db[i - COUT * CIN] = s; }
}

inline int grid_for_rows(long long rows) {
  long long g = (rows + kThreads - 1) / kThreads;          // one 32-pixel block per warp per pass
  return (int)(g < 1 ? 1 : (g < MAX_CTAS ? g : MAX_CTAS));
}

}  // namespace head
}  // namespace gin

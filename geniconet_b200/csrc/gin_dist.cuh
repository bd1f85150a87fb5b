// Point-to-mesh distance for the evaluation metric (ico_utils.py:26-44, mode 'point2mesh'): the reference calls
// kaolin 0.9.1 `kaolin.metrics.trianglemesh.point_to_mesh_distance(points[1,N,3], vertices[1,V,3], faces[F,3])` (pinned in
// Dockerfile:50-52; the package is absent here) and averages its first return value.  That value is, per point, the SQUARED
// Euclidean distance to the closest point of the closest triangle; it is restated here with the closest-point-on-triangle
// region test (vertex / edge / interior Voronoi regions).
//
// Brute force: N x F point-triangle tests (2.1e8 per level-5 mesh pair), compute-bound fp32.  One thread owns one point;
// the grid is (point blocks) x (face chunks) x (batch); a chunk's triangles are gathered ONCE per CTA into shared memory
// as corner + two edge vectors + their three dot products, every thread then walks the chunk out of broadcast smem reads.
// Chunks combine with one 64-bit atomicMin per (point, chunk) on (distance bits << 32 | face index): non-negative floats
// order like their bit patterns, so the minimum is the smallest distance and, among equal distances, the lowest face index
// -- deterministic whatever the chunk order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gin {
namespace dist {

constexpr int kThreads = 128;
constexpr int kChunk = 512;           // faces per CTA: 512 * 12 floats = 24 KB of shared memory

struct Tri { float ax, ay, az, abx, aby, abz, acx, acy, acz, ab2, ac2, abac; };

__device__ __forceinline__ float point_tri_sq(const Tri& t, float px, float py, float pz) {
  const float apx = px - t.ax, apy = py - t.ay, apz = pz - t.az;
  const float d1 = t.abx * apx + t.aby * apy + t.abz * apz;          // ab . ap
  const float d2 = t.acx * apx + t.acy * apy + t.acz * apz;          // ac . ap
  // with bp = ap - ab, cp = ap - ac the remaining dot products follow from d1, d2 and the per-triangle constants
  const float d3 = d1 - t.ab2, d4 = d2 - t.abac;                     // ab . bp, ac . bp
  const float d5 = d1 - t.abac, d6 = d2 - t.ac2;                     // ab . cp, ac . cp
  float v, w;                                                        // closest point = a + v*ab + w*ac
  const float vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
  if (d1 <= 0.f && d2 <= 0.f) { v = 0.f; w = 0.f; }                                  // vertex a
  else if (d3 >= 0.f && d4 <= d3) { v = 1.f; w = 0.f; }                              // vertex b
  else if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { v = d1 / (d1 - d3); w = 0.f; }     // edge ab
  else if (d6 >= 0.f && d5 <= d6) { v = 0.f; w = 1.f; }                              // vertex c
  else if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { v = 0.f; w = d2 / (d2 - d6); }     // edge ac
  else if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {                      // edge bc
    w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); v = 1.f - w;
  } else {                                                                           // interior
    const float den = 1.f / (va + vb + vc);
    v = vb * den; w = vc * den;
  }
  // p - closest point, component by component (the expanded quadratic form would cancel badly for points near the surface)
  const float rx = apx - v * t.abx - w * t.acx, ry = apy - v * t.aby - w * t.acy, rz = apz - v * t.abz - w * t.acz;
  return rx * rx + ry * ry + rz * rz;
}

// best[b][i] must be preset to all ones (cudaMemsetAsync 0xFF).
__global__ void __launch_bounds__(kThreads) point_mesh_kernel(const float* __restrict__ pts, const float* __restrict__ verts,
                                                               const int32_t* __restrict__ faces, unsigned long long* __restrict__ best,
                                                               int N, int V, int F) {
  __shared__ Tri tri[kChunk];
  const int b = blockIdx.z, f0 = blockIdx.y * kChunk, nf = min(kChunk, F - f0);
  const float* vb = verts + (size_t)b * V * 3;
  for (int j = threadIdx.x; j < nf; j += kThreads) {
    const int32_t* fc = faces + (size_t)(f0 + j) * 3;
    const float* A = vb + (size_t)fc[0] * 3; const float* B = vb + (size_t)fc[1] * 3; const float* C = vb + (size_t)fc[2] * 3;
    Tri t;
    t.ax = A[0]; t.ay = A[1]; t.az = A[2];
    t.abx = B[0] - t.ax; t.aby = B[1] - t.ay; t.abz = B[2] - t.az;
    t.acx = C[0] - t.ax; t.acy = C[1] - t.ay; t.acz = C[2] - t.az;
    t.ab2 = t.abx * t.abx + t.aby * t.aby + t.abz * t.abz;
    t.ac2 = t.acx * t.acx + t.acy * t.acy + t.acz * t.acz;
    t.abac = t.abx * t.acx + t.aby * t.acy + t.abz * t.acz;
    tri[j] = t;
  }
  __syncthreads();
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= N) return;
  const float* p = pts + ((size_t)b * N + i) * 3;
  const float px = p[0], py = p[1], pz = p[2];
  float bd = 3.4e38f; int bj = 0;
  for (int j = 0; j < nf; ++j) {
    const float d = point_tri_sq(tri[j], px, py, pz);
    if (d < bd) { bd = d; bj = j; }                 // strict: the lowest face index wins a tie inside the chunk
  }
  const unsigned long long key = ((unsigned long long)__float_as_uint(bd) << 32) | (unsigned)(f0 + bj);
  atomicMin(best + (size_t)b * N + i, key);
}

__global__ void unpack_kernel(const unsigned long long* __restrict__ best, float* __restrict__ d, int32_t* __restrict__ face, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = best[i];
  d[i] = __uint_as_float((unsigned)(k >> 32));
  if (face) face[i] = (int32_t)(k & 0xffffffffu);
}

}  // namespace dist
}  // namespace gin

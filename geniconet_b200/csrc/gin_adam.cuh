// Adam over a LIST of fp32 tensors in one launch -- the optimizer step of the reference's training loop
// (run.py:446 `torch.optim.Adam(model.parameters(), lr=...)`, run.py:250 `optimizer.step()`).
//
// torch's fused Adam walks 78 parameter tensors (ico2ico: 4.63 M floats) with its generic multi-tensor-apply machinery in four
// launches and reaches 1.2 TB/s inside the replayed step (profiles/r02_trace_n1.log: 111 us for 130 MB).  Here the list is a
// device table {param, grad, exp_avg, exp_avg_sq, step, n}; a CTA owns one 4096-element chunk of one tensor (binary search in
// the chunk prefix), moves 16-byte groups when the four pointers allow it, and the step counters are advanced by the last
// CTA to finish (a ticket), so no CTA ever sees a counter another one has already bumped.
//
// Arithmetic (torch/optim/adam.py `_single_tensor_adam`, fused kernel `adam_math`):
//   g' = g + weight_decay * p;  m = m + (1 - beta1) * (g' - m);  v = beta2 * v + (1 - beta2) * g'^2
//   p  = p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps),   t = step + 1
#pragma once
#include "gin_common.cuh"

namespace gin {
namespace adam {

struct Tensor {            // one row of the device table (48 bytes; mirrored by geniconet_b200/optim.py)
  float* p;
  const float* g;
  float* m;
  float* v;
  float* step;             // fp32 scalar on the device, as torch keeps it for capturable optimizers
  long long n;
};
constexpr int CHUNK = 4096, THREADS = 256;

struct Hyper { float lr, beta1, beta2, eps, weight_decay; };

GIN_DEVINL void update(float& p, float g, float& m, float& v, const Hyper& h, float step_size, float bc2_sqrt) {
  if (h.weight_decay != 0.f) g = fmaf(h.weight_decay, p, g);
  m = fmaf(1.f - h.beta1, g - m, m);
  v = fmaf(h.beta2, v, (1.f - h.beta2) * g * g);
  const float denom = sqrtf(v) / bc2_sqrt + h.eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(THREADS)
step_kernel(const Tensor* __restrict__ tab, const int32_t* __restrict__ chunk_first, int count, Hyper h, const float* __restrict__ lr_dev,
            unsigned int* __restrict__ ticket) {
  __shared__ Tensor T;
  __shared__ float s_step_size, s_bc2_sqrt;
  __shared__ long long s_e0;
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    int lo = 0, hi = count - 1;                           // the tensor whose chunk range holds blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (chunk_first[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    T = tab[lo];
    const double t = (double)(*T.step) + 1.0;
    const double bc1 = 1.0 - pow((double)h.beta1, t), bc2 = 1.0 - pow((double)h.beta2, t);
    const float lr = lr_dev ? *lr_dev : h.lr;
    s_step_size = (float)((double)lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
    s_e0 = (long long)((int)blockIdx.x - chunk_first[lo]) * CHUNK;
  }
  __syncthreads();
  const long long e0 = s_e0, e1 = (T.n - e0 < CHUNK) ? T.n : e0 + CHUNK;
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  float* p = T.p; const float* g = T.g; float* m = T.m; float* v = T.v;
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0;      // e0 is a multiple of 4
  long long done = e0;
  if (vec) {
    const long long nv = (e1 - e0) >> 2;                  // whole 4-element groups of this chunk
    for (long long q = threadIdx.x; q < nv; q += THREADS) {
      const long long i = e0 + 4 * q;
      float4 P = *reinterpret_cast<const float4*>(p + i), M = *reinterpret_cast<const float4*>(m + i), V = *reinterpret_cast<const float4*>(v + i);
      const float4 G = __ldg(reinterpret_cast<const float4*>(g + i));
      update(P.x, G.x, M.x, V.x, h, step_size, bc2_sqrt);
      update(P.y, G.y, M.y, V.y, h, step_size, bc2_sqrt);
      update(P.z, G.z, M.z, V.z, h, step_size, bc2_sqrt);
      update(P.w, G.w, M.w, V.w, h, step_size, bc2_sqrt);
      *reinterpret_cast<float4*>(p + i) = P;
      *reinterpret_cast<float4*>(m + i) = M;
      *reinterpret_cast<float4*>(v + i) = V;
    }
    done = e0 + 4 * nv;
  }
  for (long long i = done + threadIdx.x; i < e1; i += THREADS) {
    float P = p[i], M = m[i], V = v[i];
    update(P, g[i], M, V, h, step_size, bc2_sqrt);
    p[i] = P; m[i] = M; v[i] = V;
  }
  // the last CTA to get here advances every step counter: all CTAs have read theirs by then
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    for (int i = threadIdx.x; i < count; i += THREADS) {
      // several tensors may share one counter (not in torch's layout, but allowed): bump each distinct address once
      float* sp = tab[i].step;
      bool first = true;
      for (int j = 0; j < i; ++j) if (tab[j].step == sp) { first = false; break; }
      if (first) *sp += 1.f;
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

}  // namespace adam
}  // namespace gin

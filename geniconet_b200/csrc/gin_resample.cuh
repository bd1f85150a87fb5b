// Memory-bound kernels: IcoUpsampleS2S forward/backward (row a4), VAE reparameterisation (a6),
// KL divergence (a9).  All are pure streaming work: one thread per 16-byte channel vector,
// coalesced along the channel axis, grid-stride.
#pragma once
#include <cuda_bf16.h>

#include "gin_common.cuh"

namespace gin {

GIN_DEVINL float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
GIN_DEVINL float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
GIN_DEVINL float4 fma4(float w, float4 a, float4 acc) {
  return make_float4(fmaf(w, a.x, acc.x), fmaf(w, a.y, acc.y), fmaf(w, a.z, acc.z), fmaf(w, a.w, acc.w));
}

GIN_DEVINL float4 up_fetch(const float* __restrict__ xb, const int32_t* __restrict__ ring, int code, int C, int c) {
  if (code >= 0) return ld4(xb + (size_t)code * C + c);
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (code == GIN_SRC_ZERO) return z;
  const int pole = (-2 - code) & 1;
#pragma unroll
  for (int j = 0; j < 5; ++j) z = fma4(0.2f, ld4(xb + (size_t)ring[pole * 5 + j] * C + c), z);
  return z;
}

// y[b, f, :] = 0.5 * (x[b, src0(f), :] + x[b, src1(f), :]);  C % 4 == 0
__global__ void __launch_bounds__(256)
upsample_fwd_kernel(const int32_t* __restrict__ plan, const float* __restrict__ x, float* __restrict__ y, int B, int C) {
  const GinUpPlanHdr* h = reinterpret_cast<const GinUpPlanHdr*>(plan);
  const int Pc = h->Pc, Pf = h->Pf, C4 = C >> 2;
  const int32_t* src = plan + h->fwd_off;
  const int32_t* ring = plan + h->ring_off;
  const long long total = (long long)B * Pf * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const long long bp = i / C4;
    const int f = (int)(bp % Pf);
    const long long b = bp / Pf;
    const float* xb = x + (size_t)b * Pc * C;
    const int s0 = src[2 * f], s1 = src[2 * f + 1];
    float4 a = up_fetch(xb, ring, s0, C, c);
    float4 r;
    if (s0 == s1) r = a;
    else { float4 d = up_fetch(xb, ring, s1, C, c); r = make_float4(0.5f * (a.x + d.x), 0.5f * (a.y + d.y), 0.5f * (a.z + d.z), 0.5f * (a.w + d.w)); }
    *reinterpret_cast<float4*>(y + (size_t)bp * C + c) = r;
  }
}

// dx[b, c, :] = sum_e w[c][e] * dy[b, idx[c][e], :].  Idx: unsigned when B*Pc*C/4 < 2^31 (32-bit index arithmetic), else long long.
template <typename Idx>
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const int32_t* __restrict__ plan, const float* __restrict__ dy, float* __restrict__ dx, int B, int C) {
  const GinUpPlanHdr* h = reinterpret_cast<const GinUpPlanHdr*>(plan);
  const int Pf = h->Pf, deg = h->bwd_deg;
  const Idx Pc = (Idx)h->Pc, C4 = (Idx)(C >> 2);
  const int32_t* idx = plan + h->bwd_idx_off;
  const float* w = reinterpret_cast<const float*>(plan + h->bwd_w_off);
  const Idx total = (Idx)B * Pc * C4;
  for (Idx i = (Idx)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (Idx)gridDim.x * blockDim.x) {
    const Idx bp = i / C4;
    const int c = (int)(i - bp * C4) * 4;
    const Idx b = bp / Pc;
    const int p = (int)(bp - b * Pc);
    const float* dyb = dy + (size_t)b * Pf * C;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 0; e < deg; ++e) {
      const int f = idx[(size_t)p * deg + e];
      if (f < 0) break;
      acc = fma4(w[(size_t)p * deg + e], ld4(dyb + (size_t)f * C + c), acc);
    }
    *reinterpret_cast<float4*>(dx + (size_t)bp * C + c) = acc;
  }
}

// fp32 [B*P][C] -> 16-bit operand copy [B*P + 2B][C] (f16 != 0: fp16, else bf16; gin_common.cuh).  The 2B extra rows are the per-sample pole means (mean of the pole's five ring
// pixels), so that the tcgen05 producers can fetch a pole cell like any other row.  One thread = 8 channels.
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb, const int32_t* __restrict__ ring, int B, int P, int C, int f16) {
  const int C8 = C >> 3;
  const long long n_main = (long long)B * P * C8, n_all = n_main + 2LL * B * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_all; i += (long long)gridDim.x * blockDim.x) {
    float4 a, b;
    if (i < n_main) {
      a = ld4(x + i * 8);
      b = ld4(x + i * 8 + 4);
    } else {
      const long long j = i - n_main;
      const int c = (int)(j % C8) * 8;
      const int sp = (int)(j / C8), sample = sp >> 1, pole = sp & 1;
      a = make_float4(0.f, 0.f, 0.f, 0.f); b = a;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float* src = x + ((size_t)sample * P + ring[pole * 5 + k]) * C + c;
        a = fma4(0.2f, ld4(src), a);
        b = fma4(0.2f, ld4(src + 4), b);
      }
    }
    uint4 o;
    o.x = pack2_op(a.x, a.y, f16); o.y = pack2_op(a.z, a.w, f16); o.z = pack2_op(b.x, b.y, f16); o.w = pack2_op(b.z, b.w, f16);
    *reinterpret_cast<uint4*>(xb + i * 8) = o;
  }
}

// Same cast, plus the per-channel column sums of the fp32 input (the conv BIAS gradient db[c] = sum over pixels of
// dy[., c], losses.backward -> IcoConvS2S.bias.grad): the tensor is being read anyway, so db costs no extra pass.
// Needs (C/8) | 256 so that a thread keeps the same 8 channels for all of its grid-stride iterations.  Every CTA writes
// its partial sums to ws[block][C]; colsum_final_kernel adds them in a fixed order, so db is deterministic.  (A single
// "last CTA" doing that sum serially cost 40 us of pure load latency; the second kernel does it in ~3.)
constexpr int CAST_COLSUM_MAX_CTAS = 148 * 2;

__global__ void __launch_bounds__(256)
cast_bf16_colsum_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb, const int32_t* __restrict__ ring, int B, int P, int C,
                        float* __restrict__ ws, int f16) {
  __shared__ float part[256][9];
  const int C8 = C >> 3;
  const long long n_main = (long long)B * P * C8, n_all = n_main + 2LL * B * C8;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_all; i += (long long)gridDim.x * blockDim.x) {
    float4 a, b;
    if (i < n_main) {
      a = ld4(x + i * 8);
      b = ld4(x + i * 8 + 4);
      s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w; s[4] += b.x; s[5] += b.y; s[6] += b.z; s[7] += b.w;
    } else {
      const long long j = i - n_main;
      const int c = (int)(j % C8) * 8;
      const int sp = (int)(j / C8), sample = sp >> 1, pole = sp & 1;
      a = make_float4(0.f, 0.f, 0.f, 0.f); b = a;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float* src = x + ((size_t)sample * P + ring[pole * 5 + k]) * C + c;
        a = fma4(0.2f, ld4(src), a);
        b = fma4(0.2f, ld4(src + 4), b);
      }
    }
    uint4 o;
    o.x = pack2_op(a.x, a.y, f16); o.y = pack2_op(a.z, a.w, f16); o.z = pack2_op(b.x, b.y, f16); o.w = pack2_op(b.z, b.w, f16);
    *reinterpret_cast<uint4*>(xb + i * 8) = o;
  }
  // channel group of this thread: (threadIdx.x % C8) because C8 | 256 | grid stride
#pragma unroll
  for (int k = 0; k < 8; ++k) part[threadIdx.x][k] = s[k];
  __syncthreads();
  float* mine = ws + (size_t)blockIdx.x * C;
  for (int c = threadIdx.x; c < C; c += 256) {
    const int g = c >> 3, k = c & 7;
    float acc = 0.f;
    for (int t = g; t < 256; t += C8) acc += part[t][k];
    mine[c] = acc;
  }
}

// colsum[c] = sum_b ws[b][c]: one CTA per 8 columns, 32 row groups of (nblocks / 32) partials each, fixed summation order
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ ws, float* __restrict__ colsum, int C, int nblocks) {
  __shared__ float red[32][9];
  const int c = blockIdx.x * 8 + (threadIdx.x & 7), rg = threadIdx.x >> 3;
  float acc = 0.f;
  if (c < C)
    for (int b = rg; b < nblocks; b += 32) acc += __ldg(ws + (size_t)b * C + c);
  red[rg][threadIdx.x & 7] = acc;
  __syncthreads();
  if (threadIdx.x < 8 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) t += red[r][threadIdx.x];
    colsum[c] = t;
  }
}

// ---------------------------------------------------------------- Philox4x32-10 + Box-Muller
GIN_DEVINL void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

GIN_DEVINL void philox4x32_10(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4]) {
  uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

GIN_DEVINL void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u = ((float)a + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
  const float v = (float)b * 2.3283064365386963e-10f;           // [0,1)
  const float r = sqrtf(-2.0f * logf(u));
  float s, c;
  sincosf(6.283185307179586f * v, &s, &c);
  n0 = r * c; n1 = r * s;
}

__global__ void counter_inc_kernel(unsigned long long* c) { *c += 1ull; }

// z = eps * exp(0.5*logvar) + mu; one Philox call per 4 elements, counter = (element/4, offset)
__global__ void __launch_bounds__(256)
reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, float* __restrict__ eps,
                   float* __restrict__ z, long long n, uint64_t seed, uint64_t offset, const unsigned long long* __restrict__ step) {
  if (step) offset += *step;          // device-side step counter: a replayed CUDA graph still draws fresh noise every step
  const long long n4 = (n + 3) >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10(seed, (uint64_t)q, offset, r);
    float e[4];
    box_muller(r[0], r[1], e[0], e[1]);
    box_muller(r[2], r[3], e[2], e[3]);
    const long long i0 = q << 2;
    if (i0 + 3 < n) {
      float4 m = ld4(mu + i0), lv = ld4(logvar + i0);
      *reinterpret_cast<float4*>(eps + i0) = make_float4(e[0], e[1], e[2], e[3]);
      *reinterpret_cast<float4*>(z + i0) = make_float4(fmaf(e[0], expf(0.5f * lv.x), m.x), fmaf(e[1], expf(0.5f * lv.y), m.y),
                                                       fmaf(e[2], expf(0.5f * lv.z), m.z), fmaf(e[3], expf(0.5f * lv.w), m.w));
    } else {
      for (int j = 0; j < 4 && i0 + j < n; ++j) {
        eps[i0 + j] = e[j];
        z[i0 + j] = fmaf(e[j], expf(0.5f * logvar[i0 + j]), mu[i0 + j]);
      }
    }
  }
}

// dmu = dz ; dlogvar = dz * eps * 0.5 * exp(0.5*logvar)
__global__ void __launch_bounds__(256)
reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ logvar, const float* __restrict__ eps,
                   float* __restrict__ dmu, float* __restrict__ dlogvar, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = dz[i];
    dmu[i] = g;
    dlogvar[i] = g * eps[i] * 0.5f * expf(0.5f * logvar[i]);
  }
}

// KLD (losses.py:105): mean_b(-0.5 * mean_i(1 + lv - mu^2 - exp(lv))) == -0.5/n * sum_all(...)
__global__ void __launch_bounds__(256)
kld_partial_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, double* __restrict__ partial, long long n) {
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float m = mu[i], lv = logvar[i];
    s += (double)(1.0f + lv - m * m - expf(lv));
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void kld_final_kernel(const double* __restrict__ partial, int nparts, float* __restrict__ out, double scale) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(sh[0] * scale);
}

// d/dmu = dout * scale_bwd * (-2 mu) ; d/dlogvar = dout * scale_bwd * (1 - exp(lv)),  scale_bwd = -0.5/n * user scale
__global__ void __launch_bounds__(256)
kld_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, const float* __restrict__ dout, float scale,
               float* __restrict__ dmu, float* __restrict__ dlogvar, long long n) {
  const float g = dout[0] * scale;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    dmu[i] = g * (-2.0f * mu[i]);
    dlogvar[i] = g * (1.0f - expf(logvar[i]));
  }
}

}  // namespace gin

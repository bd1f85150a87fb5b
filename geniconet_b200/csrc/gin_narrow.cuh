// Warp-level fp32 kernels for the NARROW hex-conv layer (Cin <= 4: the xyz input layer, models.py:104).
//
// With K = 7 taps x 3 channels = 21 the layer is no tensor-core shape and is bound by writing (forward) or reading
// (wgrad) the 64-channel side once:  B * P * Cout * 4 bytes.  The chart padding is fused exactly as in the wide
// kernels: the plan's gather table names, per output pixel and tap, the source pixel / pole mean / zero.
//
//   forward   y[row, :]   = bias + sum_slot sum_c  x~[src[slot][row], c] * W[tap(slot)][c][:]
//   wgrad     dWp[t][c][:] += sum_row  x~[src_t[row], c] * dy[row, :] ;   db[:] += sum_row dy[row, :]
//
// One CTA = one 128-row plan tile at a time (persistent, grid-stride).  The 21 gathered input values of every row are
// staged in shared memory ([slot][c][row], read back as warp broadcasts); a lane owns Cout/32 output channels and a
// register block of rows, so weights (forward) / dy (wgrad) are fetched once per block of rows.
#pragma once
#include "gin_common.cuh"
#include "gin_gemm_simt.cuh"

namespace gin {
namespace narrow {

constexpr int TM = 128, THREADS = 256, WARPS = 8, ROWS_PER_WARP = TM / WARPS;   // 16
constexpr int MAX_CIN = 4;

GIN_DEVINL float gather_value(const RowSrc& rs, const GinSrcView& X, const int32_t* __restrict__ ring, int c) {
  if (rs.kind == 1) return __ldg(X.p + rs.off + (long long)c * X.sc);
  if (rs.kind == 2) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) s += __ldg(X.p + (long long)rs.sample * X.sb + (long long)ring[rs.pole * 5 + j] * X.sp + (long long)c * X.sc);
    return 0.2f * s;
  }
  return 0.f;
}

// stage xs[slot][c][row] for one tile; returns nothing, caller syncs
template <int CIN>
GIN_DEVINL void stage_tile(const int32_t* __restrict__ plan, const GinSide& side, const GinTileDesc* desc, const GinSrcView& X, int G, int group,
                           int B, float* xs, long long* dst_s, int t) {
  const int nslots = desc->nslots;
  const long long base_src = (long long)G * group * side.P_src, total_src = (long long)B * side.P_src;
  const int32_t* src_tab = plan + side.src_off + desc->src_off;
  const int32_t* ring = plan + side.ring_off;
  for (int i = threadIdx.x; i < nslots * TM; i += THREADS) {
    const int slot = i >> 7, r = i & (TM - 1);
    const RowSrc rs = resolve_src(__ldg(src_tab + i), base_src, total_src, G * group, B, X);
#pragma unroll
    for (int c = 0; c < CIN; ++c) xs[(slot * CIN + c) * TM + r] = gather_value(rs, X, ring, c);
  }
  if (threadIdx.x < TM) {
    const long long base_dst = (long long)G * group * side.P_dst, total_dst = (long long)B * side.P_dst;
    const int r = __ldg(plan + side.rows_off + t * TM + threadIdx.x);
    const long long d = (r >= 0) ? base_dst + r : -1;
    dst_s[threadIdx.x] = (d >= 0 && d < total_dst) ? d : -1;
  }
}

// ------------------------------------------------------------------------------------------------ forward
// CPL = output channels per lane (Cout = 32 * CPL).  W is wf[7][CIN][Cout] fp32.
// stats != null: this CTA's column sums of y and y^2 (BatchNorm statistics of the layer's output) go to stats[blockIdx][2][COUT];
// y_f16: Y is written as fp16 (the fused chain's stem output is only read by BatchNorm kernels).
template <int CIN, int CPL>
__global__ void __launch_bounds__(THREADS, 4)
fwd_kernel(const int32_t* __restrict__ plan, GinSide side, GinSrcView X, const float* __restrict__ W, const float* __restrict__ bias,
           float* __restrict__ Y, int group, int B, int total_tiles, float* __restrict__ stats, int y_f16) {
  extern __shared__ __align__(16) float smem_f[];
  constexpr int COUT = 32 * CPL;
  float* ws = smem_f;                                  // [7][CIN][COUT]
  float* xs = ws + 7 * CIN * COUT;                     // [max_slots][CIN][TM]
  long long* dst_s = reinterpret_cast<long long*>(xs + side.max_slots * CIN * TM);
  __shared__ int8_t tap_s[GIN_MAX_SLOTS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 7 * CIN * COUT; i += THREADS) ws[i] = __ldg(W + i);
  float bz[CPL], ssum[CPL], ssq[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) { bz[k] = bias ? __ldg(bias + lane * CPL + k) : 0.f; ssum[k] = 0.f; ssq[k] = 0.f; }

  for (int T = blockIdx.x; T < total_tiles; T += gridDim.x) {
    const int G = T / side.ntiles, t = T % side.ntiles;
    const GinTileDesc* desc = reinterpret_cast<const GinTileDesc*>(plan + side.tiles_off) + t;
    const int nslots = desc->nslots;
    __syncthreads();                                   // previous tile fully consumed
    if (threadIdx.x < GIN_MAX_SLOTS) tap_s[threadIdx.x] = desc->tap[threadIdx.x];
    stage_tile<CIN>(plan, side, desc, X, G, group, B, xs, dst_s, t);
    __syncthreads();
#pragma unroll
    for (int half = 0; half < 2; ++half) {             // 8 rows at a time
      const int r0 = warp * ROWS_PER_WARP + half * 8;
      float acc[8][CPL];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int k = 0; k < CPL; ++k) acc[r][k] = bz[k];
      for (int slot = 0; slot < nslots; ++slot) {
        const float* wrow = ws + tap_s[slot] * CIN * COUT + lane * CPL;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          float w[CPL];
          if (CPL == 2) { const float2 v = *reinterpret_cast<const float2*>(wrow + c * COUT); w[0] = v.x; w[CPL - 1] = v.y; }
          else if (CPL == 4) { const float4 v = *reinterpret_cast<const float4*>(wrow + c * COUT); w[0] = v.x; w[1 % CPL] = v.y; w[2 % CPL] = v.z; w[3 % CPL] = v.w; }
          else {
#pragma unroll
            for (int k = 0; k < CPL; ++k) w[k] = wrow[c * COUT + k];
          }
          const float4 xa = *reinterpret_cast<const float4*>(xs + (slot * CIN + c) * TM + r0);
          const float4 xb = *reinterpret_cast<const float4*>(xs + (slot * CIN + c) * TM + r0 + 4);
          const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < CPL; ++k) acc[r][k] = fmaf(xv[r], w[k], acc[r][k]);
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const long long d = dst_s[r0 + r];
        if (d >= 0) {
#pragma unroll
          for (int k = 0; k < CPL; ++k) { ssum[k] += acc[r][k]; ssq[k] = fmaf(acc[r][k], acc[r][k], ssq[k]); }
          if (y_f16) {
            unsigned short* yh = reinterpret_cast<unsigned short*>(Y) + (size_t)d * COUT + lane * CPL;
            if (CPL == 2) *reinterpret_cast<uint32_t*>(yh) = pack2_f16(acc[r][0], acc[r][1]);
            else if (CPL == 4) *reinterpret_cast<uint2*>(yh) = make_uint2(pack2_f16(acc[r][0], acc[r][1 % CPL]), pack2_f16(acc[r][2 % CPL], acc[r][3 % CPL]));
            else {
#pragma unroll
              for (int k = 0; k < CPL; ++k) yh[k] = cvt_op(acc[r][k], 1);
            }
          } else {
            float* yp = Y + (size_t)d * COUT + lane * CPL;
            if (CPL == 2) *reinterpret_cast<float2*>(yp) = make_float2(acc[r][0], acc[r][1]);
            else if (CPL == 4) *reinterpret_cast<float4*>(yp) = make_float4(acc[r][0], acc[r][1 % CPL], acc[r][2 % CPL], acc[r][3 % CPL]);
            else {
#pragma unroll
              for (int k = 0; k < CPL; ++k) yp[k] = acc[r][k];
            }
          }
        }
      }
    }
  }
  if (stats) {
    // lane l of every warp owns the same CPL channels: the eight warps meet in shared memory and are added in a fixed order
    __syncthreads();
    float* red = xs;                                   // [WARPS][2][COUT], the staging area is free now
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      red[(warp * 2) * COUT + lane * CPL + k] = ssum[k];
      red[(warp * 2 + 1) * COUT + lane * CPL + k] = ssq[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * COUT; i += THREADS) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) a += red[w * 2 * COUT + i];
      stats[(size_t)blockIdx.x * 2 * COUT + i] = a;
    }
  }
}

// ------------------------------------------------------------------------------------------------ wgrad (+ bias grad)
// partial [gridDim.x][7*CIN + 1][COUT]: this CTA's sums of dW ([7][CIN][COUT]) and of the bias gradient; wgrad_final_kernel adds the rows.
template <int CIN, int CPL>
__global__ void __launch_bounds__(THREADS, CPL <= 2 ? 3 : 2)
wgrad_kernel(const int32_t* __restrict__ plan, GinSide side, GinSrcView X, const float* __restrict__ dY, float* __restrict__ partial,
             int group, int B, int total_tiles) {
  extern __shared__ __align__(16) float smem_f[];
  constexpr int COUT = 32 * CPL;
  constexpr int NACC = 7 * CIN + 1;                    // + bias
  float* red = smem_f;                                 // [NACC][COUT] block reduction
  float* xs = red + NACC * COUT;                       // [max_slots][CIN][TM]
  long long* dst_s = reinterpret_cast<long long*>(xs + side.max_slots * CIN * TM);
  __shared__ int8_t tap_s[GIN_MAX_SLOTS];
  __shared__ int slot_of_tap[7];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[7][CIN][CPL], bsum[CPL];
#pragma unroll
  for (int t = 0; t < 7; ++t)
#pragma unroll
    for (int c = 0; c < CIN; ++c)
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[t][c][k] = 0.f;
#pragma unroll
  for (int k = 0; k < CPL; ++k) bsum[k] = 0.f;

  for (int T = blockIdx.x; T < total_tiles; T += gridDim.x) {
    const int G = T / side.ntiles, t = T % side.ntiles;
    const GinTileDesc* desc = reinterpret_cast<const GinTileDesc*>(plan + side.tiles_off) + t;
    const int nslots = desc->nslots;
    __syncthreads();
    if (threadIdx.x < GIN_MAX_SLOTS) tap_s[threadIdx.x] = desc->tap[threadIdx.x];
    if (threadIdx.x < 7) {                             // first slot of every tap (bank 0); later banks take the slow path
      int s = -1;
      for (int q = 0; q < nslots; ++q) if (desc->tap[q] == threadIdx.x) { s = q; break; }
      slot_of_tap[threadIdx.x] = s;
    }
    stage_tile<CIN>(plan, side, desc, X, G, group, B, xs, dst_s, t);
    __syncthreads();
    int first[7], ndistinct = 0;
#pragma unroll
    for (int tp = 0; tp < 7; ++tp) { first[tp] = slot_of_tap[tp]; ndistinct += first[tp] >= 0; }
    const bool has_extra = nslots > ndistinct;
#pragma unroll 1
    for (int q4 = 0; q4 < ROWS_PER_WARP; q4 += 4) {
      const int r0 = warp * ROWS_PER_WARP + q4;
      float dv[4][CPL];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const long long d = dst_s[r0 + r];
#pragma unroll
        for (int k = 0; k < CPL; ++k) dv[r][k] = 0.f;
        if (d >= 0) {
          const float* p = dY + (size_t)d * COUT + lane * CPL;
          if (CPL == 2) { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); dv[r][0] = v.x; dv[r][CPL - 1] = v.y; }
          else if (CPL == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); dv[r][0] = v.x; dv[r][1 % CPL] = v.y; dv[r][2 % CPL] = v.z; dv[r][3 % CPL] = v.w; }
          else {
#pragma unroll
            for (int k = 0; k < CPL; ++k) dv[r][k] = __ldg(p + k);
          }
        }
#pragma unroll
        for (int k = 0; k < CPL; ++k) bsum[k] += dv[r][k];
      }
#pragma unroll
      for (int tp = 0; tp < 7; ++tp) {
        if (first[tp] < 0) continue;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + (first[tp] * CIN + c) * TM + r0);
          const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < CPL; ++k) acc[tp][c][k] = fmaf(x4[r], dv[r][k], acc[tp][c][k]);
        }
      }
      // slots of banks >= 1 (a tap used twice by one row: only at the stitched corners)
      for (int slot = 0; has_extra && slot < nslots; ++slot) {
        const int tp = tap_s[slot];
        if (slot_of_tap[tp] == slot) continue;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + (slot * CIN + c) * TM + r0);
          const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int t2 = 0; t2 < 7; ++t2)
            if (t2 == tp) {
#pragma unroll
              for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < CPL; ++k) acc[t2][c][k] = fmaf(x4[r], dv[r][k], acc[t2][c][k]);
            }
        }
      }
    }
  }
  // block reduction through shared memory, warp after warp in a fixed order, then this CTA's sums go to its own row of
  // `partial` (wgrad_final_kernel adds the rows, again in a fixed order): no atomics, reproducible run to run
  __syncthreads();
  for (int i = threadIdx.x; i < NACC * COUT; i += THREADS) red[i] = 0.f;
  __syncthreads();
#pragma unroll 1
  for (int w = 0; w < WARPS; ++w) {
    if (warp == w) {
#pragma unroll
      for (int t = 0; t < 7; ++t)
#pragma unroll
        for (int c = 0; c < CIN; ++c)
#pragma unroll
          for (int k = 0; k < CPL; ++k) red[(t * CIN + c) * COUT + lane * CPL + k] += acc[t][c][k];
#pragma unroll
      for (int k = 0; k < CPL; ++k) red[7 * CIN * COUT + lane * CPL + k] += bsum[k];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < NACC * COUT; i += THREADS) partial[(size_t)blockIdx.x * NACC * COUT + i] = red[i];
}

// dWp[i] / db[i]: sum over the CTAs' rows.  One WARP per output: lane l adds rows l, l+32, ... in order, then a fixed shuffle tree.
__global__ void __launch_bounds__(256) wgrad_final_kernel(const float* __restrict__ partial, int nparts, int n_w, int n_b, float* __restrict__ dWp,
                                                          float* __restrict__ db) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, n = n_w + n_b;
  if (i >= n) return;
  float s = 0.f;
  for (int k = lane; k < nparts; k += 32) s += partial[(size_t)k * n + i];
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (lane == 0) { if (i < n_w) dWp[i] = s; else if (db) db[i - n_w] = s; }
}

constexpr int WGRAD_MAX_CTAS = 148 * 3;
inline size_t wgrad_partial_bytes(int cin, int cout) { return (size_t)WGRAD_MAX_CTAS * (7 * cin + 1) * cout * 4; }

inline size_t fwd_smem(int cin, int cout, int max_slots) { return (size_t)(7 * cin * cout + max_slots * cin * TM) * 4 + TM * 8; }
inline size_t wgrad_smem(int cin, int cout, int max_slots) { return (size_t)((7 * cin + 1) * cout + max_slots * cin * TM) * 4 + TM * 8; }

}  // namespace narrow

inline bool narrow_supported(int Cin, int Cout, const GinSide& side) {
  return Cin == 3 && (Cout == 64 || Cout == 128) && side.max_slots <= GIN_MAX_SLOTS &&
         narrow::wgrad_smem(Cin, Cout, side.max_slots) <= 200 * 1024;
}

template <typename K>
inline int narrow_config(K kern, size_t smem) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess ? 0 : -3;
}

// stats (optional): [grid][2][Cout] per-CTA column sums of y and y^2, *nparts receives the grid size; y_f16: Y is fp16
inline int launch_narrow_fwd(const int32_t* plan_dev, const GinSide& side, int group, GinSrcView X, const float* W, const float* bias, float* Y,
                             int B, int Cin, int Cout, cudaStream_t st, float* stats = nullptr, int* nparts = nullptr, int y_f16 = 0) {
  const int groups = (B + group - 1) / group, total = groups * side.ntiles;
  const size_t smem = narrow::fwd_smem(Cin, Cout, side.max_slots);
  const int grid = total < 148 * 4 ? total : 148 * 4;      // several CTAs per SM hide the table -> value load chain of a tile
  if (nparts) *nparts = stats ? grid : 0;
  if (Cout == 64) {
    auto k = narrow::fwd_kernel<3, 2>;
    if (narrow_config(k, smem)) return -3;
    k<<<grid, narrow::THREADS, smem, st>>>(plan_dev, side, X, W, bias, Y, group, B, total, stats, y_f16);
  } else {
    auto k = narrow::fwd_kernel<3, 4>;
    if (narrow_config(k, smem)) return -3;
    k<<<grid, narrow::THREADS, smem, st>>>(plan_dev, side, X, W, bias, Y, group, B, total, stats, y_f16);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

// partial: narrow::wgrad_partial_bytes(Cin, Cout) of workspace.  dWp [7][Cin][Cout] and db (or null) are plainly overwritten.
inline int launch_narrow_wgrad(const int32_t* plan_dev, const GinSide& side, int group, GinSrcView X, const float* dY, float* dWp, float* db,
                               float* partial, int B, int Cin, int Cout, cudaStream_t st) {
  const int groups = (B + group - 1) / group, total = groups * side.ntiles;
  const size_t smem = narrow::wgrad_smem(Cin, Cout, side.max_slots);
  const int grid = total < narrow::WGRAD_MAX_CTAS ? total : narrow::WGRAD_MAX_CTAS;
  if (Cout == 64) {
    auto k = narrow::wgrad_kernel<3, 2>;
    if (narrow_config(k, smem)) return -3;
    k<<<grid, narrow::THREADS, smem, st>>>(plan_dev, side, X, dY, partial, group, B, total);
  } else {
    auto k = narrow::wgrad_kernel<3, 4>;
    if (narrow_config(k, smem)) return -3;
    k<<<grid, narrow::THREADS, smem, st>>>(plan_dev, side, X, dY, partial, group, B, total);
  }
  if (cudaGetLastError() != cudaSuccess) return -3;
  const int n_w = 7 * Cin * Cout;
  narrow::wgrad_final_kernel<<<(n_w + Cout + 7) / 8, 256, 0, st>>>(partial, grid, n_w, Cout, dWp, db);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace gin

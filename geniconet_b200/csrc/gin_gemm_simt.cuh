// fp32 CUDA-core gather-GEMM: the exact-fp32 implementation of the hex-conv forward / dgrad
// (rows a1+a2+a3+a5 of SURVEY 8a) and of wgrad.  Used for the narrow xyz layer (Cin = 3: the
// contraction is K = 21, no tensor-core shape) and as the fp32 cross-check of the tcgen05 path.
//
//   dst[row, :] = bias + sum_slots  X[src[slot][row], :] * W[tap(slot)]        (W[t] is [K][N])
//
// The chart padding, pole averaging, stride-2 sub-lattice and the adjoint of all three are
// entirely in the plan tables (gin_host.cpp); this kernel only gathers rows.
#pragma once
#include "gin_common.cuh"

namespace gin {

constexpr int SIMT_TM = 128, SIMT_TN = 64, SIMT_TK = 16, SIMT_THREADS = 256;

struct RowSrc {      // resolved gather source of one row
  long long off;     // element offset of channel 0 (valid when kind == 1)
  int kind;          // 0 zero, 1 pixel, 2 pole mean
  int sample, pole;
};

GIN_DEVINL RowSrc resolve_src(int code, long long base_src, long long total_src, int group_sample0, int B,
                              const GinSrcView& X) {
  RowSrc r;
  r.kind = 0; r.off = 0; r.sample = 0; r.pole = 0;
  if (code >= 0) {
    long long gp = base_src + code;
    if (gp < total_src) {
      long long b = gp / X.P, p = gp - b * X.P;
      r.kind = 1;
      r.off = b * X.sb + p * X.sp;
    }
  } else if (code <= -2) {
    int q = -2 - code;
    int sample = group_sample0 + (q >> 1);
    if (sample < B) { r.kind = 2; r.sample = sample; r.pole = q & 1; }
  }
  return r;
}

template <bool VEC>
GIN_DEVINL void load_row8(const RowSrc& rs, const GinSrcView& X, const int32_t* __restrict__ ring, int c0, int K,
                          float v[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (rs.kind == 1) {
    if (VEC) {
      if (c0 < K) {
        const float4* p = reinterpret_cast<const float4*>(X.p + rs.off + c0);
        float4 a = __ldg(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        if (c0 + 4 < K) { float4 b = __ldg(p + 1); v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c0 + i < K) v[i] = __ldg(X.p + rs.off + (long long)(c0 + i) * X.sc);
    }
  } else if (rs.kind == 2) {
    for (int j = 0; j < 5; ++j) {
      long long off = (long long)rs.sample * X.sb + (long long)ring[rs.pole * 5 + j] * X.sp;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c0 + i < K) v[i] += __ldg(X.p + off + (long long)(c0 + i) * X.sc);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= 0.2f;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(SIMT_THREADS)
gather_gemm_simt_kernel(const int32_t* __restrict__ plan, GinSide side, GinSrcView X, const float* __restrict__ W,
                        const float* __restrict__ bias, float* __restrict__ Y, int group, int B, int K, int N) {
  __shared__ __align__(16) float As[SIMT_TK][SIMT_TM + 4];
  __shared__ __align__(16) float Bs[SIMT_TK][SIMT_TN];
  __shared__ long long dst_s[SIMT_TM];

  const int tid = threadIdx.x;
  const int G = blockIdx.x / side.ntiles, t = blockIdx.x % side.ntiles;
  const int n0 = blockIdx.y * SIMT_TN;
  const long long base_src = (long long)G * group * side.P_src, total_src = (long long)B * side.P_src;
  const long long base_dst = (long long)G * group * side.P_dst, total_dst = (long long)B * side.P_dst;
  const GinTileDesc* desc = reinterpret_cast<const GinTileDesc*>(plan + side.tiles_off) + t;
  const int32_t* src_tab = plan + side.src_off + desc->src_off;
  const int32_t* ring = plan + side.ring_off;

  if (tid < SIMT_TM) {
    int r = plan[side.rows_off + t * SIMT_TM + tid];
    long long d = (r >= 0) ? base_dst + r : -1;
    dst_s[tid] = (d >= 0 && d < total_dst) ? d : -1;
  }

  const int lrow = tid >> 1, lk = (tid & 1) * 8;       // A-load mapping
  const int brow = tid >> 4, bcol = (tid & 15) * 4;    // B-load mapping
  const int ty = tid >> 4, tx = tid & 15;              // compute mapping: rows ty*8.., cols tx*4..
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nslots = desc->nslots;
  for (int slot = 0; slot < nslots; ++slot) {
    const int tap = desc->tap[slot];
    const RowSrc rs = resolve_src(src_tab[slot * SIMT_TM + lrow], base_src, total_src, G * group, B, X);
    const float* Wt = W + (size_t)tap * K * N;
    for (int k0 = 0; k0 < K; k0 += SIMT_TK) {
      float av[8];
      load_row8<VEC>(rs, X, ring, k0 + lk, K, av);
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      {
        int kk = k0 + brow, nn = n0 + bcol;
        if (kk < K) {
          if (nn + 3 < N && (N & 3) == 0) {
            bv = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)kk * N + nn));
          } else {
            if (nn + 0 < N) bv.x = __ldg(Wt + (size_t)kk * N + nn + 0);
            if (nn + 1 < N) bv.y = __ldg(Wt + (size_t)kk * N + nn + 1);
            if (nn + 2 < N) bv.z = __ldg(Wt + (size_t)kk * N + nn + 2);
            if (nn + 3 < N) bv.w = __ldg(Wt + (size_t)kk * N + nn + 3);
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) As[lk + i][lrow] = av[i];
      *reinterpret_cast<float4*>(&Bs[brow][bcol]) = bv;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SIMT_TK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
    }
  }
  __syncthreads();
  float bsv[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + tx * 4 + j < N) bsv[j] = __ldg(bias + n0 + tx * 4 + j);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long d = dst_s[ty * 8 + i];
    if (d < 0) continue;
    float* yp = Y + d * N + n0 + tx * 4;
    if ((N & 3) == 0 && n0 + tx * 4 + 3 < N) {
      *reinterpret_cast<float4*>(yp) =
          make_float4(acc[i][0] + bsv[0], acc[i][1] + bsv[1], acc[i][2] + bsv[2], acc[i][3] + bsv[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + tx * 4 + j < N) yp[j] = acc[i][j] + bsv[j];
    }
  }
}

// ---------------------------------------------------------------------------------- wgrad
// dWp[t][ci][co] += sum_rows X[src_t[row], ci] * dY[row, co]   over this CTA's slice of tiles.
// grid = (row slices, 7 taps, ci-blocks * co-blocks); 64x64 output block per CTA, fp32 atomics.
constexpr int WG_TC = 64, WG_ROWS = 16;

template <bool VEC>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const int32_t* __restrict__ plan, GinSide side, GinSrcView X, const float* __restrict__ dY,
                  float* __restrict__ dWp, int group, int B, int Cin, int Cout, int tiles_per_cta, int total_tiles) {
  __shared__ __align__(16) float As[WG_ROWS][WG_TC];
  __shared__ __align__(16) float Bs[WG_ROWS][WG_TC];
  const int tid = threadIdx.x;
  const int tap = blockIdx.y;
  const int ncb = (Cout + WG_TC - 1) / WG_TC;
  const int ci0 = (blockIdx.z / ncb) * WG_TC, co0 = (blockIdx.z % ncb) * WG_TC;
  const int32_t* ring = plan + side.ring_off;
  const long long total_src = (long long)B * side.P_src, total_dst = (long long)B * side.P_dst;
  const int lr = tid >> 4, lc = (tid & 15) * 4;   // loads: 16 rows x 16 float4
  const int ty = tid >> 4, tx = tid & 15;         // compute: ci ty*4.., co tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int T0 = blockIdx.x * tiles_per_cta, T1 = min(T0 + tiles_per_cta, total_tiles);
  for (int T = T0; T < T1; ++T) {
    const int G = T / side.ntiles, t = T % side.ntiles;
    const GinTileDesc* desc = reinterpret_cast<const GinTileDesc*>(plan + side.tiles_off) + t;
    int slot = -1;
    for (int s = 0; s < desc->nslots; ++s)
      if (desc->tap[s] == tap) { slot = s; break; }
    if (slot < 0) continue;
    const int32_t* src_tab = plan + side.src_off + desc->src_off + slot * GIN_TILE_M;
    const long long base_src = (long long)G * group * side.P_src, base_dst = (long long)G * group * side.P_dst;
    for (int r0 = 0; r0 < GIN_TILE_M; r0 += WG_ROWS) {
      const int row = r0 + lr;
      int drow = plan[side.rows_off + t * GIN_TILE_M + row];
      long long d = (drow >= 0) ? base_dst + drow : -1;
      if (d >= total_dst) d = -1;
      // A: 4 channels ci0+lc.. of the gathered row
      float a4[4] = {0.f, 0.f, 0.f, 0.f};
      if (d >= 0) {
        RowSrc rs = resolve_src(src_tab[row], base_src, total_src, G * group, B, X);
        const int c0 = ci0 + lc;
        if (rs.kind == 1) {
          if (VEC) {
            if (c0 < Cin) { float4 a = __ldg(reinterpret_cast<const float4*>(X.p + rs.off + c0)); a4[0] = a.x; a4[1] = a.y; a4[2] = a.z; a4[3] = a.w; }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (c0 + i < Cin) a4[i] = __ldg(X.p + rs.off + (long long)(c0 + i) * X.sc);
          }
        } else if (rs.kind == 2) {
          for (int j = 0; j < 5; ++j) {
            long long off = (long long)rs.sample * X.sb + (long long)ring[rs.pole * 5 + j] * X.sp;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (c0 + i < Cin) a4[i] += 0.2f * __ldg(X.p + off + (long long)(c0 + i) * X.sc);
          }
        }
      }
      float b4[4] = {0.f, 0.f, 0.f, 0.f};
      if (d >= 0) {
        const int c0 = co0 + lc;
        if ((Cout & 3) == 0 && c0 + 3 < Cout) {
          float4 b = __ldg(reinterpret_cast<const float4*>(dY + d * Cout + c0));
          b4[0] = b.x; b4[1] = b.y; b4[2] = b.z; b4[3] = b.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (c0 + i < Cout) b4[i] = __ldg(dY + d * Cout + c0 + i);
        }
      }
      __syncthreads();
      *reinterpret_cast<float4*>(&As[lr][lc]) = make_float4(a4[0], a4[1], a4[2], a4[3]);
      *reinterpret_cast<float4*>(&Bs[lr][lc]) = make_float4(b4[0], b4[1], b4[2], b4[3]);
      __syncthreads();
#pragma unroll
      for (int r = 0; r < WG_ROWS; ++r) {
        float4 a = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
        float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ci = ci0 + ty * 4 + i, co = co0 + tx * 4 + j;
      if (ci < Cin && co < Cout && acc[i][j] != 0.f) atomicAdd(dWp + ((size_t)tap * Cin + ci) * Cout + co, acc[i][j]);
    }
}

// db[co] = sum_rows dY[row, co]
__global__ void __launch_bounds__(256)
bias_grad_kernel(const float* __restrict__ dY, float* __restrict__ db, long long rows, int C, int rows_per_cta) {
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(r0 + (long long)rows_per_cta, rows);
  const int cpt = min(C, 256), lanes = 256 / cpt;
  const int cl = threadIdx.x % cpt, lane = threadIdx.x / cpt;
  if (lane >= lanes) return;
  for (int c = cl; c < C; c += cpt) {
    float s = 0.f;
    for (long long r = r0 + lane; r < r1; r += lanes) s += __ldg(dY + r * C + c);
    atomicAdd(db + c, s);
  }
}

// weight [Cout][Cin][7] -> wf[7][Cin][Cout], wd[7][Cout][Cin] (fp32) and bf16 K-major copies
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wd,
                                    unsigned short* __restrict__ bf, unsigned short* __restrict__ bd, int Cin, int Cout, int fwd_f16) {
  const long long n = (long long)Cin * Cout * 7;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i % 7);
    long long r = i / 7;
    int ci = (int)(r % Cin), co = (int)(r / Cin);
    float v = w[i];
    const unsigned short h = cvt_op(v, 0);                      // dgrad tiles: bf16
    const unsigned short hf = cvt_op(v, fwd_f16);               // forward tiles: the forward operand format
    wf[((size_t)t * Cin + ci) * Cout + co] = v;
    wd[((size_t)t * Cout + co) * Cin + ci] = v;
    // tcgen05 B operand tiles, stored exactly as the smem image: [tap][k/64][n][64] with the 16-byte chunk index
    // XOR-swizzled by (n & 7), so one contiguous bulk copy of N_TILE*128 bytes lands a ready SWIZZLE_128B tile.
    if ((Cin & 63) == 0 && (Cout & 63) == 0) {
      {  // forward: n = co, k = ci
        const int kc = ci >> 6, c = (ci >> 3) & 7, e = ci & 7;
        bf[((((size_t)t * (Cin >> 6) + kc) * Cout + co) << 6) + (((c ^ (co & 7)) << 3) | e)] = hf;
      }
      {  // dgrad: n = ci, k = co
        const int kc = co >> 6, c = (co >> 3) & 7, e = co & 7;
        bd[((((size_t)t * (Cout >> 6) + kc) * Cin + ci) << 6) + (((c ^ (ci & 7)) << 3) | e)] = h;
      }
    }
  }
}

// Only the tcgen05 tile images (see pack_weights_kernel) of the weight [w0; w1] concatenated along Cout -- what the fused
// chains need: two sibling convolutions become one GEMM without materialising the concatenation.
__global__ void pack_weights_bf16_kernel(const float* __restrict__ w0, int Cout0, const float* __restrict__ w1, unsigned short* __restrict__ bf,
                                         unsigned short* __restrict__ bd, int Cin, int Cout, int fwd_f16) {
  GIN_PDL_SYNC();
  const long long n = (long long)Cin * Cout * 7;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % 7);
    const long long r = i / 7;
    const int ci = (int)(r % Cin), co = (int)(r / Cin);
    const float v = co < Cout0 ? w0[i] : w1[i - (long long)Cout0 * Cin * 7];
    const unsigned short h = cvt_op(v, 0), hf = cvt_op(v, fwd_f16);
    {
      const int kc = ci >> 6, c = (ci >> 3) & 7, e = ci & 7;
      bf[((((size_t)t * (Cin >> 6) + kc) * Cout + co) << 6) + (((c ^ (co & 7)) << 3) | e)] = hf;
    }
    {
      const int kc = co >> 6, c = (co >> 3) & 7, e = co & 7;
      bd[((((size_t)t * (Cout >> 6) + kc) * Cin + ci) << 6) + (((c ^ (ci & 7)) << 3) | e)] = h;
    }
  }
}

// The same for up to PACK_MAX_JOBS weights in ONE launch (a fused chain packs ~12 weights per step; launched one by one they
// cost ~8 us each of pure launch latency).  Job j owns blocks [first[j], first[j+1]); descriptors travel as kernel parameters.
constexpr int PACK_MAX_JOBS = 24;
struct PackJobs {
  const float* w0[PACK_MAX_JOBS];
  const float* w1[PACK_MAX_JOBS];
  unsigned short* bf[PACK_MAX_JOBS];
  unsigned short* bd[PACK_MAX_JOBS];
  int cin[PACK_MAX_JOBS], cout0[PACK_MAX_JOBS], cout[PACK_MAX_JOBS];
  int first[PACK_MAX_JOBS + 1];
  int n, fwd_f16;
};
__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const PackJobs J) {
  GIN_PDL_SYNC();
  int j = 0;
  while (j + 1 < J.n && (int)blockIdx.x >= J.first[j + 1]) ++j;
  const int Cin = J.cin[j], Cout = J.cout[j], Cout0 = J.cout0[j];
  const float* __restrict__ w0 = J.w0[j];
  const float* __restrict__ w1 = J.w1[j];
  // One thread = one 16-byte chunk (8 consecutive K elements) of a tile image, threads in the order of the image: the stores are
  // whole 128-byte rows; the reads are strided gathers from the fp32 weight, which is a few MB and L2-resident.
  //   forward image  [t][Cin/64][Cout][8 chunks]: K = ci, row n = co;   dgrad image [t][Cout/64][Cin][8 chunks]: K = co, row n = ci
  const long long half = 7LL * Cin * Cout / 8, n = 2 * half;
  const int nb = J.first[j + 1] - J.first[j];
  for (long long i = (long long)(blockIdx.x - J.first[j]) * 256 + threadIdx.x; i < n; i += (long long)nb * 256) {
    const bool dg = i >= half;
    long long pos = dg ? i - half : i;
    const int cpos = (int)(pos & 7);
    pos >>= 3;
    const int N = dg ? Cin : Cout, K = dg ? Cout : Cin;
    const int row = (int)(pos % N);
    pos /= N;
    const int kc = (int)(pos % (K >> 6)), t = (int)(pos / (K >> 6));
    const int k0 = kc * 64 + ((cpos ^ (row & 7)) << 3);       // first of the 8 K elements this physical chunk holds
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int co = dg ? k0 + e : row, ci = dg ? row : k0 + e;
      const long long src = ((long long)co * Cin + ci) * 7 + t;
      v[e] = co < Cout0 ? __ldg(w0 + src) : __ldg(w1 + (src - (long long)Cout0 * Cin * 7));
    }
    const int f16 = dg ? 0 : J.fwd_f16;
    uint4 o;
    o.x = pack2_op(v[0], v[1], f16); o.y = pack2_op(v[2], v[3], f16); o.z = pack2_op(v[4], v[5], f16); o.w = pack2_op(v[6], v[7], f16);
    unsigned short* img = dg ? J.bd[j] : J.bf[j];
    *reinterpret_cast<uint4*>(img + (dg ? i - half : i) * 8) = o;
  }
}

// dWp[7][Cin][Cout] -> dW[Cout][Cin][7]
__global__ void unpack_wgrad_kernel(const float* __restrict__ dWp, float* __restrict__ dW, int Cin, int Cout) {
  const long long n = (long long)Cin * Cout * 7;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i % 7);
    long long r = i / 7;
    int ci = (int)(r % Cin), co = (int)(r / Cin);
    dW[i] = dWp[((size_t)t * Cin + ci) * Cout + co];
  }
}

}  // namespace gin

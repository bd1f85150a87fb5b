// Loss path (rows a7, a8 of SURVEY 8a): pole averaging / grid -> vertex list (losses.py:22-31,
// 49-51), area-weighted vertex normals (generate.py:20-43 recipe), uniform Laplacian, and the
// fused Point2Point loss (losses.py:47-82) with its backward.
//
// Everything is driven by the one-ring table of the loss plan: ring r_0..r_{d-1} of vertex v,
// counter-clockwise seen from outside.  Because the ring is closed,
//      sum over incident faces of cross(v1-v0, v2-v0)  ==  sum_i  r_i x r_{i+1}
// so the un-normalised vertex normal needs no face list and no atomics.
#pragma once
#include "gin_common.cuh"

namespace gin {

struct F3 { float x, y, z; };
GIN_DEVINL F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
GIN_DEVINL F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
GIN_DEVINL F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
GIN_DEVINL F3 operator*(float s, F3 a) { return f3(s * a.x, s * a.y, s * a.z); }
GIN_DEVINL float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GIN_DEVINL F3 cross(F3 a, F3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GIN_DEVINL F3 ldv(const float* v, long long i) { return f3(v[3 * i], v[3 * i + 1], v[3 * i + 2]); }
GIN_DEVINL void stv(float* v, long long i, F3 a) { v[3 * i] = a.x; v[3 * i + 1] = a.y; v[3 * i + 2] = a.z; }

constexpr float kNormalEps = 1e-10f;  // generate.py:20 eps
constexpr float kCosEps = 1e-8f;      // torch.nn.CosineSimilarity default eps (losses.py:19)

// v[b][u][c] : pixels copied, poles = mean of their 5 ring pixels.  C channels (3 for xyz).
__global__ void __launch_bounds__(256)
pole_vertices_fwd_kernel(const int32_t* __restrict__ plan, GinSrcView X, float* __restrict__ v, int B, int C) {
  const GinLossPlanHdr* h = reinterpret_cast<const GinLossPlanHdr*>(plan);
  const int P = h->P, V = h->V;
  const int32_t* ring = plan + h->pole_off;
  const long long total = (long long)B * V * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bu = i / C;
    const int u = (int)(bu % V);
    const long long b = bu / V;
    const float* xb = X.p + b * X.sb + (long long)c * X.sc;
    float r;
    if (u < P) r = __ldg(xb + (long long)u * X.sp);
    else {
      r = 0.f;
#pragma unroll
      for (int j = 0; j < 5; ++j) r += __ldg(xb + (long long)ring[(u - P) * 5 + j] * X.sp);
      r *= 0.2f;
    }
    v[i] = r;
  }
}

// dx[b][c][p] = dv[b][p][c] + (p in north ring) dv[b][P][c]/5 + (p in south ring) dv[b][P+1][c]/5
__global__ void __launch_bounds__(256)
pole_vertices_bwd_kernel(const int32_t* __restrict__ plan, const float* __restrict__ dv, float* __restrict__ dx,
                         long long sb, long long sp, long long sc, int B, int C) {
  const GinLossPlanHdr* h = reinterpret_cast<const GinLossPlanHdr*>(plan);
  const int P = h->P, V = h->V;
  const int32_t* flag = plan + h->flag_off;
  const long long total = (long long)B * P * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // iterate with the pixel index fastest when the destination is planar (sp == 1), channel fastest otherwise
    int c, p; long long b;
    if (sp == 1) { p = (int)(i % P); long long bc = i / P; c = (int)(bc % C); b = bc / C; }
    else { c = (int)(i % C); long long bp = i / C; p = (int)(bp % P); b = bp / P; }
    const float* dvb = dv + (size_t)b * V * C;
    float g = dvb[(size_t)p * C + c];
    const int fl = flag[p];
    if (fl & 1) g += 0.2f * dvb[(size_t)P * C + c];
    if (fl & 2) g += 0.2f * dvb[(size_t)(P + 1) * C + c];
    dx[b * sb + (long long)p * sp + (long long)c * sc] = g;
  }
}

struct RingGeom {
  F3 nraw;      // sum_i r_i x r_{i+1}
  F3 mean;      // mean of the ring
  int deg;
};

GIN_DEVINL RingGeom ring_geom(const float* __restrict__ vb, const int32_t* __restrict__ nb) {
  RingGeom g;
  F3 r[6];
  int deg = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    const int id = nb[a];
    if (id >= 0) { r[a] = ldv(vb, id); deg = a + 1; } else r[a] = f3(0.f, 0.f, 0.f);
  }
  F3 acc = f3(0.f, 0.f, 0.f), sum = f3(0.f, 0.f, 0.f);
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    if (a < deg) {
      const F3 nx = (a + 1 < deg) ? r[(a + 1) % 6] : r[0];
      acc = acc + cross(r[a], nx);
      sum = sum + r[a];
    }
  }
  g.nraw = acc; g.mean = (1.0f / (float)deg) * sum; g.deg = deg;
  return g;
}

// compute_vertex_normals (losses.py:54) and compute_laplacian_batch (losses.py:57) as plain ops
__global__ void __launch_bounds__(256)
normals_laplacian_kernel(const int32_t* __restrict__ plan, const float* __restrict__ v, float* __restrict__ nrm,
                         float* __restrict__ lap, int B) {
  const GinLossPlanHdr* h = reinterpret_cast<const GinLossPlanHdr*>(plan);
  const int V = h->V;
  const int32_t* nbt = plan + h->ring_off;
  const long long total = (long long)B * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % V);
    const float* vb = v + (size_t)(i / V) * V * 3;
    const RingGeom g = ring_geom(vb, nbt + (size_t)u * 6);
    if (nrm) {
      const float m = fmaxf(sqrtf(dot(g.nraw, g.nraw)), kNormalEps);
      stv(nrm, i, (1.0f / m) * g.nraw);
    }
    if (lap) stv(lap, i, g.mean - ldv(vb, u));
  }
}

// partial sums of the three loss terms; target is [B][9][V] (generate.py:200-203)
__global__ void __launch_bounds__(256)
p2p_fwd_kernel(const int32_t* __restrict__ plan, const float* __restrict__ v, const float* __restrict__ target,
               double* __restrict__ partial, int B) {
  const GinLossPlanHdr* h = reinterpret_cast<const GinLossPlanHdr*>(plan);
  const int V = h->V;
  const int32_t* nbt = plan + h->ring_off;
  const long long total = (long long)B * V;
  double s_pos = 0.0, s_nor = 0.0, s_lap = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % V);
    const long long b = i / V;
    const float* vb = v + (size_t)b * V * 3;
    const float* tb = target + (size_t)b * 9 * V;
    const RingGeom g = ring_geom(vb, nbt + (size_t)u * 6);
    const F3 p = ldv(vb, u);
    const F3 tp = f3(tb[u], tb[V + u], tb[2 * V + u]);
    const F3 tn = f3(tb[3 * V + u], tb[4 * V + u], tb[5 * V + u]);
    const F3 tl = f3(tb[6 * V + u], tb[7 * V + u], tb[8 * V + u]);
    const F3 dp = p - tp;
    s_pos += (double)dot(dp, dp);
    const float m = fmaxf(sqrtf(dot(g.nraw, g.nraw)), kNormalEps);
    const F3 n = (1.0f / m) * g.nraw;
    const float an = fmaxf(sqrtf(dot(n, n)), kCosEps), at = fmaxf(sqrtf(dot(tn, tn)), kCosEps);
    s_nor += (double)(1.0f - dot(n, tn) / (an * at));
    const F3 dl = (g.mean - p) - tl;
    s_lap += (double)dot(dl, dl);
  }
  __shared__ double sh[3][256];
  sh[0][threadIdx.x] = s_pos; sh[1][threadIdx.x] = s_nor; sh[2][threadIdx.x] = s_lap;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) partial[(size_t)blockIdx.x * 3 + threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void p2p_final_kernel(const double* __restrict__ partial, int nparts, float* __restrict__ out, double inv_bv,
                                 float f_pos, float f_nor, float f_lap) {
  __shared__ double sh[3][256];
  double s[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < nparts; i += blockDim.x)
    for (int k = 0; k < 3; ++k) s[k] += partial[(size_t)i * 3 + k];
  for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] = s[k];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float l_pos = (float)(sh[0][0] * inv_bv / 3.0), l_nor = (float)(sh[1][0] * inv_bv), l_lap = (float)(sh[2][0] * inv_bv / 3.0);
    out[0] = l_pos; out[1] = l_nor; out[2] = l_lap;
    out[3] = f_pos * l_pos + f_nor * l_nor + f_lap * l_lap;
  }
}

// backward pass A: per-vertex cotangents  g[b][u][0:3] = dL/d(nraw_u),  g[b][u][3:6] = dL/d(lap_u),
// and the direct position term into dv.
__global__ void __launch_bounds__(256)
p2p_bwd_vertex_kernel(const int32_t* __restrict__ plan, const float* __restrict__ v, const float* __restrict__ target,
                      const float* __restrict__ dout, float f_pos, float f_nor, float f_lap,
                      float* __restrict__ g, float* __restrict__ dv, int B) {
  const GinLossPlanHdr* h = reinterpret_cast<const GinLossPlanHdr*>(plan);
  const int V = h->V;
  const int32_t* nbt = plan + h->ring_off;
  const long long total = (long long)B * V;
  const float go = dout[0];
  const float c_pos = go * f_pos * 2.0f / (3.0f * (float)total);
  const float c_nor = go * f_nor / (float)total;
  const float c_lap = go * f_lap * 2.0f / (3.0f * (float)total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % V);
    const long long b = i / V;
    const float* vb = v + (size_t)b * V * 3;
    const float* tb = target + (size_t)b * 9 * V;
    const RingGeom rg = ring_geom(vb, nbt + (size_t)u * 6);
    const F3 p = ldv(vb, u);
    const F3 tp = f3(tb[u], tb[V + u], tb[2 * V + u]);
    const F3 tn = f3(tb[3 * V + u], tb[4 * V + u], tb[5 * V + u]);
    const F3 tl = f3(tb[6 * V + u], tb[7 * V + u], tb[8 * V + u]);
    // laplacian term: dL/d lap_u ; lap_u = mean(ring) - p_u
    const F3 gl = c_lap * ((rg.mean - p) - tl);
    // normal term
    const float len = sqrtf(dot(rg.nraw, rg.nraw));
    const float m = fmaxf(len, kNormalEps);
    const F3 n = (1.0f / m) * rg.nraw;
    const float a = sqrtf(dot(n, n)), A = fmaxf(a, kCosEps), T = fmaxf(sqrtf(dot(tn, tn)), kCosEps);
    // d(1 - cos)/dn
    F3 gn = (-1.0f / (A * T)) * tn;
    if (a > kCosEps) gn = gn + (dot(n, tn) / (A * A * T * a)) * n;
    gn = c_nor * gn;
    // through n = nraw / max(|nraw|, eps)
    F3 graw;
    if (len > kNormalEps) graw = (1.0f / len) * (gn - dot(gn, n) * n);
    else graw = (1.0f / kNormalEps) * gn;
    float* gi = g + (size_t)i * 6;
    gi[0] = graw.x; gi[1] = graw.y; gi[2] = graw.z; gi[3] = gl.x; gi[4] = gl.y; gi[5] = gl.z;
    stv(dv, i, (c_pos * (p - tp)) - gl);
  }
}

// Backward of compute_vertex_normals / compute_laplacian_batch as plain ops (mesh.utils shim): upstream cotangents gn [B][V][3]
// (w.r.t. the unit normals, may be null) and gl [B][V][3] (w.r.t. the Laplacian, may be null) -> per-vertex cotangents in the
// layout p2p_bwd_gather_kernel reads, and the direct term -gl into dv.  The gather pass then finishes dv.
__global__ void __launch_bounds__(256)
ring_ops_bwd_vertex_kernel(const int32_t* __restrict__ plan, const float* __restrict__ v, const float* __restrict__ gn_in,
                           const float* __restrict__ gl_in, float* __restrict__ g, float* __restrict__ dv, int B) {
  const GinLossPlanHdr* h = reinterpret_cast<const GinLossPlanHdr*>(plan);
  const int V = h->V;
  const int32_t* nbt = plan + h->ring_off;
  const long long total = (long long)B * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % V);
    const float* vb = v + (size_t)(i / V) * V * 3;
    F3 graw = f3(0.f, 0.f, 0.f), gl = f3(0.f, 0.f, 0.f);
    if (gn_in) {
      const RingGeom rg = ring_geom(vb, nbt + (size_t)u * 6);
      const F3 gn = ldv(gn_in, i);
      const float len = sqrtf(dot(rg.nraw, rg.nraw));
      if (len > kNormalEps) {                       // n = nraw / |nraw|
        const F3 n = (1.0f / len) * rg.nraw;
        graw = (1.0f / len) * (gn - dot(gn, n) * n);
      } else graw = (1.0f / kNormalEps) * gn;       // clamped: n = nraw / eps
    }
    if (gl_in) gl = ldv(gl_in, i);
    float* gi = g + (size_t)i * 6;
    gi[0] = graw.x; gi[1] = graw.y; gi[2] = graw.z; gi[3] = gl.x; gi[4] = gl.y; gi[5] = gl.z;
    stv(dv, i, -1.0f * gl);
  }
}

// backward pass B: gather the ring contributions.
//   nraw_w = sum_i r_i x r_{i+1}  =>  d/d(u) over all w with u in ring(w):  sum_i g_{s_i} x (s_{i+1} - s_{i-1}),  s = ring(u)
//   lap_w  = mean(ring(w)) - w    =>  + sum_i gl_{s_i} / deg(s_i)
__global__ void __launch_bounds__(256)
p2p_bwd_gather_kernel(const int32_t* __restrict__ plan, const float* __restrict__ v, const float* __restrict__ g,
                      float* __restrict__ dv, int B) {
  const GinLossPlanHdr* h = reinterpret_cast<const GinLossPlanHdr*>(plan);
  const int V = h->V;
  const int32_t* nbt = plan + h->ring_off;
  const long long total = (long long)B * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % V);
    const long long b = i / V;
    const float* vb = v + (size_t)b * V * 3;
    const float* gb = g + (size_t)b * V * 6;
    const int32_t* nb = nbt + (size_t)u * 6;
    int ids[6];
    int deg = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) { ids[a] = nb[a]; if (ids[a] >= 0) deg = a + 1; }
    F3 acc = ldv(dv, i);
    for (int a = 0; a < deg; ++a) {
      const int w = ids[a];
      const int nx = ids[(a + 1 == deg) ? 0 : a + 1], pv = ids[(a == 0) ? deg - 1 : a - 1];
      const F3 gw = f3(gb[(size_t)w * 6], gb[(size_t)w * 6 + 1], gb[(size_t)w * 6 + 2]);
      const F3 glw = f3(gb[(size_t)w * 6 + 3], gb[(size_t)w * 6 + 4], gb[(size_t)w * 6 + 5]);
      // in w's ring, u is preceded by u's next neighbour and followed by u's previous neighbour
      acc = acc + cross(gw, ldv(vb, nx) - ldv(vb, pv));
      int degw = 0;
#pragma unroll
      for (int q = 0; q < 6; ++q) if (nbt[(size_t)w * 6 + q] >= 0) degw = q + 1;
      acc = acc + (1.0f / (float)degw) * glw;
    }
    stv(dv, i, acc);
  }
}

}  // namespace gin

// tcgen05 implicit-GEMM hex-conv (forward and dgrad; rows a1+a2+a3+a5 of SURVEY 8a).
//
//   dst[row, n] = bias[n] + sum_{slot} sum_k  bf16(X[src[slot][row], k]) * bf16(W[tap(slot)][n][k])
//
// One CTA owns one 128-row tile x N_TILE output channels.  The accumulator lives in TMEM
// (128 lanes x N_TILE fp32 columns); the A operand is GATHERED: eight producer warps read the
// source rows named by the plan (chart padding, pole means, stride-2 lattice and the adjoint
// tables are all just row indices), convert fp32 -> bf16 in registers and write the
// 128-byte-swizzled K-major tile that the UMMA smem descriptor expects.  The B operand is the
// pre-packed bf16 weight slice W[tap][n0:n0+N_TILE][k0:k0+64].  A single elected thread issues
// tcgen05.mma; stages are recycled through mbarriers signalled by tcgen05.commit.
//
// The K loop runs k-chunk outer / tap inner, so the seven taps of one 64-channel chunk re-read
// the same few source rows back to back (L1-resident), and HBM/L2 see each row about once.
#pragma once
#include <cuda_bf16.h>

#include "gin_common.cuh"

namespace gin {
namespace tc {

constexpr int BM = 128;              // rows per tile (UMMA M)
constexpr int BK = 64;               // channels per stage: 64 bf16 = one 128-byte swizzle row
constexpr int PRODUCER_WARPS = 8;
constexpr int PRODUCER_THREADS = PRODUCER_WARPS * 32;
constexpr int THREADS = PRODUCER_THREADS + 32;   // + the MMA / TMEM warp
constexpr int A_BYTES = BM * BK * 2;             // 16 KB

// ------------------------------------------------------------------ PTX wrappers
GIN_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

GIN_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
GIN_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
GIN_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
GIN_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
GIN_DEVINL void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
GIN_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
GIN_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

GIN_DEVINL void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
GIN_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread
GIN_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
GIN_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t = TMEM lane base+t)
GIN_DEVINL void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
GIN_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B (SBO), LBO unused.
GIN_DEVINL uint64_t make_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address
  d |= (uint64_t)0 << 16;                         // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;              // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> f32, both operands K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

GIN_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// byte offset of 16-byte chunk `c` of row `r` inside a 128B-swizzled tile whose base is 1024-aligned
GIN_DEVINL uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

struct Params {
  const int32_t* plan;
  GinSide side;
  int group, B, K, N;
  const float* X;            // [B*P_src][K] fp32
  const __nv_bfloat16* W;    // [7][N][K] bf16 (K contiguous)
  const float* bias;         // [N] or null
  float* Y;                  // [B*P_dst][N] fp32
};

template <int N_TILE, int STAGES>
struct Smem {
  static constexpr int B_BYTES = N_TILE * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;                       // full[STAGES], empty[STAGES], accum
  static constexpr int SRC_OFF = BAR_OFF + (2 * STAGES + 1) * 8 + 8;         // + tmem ptr
  static constexpr int TOTAL = SRC_OFF + GIN_MAX_SLOTS * BM * 4 + BM * 8 + 1024;  // + src table + dst rows + align slack
};

template <int N_TILE, int STAGES>
__global__ void __launch_bounds__(THREADS, (N_TILE <= 128 ? 2 : 1)) gather_gemm_tc_kernel(const Params p) {
  using L = Smem<N_TILE, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  int32_t* src_s = reinterpret_cast<int32_t*>(smem + L::SRC_OFF);            // [nslots][128] global src pixel / code
  long long* dst_s = reinterpret_cast<long long*>(src_s + GIN_MAX_SLOTS * BM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = blockIdx.x / p.side.ntiles, t = blockIdx.x % p.side.ntiles;
  const int n0 = blockIdx.y * N_TILE;
  const GinTileDesc* desc = reinterpret_cast<const GinTileDesc*>(p.plan + p.side.tiles_off) + t;
  const int nslots = desc->nslots;
  const int kchunks = p.K / BK;
  const int total_stages = nslots * kchunks;
  const int32_t* ring = p.plan + p.side.ring_off;

  // ---- one-time setup
  if (warp == PRODUCER_WARPS) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], PRODUCER_THREADS); mbar_init(&empty_bar[s], 1); }
      mbar_init(accum_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, N_TILE < 32 ? 32 : N_TILE);
  } else {
    // resolve the gather table of this tile once: global source pixel, -1 = zero, <= -2 = pole (absolute sample)
    const long long base_src = (long long)G * p.group * p.side.P_src, total_src = (long long)p.B * p.side.P_src;
    const int32_t* src_tab = p.plan + p.side.src_off + desc->src_off;
    for (int i = tid; i < nslots * BM; i += PRODUCER_THREADS) {
      const int code = src_tab[i];
      int v = -1;
      if (code >= 0) { long long gp = base_src + code; v = (gp < total_src) ? (int)gp : -1; }
      else if (code <= -2) { int q = -2 - code; int sample = G * p.group + (q >> 1); v = (sample < p.B) ? -2 - (2 * sample + (q & 1)) : -1; }
      src_s[i] = v;
    }
    if (tid < BM) {
      const long long base_dst = (long long)G * p.group * p.side.P_dst, total_dst = (long long)p.B * p.side.P_dst;
      const int r = p.plan[p.side.rows_off + t * BM + tid];
      const long long d = (r >= 0) ? base_dst + r : -1;
      dst_s[tid] = (d >= 0 && d < total_dst) ? d : -1;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PRODUCER_WARPS) {
    // =========================================================== producers: gather + convert + swizzled store
    const int sub = lane >> 3;        // which of the 4 rows this warp touches per pass
    const int c8 = lane & 7;          // 16-byte bf16 chunk (8 channels) of the row
    for (int it = 0; it < total_stages; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      const int kc = it / nslots, slot = it - kc * nslots;
      const int k0 = kc * BK;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      uint8_t* a_tile = smem + s * L::STAGE_BYTES;
      uint8_t* b_tile = a_tile + A_BYTES;
      // A: 128 gathered rows, 4 passes of 32 rows (4 per warp)
#pragma unroll
      for (int pass = 0; pass < BM / (PRODUCER_WARPS * 4); ++pass) {
        const int r = pass * (PRODUCER_WARPS * 4) + warp * 4 + sub;
        const int v = src_s[slot * BM + r];
        float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
        if (v >= 0) {
          const float4* src = reinterpret_cast<const float4*>(p.X + (size_t)v * p.K + k0 + c8 * 8);
          f0 = __ldg(src); f1 = __ldg(src + 1);
        } else if (v <= -2) {
          const int q = -2 - v, sample = q >> 1, pole = q & 1;
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const float4* src = reinterpret_cast<const float4*>(
                p.X + ((size_t)sample * p.side.P_src + ring[pole * 5 + j]) * p.K + k0 + c8 * 8);
            const float4 a = __ldg(src), b = __ldg(src + 1);
            f0.x += a.x; f0.y += a.y; f0.z += a.z; f0.w += a.w; f1.x += b.x; f1.y += b.y; f1.z += b.z; f1.w += b.w;
          }
          f0.x *= 0.2f; f0.y *= 0.2f; f0.z *= 0.2f; f0.w *= 0.2f; f1.x *= 0.2f; f1.y *= 0.2f; f1.z *= 0.2f; f1.w *= 0.2f;
        }
        uint4 o;
        o.x = pack_bf16x2(f0.x, f0.y); o.y = pack_bf16x2(f0.z, f0.w); o.z = pack_bf16x2(f1.x, f1.y); o.w = pack_bf16x2(f1.z, f1.w);
        *reinterpret_cast<uint4*>(a_tile + swz(r, c8)) = o;
      }
      // B: N_TILE weight rows of this (tap, k-chunk)
      const int tap = desc->tap[slot];
      const __nv_bfloat16* wsrc = p.W + ((size_t)tap * p.N + n0) * p.K + k0;
#pragma unroll
      for (int pass = 0; pass < N_TILE / (PRODUCER_WARPS * 4); ++pass) {
        const int r = pass * (PRODUCER_WARPS * 4) + warp * 4 + sub;
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(wsrc + (size_t)r * p.K + c8 * 8));
        *reinterpret_cast<uint4*>(b_tile + swz(r, c8)) = w;
      }
      fence_async_smem();            // generic-proxy stores -> visible to the tensor-core (async) proxy
      mbar_arrive(&full_bar[s]);
    }
    // =========================================================== epilogue: TMEM -> registers -> global
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3;                         // TMEM lane quarter this warp may touch
    const int half = warp >> 2;                     // column half
    const int row = q * 32 + lane;
    const long long d = dst_s[row];
    constexpr int COLS_PER_WARP = N_TILE / 2;
#pragma unroll 1
    for (int cb = 0; cb < COLS_PER_WARP; cb += 32) {
      const int col = half * COLS_PER_WARP + cb;
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
      tmem_ld_wait();
      if (d >= 0) {
        float* yp = p.Y + (size_t)d * p.N + n0 + col;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          if (p.bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + col + j));
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          }
          *reinterpret_cast<float4*>(yp + j) = o;
        }
      }
    }
    tc_fence_before();
  } else {
    // =========================================================== MMA issuer (one elected thread)
    constexpr uint32_t idesc = make_idesc_bf16(N_TILE);
    if (lane == 0) {
      for (int it = 0; it < total_stages; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
        const uint64_t da = make_desc_kmajor_sw128(a_addr), db = make_desc_kmajor_sw128(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)   // UMMA_K = 16 bf16 = 32 bytes inside the swizzle atom
          umma_bf16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it | k) != 0);
        umma_commit(&empty_bar[s]);          // frees the stage once these MMAs have read it
      }
      umma_commit(accum_bar);                // accumulator complete
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == PRODUCER_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, N_TILE < 32 ? 32 : N_TILE);
  }
}

template <int N_TILE, int STAGES>
int launch(const Params& p, int ntiles, cudaStream_t st) {
  using L = Smem<N_TILE, STAGES>;
  auto kern = gather_gemm_tc_kernel<N_TILE, STAGES>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) return -3;
    configured = true;
  }
  dim3 grid((unsigned)ntiles, (unsigned)(p.N / N_TILE));
  kern<<<grid, THREADS, L::TOTAL, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace tc

inline bool tc_supported(int K, int N) { return K % 64 == 0 && N % 64 == 0 && K >= 64 && N >= 64; }

inline int launch_gather_gemm_tc(const int32_t* plan_dev, const GinSide& side, int group, const float* X, const void* Wb,
                                 const float* bias, float* Y, int B, int K, int N, int ntiles, cudaStream_t st) {
  tc::Params p;
  p.plan = plan_dev; p.side = side; p.group = group; p.B = B; p.K = K; p.N = N;
  p.X = X; p.W = reinterpret_cast<const __nv_bfloat16*>(Wb); p.bias = bias; p.Y = Y;
  if ((long long)B * side.P_src >= 0x7fffffffLL) return -4;
  if (N % 256 == 0) return tc::launch<256, 3>(p, ntiles, st);
  if (N % 128 == 0) return tc::launch<128, 3>(p, ntiles, st);
  return tc::launch<64, 4>(p, ntiles, st);
}

// ======================================================================================= wgrad
//   dWp[tap][ci][co] += sum_rows bf16(X[src_tap[row], ci]) * bf16(dY[row, co])
//
// GEMM with K = pixels.  Both operands are "MN-major" for the tensor core: a staged tile is
// [pixel row][64 channels = 128 swizzled bytes] -- the very image the forward producer writes --
// read by UMMA with K running down the rows.  M = 128 is a PAIR of 64-channel atoms
// (tap, ci-block): two ci-blocks of one tap when Cin >= 128, two taps when Cin == 64.
// A CTA owns `npairs` pairs x N_BLK output channels (npairs * N_BLK <= 512 TMEM columns) and a
// slice of the pixel tiles; partial sums are added to dWp with fp32 atomics.
namespace tcw {
using namespace tc;

struct Params {
  const int32_t* plan;
  GinSide side;
  int group, B, Cin, Cout;
  const float* X;     // [B*P_src][Cin]
  const float* dY;    // [B*P_dst][Cout]
  float* dWp;         // [7][Cin][Cout]
  int total_tiles, tiles_per_cta, units_m;   // units_m: number of pair-groups along M
};

constexpr int ATOM_BYTES = BM * 128;          // 128 pixel rows x 64 bf16

template <int N_BLK, int NPAIRS, int STAGES>
struct Smem {
  static constexpr int A_STAGE = 2 * ATOM_BYTES;                 // one pair
  static constexpr int B_BYTES_ = (N_BLK / 64) * ATOM_BYTES;     // dY tile, double buffered
  static constexpr int B_OFF = STAGES * A_STAGE;
  static constexpr int BAR_OFF = B_OFF + 2 * B_BYTES_;
  // a_full[STAGES], a_empty[STAGES], b_full[2], b_empty[2], accum, tmem slot
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 5) * 8 + 16 + 1024;
};

GIN_DEVINL uint64_t make_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;          // stride between 64-element atoms along M/N
  d |= (uint64_t)(1024u >> 4) << 32;              // stride between 8-row groups along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// gather 128 rows x 64 channels (fp32 -> bf16) into a swizzled atom; rows come from `codes` (plan src codes)
// or, when codes == nullptr, from the tile's own dst rows (the dY side).
GIN_DEVINL void stage_atom(uint8_t* atom, const float* __restrict__ base, int C, int c0, const int32_t* __restrict__ codes,
                           const int32_t* __restrict__ drows, long long base_src, long long total_src, long long base_dst,
                           long long total_dst, int sample0, int B, int P_src, const int32_t* __restrict__ ring, int warp,
                           int lane, bool zero_all) {
  const int sub = lane >> 3, c8 = lane & 7;
#pragma unroll
  for (int pass = 0; pass < BM / (PRODUCER_WARPS * 4); ++pass) {
    const int r = pass * (PRODUCER_WARPS * 4) + warp * 4 + sub;
    float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
    const int dr = drows[r];
    const long long d = (dr >= 0) ? base_dst + dr : -1;
    const bool row_ok = !zero_all && d >= 0 && d < total_dst;
    if (row_ok) {
      if (codes == nullptr) {
        const float4* src = reinterpret_cast<const float4*>(base + (size_t)d * C + c0 + c8 * 8);
        f0 = __ldg(src); f1 = __ldg(src + 1);
      } else {
        const int code = codes[r];
        if (code >= 0) {
          const long long gp = base_src + code;
          if (gp < total_src) {
            const float4* src = reinterpret_cast<const float4*>(base + (size_t)gp * C + c0 + c8 * 8);
            f0 = __ldg(src); f1 = __ldg(src + 1);
          }
        } else if (code <= -2) {
          const int q = -2 - code, sample = sample0 + (q >> 1), pole = q & 1;
          if (sample < B) {
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              const float4* src = reinterpret_cast<const float4*>(base + ((size_t)sample * P_src + ring[pole * 5 + j]) * C + c0 + c8 * 8);
              const float4 a = __ldg(src), b = __ldg(src + 1);
              f0.x += a.x; f0.y += a.y; f0.z += a.z; f0.w += a.w; f1.x += b.x; f1.y += b.y; f1.z += b.z; f1.w += b.w;
            }
            f0.x *= 0.2f; f0.y *= 0.2f; f0.z *= 0.2f; f0.w *= 0.2f; f1.x *= 0.2f; f1.y *= 0.2f; f1.z *= 0.2f; f1.w *= 0.2f;
          }
        }
      }
    }
    uint4 o;
    o.x = pack_bf16x2(f0.x, f0.y); o.y = pack_bf16x2(f0.z, f0.w); o.z = pack_bf16x2(f1.x, f1.y); o.w = pack_bf16x2(f1.z, f1.w);
    *reinterpret_cast<uint4*>(atom + swz(r, c8)) = o;
  }
}

// atom index a in [0, 7*Cin/64): tap = a / (Cin/64), ci-block = a % (Cin/64)  (Cin >= 128: consecutive atoms pair up
// inside one tap because Cin/64 is even; Cin == 64: consecutive taps pair up, atom 7 is an all-zero dummy)
template <int N_BLK, int NPAIRS, int STAGES>
__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const Params p) {
  using L = Smem<N_BLK, NPAIRS, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* a_empty = a_full + STAGES;
  uint64_t* b_full = a_empty + STAGES;
  uint64_t* b_empty = b_full + 2;
  uint64_t* accum_bar = b_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cblocks = p.Cin / 64, natoms = 7 * cblocks;
  const int unit_m = blockIdx.y % p.units_m, unit_n = blockIdx.y / p.units_m;
  const int pair0 = unit_m * NPAIRS;                       // first pair of this CTA
  const int total_pairs = (natoms + 1) / 2;
  const int npairs = min(NPAIRS, total_pairs - pair0);
  const int n0 = unit_n * N_BLK;
  const int T0 = blockIdx.x * p.tiles_per_cta, T1 = min(T0 + p.tiles_per_cta, p.total_tiles);
  const int ntile = T1 - T0;
  constexpr int TM_COLS = (NPAIRS * N_BLK <= 32) ? 32 : (NPAIRS * N_BLK <= 64) ? 64 : (NPAIRS * N_BLK <= 128) ? 128 : (NPAIRS * N_BLK <= 256) ? 256 : 512;

  if (warp == PRODUCER_WARPS) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&a_full[s], PRODUCER_THREADS); mbar_init(&a_empty[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&b_full[s], PRODUCER_THREADS); mbar_init(&b_empty[s], 1); }
      mbar_init(accum_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (ntile <= 0 || npairs <= 0) {
    __syncthreads();
    if (warp == PRODUCER_WARPS) tmem_dealloc(tmem_base, TM_COLS);
    return;
  }

  if (warp < PRODUCER_WARPS) {
    const int32_t* ring = p.plan + p.side.ring_off;
    const long long total_src = (long long)p.B * p.side.P_src, total_dst = (long long)p.B * p.side.P_dst;
    int it = 0;
    for (int ti = 0; ti < ntile; ++ti) {
      const int T = T0 + ti, G = T / p.side.ntiles, t = T % p.side.ntiles;
      const GinTileDesc* desc = reinterpret_cast<const GinTileDesc*>(p.plan + p.side.tiles_off) + t;
      const int32_t* src_tab = p.plan + p.side.src_off + desc->src_off;
      const int32_t* drows = p.plan + p.side.rows_off + t * BM;
      const long long base_src = (long long)G * p.group * p.side.P_src, base_dst = (long long)G * p.group * p.side.P_dst;
      // dY tile of this pixel block (double buffered)
      {
        const int bs = ti & 1;
        mbar_wait(&b_empty[bs], (((uint32_t)(ti >> 1)) & 1u) ^ 1u);
        uint8_t* bt = smem + L::B_OFF + bs * L::B_BYTES_;
#pragma unroll
        for (int j = 0; j < N_BLK / 64; ++j)
          stage_atom(bt + j * ATOM_BYTES, p.dY, p.Cout, n0 + j * 64, nullptr, drows, 0, 0, base_dst, total_dst, 0, p.B, 0, ring, warp, lane, false);
        fence_async_smem();
        mbar_arrive(&b_full[bs]);
      }
      for (int pr = 0; pr < npairs; ++pr, ++it) {
        const int s = it % STAGES;
        mbar_wait(&a_empty[s], (((uint32_t)(it / STAGES)) & 1u) ^ 1u);
        uint8_t* at = smem + s * L::A_STAGE;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int atom = (pair0 + pr) * 2 + h;
          const bool dummy = atom >= natoms;
          const int tap = dummy ? 0 : atom / cblocks, cb = dummy ? 0 : atom % cblocks;
          int slot = -1;
          for (int q = 0; q < desc->nslots; ++q) if (desc->tap[q] == tap) { slot = q; break; }
          stage_atom(at + h * ATOM_BYTES, p.X, p.Cin, cb * 64, src_tab + (slot < 0 ? 0 : slot) * BM, drows, base_src, total_src, base_dst,
                     total_dst, G * p.group, p.B, p.side.P_src, ring, warp, lane, dummy || slot < 0);
        }
        fence_async_smem();
        mbar_arrive(&a_full[s]);
      }
    }
    // ---- epilogue: TMEM -> atomics into dWp
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;                 // row of the pair: atom = row / 64, channel = row % 64
    for (int pr = 0; pr < npairs; ++pr) {
      const int atom = (pair0 + pr) * 2 + (row >> 6);
      const bool ok = atom < natoms;
      const int tap = ok ? atom / cblocks : 0, ci = ok ? (atom % cblocks) * 64 + (row & 63) : 0;
      float* dst = p.dWp + ((size_t)tap * p.Cin + ci) * p.Cout + n0;
#pragma unroll 1
      for (int cb = 0; cb < N_BLK / 2; cb += 32) {
        const int col = half * (N_BLK / 2) + cb;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pr * N_BLK + col), v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + col + j, __uint_as_float(v[j]));
        }
      }
    }
    tc_fence_before();
  } else {
    constexpr uint32_t idesc = make_idesc_bf16(N_BLK, 1, 1);
    if (lane == 0) {
      int it = 0;
      for (int ti = 0; ti < ntile; ++ti) {
        const int bs = ti & 1;
        mbar_wait(&b_full[bs], ((uint32_t)(ti >> 1)) & 1u);
        tc_fence_after();
        const uint32_t b_addr = smem_u32(smem + L::B_OFF + bs * L::B_BYTES_);
        for (int pr = 0; pr < npairs; ++pr, ++it) {
          const int s = it % STAGES;
          mbar_wait(&a_full[s], ((uint32_t)(it / STAGES)) & 1u);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * L::A_STAGE);
#pragma unroll
          for (int k = 0; k < BM / 16; ++k) {      // 16 pixel rows per MMA = 2048 bytes down the tile
            const uint64_t da = make_desc_mnmajor_sw128(a_addr + k * 2048, ATOM_BYTES);
            const uint64_t db = make_desc_mnmajor_sw128(b_addr + k * 2048, ATOM_BYTES);
            umma_bf16(tmem_base + (uint32_t)(pr * N_BLK), da, db, idesc, (ti | k) != 0);
          }
          umma_commit(&a_empty[s]);
        }
        umma_commit(&b_empty[bs]);
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == PRODUCER_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

template <int N_BLK, int NPAIRS, int STAGES>
int launch(Params p, cudaStream_t st) {
  using L = Smem<N_BLK, NPAIRS, STAGES>;
  auto kern = wgrad_tc_kernel<N_BLK, NPAIRS, STAGES>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) return -3;
    configured = true;
  }
  const int total_pairs = (7 * (p.Cin / 64) + 1) / 2;
  p.units_m = (total_pairs + NPAIRS - 1) / NPAIRS;
  const int units = p.units_m * (p.Cout / N_BLK);
  int slices = (148 + units - 1) / units;
  if (slices > p.total_tiles) slices = p.total_tiles;
  if (slices < 1) slices = 1;
  p.tiles_per_cta = (p.total_tiles + slices - 1) / slices;
  dim3 grid((unsigned)((p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta), (unsigned)units);
  kern<<<grid, THREADS, L::TOTAL, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
}  // namespace tcw

inline bool tc_wgrad_supported(int Cin, int Cout) { return Cin % 64 == 0 && Cout % 64 == 0 && Cin >= 64 && Cout >= 64; }

inline int launch_wgrad_tc(const int32_t* plan_dev, const GinSide& side, int group, const float* X, const float* dY, float* dWp,
                           int B, int Cin, int Cout, int total_tiles, cudaStream_t st) {
  tcw::Params p;
  p.plan = plan_dev; p.side = side; p.group = group; p.B = B; p.Cin = Cin; p.Cout = Cout;
  p.X = X; p.dY = dY; p.dWp = dWp; p.total_tiles = total_tiles; p.tiles_per_cta = 1; p.units_m = 1;
  if ((long long)B * side.P_src >= 0x7fffffffLL) return -4;
  if (Cout % 256 == 0) return tcw::launch<256, 2, 2>(p, st);
  if (Cout % 128 == 0) return tcw::launch<128, 4, 3>(p, st);
  return tcw::launch<64, 7, 4>(p, st);
}

}  // namespace gin

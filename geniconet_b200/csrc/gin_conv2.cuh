// Patch-mode tcgen05 hex-conv, second generation (forward and the in-chart part of dgrad of stride-1 layers).
//
// What changed against gin_gemm_tcp.cuh, each item with the measurement that drove it (gpurun_out/dbg_matrix*.log,
// profiles/r01b_*):
//   * ONE shared-memory copy of the padded patch instead of three column-shifted ones.  tools/exp/umma_shift_test.cu
//     shows that the SWIZZLE_128B pattern of a UMMA operand is a pure function of the shared-memory ADDRESS (bits 4-6 ^=
//     bits 7-9): a descriptor may start at any 128-byte row and its 8-row groups may be any distance apart (SBO) as long
//     as the data was written with the same address-based XOR.  The patch is stored densely in plan order (row u of the
//     plan's source table = 128 bytes at u*128) and tap (di, dj) is the SAME image read from start row
//     (1+di)*Q*10 + (1+dj) with a group stride of 10 rows.  A stage is <= 30 KB instead of 72 KB.
//   * The MMA issuer loop runs WARP-UNIFORM (role index via shuffle, all 32 lanes in lock step, one elected lane issues):
//     descriptors and counters then live in uniform registers, which is what UTCHMMA takes.  With the loop under
//     `if (lane == 0)` every tcgen05.mma cost five R2UR conversions and the issuing thread, not the tensor pipe, set the pace
//     (~140 clk per MMA against 32 clk of tensor work at N = 64).
//   * Weights stay RESIDENT in shared memory when the [7][K/64] x N_TILE slice fits (<= 112 KB: the 64-wide layers);
//     otherwise they stream through a ring that holds >= 96 KB (2-4 tiles in flight could not cover the L2 latency).
//   * A dedicated TABLE warp resolves the plan's gather tables several tiles ahead into a shared-memory ring.  A global
//     load costs ~1 us on B200; with the table fetched one tile ahead by the producers themselves that latency was paid
//     once per tile (24 of 76 us).
//   * The epilogue transposes through shared memory so that every global store instruction writes whole 256-byte rows
//     (the direct TMEM-lane-per-row stores touched 32 different rows per instruction and bounded the kernel), and the
//     output row of a tile is computed from Q base pixels per tile (kept in shared memory for all tiles), not looked up.
//   * cp.async.cg: the patch is read once, and L1 is only ~20 KB next to 200 KB of shared memory.
//
// Warp roles (608 threads, one CTA per SM, persistent, every CTA keeps one n-block):
//   8 producer warps (cp.async patch gather) | 1 MMA warp | 1 weight warp (cp.async.bulk) | 1 table warp |
//   8 epilogue warps (tcgen05.ld -> smem transpose -> +bias -> coalesced fp32 stores).
// Two TMEM accumulators of N_TILE columns: the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "gin_gemm_tc.cuh"

namespace gin {
namespace cv2 {
using namespace tc;

constexpr int PROD_WARPS = 8, PROD_THREADS = PROD_WARPS * 32;
constexpr int EPI_WARPS = 8;                 // two per TMEM lane quarter: each takes every other 32-column slab
constexpr int W_MMA = PROD_WARPS, W_WEIGHT = PROD_WARPS + 1, W_TABLE = PROD_WARPS + 2, W_EPI0 = PROD_WARPS + 3;
constexpr int NWARPS = W_EPI0 + EPI_WARPS;
constexpr int NTHREADS = NWARPS * 32;
constexpr int MAX_ITEMS = 8;                 // patch rows per producer thread: ceil(256 / 32)
constexpr int MAX_A_STAGES = 4, MAX_B_STAGES = 8;
constexpr int TAB_SLOTS = 8, TAB_ROWS = 256, TAB_BATCH = 4;
constexpr int STAGE_PITCH = 144;             // epilogue staging: 32 fp32 + 16 bytes of padding per row (conflict-free both ways)
constexpr int STAGE_BYTES = EPI_WARPS * 32 * STAGE_PITCH;
constexpr int SMEM_LIMIT = 227 * 1024;

struct Params {
  const int32_t* plan;
  GinPSide ps;
  int group, B, K, N, P, W;   // P = pixels per sample (stride 1: same on both sides), W = 2n pixels per chart row
  const __nv_bfloat16* X;     // [B*P + 2B][K] bf16 (pixels, then pole-mean rows)
  const __nv_bfloat16* Wt;    // pre-swizzled bf16 tiles [7][K/64][N][64]
  const float* bias;
  float* Y;                   // [B*P][N] fp32
  int mirror;                 // 0 forward: tap (di,dj) reads cell (+di,+dj);  1 dgrad: reads (-di,-dj)
  int total_tiles, n_blocks;  // CTA c owns n-block c % n_blocks and tiles c / n_blocks + k * (gridDim.x / n_blocks)
  int a_stage_bytes, a_stages, b_stages, resident;
  int dbg;                    // GIN_DBG bit mask (experiments only): 1 no epilogue stores, 2 no patch loads, 4 no MMAs
};

__device__ __constant__ int8_t kDi[7] = {0, -1, 1, 0, 0, -1, 1};
__device__ __constant__ int8_t kDj[7] = {0, 0, 0, -1, 1, 1, -1};

GIN_DEVINL uint64_t desc_kmajor(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B, base offset 0 (address-based pattern)
  return d;
}

// barrier block: a_full[4] a_empty[4] b_full[8] b_empty[8] acc_full[2] acc_empty[2] tab_full[8] tab_empty[8] | tmem slot
constexpr int NBARS = 2 * MAX_A_STAGES + 2 * MAX_B_STAGES + 4 + 2 * TAB_SLOTS;
constexpr int BAR_BYTES = NBARS * 8 + 16;

template <int N_TILE, bool RESIDENT>
__global__ void __launch_bounds__(NTHREADS, 1) patch_conv_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int B_TILE = N_TILE * 128;
  const int kchunks = p.K / BK;
  const int b_tiles = RESIDENT ? 7 * kchunks : p.b_stages;
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + p.a_stages * p.a_stage_bytes;
  uint8_t* stage_smem = b_smem + (size_t)b_tiles * B_TILE;
  int32_t* tab = reinterpret_cast<int32_t*>(stage_smem + STAGE_BYTES);           // [TAB_SLOTS][TAB_ROWS] resolved source rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab + TAB_SLOTS * TAB_ROWS);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + MAX_A_STAGES;
  uint64_t* b_full = a_empty + MAX_A_STAGES;
  uint64_t* b_empty = b_full + MAX_B_STAGES;
  uint64_t* acc_full = b_empty + MAX_B_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* tab_full = acc_empty + 2;
  uint64_t* tab_empty = tab_full + TAB_SLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tab_empty + TAB_SLOTS);
  int32_t* tile_base = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES);   // [ntiles][Q] first pixel of each octet column

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-UNIFORM role index: keeps the issuer loops in uniform registers
  const int U = p.ps.U, Q = p.ps.Q;
  const int AS = p.a_stages, BS = p.b_stages;
  constexpr uint32_t TM_COLS = (2 * N_TILE <= 32) ? 32 : 2 * N_TILE;
  const int nb = blockIdx.x % p.n_blocks, n0 = nb * N_TILE;
  const int t_first = blockIdx.x / p.n_blocks, t_step = gridDim.x / p.n_blocks;

  if (warp == W_MMA) {
    if (lane == 0) {
      for (int s = 0; s < MAX_A_STAGES; ++s) { mbar_init(&a_full[s], PROD_THREADS); mbar_init(&a_empty[s], 1); }
      for (int s = 0; s < MAX_B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_WARPS * 32); }
      for (int s = 0; s < TAB_SLOTS; ++s) { mbar_init(&tab_full[s], 1); mbar_init(&tab_empty[s], PROD_THREADS); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TM_COLS);
  }
  for (int i = tid; i < p.ps.ntiles * Q; i += NTHREADS)
    tile_base[i] = __ldg(p.plan + p.ps.rows_off + (i / Q) * BM + (i % Q) * 8);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PROD_WARPS) {
    // =========================================================== producers: asynchronous patch gather (one copy)
    const int sub = lane >> 3, c8 = lane & 7;
    const __nv_bfloat16* __restrict__ Xc = p.X + c8 * 8;
    int s = 0, ts = 0;
    uint32_t ph = 0, tph = 0;
    for (int T = t_first; T < p.total_tiles; T += t_step) {
      int v[MAX_ITEMS];
      mbar_wait(&tab_full[ts], tph);
#pragma unroll
      for (int it = 0; it < MAX_ITEMS; ++it) v[it] = tab[ts * TAB_ROWS + it * 32 + warp * 4 + sub];
      mbar_arrive(&tab_empty[ts]);                   // the row ids are in registers now
      if (++ts == TAB_SLOTS) { ts = 0; tph ^= 1u; }
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&a_empty[s], ph ^ 1u);
        const uint32_t st = smem_u32(a_smem + s * p.a_stage_bytes);
#pragma unroll
        for (int it = 0; it < MAX_ITEMS; ++it) {
          const int u = it * 32 + warp * 4 + sub;
          if (u < U && !(p.dbg & 2)) {
            const bool ok = v[it] >= 0;
            cp_async16_cg(st + swz(u, c8), Xc + (size_t)(ok ? v[it] : 0) * p.K + kc * BK, ok);
          }
        }
        cp_async_arrive(&a_full[s]);
        if (++s == AS) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == W_MMA) {
    // =========================================================== MMA issuer
    // The whole warp runs this loop in lock step (addresses, descriptors and counters stay in uniform registers);
    // only the elected lane issues tcgen05.mma / tcgen05.commit.
    constexpr uint32_t idesc = make_idesc_bf16(N_TILE);
    const bool leader = elect_one();
    int s = 0, bs = 0;
    uint32_t ph = 0, bph = 0, wc = 0;
    uint32_t tap_off[7];
#pragma unroll
    for (int tap = 0; tap < 7; ++tap) {
      int di = kDi[tap], dj = kDj[tap];
      if (p.mirror) { di = -di; dj = -dj; }
      tap_off[tap] = (uint32_t)(((1 + di) * Q * 10 + (1 + dj)) * 128);
    }
    if (RESIDENT) mbar_wait(&b_full[0], 0);          // the whole weight slice, loaded once
    for (int T = t_first; T < p.total_tiles; T += t_step, ++wc) {
      const uint32_t ab = wc & 1;
      mbar_wait(&acc_empty[ab], ((wc >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * N_TILE;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&a_full[s], ph);
        fence_async_smem();                          // cp.async (generic proxy) writes -> visible to the async proxy
        tc_fence_after();
        const uint32_t a_addr = smem_u32(a_smem + s * p.a_stage_bytes);
#pragma unroll
        for (int tap = 0; tap < 7; ++tap) {
          uint32_t b_addr;
          if (RESIDENT) b_addr = smem_u32(b_smem + (size_t)(tap * kchunks + kc) * B_TILE);
          else {
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            b_addr = smem_u32(b_smem + (size_t)bs * B_TILE);
          }
          const uint64_t da = desc_kmajor(a_addr + tap_off[tap], 1280), db = desc_kmajor(b_addr, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            if (leader && !(p.dbg & 4)) umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kc | tap | k) != 0);
          if (!RESIDENT) {
            if (leader) umma_commit(&b_empty[bs]);
            if (++bs == BS) { bs = 0; bph ^= 1u; }
          }
        }
        if (leader) umma_commit(&a_empty[s]);
        __syncwarp();
        if (++s == AS) { s = 0; ph ^= 1u; }
      }
      if (leader) umma_commit(&acc_full[ab]);
      __syncwarp();
    }
  } else if (warp == W_WEIGHT) {
    // =========================================================== weight tiles
    if (lane == 0) {
      if (RESIDENT) {
        mbar_arrive_expect_tx(&b_full[0], (uint32_t)(7 * kchunks * B_TILE));
        for (int i = 0; i < 7 * kchunks; ++i)        // i = tap * kchunks + kc
          bulk_g2s(b_smem + (size_t)i * B_TILE, p.Wt + ((size_t)i * p.N + n0) * BK, B_TILE, &b_full[0]);
      } else {
        int bs = 0;
        uint32_t bph = 0;
        for (int T = t_first; T < p.total_tiles; T += t_step)
          for (int kc = 0; kc < kchunks; ++kc)
            for (int tap = 0; tap < 7; ++tap) {
              mbar_wait(&b_empty[bs], bph ^ 1u);
              mbar_arrive_expect_tx(&b_full[bs], B_TILE);
              bulk_g2s(b_smem + (size_t)bs * B_TILE, p.Wt + (((size_t)tap * kchunks + kc) * p.N + n0) * BK, B_TILE, &b_full[bs]);
              if (++bs == BS) { bs = 0; bph ^= 1u; }
            }
      }
    }
    __syncwarp();
  } else if (warp == W_TABLE) {
    // =========================================================== gather tables, resolved TAB_BATCH tiles at a time, up to 8 ahead
    const long long total_pix = (long long)p.B * p.P;
    int ts = 0;
    uint32_t tph = 0;
    for (int T0 = t_first; T0 < p.total_tiles; T0 += TAB_BATCH * t_step) {
      int code[TAB_BATCH][MAX_ITEMS];
#pragma unroll
      for (int j = 0; j < TAB_BATCH; ++j) {
        const int T = T0 + j * t_step;
        if (T < p.total_tiles) {
          const int32_t* __restrict__ src_tab = p.plan + p.ps.src_off + (size_t)(T % p.ps.ntiles) * U;
#pragma unroll
          for (int it = 0; it < MAX_ITEMS; ++it) {
            const int u = it * 32 + lane;
            code[j][it] = (u < U) ? __ldg(src_tab + u) : GIN_SRC_ZERO;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < TAB_BATCH; ++j) {
        const int T = T0 + j * t_step;
        if (T < p.total_tiles) {
          const int G = T / p.ps.ntiles;
          const long long base = (long long)G * p.group * p.P;
          mbar_wait(&tab_empty[ts], tph ^ 1u);
#pragma unroll
          for (int it = 0; it < MAX_ITEMS; ++it)
            tab[ts * TAB_ROWS + it * 32 + lane] = resolve_row(code[j][it], base, total_pix, G * p.group, p.B);
          __syncwarp();
          if (lane == 0) mbar_arrive(&tab_full[ts]);
          if (++ts == TAB_SLOTS) { ts = 0; tph ^= 1u; }
        }
      }
    }
  } else {
    // =========================================================== epilogue: TMEM -> smem transpose -> coalesced global stores
    const int e = warp - W_EPI0;
    const int q = warp & 3;                           // TMEM lane quarter this warp may touch
    const int hslab = e >> 2;                         // which of the two warps of this quarter: takes 32-column slabs hslab, hslab+2, ...
    const int row = q * 32 + lane;                    // tile row of this thread: group g = row / 8 = r * Q + oq, pixel px = row % 8
    const int g = row >> 3, r_in = g / Q, oq = g - r_in * Q;
    const int row_off = r_in * p.W + (row & 7);
    uint8_t* my_stage = stage_smem + (size_t)e * 32 * STAGE_PITCH;
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;  // store phase: one instruction = four whole 128-byte row segments
    const long long total_pix = (long long)p.B * p.P;
    uint32_t wc = 0;
    for (int T = t_first; T < p.total_tiles; T += t_step, ++wc) {
      const int G = T / p.ps.ntiles, t = T - G * p.ps.ntiles;
      const long long gdl = (long long)G * p.group * p.P + tile_base[t * Q + oq] + row_off;
      const int gd = gdl < total_pix ? (int)gdl : -1; // B*P < 2^31 is checked by the launcher
      const uint32_t ab = wc & 1;
      mbar_wait(&acc_full[ab], (wc >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int slab = hslab * 32; slab < N_TILE; slab += 64) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * N_TILE + slab), v);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(my_stage + lane * STAGE_PITCH);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        if (slab + 64 >= N_TILE) {                    // this warp has read its last columns of the accumulator
          tc_fence_before();
          mbar_arrive(&acc_empty[ab]);
        }
        __syncwarp();
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + slab + c4));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r2 = 4 * i + rsub;
          const int gd2 = __shfl_sync(0xffffffffu, gd, r2);
          float4 o = *reinterpret_cast<const float4*>(my_stage + r2 * STAGE_PITCH + c4 * 4);
          o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          if (gd2 >= 0 && !(p.dbg & 1)) *reinterpret_cast<float4*>(p.Y + (size_t)gd2 * p.N + n0 + slab + c4) = o;
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

// shared-memory plan for one launch; returns false when nothing fits
inline bool plan_smem(int n_tile, int K, const GinPSide& ps, Params& p, int& total) {
  const int b_tile = n_tile * 128, kchunks = K / 64;
  p.a_stage_bytes = ((ps.U * 128 + 1023) / 1024) * 1024;
  const int fixed = STAGE_BYTES + TAB_SLOTS * TAB_ROWS * 4 + BAR_BYTES + ps.ntiles * ps.Q * 4 + 16;
  const int budget = SMEM_LIMIT - 1024 - fixed;
  const int resident_bytes = 7 * kchunks * b_tile;
  p.resident = resident_bytes <= 114688 && resident_bytes + 2 * p.a_stage_bytes <= budget;
  int b_bytes;
  if (p.resident) { p.b_stages = 1; b_bytes = resident_bytes; }
  else {
    p.b_stages = 98304 / b_tile;                    // >= 96 KB of weights in flight
    if (p.b_stages > MAX_B_STAGES) p.b_stages = MAX_B_STAGES;
    if (p.b_stages < 2) p.b_stages = 2;
    b_bytes = p.b_stages * b_tile;
  }
  p.a_stages = (budget - b_bytes) / p.a_stage_bytes;
  if (p.a_stages > MAX_A_STAGES) p.a_stages = MAX_A_STAGES;
  if (p.a_stages < 2) return false;
  total = p.a_stages * p.a_stage_bytes + b_bytes + fixed + 1024;
  return true;
}

template <int N_TILE>
int launch(Params p, cudaStream_t st) {
  int smem_total = 0;
  if (!plan_smem(N_TILE, p.K, p.ps, p, smem_total)) return -4;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(patch_conv_kernel<N_TILE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess ||
        cudaFuncSetAttribute(patch_conv_kernel<N_TILE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -3;
    configured = true;
  }
  p.n_blocks = p.N / N_TILE;
  const long long items = (long long)p.total_tiles * p.n_blocks;
  int grid = (int)(items < 148 ? items : 148);
  grid -= grid % p.n_blocks;                         // every CTA keeps one n-block
  if (grid < p.n_blocks) grid = p.n_blocks;
  if (p.resident) patch_conv_kernel<N_TILE, true><<<grid, NTHREADS, smem_total, st>>>(p);
  else patch_conv_kernel<N_TILE, false><<<grid, NTHREADS, smem_total, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace cv2

inline bool cv2_supported(const GinPSide& ps, int K, int N) {
  cv2::Params tmp;
  int total;
  return ps.ntiles > 0 && ps.U <= cv2::TAB_ROWS && tc_supported(K, N) && cv2::plan_smem(64, K, ps, tmp, total);
}

// N tile: minimise (rounds of 148 CTAs) x (MMA time per item); a 64-wide tile is shared-memory-bandwidth bound (A 128 B/clk +
// B 64 B/clk against 128 B/clk), hence the 1.5 penalty
inline int cv2_pick_ntile(long long tiles, int N) {
  int best = 0;
  double best_cost = 0;
  for (int nt : {256, 128, 64}) {
    if (N % nt) continue;
    const long long items = tiles * (N / nt);
    int ctas = (int)(items < 148 ? items : 148);
    ctas -= ctas % (N / nt);
    if (ctas < N / nt) ctas = N / nt;
    const long long rounds = (items + ctas - 1) / ctas;
    const double cost = (double)rounds * nt * (nt == 64 ? 1.5 : 1.0);
    if (!best || cost < best_cost * 0.999) { best = nt; best_cost = cost; }
  }
  return best;
}

inline int launch_patch_conv2(const int32_t* plan_dev, const GinPSide& ps, int group, int P, int W, const void* Xb, const void* Wb,
                              const float* bias, float* Y, int B, int K, int N, int mirror, cudaStream_t st) {
  cv2::Params p;
  p.plan = plan_dev; p.ps = ps; p.group = group; p.B = B; p.K = K; p.N = N; p.P = P; p.W = W;
  p.X = reinterpret_cast<const __nv_bfloat16*>(Xb); p.Wt = reinterpret_cast<const __nv_bfloat16*>(Wb); p.bias = bias; p.Y = Y; p.mirror = mirror;
  if ((long long)B * P + 2LL * B >= 0x7fffffffLL) return -4;
  const int groups = (B + group - 1) / group;
  p.total_tiles = groups * ps.ntiles;
  { const char* e = getenv("GIN_DBG"); p.dbg = e ? atoi(e) : 0; }
  switch (cv2_pick_ntile(p.total_tiles, N)) {
    case 256: return cv2::launch<256>(p, st);
    case 128: return cv2::launch<128>(p, st);
    default: return cv2::launch<64>(p, st);
  }
}

}  // namespace gin

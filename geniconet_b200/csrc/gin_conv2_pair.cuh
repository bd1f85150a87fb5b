// Tile-PAIR variant of the patch conv kernel (gin_conv2.cuh) for layers whose weights do not fit in shared memory.
//
// Measured on the single-tile kernel (profiles/r01_bottleneck_matrix_v2_kernel.log, fwd 256->128 @ I4): with patch loads, MMAs
// and stores all switched off the kernel still takes 31 of its 40 us -- the time to stream every tap's weight tile from L2 once
// per 128-row tile (322 MB per launch, ~10 TB/s).  Here a CTA works on TWO 128-row tiles at a time and applies every weight
// tile to both before releasing it: half the weight stream, half the weight barriers, eight MMAs per barrier instead of four.
// Everything else is the single-tile kernel: one shared-memory image per (tile, 64-channel chunk), taps = descriptor start rows,
// warp-uniform issue, table warp, transposed epilogue, optional BatchNorm statistics.  TMEM: 2 buffers x 2 tiles x N_TILE columns
// (N_TILE <= 128).  Segments (`nplanes`) accumulate into one output per tile (flush_each is not supported here).
#pragma once
#include "gin_conv2.cuh"

namespace gin {
namespace cv2 {

template <int N_TILE, bool STATS>
__global__ void __launch_bounds__(NTHREADS, 1) patch_conv_pair_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int B_TILE = N_TILE * 128;
  const int kchunks = p.K / BK;
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + p.a_stages * p.a_stage_bytes;
  uint8_t* stage_smem = b_smem + (size_t)p.b_stages * B_TILE;
  int32_t* tab = reinterpret_cast<int32_t*>(stage_smem + STAGE_BYTES);
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
  int32_t* tile_base = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int U = p.U, Q = p.Q, NP = p.nplanes;
  const int AS = p.a_stages, BS = p.b_stages;      // AS is even
  constexpr uint32_t TM_COLS = 4 * N_TILE;
  const int nb = blockIdx.x % p.n_blocks, n0 = nb * N_TILE;
  const int npairs = (p.total_tiles + 1) / 2;
  const int q_first = blockIdx.x / p.n_blocks, q_step = gridDim.x / p.n_blocks;   // this CTA's pairs: q_first, q_first + q_step, ...

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
  for (int i = tid; i < p.ntiles * Q; i += NTHREADS)
    tile_base[i] = __ldg(p.plan + p.base_off + (i / Q) * p.base_tstride + (i % Q) * p.base_qstride);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  GIN_PDL_SYNC();

  if (warp < PROD_WARPS) {
    // =========================================================== producers: per (pair, segment): tables of both tiles, then per chunk one
    // stage per tile (stage s: tile 0, stage s+1: tile 1)
    const int sub = lane >> 3, c8 = lane & 7;
    const __nv_bfloat16* __restrict__ Xc = p.X + c8 * 8;
    int s = 0, ts = 0;
    uint32_t ph = 0, tph = 0;
    for (int pq = q_first; pq < npairs; pq += q_step)
      for (int pl = 0; pl < NP; ++pl) {
        int v[2][MAX_ITEMS];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&tab_full[ts], tph);
#pragma unroll
          for (int it = 0; it < MAX_ITEMS; ++it) v[h][it] = tab[ts * TAB_ROWS + it * 32 + warp * 4 + sub];
          mbar_arrive(&tab_empty[ts]);
          if (++ts == TAB_SLOTS) { ts = 0; tph ^= 1u; }
        }
        for (int kc = 0; kc < kchunks; ++kc)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&a_empty[s], ph ^ 1u);
            const uint32_t st = smem_u32(a_smem + s * p.a_stage_bytes);
#pragma unroll
            for (int it = 0; it < MAX_ITEMS; ++it) {
              const int u = it * 32 + warp * 4 + sub;
              if (u < U) {
                const bool ok = v[h][it] >= 0;
                cp_async16_cg(st + swz(u, c8), Xc + (size_t)(ok ? v[h][it] : 0) * p.K + kc * BK, ok);
              }
            }
            cp_async_arrive(&a_full[s]);
            if (++s == AS) { s = 0; ph ^= 1u; }
          }
      }
  } else if (warp == W_MMA) {
    // =========================================================== MMA issuer: every weight tile is applied to both tiles of the pair
    const uint32_t idesc = make_idesc_f16kind(N_TILE) | p.fmt;
    const bool leader = elect_one();
    int s = 0, bs = 0;
    uint32_t ph = 0, bph = 0, wc = 0;
    for (int pq = q_first; pq < npairs; pq += q_step, ++wc) {
      const uint32_t ab = wc & 1;
      mbar_wait(&acc_empty[ab], ((wc >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + ab * (2 * N_TILE), d1 = d0 + N_TILE;
      uint32_t fresh = 1;
      for (int pl = 0; pl < NP; ++pl) {
        const int nt = p.ntaps[pl];
        for (int kc = 0; kc < kchunks; ++kc) {
          const int s1 = s + 1;
          mbar_wait(&a_full[s], ph);
          mbar_wait(&a_full[s1], ph);
          fence_async_smem();
          tc_fence_after();
          const uint32_t a0 = smem_u32(a_smem + s * p.a_stage_bytes), a1 = smem_u32(a_smem + s1 * p.a_stage_bytes);
          for (int j = 0; j < nt; ++j) {
            const int tap = p.tap_id[pl][j];
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(b_smem + (size_t)bs * B_TILE);
            const uint32_t row = (uint32_t)p.tap_row[pl][j] * 128u;
            const uint64_t da0 = desc_kmajor(a0 + row, (uint32_t)p.group_bytes), da1 = desc_kmajor(a1 + row, (uint32_t)p.group_bytes);
            const uint64_t db = desc_kmajor(b_addr, 1024);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              if (leader) umma_bf16(d0, da0 + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (fresh && k == 0) ? 0u : 1u);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              if (leader) umma_bf16(d1, da1 + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (fresh && k == 0) ? 0u : 1u);
            fresh = 0;
            if (leader) umma_commit(&b_empty[bs]);
            if (++bs == BS) { bs = 0; bph ^= 1u; }
            (void)tap;
          }
          if (leader) { umma_commit(&a_empty[s]); umma_commit(&a_empty[s1]); }
          __syncwarp();
          s += 2;
          if (s == AS) { s = 0; ph ^= 1u; }
        }
      }
      if (leader) umma_commit(&acc_full[ab]);
      __syncwarp();
    }
  } else if (warp == W_WEIGHT) {
    // =========================================================== weight tiles: once per PAIR
    if (lane == 0) {
      int bs = 0;
      uint32_t bph = 0;
      for (int pq = q_first; pq < npairs; pq += q_step)
        for (int pl = 0; pl < NP; ++pl)
          for (int kc = 0; kc < kchunks; ++kc)
            for (int j = 0; j < p.ntaps[pl]; ++j) {
              const int tap = p.tap_id[pl][j];
              mbar_wait(&b_empty[bs], bph ^ 1u);
              mbar_arrive_expect_tx(&b_full[bs], B_TILE);
              bulk_g2s(b_smem + (size_t)bs * B_TILE, p.Wt + (((size_t)tap * kchunks + kc) * p.N + n0) * BK, B_TILE, &b_full[bs]);
              if (++bs == BS) { bs = 0; bph ^= 1u; }
            }
    }
    __syncwarp();
  } else if (warp == W_TABLE) {
    // =========================================================== gather tables in the producers' order: (pair, segment, tile of the pair)
    const long long total_src = (long long)p.B * p.P_src;
    int ts = 0;
    uint32_t tph = 0;
    int pq = q_first, pl = 0, h = 0;
    auto advance = [&]() { if (++h == 2) { h = 0; if (++pl == NP) { pl = 0; pq += q_step; } } };
    while (pq < npairs) {
      int code[TAB_BATCH][MAX_ITEMS], tile_of[TAB_BATCH];
#pragma unroll
      for (int j = 0; j < TAB_BATCH; ++j) {
        tile_of[j] = -2;                                    // -2: past the end of this CTA's work
        if (pq < npairs) {
          const int T = 2 * pq + h;
          tile_of[j] = T < p.total_tiles ? T : -1;          // -1: the ghost half of an odd last pair (all rows zero)
          if (T < p.total_tiles) {
            const int32_t* __restrict__ src_tab = p.plan + p.tab_off + (size_t)(T % p.ntiles) * p.tab_tstride + (size_t)pl * p.tab_pstride;
#pragma unroll
            for (int it = 0; it < MAX_ITEMS; ++it) {
              const int u = it * 32 + lane;
              code[j][it] = (u < U) ? __ldg(src_tab + u) : GIN_SRC_ZERO;
            }
          }
          advance();
        }
      }
#pragma unroll
      for (int j = 0; j < TAB_BATCH; ++j) {
        if (tile_of[j] == -2) continue;
        const int T = tile_of[j];
        const int G = T >= 0 ? T / p.ntiles : 0;
        const long long base = (long long)G * p.group * p.P_src;
        mbar_wait(&tab_empty[ts], tph ^ 1u);
#pragma unroll
        for (int it = 0; it < MAX_ITEMS; ++it)
          tab[ts * TAB_ROWS + it * 32 + lane] = T >= 0 ? resolve_row(code[j][it], base, total_src, G * p.group, p.B) : -1;
        __syncwarp();
        if (lane == 0) mbar_arrive(&tab_full[ts]);
        if (++ts == TAB_SLOTS) { ts = 0; tph ^= 1u; }
      }
    }
  } else {
    // =========================================================== epilogue: both accumulators of a pair, then hand the buffer back
    const int e = warp - W_EPI0;
    const int q = warp & 3;
    const int hslab = e >> 2;
    const int row = q * 32 + lane;
    const int g = row >> 3, r_in = g / Q, oq = g - r_in * Q;
    const int row_off = r_in * p.dst_row_stride + (row & 7) * p.dst_px_stride;
    uint8_t* my_stage = stage_smem + (size_t)e * 32 * STAGE_PITCH;
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    const long long total_pix = (long long)p.B * p.P_dst;
    constexpr int NS = N_TILE / 64;
    float ssum[STATS ? NS : 1][4], ssq[STATS ? NS : 1][4];
#pragma unroll
    for (int si = 0; si < (STATS ? NS : 1); ++si)
#pragma unroll
      for (int k = 0; k < 4; ++k) { ssum[si][k] = 0.f; ssq[si][k] = 0.f; }
    uint32_t wc = 0;
    for (int pq = q_first; pq < npairs; pq += q_step, ++wc) {
      const uint32_t ab = wc & 1;
      mbar_wait(&acc_full[ab], (wc >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int T = 2 * pq + h;
        int gd = -1;
        if (T < p.total_tiles) {
          const int G = T / p.ntiles, t = T - G * p.ntiles;
          const long long gdl = (long long)G * p.group * p.P_dst + tile_base[t * Q + oq] + row_off;
          gd = gdl < total_pix ? (int)gdl : -1;
        }
#pragma unroll
        for (int si = 0; si < NS; ++si) {
          const int slab = hslab * 32 + si * 64;
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 2 * N_TILE + h * N_TILE + slab), v);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(my_stage + lane * STAGE_PITCH);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          if (h == 1 && si == NS - 1) {                   // this warp has read its last columns of both accumulators
            tc_fence_before();
            mbar_arrive(&acc_empty[ab]);
          }
          __syncwarp();
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) {
          const int col = n0 + slab + c4;
          bb = __ldg(reinterpret_cast<const float4*>((p.bias1 && col >= p.bias_split) ? p.bias1 + (col - p.bias_split) : p.bias + col));
        }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r2 = 4 * i + rsub;
            const int gd2 = __shfl_sync(0xffffffffu, gd, r2);
            float4 o = *reinterpret_cast<const float4*>(my_stage + r2 * STAGE_PITCH + c4 * 4);
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            if (gd2 >= 0) {
              if (p.y_f16) *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.Y) + (size_t)gd2 * p.N + n0 + slab + c4) = make_uint2(pack2_f16(o.x, o.y), pack2_f16(o.z, o.w));
              else *reinterpret_cast<float4*>(p.Y + (size_t)gd2 * p.N + n0 + slab + c4) = o;
              if (STATS) {
                ssum[si][0] += o.x; ssum[si][1] += o.y; ssum[si][2] += o.z; ssum[si][3] += o.w;
                ssq[si][0] = fmaf(o.x, o.x, ssq[si][0]); ssq[si][1] = fmaf(o.y, o.y, ssq[si][1]);
                ssq[si][2] = fmaf(o.z, o.z, ssq[si][2]); ssq[si][3] = fmaf(o.w, o.w, ssq[si][3]);
              }
            }
          }
          __syncwarp();
        }
      }
    }
    if (STATS) {
      float* sred = reinterpret_cast<float*>(stage_smem);
      // one slot per (TMEM quarter, column): every slot is written exactly once, the four quarters are then added in a fixed
      // order -- no atomics, so the statistics (and everything downstream of them) are reproducible run to run
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
#pragma unroll
      for (int si = 0; si < NS; ++si)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float a = ssum[si][k], b = ssq[si][k];
          a += __shfl_xor_sync(0xffffffffu, a, 8); a += __shfl_xor_sync(0xffffffffu, a, 16);
          b += __shfl_xor_sync(0xffffffffu, b, 8); b += __shfl_xor_sync(0xffffffffu, b, 16);
          if (lane < 8) {
            const int col = hslab * 32 + si * 64 + c4 + k;
            sred[(q * 2) * N_TILE + col] = a;
            sred[(q * 2 + 1) * N_TILE + col] = b;
          }
        }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      float* out = p.stats + (size_t)(blockIdx.x / p.n_blocks) * 2 * p.N + n0;
      for (int i = e * 32 + lane; i < 2 * N_TILE; i += EPI_WARPS * 32)
        out[(i / N_TILE) * p.N + (i % N_TILE)] = (sred[i] + sred[2 * N_TILE + i]) + (sred[4 * N_TILE + i] + sred[6 * N_TILE + i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

// pair mode: streamed weights, arithmetic destinations, plain stores, segments accumulating into one output, N_TILE <= 128
template <int N_TILE>
int launch_pair(Params p, cudaStream_t st) {
  // shared memory: 4 patch stages (two chunk pairs in flight) + a 64 KB weight ring (a tile now feeds 8 MMAs) + the fixed part
  p.a_stage_bytes = ((p.U * 128 + 1023) / 1024) * 1024;
  p.a_stages = 4;
  p.b_stages = 65536 / (N_TILE * 128);
  p.resident = 0;
  const int fixed = STAGE_BYTES + TAB_SLOTS * TAB_ROWS * 4 + BAR_BYTES + p.ntiles * p.Q * 4 + 16;
  const int smem_total = p.a_stages * p.a_stage_bytes + p.b_stages * N_TILE * 128 + fixed + 1024;
  if (smem_total > SMEM_LIMIT) return -4;
  static PerDeviceFlag configured_on;
  bool& configured = configured_on.here();
  if (!configured) {
    if (cudaFuncSetAttribute(patch_conv_pair_kernel<N_TILE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess ||
        cudaFuncSetAttribute(patch_conv_pair_kernel<N_TILE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -3;
    configured = true;
  }
  p.n_blocks = p.N / N_TILE;
  const long long items = (long long)((p.total_tiles + 1) / 2) * p.n_blocks;
  const int sms = num_sms();
  int grid = (int)(items < sms ? items : sms);
  grid -= grid % p.n_blocks;
  if (grid < p.n_blocks) grid = p.n_blocks;
  const bool stats = p.stats != nullptr;
  if (p.stats_parts) *p.stats_parts = stats ? grid / p.n_blocks : 0;
  if (stats) launch_pdl(patch_conv_pair_kernel<N_TILE, true>, dim3(grid), dim3(NTHREADS), smem_total, st, p);
  else launch_pdl(patch_conv_pair_kernel<N_TILE, false>, dim3(grid), dim3(NTHREADS), smem_total, st, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace cv2
}  // namespace gin

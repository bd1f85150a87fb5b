// Experiment for round 2: what is the per-SM ceiling of SS-mode tcgen05 MMAs on the conv kernels' tile shapes when NOTHING else
// runs (operands resident in shared memory, no gathers, no epilogue), single CTA (M = 128, N = 64 / 128 / 256) against a CTA pair
// (cta_group::2, M = 256, each CTA holding its own 128 A rows and HALF of the B columns)?  profiles/r01_ncu_fused_problems.txt
// shows the tensor pipe 59 % busy with the tensor-core shared-memory reads at 44 % + LSU 19 %: if the single-CTA loop below
// already reaches ~100 % of the burst peak, shared-memory operand bandwidth is NOT the limiter and cta_group::2 is not the
// next step; if it saturates near 60 %, it is.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I geniconet_b200/csrc -o tools/exp/umma_peak_test tools/exp/umma_peak_test.cu
//   run:   timeout 60 tools/exp/umma_peak_test            (mode 2 = CTA pair is UNTESTED code: keep the timeout)
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "gin_gemm_tc.cuh"

using namespace gin::tc;

constexpr int PK_A_BYTES = 128 * 128;            // one A k-chunk: 128 rows x 64 bf16
constexpr int TAPS = 7;                       // distinct B tiles cycled through, like the 7 taps of a conv k-chunk

GIN_DEVINL uint32_t idesc_m(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

GIN_DEVINL void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
GIN_DEVINL void umma_commit_2cta(uint64_t* bar) {       // arrives on the barrier at the same offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
GIN_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
GIN_DEVINL uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

// One CTA (or CTA pair) per SM (pair of SMs); thread 0 of the leader issues `iters` rounds of TAPS x 4 MMAs (K = 64 per tap),
// alternating between two accumulators, committing every round; everything else idles.  N_CTA = columns of B held by ONE CTA.
template <int N, bool PAIR>
__global__ void __launch_bounds__(128) peak_kernel(int iters, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int N_CTA = PAIR ? N / 2 : N;
  constexpr int B_BYTES = N_CTA * 128;
  uint8_t* A = smem;                                  // 128 x 64 bf16, SWIZZLE_128B K-major
  uint8_t* Bt = smem + PK_A_BYTES;                       // TAPS tiles of N_CTA x 64 bf16
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < (PK_A_BYTES + TAPS * B_BYTES) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  __syncthreads();
  constexpr uint32_t COLS = 2 * N <= 32 ? 32 : 2 * N;  // two accumulators of N fp32 columns (per CTA: 128 lanes x N columns)
  if (tid < 32) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else tmem_alloc(&tmem_slot, COLS);
  }
  fence_async_smem();
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const bool leader = !PAIR || cluster_rank() == 0;
  long long t0 = 0, t1 = 0;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);            // warp-uniform: the issue loop stays in uniform registers
  if (warp == 0 && leader) {
    const uint32_t idesc = idesc_m(PAIR ? 256 : 128, N);
    const uint32_t a0 = smem_u32(A), b0 = smem_u32(Bt);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t acc = tmem + (uint32_t)((it & 1) * N);
      if (it >= 2) mbar_wait(&bar[it & 1], ((it >> 1) - 1) & 1u);     // the accumulator's previous round has completed
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < TAPS; ++t)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_desc_kmajor_sw128(a0 + k * 32), db = make_desc_kmajor_sw128(b0 + t * B_BYTES + k * 32);
            if (PAIR) umma_bf16_2cta(acc, da, db, idesc, (t | k) != 0);
            else umma_bf16(acc, da, db, idesc, (t | k) != 0);
          }
        if (PAIR) umma_commit_2cta(&bar[it & 1]); else umma_commit(&bar[it & 1]);
      }
      __syncwarp();
    }
    for (int it = iters - 2 > 0 ? iters - 2 : 0; it < iters; ++it) mbar_wait(&bar[it & 1], (it >> 1) & 1u);
    t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) *cycles = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (tid < 32) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(COLS) : "memory");
    else tmem_dealloc(tmem, COLS);
  }
}

template <int N, bool PAIR>
static void run(int iters) {
  constexpr int N_CTA = PAIR ? N / 2 : N;
  const size_t smem = PK_A_BYTES + TAPS * N_CTA * 128 + 1024;
  auto k = peak_kernel<N, PAIR>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  unsigned long long* cyc;
  cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = PAIR ? (sms / 2) * 2 : sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  if (PAIR) { attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1; cfg.attrs = attr; cfg.numAttrs = 1; }
  for (int rep = 0; rep < 2; ++rep) {                 // first launch warms up
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, k, iters, cyc);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { printf("N=%d pair=%d: launch failed: %s\n", N, (int)PAIR, cudaGetErrorString(cudaGetLastError())); return; }
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long c = 0;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double units = PAIR ? grid / 2 : grid;        // issuing CTAs
  const double flop = units * (double)iters * TAPS * 4 * 2.0 * (PAIR ? 256 : 128) * N * 16;
  printf("M=%3d N=%3d %-8s grid %3d: %8.3f ms  %7.1f TFLOP/s  (%.1f cycles per MMA on CTA 0; smem operand bytes per MMA per SM: %d)\n", PAIR ? 256 : 128, N,
         PAIR ? "CTA pair" : "1 CTA", grid, ms, flop / (ms * 1e-3) / 1e12, (double)c / ((double)iters * TAPS * 4), 128 * 32 + N_CTA * 32);
  cudaFree(cyc);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 4000;
  const int mode = argc > 2 ? atoi(argv[2]) : 3;      // bit 0: single CTA, bit 1: CTA pair
  if (mode & 1) { run<64, false>(iters); run<128, false>(iters); run<256, false>(iters); }
  if (mode & 2) { run<128, true>(iters); run<256, true>(iters); }
  return 0;
}

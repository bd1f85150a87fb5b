// Experiment: can a SWIZZLE_128B K-major UMMA operand start at an arbitrary 128-byte ROW offset (not 1024-aligned),
// and can its 8-row groups be spaced by an SBO that is not a multiple of 1024?  Decides whether the conv kernels can keep ONE
// copy of the padded patch in shared memory and express the (di, dj) taps purely as descriptor start addresses.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I geniconet_b200/csrc -o tools/exp/umma_shift_test tools/exp/umma_shift_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "gin_gemm_tc.cuh"

using namespace gin::tc;

constexpr int NPIX = 320;     // pixel rows of 128 bytes in the A image

__device__ __host__ inline float aval(int p, int k) { return (float)(((p * 7 + k * 3) % 13) - 6); }

GIN_DEVINL uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

// mode 0: K-major (rows = M, 64 K per 128-byte row).  mode 1: MN-major (rows = K, 64 M per row), A^T semantic:
//   D[m][n] = sum_k A[(shift + k-row)][m] * B[k][n]
__global__ void __launch_bounds__(128) test_kernel(float* out, int shift, int pitch, int use_base_off, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* A = smem;                         // NPIX * 128 bytes
  uint8_t* Bt = smem + NPIX * 128;           // B: 64 rows x 128 bytes (1024-aligned since NPIX*128 % 1024 == 0)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < NPIX * 64; i += 128) {
    const int p = i >> 6, k = i & 63;
    const uint32_t off = p * 128 + ((((k >> 3) ^ (p & 7)) << 4) | ((k & 7) << 1));
    *reinterpret_cast<__nv_bfloat16*>(A + off) = __float2bfloat16(aval(p, k));
  }
  for (int i = tid; i < 64 * 64; i += 128) {   // identity, K-major rows (n) x k
    const int n = i >> 6, k = i & 63;
    const uint32_t off = n * 128 + ((((k >> 3) ^ (n & 7)) << 4) | ((k & 7) << 1));
    *reinterpret_cast<__nv_bfloat16*>(Bt + off) = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (tid < 32) tmem_alloc(&tmem_slot, 64);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(A) + shift * 128, b0 = smem_u32(Bt);
    const uint32_t sbo = pitch * 128;
    const uint32_t bo = use_base_off ? ((a0 >> 7) & 7) : 0;
    if (mode == 0) {
      const uint32_t idesc = make_idesc_bf16(64);
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, desc_sw128(a0 + k * 32, sbo, bo), desc_sw128(b0 + k * 32, 1024, 0), idesc, k != 0);
    } else {
      // A MN-major: M = 64 channels of a row... M must be 128: two atoms -> second atom = the same rows shifted by `pitch` rows (LBO)
      const uint32_t idesc = make_idesc_bf16(64, 1, 0);
      for (int k = 0; k < 4; ++k) {     // K = 64 pixel rows: 4 MMAs of 16 rows (2048 bytes)
        uint64_t da = desc_sw128(a0 + k * 2048, 1024, use_base_off ? (((a0 + k * 2048) >> 7) & 7) : 0);
        da |= (uint64_t)((pitch * 128) >> 4) << 16;     // LBO: distance between the two 64-wide M atoms
        umma_bf16(tmem, da, desc_sw128(b0 + k * 32, 1024, 0), idesc, k != 0);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  for (int cb = 0; cb < 64; cb += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cb, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + cb + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 64 * 4);
  const int smem = NPIX * 128 + 64 * 128 + 2048;
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> h(128 * 64);
  const int shifts[] = {0, 1, 2, 3, 7, 9, 16};
  const int pitches[] = {8, 10, 16, 20};
  for (int mode = 0; mode < 2; ++mode)
    for (int ubo = 0; ubo < 2; ++ubo)
      for (int pitch : pitches)
        for (int shift : shifts) {
          test_kernel<<<1, 128, smem>>>(d_out, shift, pitch, ubo, mode);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("mode %d baseoff %d pitch %d shift %d: CUDA error %s\n", mode, ubo, pitch, shift, cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
          int bad = 0, first = -1;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
              float want;
              if (mode == 0) want = aval(shift + (m / 8) * pitch + (m % 8), n);          // row m of A, column n (B = identity)
              else { /* D[m][n] = A[(shift + n)-th k row][channel m%64 of atom m/64] */ want = aval(shift + n + (m / 64) * pitch, m % 64); }
              if (h[m * 64 + n] != want) { ++bad; if (first < 0) first = m * 64 + n; }
            }
          printf("mode %s base_off %d pitch %2d shift %2d : %s", mode ? "MN" : "K ", ubo, pitch, shift, bad ? "FAIL" : "ok");
          if (bad) printf(" (%d bad, first at m=%d n=%d got %g)", bad, first / 64, first % 64, h[first]);
          printf("\n");
        }
  return 0;
}

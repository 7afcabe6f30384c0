// Can a K-major SWIZZLE_128B UMMA A-operand start at an ARBITRARY 128-byte row of a TMA-loaded tile (not at a 1024-byte
// swizzle atom)?  That is what a 3x3 convolution needs to read its nine taps as shifted views of ONE halo tile in shared
// memory instead of nine im2col gathers.  A_src[r][c] = (c == r % 64) ? r + 1 : 0 (256 rows x 64 bf16), B = identity:
// D[m][n] = A_view[m][n], so row m of the result must hold (m + d + 1) in column (m + d) % 64 for a view shifted by d rows.
// Tried with the descriptor's base-offset field = 0 and = (start address >> 7) & 7.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../mvfnet_b200/csrc -o umma_shift umma_shift.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"
using namespace mvfb;

__global__ void __launch_bounds__(128, 1)
shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int d, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // 256 rows x 128 B
  uint8_t* sB = smem + 256 * 128;           // 64 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], 256 * 128 + 64 * 128);
    tma_load_2d(sA, &tmA, &bars[0], 0, 0);
    tma_load_2d(sB, &tmB, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    const uint32_t a0 = smem_u32(sA) + (uint32_t)d * 128u, b0 = smem_u32(sB);
    for (int kk = 0; kk < 4; ++kk) {
      uint64_t ad = umma_smem_desc_sw128(a0 + kk * 32, 0, 1024);
      if (mode == 1) ad |= (uint64_t)((a0 >> 7) & 7u) << 49;
      const uint64_t bd = umma_smem_desc_sw128(b0 + kk * 32, 0, 1024);
      umma_f16(tmem, ad, bd, idesc, kk != 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int q = 0; q < 32; ++q) out[row * 64 + c0 + q] = __uint_as_float(v[q]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  std::vector<__nv_bfloat16> hA(256 * 64), hB(64 * 64);
  for (int r = 0; r < 256; ++r) for (int c = 0; c < 64; ++c) hA[r * 64 + c] = __float2bfloat16(c == r % 64 ? (float)(r + 1) : 0.f);
  for (int r = 0; r < 64; ++r) for (int c = 0; c < 64; ++c) hB[r * 64 + c] = __float2bfloat16(r == c ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB; float* dout;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  cuuint64_t dimsA[2] = {64, 256}, dimsB[2] = {64, 64}, strides[1] = {128};
  cuuint32_t boxA[2] = {64, 256}, boxB[2] = {64, 64}, es[2] = {1, 1};
  if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, strides, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
      enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, strides, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("encode failed\n"); return 1;
  }
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> h(128 * 64);
  const int ds[] = {0, 1, 2, 3, 5, 7, 8, 9, 57, 58, 59, 64, 100};
  for (int mode = 0; mode < 2; ++mode)
    for (int d : ds) {
      cudaMemset(dout, 0xff, 128 * 64 * 4);
      shift_kernel<<<1, 128, 1024 + 256 * 128 + 64 * 128 + 64>>>(tmA, tmB, dout, d, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d d %3d: CUDA error %s\n", mode, d, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          const float want = (n == (m + d) % 64) ? (float)(m + d + 1) : 0.f;
          if (h[m * 64 + n] != want) { if (first < 0) first = m * 64 + n; ++bad; }
        }
      printf("base_offset %s, shift %3d rows: %s", mode ? "(addr>>7)&7" : "0          ", d, bad ? "MISMATCH" : "exact");
      if (bad) printf("  (%d wrong, first at row %d col %d: got %g)", bad, first / 64, first % 64, h[first]);
      printf("\n");
    }
  return 0;
}

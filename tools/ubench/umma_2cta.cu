// Feasibility probe for tcgen05 cta_group::2 (a CTA pair computing one M = 256 tile: each CTA holds 128 rows of A and HALF
// of B; the leader issues the MMA, both read their 128 accumulator rows).  A[m][k] = (k == m % 64) ? m + 1 : 0,
// B[n][k] = (k == n % 64): D[m][n] = (m % 64 == n % 64) ? m + 1 : 0.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../mvfnet_b200/csrc -o umma_2cta umma_2cta.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <vector>
#include "ptx.cuh"
using namespace mvfb;

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // 128 rows x 128 B (this CTA's half of the M = 256 tile)
  uint8_t* sB = smem + 128 * 128;           // 128 rows x 128 B (this CTA's half of the N = 256 columns)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 128 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 3);
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); fence_barrier_init(); }
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  // both CTAs' loads complete on the LEADER's barrier (cp.async.bulk.tensor ... .cta_group::2 with the barrier address mapped
  // into CTA 0 by mapa); the peer also arrives remotely on the leader's third barrier
  uint32_t lead_full, lead_extra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lead_full) : "r"(smem_u32(&bars[0])), "r"(0));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lead_extra) : "r"(smem_u32(&bars[2])), "r"(0));
  if (threadIdx.x == 0) {
    if (rank == 0) mbar_arrive_expect_tx(&bars[0], 4 * 128 * 128);       // A and B halves of BOTH CTAs
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sA)), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(lead_full), "r"(0), "r"((int)rank * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sB)), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(lead_full), "r"(0), "r"((int)rank * 128) : "memory");
    if (rank == 1) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(lead_extra) : "memory");
  }
  if (rank == 0 && threadIdx.x == 0) {
    mbar_wait(&bars[0], 0);                  // all four boxes have landed (two of them in the peer's shared memory)
    mbar_wait(&bars[2], 0);                  // the peer's remote arrive
  }
  if (rank == 0 && threadIdx.x == 0) {
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(256, 256, 0, 0);
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t ad = umma_smem_desc_sw128(smem_u32(sA) + kk * 32, 0, 1024);
      const uint64_t bd = umma_smem_desc_sw128(smem_u32(sB) + kk * 32, 0, 1024);
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(kk != 0)) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bars[1])), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int q = 0; q < 32; ++q) out[(size_t)(rank * 128 + row) * 256 + c0 + q] = __uint_as_float(v[q]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  std::vector<__nv_bfloat16> hA(256 * 64), hB(256 * 64);
  for (int r = 0; r < 256; ++r) for (int c = 0; c < 64; ++c) {
    hA[r * 64 + c] = __float2bfloat16(c == r % 64 ? (float)(r + 1) : 0.f);
    hB[r * 64 + c] = __float2bfloat16(c == r % 64 ? 1.f : 0.f);
  }
  __nv_bfloat16 *dA, *dB; float* dout;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dout, 256 * 256 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xff, 256 * 256 * 4);
  CUtensorMap tmA, tmB;
  cuuint64_t dims[2] = {64, 256}, strides[1] = {128};
  cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
  if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
      enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("encode failed\n"); return 1;
  }
  cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  pair_kernel<<<2, 128, 1024 + 2 * 128 * 128 + 64>>>(tmA, tmB, dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> h(256 * 256);
  cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0, first = -1;
  for (int m = 0; m < 256; ++m) for (int n = 0; n < 256; ++n) {
    const float want = (m % 64 == n % 64) ? (float)(m + 1) : 0.f;
    if (h[m * 256 + n] != want) { if (first < 0) first = m * 256 + n; ++bad; }
  }
  printf("cta_group::2 M=256 N=256 K=64: %s", bad ? "MISMATCH" : "exact");
  if (bad) printf(" (%d wrong, first at row %d col %d: got %g)", bad, first / 256, first % 256, h[first]);
  printf("\n");
  return 0;
}

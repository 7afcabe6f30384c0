// Cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16, bf16) as a function of N, with both operands already
// in shared memory (K-major SWIZZLE_128B tiles, the layout gemm_tn.cu / conv_halo.cu use): one thread issues `iters` x 4
// MMAs into one accumulator, commits, waits.  Is the conv kernels' ~150-260 clocks per MMA the hardware rate for this
// operand layout, or something in their pipelines?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../mvfnet_b200/csrc -o umma_rate umma_rate.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "ptx.cuh"
using namespace mvfb;

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int a_shift_rows, int two_acc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // 256 rows x 128 B (garbage values: timing only)
  uint8_t* sB = smem + 256 * 128;           // N rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 256 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (256 + 256) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = smem_u32(sA) + (uint32_t)a_shift_rows * 128u, b0 = smem_u32(sB);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + ((two_acc && (i & 1)) ? 256u : 0u);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_f16(d, umma_smem_desc_sw128(a0 + kk * 32, 0, 1024), umma_smem_desc_sw128(b0 + kk * 32, 0, 1024), idesc, 1);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N>
void run(int shift, int two_acc) {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int iters = 2000;
  for (int grid : {1, 148}) {
    rate_kernel<N><<<grid, 128, 1024 + 512 * 128 + 64>>>(d, iters, shift, two_acc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
    printf("N=%3d  A start +%d rows, %s, %3d CTAs: %.1f clocks per MMA (ideal %d)  %s\n", N, shift, two_acc ? "two accumulators" : "one accumulator ",
           grid, (double)mx / (iters * 4), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  cudaFree(d);
}
int main() {
  run<64>(0, 0); run<64>(3, 0); run<64>(0, 1);
  run<128>(0, 0); run<128>(3, 0);
  run<256>(0, 0); run<256>(3, 0); run<256>(0, 1);
  return 0;
}

// TMA delivery rate vs. the width of a box row (B200): every CTA streams (Cg channels x 16 x 16 pixel) boxes of an NHWC
// bf16 tensor (C = 1024 channels per pixel, the MVF slab = its first 128) through a shared-memory ring; one thread issues
// the loads, one waits for them.  Same total bytes for every Cg: the time is the TMA unit's / L2's floor for that row
// width.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rows tma_rows.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma4(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(s32(dst)), "l"(m), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__global__ void __launch_bounds__(64, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm, int Cg, int ngroups, int frames, int R, int slot_bytes, int hw) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 32;
  uint8_t* slots = smem + 1024;
  const int cg = blockIdx.x % ngroups, p = blockIdx.x / ngroups, P = gridDim.x / ngroups;
  if (p >= P) return;
  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int mine = p < frames ? (frames - p + P - 1) / P : 0;
  if (threadIdx.x == 0) {                                        // producer
    int s = 0; uint32_t ph = 0;
    for (int k = 0; k < mine; ++k) {
      if (k >= R) wait(&empty[s], ph ^ 1);
      expect_tx(&full[s], (uint32_t)(Cg * 2 * hw * hw));
      tma4(slots + (size_t)s * slot_bytes, &tm, &full[s], cg * Cg, -1, -1, p + k * P);
      if (++s == R) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {                                // consumer
    int s = 0; uint32_t ph = 0;
    for (int k = 0; k < mine; ++k) {
      wait(&full[s], ph);
      arrive(&empty[s]);
      if (++s == R) { s = 0; ph ^= 1; }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int F = 1280, H = 14, W = 14, C = 1024, Cs = 128;
  void* x;
  cudaMalloc(&x, (size_t)F * H * W * C * 2);
  cudaMemset(x, 0, (size_t)F * H * W * C * 2);
  void* flush; cudaMalloc(&flush, 512u << 20);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int hws[2] = {16, 9};
  for (int hi = 0; hi < 1; ++hi)
  for (int Cg = 16; Cg <= 128; Cg *= 2) {
    for (int inflight_kb = 64; inflight_kb <= 192; inflight_kb *= 2) {
      const int hw = hws[hi];
      CUtensorMap tm;
      cuuint64_t dims[4] = {(cuuint64_t)Cs, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
      cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
      cuuint32_t box[4] = {(cuuint32_t)Cg, (cuuint32_t)hw, (cuuint32_t)hw, 1}, es[4] = {1, 1, 1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const int slot = Cg * 2 * hw * hw;
      int R = inflight_kb * 1024 / slot;
      if (R < 2) R = 2;
      if (R > 32) R = 32;
      if ((size_t)R * slot + 1024 > 220 * 1024) continue;
      const int ngroups = Cs / Cg;
      const int grid = 148 / ngroups * ngroups;
      float best = 1e9f;
      for (int it = 0; it < 5; ++it) {
        cudaMemsetAsync(flush, it, 512u << 20);
        cudaEventRecord(e0);
        stream_kernel<<<grid, 64, 1024 + (size_t)R * slot>>>(tm, Cg, ngroups, F, R, slot, hw);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
      }
      const double bytes = (double)F * H * W * Cs * 2;
      printf("Cg=%3d (row %3d B) ring %2d x %5d B  grid %3d: %7.1f us  %6.0f GB/s algorithmic  (%s)\n", Cg, Cg * 2, R, slot, grid,
             best * 1e3, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}

// Micro-benchmark (B200): the sm_100 mixed-precision FMA  fma.rn.f32.bf16  (SASS FHFMA.BF16: fp32 accumulate,
// bf16 operands read straight from either half of a 32-bit register -- no unpack instruction), against FFMA / FFMA2,
// alone and interleaved with FFMA (do they share a pipe?), plus the dependent-chain latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fhfma fhfma.cu && ./fhfma
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ float fh_lo(float c, uint32_t a, uint32_t b) {
  float r;
  asm("{.reg .b16 l,h,wl,wh;\n\t mov.b32 {l,h}, %1;\n\t mov.b32 {wl,wh}, %2;\n\t fma.rn.f32.bf16 %0, l, wl, %3;}"
      : "=f"(r) : "r"(a), "r"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fh_hi(float c, uint32_t a, uint32_t b) {
  float r;
  asm("{.reg .b16 l,h,wl,wh;\n\t mov.b32 {l,h}, %1;\n\t mov.b32 {wl,wh}, %2;\n\t fma.rn.f32.bf16 %0, h, wh, %3;}"
      : "=f"(r) : "r"(a), "r"(b), "f"(c));
  return r;
}

// MODE 0: 16 FFMA   1: 8 FFMA2   2: 16 FHFMA   3: 8 FHFMA + 8 FFMA   4: 8 FHFMA + 4 FFMA2   5: 1 dependent FHFMA chain
// MODE 6: 1 dependent FFMA chain
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float a, float b, uint32_t xa, uint32_t wb) {
  float2 r[8];
  uint32_t x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    r[j] = make_float2(threadIdx.x * 0.001f + j, j * 0.5f);
    x[j] = xa + 0x00010001u * (threadIdx.x + j);
  }
  const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.999f);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) { r[j].x = fmaf(r[j].x, a, b); r[j].y = fmaf(r[j].y, a, b); }
      if (MODE == 1) { r[j] = __ffma2_rn(r[j], A, B); }
      if (MODE == 2) { r[j].x = fh_lo(r[j].x, x[j], wb); r[j].y = fh_hi(r[j].y, x[j], wb); }
      if (MODE == 3) { r[j].x = fh_lo(r[j].x, x[j], wb); r[j].y = fmaf(r[j].y, a, b); }
      if (MODE == 4) { r[j].x = fh_lo(r[j].x, x[j], wb); if (j & 1) r[j - 1] = __ffma2_rn(make_float2(r[j - 1].y, r[j].y), A, B); }
      if (MODE == 5 && j == 0) { r[0].x = fh_lo(r[0].x, x[0], wb); }
      if (MODE == 6 && j == 0) { r[0].x = fmaf(r[0].x, a, b); }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += r[j].x + r[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int iters = 2048;
  for (int warps = 1; warps <= 32; warps *= 2) {
    if (MODE >= 5 && warps > 1) break;
    k<MODE><<<148, warps * 32>>>(out, cyc, 16, 1.0001f, 0.5f, 0x3f803f80u, 0x3f813f7fu);
    k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0001f, 0.5f, 0x3f803f80u, 0x3f813f7fu);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += (double)h[i];
    c /= 148;
    printf("%-16s warps/SM=%2d  cycles=%9.0f  warp-instr/clk/SM=%.2f  cycles/instr/warp=%.2f\n", name, warps, c,
           (double)warps * iters * ops / c, c / ((double)iters * ops));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA", 16);
  run<1>("FFMA2", 8);
  run<2>("FHFMA", 16);
  run<3>("FHFMA+FFMA", 16);
  run<4>("FHFMA+FFMA2", 12);
  run<5>("FHFMA dep-chain", 1);
  run<6>("FFMA dep-chain", 1);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}

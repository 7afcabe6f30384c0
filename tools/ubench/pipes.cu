// Instruction-throughput micro-benchmark (B200): FFMA vs FFMA2 vs the bf16-unpack ops, 8 independent chains/thread.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 r[8];
  uint32_t u[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { r[j] = make_float2(threadIdx.x * 0.001f + j, j * 0.5f); u[j] = threadIdx.x * 2654435761u + j; }
  const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.999f);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) { r[j].x = fmaf(r[j].x, a, b); r[j].y = fmaf(r[j].y, a, b); }          // 2 scalar FFMA
      if (MODE == 1) { r[j] = __ffma2_rn(r[j], A, B); }                                       // 1 FFMA2
      if (MODE == 2) { u[j] = (u[j] << 16) ^ 0x3f80u; }                                       // shift (+xor keeps it live)
      if (MODE == 3) { u[j] = (u[j] & 0xffff0000u) + 0x10000u; }                              // and (+add)
      if (MODE == 4) { u[j] = __byte_perm(u[j], 0x3f80u, 0x1054); }                           // PRMT
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += r[j].x + r[j].y + __uint_as_float(u[j] | 0x3f800000u);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int ops_per_iter_per_thread) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 32; warps *= 2) {
    k<MODE><<<148, warps * 32>>>(out, 16, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = (double)148 * warps * iters * ops_per_iter_per_thread;   // per SM: warps*iters*ops
    double per_sm_per_clk = warp_instr / 148 / (ms * 1e-3 * 1.9e9);
    printf("%-8s warps/SM=%2d  %.3f ms  warp-instr/clk/SM (at 1.9 GHz) = %.2f\n", name, warps, ms, per_sm_per_clk);
  }
  cudaFree(out);
}
int main() {
  run<0>("FFMA", 16); run<1>("FFMA2", 8); run<2>("SHL+XOR", 16); run<3>("AND+ADD", 16); run<4>("PRMT", 8);
  return 0;
}

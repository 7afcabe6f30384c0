// What the MVF slab's LAYOUT allows (B200): the slab is the first Cs = C/8 channels of every pixel of an NHWC bf16
// tensor, i.e. runs of 2*Cs bytes at a stride of 2*C bytes.  Plain LDG.128 / STG.128 kernels at full occupancy:
//   read   : sum the slab (no write)              copy_s2c: strided slab -> contiguous slab (the MVF forward's pattern)
//   copy_c2c: contiguous -> contiguous (the peak the bench's roofline uses)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o slab_copy slab_copy.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

// chunks of 16 B: pixel p, chunk q of Q = Cs/8
__global__ void __launch_bounds__(256) k_read(const uint4* __restrict__ x, long long pixels, int Q, int pitch16, unsigned* sink) {
  const long long total = pixels * Q, stride = (long long)gridDim.x * 256;
  unsigned acc = 0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += stride) {
    const long long p = i / Q; const int q = (int)(i - p * Q);
    const uint4 v = x[p * pitch16 + q];
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}
__global__ void __launch_bounds__(256) k_copy(const uint4* __restrict__ x, uint4* __restrict__ y, long long pixels, int Q, int pin, int pout) {
  const long long total = pixels * Q, stride = (long long)gridDim.x * 256;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += stride) {
    const long long p = i / Q; const int q = (int)(i - p * Q);
    y[p * pout + q] = x[p * pin + q];
  }
}
int main() {
  void* flush; cudaMalloc(&flush, 512u << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  unsigned* sink; cudaMalloc(&sink, 4);
  const int shapes[3][3] = {{512, 28, 64}, {1024, 14, 128}, {2048, 7, 256}};
  for (int si = 0; si < 3; ++si) {
    const int C = shapes[si][0], H = shapes[si][1], Cs = shapes[si][2];
    const long long pixels = 1280LL * H * H;
    uint4 *x, *y, *xc;
    cudaMalloc(&x, pixels * C * 2); cudaMalloc(&y, pixels * C * 2); cudaMalloc(&xc, pixels * Cs * 2);
    cudaMemset(x, 1, pixels * C * 2); cudaMemset(y, 0, pixels * C * 2); cudaMemset(xc, 1, pixels * Cs * 2);
    const int Q = Cs / 8, P = C / 8;
    const double slab = (double)pixels * Cs * 2;
    for (int mode = 0; mode < 5; ++mode) {
      for (int ctas = 148 * 4; ctas <= 148 * 16; ctas *= 2) {
        float best = 1e9f;
        for (int it = 0; it < 6; ++it) {
          cudaMemsetAsync(flush, it, 512u << 20);
          cudaEventRecord(e0);
          if (mode == 0) k_read<<<ctas, 256>>>(x, pixels, Q, P, sink);
          if (mode == 1) k_copy<<<ctas, 256>>>(x, xc, pixels, Q, P, Q);     // strided -> contiguous
          if (mode == 2) k_copy<<<ctas, 256>>>(xc, y, pixels, Q, Q, P);     // contiguous -> strided
          if (mode == 3) k_copy<<<ctas, 256>>>(x, y, pixels, Q, P, P);      // strided -> strided
          if (mode == 4) k_copy<<<ctas, 256>>>(xc, (uint4*)flush, pixels, Q, Q, Q);   // contiguous -> contiguous
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (it > 0 && ms < best) best = ms;
        }
        const char* names[5] = {"read strided      ", "copy strided->cont", "copy cont->strided", "copy strided->strd", "copy cont->cont   "};
        const double bytes = mode == 0 ? slab : 2 * slab;
        printf("C=%4d %2dx%2d Cs=%3d  %s ctas %4d: %7.1f us  %6.0f GB/s\n", C, H, H, Cs, names[mode], ctas, best * 1e3, bytes / best / 1e6);
      }
    }
    cudaFree(x); cudaFree(y); cudaFree(xc);
  }
  return 0;
}

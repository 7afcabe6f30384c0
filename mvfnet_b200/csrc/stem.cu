// ResNet stem (backbones/resnet.py:424-431, 481-484): Conv2d(3, 64, 7, stride 2, pad 3) and MaxPool2d(3, 2, 1).
//
// The step's launch list (profiles/r01_step_launches_final.txt) had the cuDNN stem convolution at 9.6 % of the step
// (one launch, 11 ms for 1024 frames: 3 input channels do not feed an implicit-GEMM kernel), its weight gradient at
// 2.6 % and ATen's max-pool pair at 8.9 %.  Here:
//   * stem_im2col: the 7x7x3 patches of every output pixel as rows of a (F*Ho*Wo, 192) bf16 matrix, column
//     k = kh*24 + kw*3 + c -- in an NHWC image with 3 channels the 21 values of one kernel row are 21 CONTIGUOUS
//     elements; each kernel row is padded to 24 columns (three aligned 16-byte chunks) and the matrix to 192 columns
//     so that the tcgen05 GEMMs of this library take it as a 1x1 convolution:
//     conv1x1_gemm gives the convolution (with the BatchNorm statistics in its epilogue), conv1x1_wgrad its
//     weight gradient.  The input needs no gradient.
//   * maxpool3x3s2_fwd / _bwd: NHWC bf16, 8 channels per thread; the forward records the arg-max position inside the
//     window (0..8, one byte per element, first maximum in (kh, kw) scan order exactly as ATen's kernels), the
//     backward GATHERS: every input pixel looks at the <= 4 windows that contain it.  No atomics, no int64 indices.
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"
#include "mvf_internal.cuh"
#include "ptx.cuh"

namespace mvfb {

namespace {

constexpr int kStemKp = 192;       // row length of the patch matrix: 8 groups of 24 columns
constexpr int kRowK = 24;          // one kernel row kh = 21 values (kw, c) + 3 zeros: 48 bytes = three aligned 16-byte chunks

// x: (F, H, W, 3) bf16 NHWC.  One thread per (output pixel, 16-byte chunk of the 192-wide row); column
// k = kh*24 + kw*3 + c (kw*3 + c < 21), the rest zero.  The 8 values of a chunk are 8 CONSECUTIVE elements of input
// row ih = 2*oh - 3 + kh starting at element (2*ow - 3)*3 + 8*j -- an odd element index, i.e. a 2-byte-aligned
// address: the interior fast path reads the five aligned 32-bit words around them and funnel-shifts (5 loads and
// 4 shifts per 16 bytes written instead of 8 two-byte loads and 7 merges; the first version gathered element by
// element and ran at 23 % of the HBM rate, profiles/r02_bench_a_b160.json).  Image borders take the element loop.
// I = index type of the flattened work item: unsigned 32-bit whenever the launch has fewer than 2^31 items (64-bit
// division / modulo costs ~100 instructions each).
template <typename I>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ a, int H, int W, int Ho, int Wo,
                   long long total) {
  const I i = (I)blockIdx.x * (I)blockDim.x + threadIdx.x;
  if ((long long)i >= total) return;
  const int chunk = (int)(i % (I)(kStemKp / 8));
  const I row = i / (I)(kStemKp / 8);
  const int ow = (int)(row % (I)Wo);
  const I t = row / (I)Wo;
  const int oh = (int)(t % (I)Ho);
  const long long f = (long long)(t / (I)Ho);
  const int rowlen = W * 3;
  const int kh = chunk / 3, j = chunk - kh * 3;
  const int ih = 2 * oh - 3 + kh;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (kh < 7 && ih >= 0 && ih < H) {
    const int off = (2 * ow - 3) * 3 + 8 * j;                    // first element of the chunk inside image row ih
    const unsigned short* src = reinterpret_cast<const unsigned short*>(x) + (f * H + ih) * (long long)rowlen;
    if (!(W & 1) && off >= 1 && off + 9 <= rowlen) {
      // off is odd: the elements off-1 .. off+8 are five aligned words
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(src + off - 1);
      const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3), w4 = __ldg(wp + 4);
      o.x = __funnelshift_r(w0, w1, 16);
      o.y = __funnelshift_r(w1, w2, 16);
      o.z = __funnelshift_r(w2, w3, 16);
      o.w = __funnelshift_r(w3, w4, 16);
      if (j == 2) { o.z &= 0xffffu; o.w = 0u; }                  // values 21..23 of the kernel row do not exist
    } else {
      unsigned short v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int r = 8 * j + e, q = off + e;
        v[e] = (r < 21 && q >= 0 && q < rowlen) ? __ldg(src + q) : (unsigned short)0;
      }
      o.x = v[0] | ((uint32_t)v[1] << 16); o.y = v[2] | ((uint32_t)v[3] << 16);
      o.z = v[4] | ((uint32_t)v[5] << 16); o.w = v[6] | ((uint32_t)v[7] << 16);
    }
  }
  reinterpret_cast<uint4*>(a)[i] = o;
}

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// x: (F, H, W, C) bf16 NHWC -> y: (F, Ho, Wo, C), idx: (F, Ho, Wo, C) bytes.  One thread per (output pixel, 8 channels).
template <typename I>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, uint2* __restrict__ idx, int H, int W, int Ho,
                   int Wo, int C8, long long total) {
  const I i = (I)blockIdx.x * (I)blockDim.x + threadIdx.x;
  if ((long long)i >= total) return;
  const int c8 = (int)(i % (I)C8);
  I t = i / (I)C8;
  const int ow = (int)(t % (I)Wo); t /= (I)Wo;
  const int oh = (int)(t % (I)Ho);
  const long long f = (long long)(t / (I)Ho);
  float best[8];
  int arg[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; arg[e] = 0; }
  bool any = false;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int ih = 2 * oh - 1 + kh;
    if (ih < 0 || ih >= H) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int iw = 2 * ow - 1 + kw;
      if (iw < 0 || iw >= W) continue;
      const uint4 v = __ldg(x + ((f * H + ih) * W + iw) * C8 + c8);
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float lo = bf_lo(w4[q]), hi = bf_hi(w4[q]);
        // ATen: (val > maxval) || isnan(val); the first valid element always replaces -inf
        if (!any || lo > best[2 * q] || lo != lo) { best[2 * q] = lo; arg[2 * q] = kh * 3 + kw; }
        if (!any || hi > best[2 * q + 1] || hi != hi) { best[2 * q + 1] = hi; arg[2 * q + 1] = kh * 3 + kw; }
      }
      any = true;
    }
  }
  uint4 o;
  o.x = pack_bf16(best[0], best[1]); o.y = pack_bf16(best[2], best[3]);
  o.z = pack_bf16(best[4], best[5]); o.w = pack_bf16(best[6], best[7]);
  y[i] = o;
  uint2 k;
  k.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
  k.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
  idx[i] = k;
}

// dx[f, ih, iw, c] = sum over the windows (oh, ow) that contain (ih, iw) of g[f, oh, ow, c] * [idx[f, oh, ow, c] == position]
template <typename I>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const uint4* __restrict__ g, const uint2* __restrict__ idx, uint4* __restrict__ dx, int H, int W,
                   int Ho, int Wo, int C8, long long total) {
  const I i = (I)blockIdx.x * (I)blockDim.x + threadIdx.x;
  if ((long long)i >= total) return;
  const int c8 = (int)(i % (I)C8);
  I t = i / (I)C8;
  const int iw = (int)(t % (I)W); t /= (I)W;
  const int ih = (int)(t % (I)H);
  const long long f = (long long)(t / (I)H);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  // windows: 2*oh - 1 <= ih <= 2*oh + 1
  const int oh0 = ih >> 1, oh1 = (ih + 1) >> 1;                 // equal when ih is even ... (ih odd: two windows)
  const int ow0 = iw >> 1, ow1 = (iw + 1) >> 1;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int oh = a == 0 ? oh0 : oh1;
    if (a == 1 && oh1 == oh0) continue;
    if (oh >= Ho) continue;
    const int kh = ih - (2 * oh - 1);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ow = b == 0 ? ow0 : ow1;
      if (b == 1 && ow1 == ow0) continue;
      if (ow >= Wo) continue;
      const int pos = kh * 3 + (iw - (2 * ow - 1));
      const long long o = ((f * Ho + oh) * Wo + ow) * C8 + c8;
      const uint2 k = __ldg(idx + o);
      const uint4 v = __ldg(g + o);
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t kk = q < 2 ? k.x : k.y;
        const int sh = (q & 1) * 16;
        if ((int)((kk >> sh) & 0xff) == pos) acc[2 * q] += bf_lo(w4[q]);
        if ((int)((kk >> (sh + 8)) & 0xff) == pos) acc[2 * q + 1] += bf_hi(w4[q]);
      }
    }
  }
  uint4 o;
  o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]);
  o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
  dx[i] = o;
}

}  // namespace

}  // namespace mvfb

using namespace mvfb;

extern "C" {

int stem_im2col(const void* x, void* a, long long F, int H, int W, mvfb_stream_t stream) {
  MVFB_CHECK(x && a && F > 0 && H > 0 && W > 0, MVFB_ERR_ARG, "stem_im2col: bad arguments");
  MVFB_CHECK(!(reinterpret_cast<uintptr_t>(a) & 15), MVFB_ERR_ARG, "stem_im2col: output must be 16-byte aligned");
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  const long long total = F * Ho * Wo * (kStemKp / 8);
  const unsigned blocks = (unsigned)ceil_div_ll(total, 256);
  if (total < (1LL << 31))
    stem_im2col_kernel<uint32_t><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)a, H, W, Ho,
                                                                           Wo, total);
  else
    stem_im2col_kernel<long long><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)a, H, W, Ho,
                                                                            Wo, total);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int maxpool3x3s2_fwd(const void* x, void* y, void* idx, long long F, int H, int W, int C, mvfb_stream_t stream) {
  MVFB_CHECK(x && y && idx && F > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, MVFB_ERR_ARG, "maxpool3x3s2_fwd: bad arguments");
  MVFB_CHECK(!((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) && !(reinterpret_cast<uintptr_t>(idx) & 7),
             MVFB_ERR_ARG, "maxpool3x3s2_fwd: misaligned tensors");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = F * Ho * Wo * (C / 8);
  if (total < (1LL << 31))
    maxpool_fwd_kernel<uint32_t><<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)x, (uint4*)y, (uint2*)idx, H, W, Ho, Wo, C / 8, total);
  else
    maxpool_fwd_kernel<long long><<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)x, (uint4*)y, (uint2*)idx, H, W, Ho, Wo, C / 8, total);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int maxpool3x3s2_bwd(const void* g, const void* idx, void* dx, long long F, int H, int W, int C, mvfb_stream_t stream) {
  MVFB_CHECK(g && dx && idx && F > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, MVFB_ERR_ARG, "maxpool3x3s2_bwd: bad arguments");
  MVFB_CHECK(!((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(dx)) & 15) && !(reinterpret_cast<uintptr_t>(idx) & 7),
             MVFB_ERR_ARG, "maxpool3x3s2_bwd: misaligned tensors");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = F * H * W * (C / 8);
  if (total < (1LL << 31))
    maxpool_bwd_kernel<uint32_t><<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)g, (const uint2*)idx, (uint4*)dx, H, W, Ho, Wo, C / 8, total);
  else
    maxpool_bwd_kernel<long long><<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)g, (const uint2*)idx, (uint4*)dx, H, W, Ho, Wo, C / 8, total);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

}  // extern "C"

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

// The same matrix, one CTA per output row (f, oh): the seven input rows it reads are staged ONCE in shared memory
// (16-byte global loads, zero padding for the image border written explicitly), then every thread assembles 16-byte
// chunks from four aligned 32-bit shared-memory loads.  The gather above issues five 4-byte GLOBAL loads per chunk and
// each warp-level load touches up to eight input rows: L1 throughput 82 %, 1.9 ms for 6.1 GB (profiles/r02_stem_pool_ncu.txt).
// Element e of an input row lives at byte kRowLead + 2*e of its shared-memory row; a chunk starts at the odd element
// 6*ow - 9 + 8*j, i.e. at byte 16 + 12*ow + 16*j: always word-aligned.  Needs W % 8 == 0 (16-byte aligned rows).
constexpr int kRowLead = 34;       // bytes before element 0: 16 + the 9 elements (3 pixels) of left padding
__global__ void __launch_bounds__(256)
stem_im2col_rows_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ a, int H, int W, int Ho, int Wo,
                        int pitch) {
  extern __shared__ __align__(16) uint8_t srow[];                // 7 rows of `pitch` bytes
  const int oh = blockIdx.x % Ho;
  const long long f = blockIdx.x / Ho;
  const int rowbytes = W * 6;
  // zero the padding of every row (bytes [0, kRowLead) and [kRowLead + rowbytes, pitch)), and whole rows outside the image
  const int tailw = (pitch - kRowLead - rowbytes) / 2;
  for (int i = threadIdx.x; i < 7 * (kRowLead / 2 + tailw); i += 256) {
    const int r = i / (kRowLead / 2 + tailw), k = i - r * (kRowLead / 2 + tailw);
    const int off = k < kRowLead / 2 ? 2 * k : kRowLead + rowbytes + 2 * (k - kRowLead / 2);
    *reinterpret_cast<unsigned short*>(srow + r * pitch + off) = 0;
  }
  const int vec = rowbytes / 16;
  for (int i = threadIdx.x; i < 7 * vec; i += 256) {
    const int r = i / vec, v = i - r * vec;
    const int ih = 2 * oh - 3 + r;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (ih >= 0 && ih < H) q = __ldg(reinterpret_cast<const uint4*>(x + (f * H + ih) * (long long)(W * 3)) + v);
    uint8_t* d = srow + r * pitch + kRowLead + 16 * v;           // 2 (mod 4): halfword, three words, halfword
    *reinterpret_cast<unsigned short*>(d) = (unsigned short)(q.x & 0xffffu);
    *reinterpret_cast<uint32_t*>(d + 2) = __funnelshift_r(q.x, q.y, 16);
    *reinterpret_cast<uint32_t*>(d + 6) = __funnelshift_r(q.y, q.z, 16);
    *reinterpret_cast<uint32_t*>(d + 10) = __funnelshift_r(q.z, q.w, 16);
    *reinterpret_cast<unsigned short*>(d + 14) = (unsigned short)(q.w >> 16);
  }
  __syncthreads();
  uint4* out = reinterpret_cast<uint4*>(a) + ((f * Ho + oh) * (long long)Wo) * (kStemKp / 8);
  const int chunks = Wo * (kStemKp / 8);
  for (int i = threadIdx.x; i < chunks; i += 256) {
    const int ow = i / (kStemKp / 8), c = i - ow * (kStemKp / 8);
    const int kh = c / 3, j = c - kh * 3;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (kh < 7) {
      const uint32_t* p = reinterpret_cast<const uint32_t*>(srow + kh * pitch + 16 + 12 * ow + 16 * j);
      o.x = p[0]; o.y = p[1];
      if (j < 2) { o.z = p[2]; o.w = p[3]; } else { o.z = p[2] & 0xffffu; }   // values 21..23 of a kernel row do not exist
    }
    out[i] = o;
  }
}

// ---- packed bf16 helpers of the pooling kernels
__device__ __forceinline__ uint32_t hmax2_nan(uint32_t a, uint32_t b) {        // NaN-propagating maximum of both halves
  uint32_t d;
  asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t heq2(uint32_t a, uint32_t b) {             // 0xffff per half where a == b
  uint32_t d;
  asm("set.eq.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ bool has_nan2(uint32_t v) {                          // either half is a NaN
  return (((v & 0x7fff7fffu) + 0x007f007fu) & 0x80008000u) != 0u;
}

// ATen's rule restated literally for a window that holds a NaN: `(val > maxval) || isnan(val)`, the first valid tap
// replaces -inf.  Out of line and self-contained (it reloads the taps) so that the common path keeps them in registers.
struct PoolOut { uint32_t m[4], a2[4]; };
__device__ __noinline__ PoolOut maxpool_window_scalar(const uint4* px, int rowp, int C8, uint32_t okmask) {
  float best[8];
  int arg[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; arg[e] = 0; }
  bool any = false;
  for (int tp = 0; tp < 9; ++tp) {
    if (!((okmask >> tp) & 1u)) continue;
    const uint4 q4 = __ldg(px + (tp / 3) * rowp + (tp % 3) * C8);
    const uint32_t w4[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float lo = bf_lo(w4[q]), hi = bf_hi(w4[q]);
      if (!any || lo > best[2 * q] || lo != lo) { best[2 * q] = lo; arg[2 * q] = tp; }
      if (!any || hi > best[2 * q + 1] || hi != hi) { best[2 * q + 1] = hi; arg[2 * q + 1] = tp; }
    }
    any = true;
  }
  PoolOut o;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    o.m[q] = pack_bf16(best[2 * q], best[2 * q + 1]);
    o.a2[q] = (uint32_t)arg[2 * q] | ((uint32_t)arg[2 * q + 1] << 16);
  }
  return o;
}

// x: (F, H, W, C) bf16 NHWC -> y: (F, Ho, Wo, C), idx: (F, Ho, Wo, C) bytes.  One thread per (output pixel, 8 channels).
// The maximum is taken on packed bf16 pairs (3-input VHMNMX), the arg-max is the FIRST tap in scan order equal to it
// (ATen's `val > maxval` rule: ties keep the earlier tap); a window that holds a NaN takes the scalar path above.  The nine
// taps sit at fixed offsets from one 64-bit pointer per thread (first version: 450 instructions per thread, a third of them
// 64-bit index arithmetic and per-channel compares; 74 % issue-bound -- profiles/r02_stem_pool_ncu.txt).
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
  const int ih0 = 2 * oh - 1, iw0 = 2 * ow - 1;
  const uint4* px = x + ((f * H + ih0) * W + iw0) * C8 + c8;    // tap (0, 0); only dereferenced where it is inside the image
  const int rowp = W * C8;
  uint32_t okmask = 0;
  uint32_t v[9][4];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int tp = kh * 3 + kw;
      const bool ok = (unsigned)(ih0 + kh) < (unsigned)H && (unsigned)(iw0 + kw) < (unsigned)W;
      uint4 q = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // -inf: never wins
      if (ok) { q = __ldg(px + kh * rowp + kw * C8); okmask |= 1u << tp; }
      v[tp][0] = q.x; v[tp][1] = q.y; v[tp][2] = q.z; v[tp][3] = q.w;
    }
  }
  uint32_t m[4], a2[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    m[q] = v[0][q];
#pragma unroll
    for (int tp = 1; tp < 9; ++tp) m[q] = hmax2_nan(m[q], v[tp][q]);
  }
  // Scan from the last tap to the first, so that the earliest tap equal to the maximum is the one written last; padded
  // taps (-inf) are masked out, and tap (1, 1) -- inside the image for every window -- is the default.
#pragma unroll
  for (int q = 0; q < 4; ++q) a2[q] = 0x00040004u;
#pragma unroll
  for (int tp = 8; tp >= 0; --tp) {
    const uint32_t on = (okmask >> tp) & 1u ? 0xffffffffu : 0u;
    const uint32_t code = (uint32_t)tp * 0x00010001u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t e = heq2(v[tp][q], m[q]) & on;
      a2[q] = (e & code) | (~e & a2[q]);
    }
  }
  if (has_nan2(m[0]) || has_nan2(m[1]) || has_nan2(m[2]) || has_nan2(m[3])) {
    const PoolOut o = maxpool_window_scalar(px, rowp, C8, okmask);
#pragma unroll
    for (int q = 0; q < 4; ++q) { m[q] = o.m[q]; a2[q] = o.a2[q]; }
  }
  y[i] = make_uint4(m[0], m[1], m[2], m[3]);
  uint2 k;                                                       // bytes 0 and 2 of each pair register
  k.x = __byte_perm(a2[0], a2[1], 0x6420);
  k.y = __byte_perm(a2[2], a2[3], 0x6420);
  idx[i] = k;
}

// dx[f, ih, iw, c] = sum over the windows (oh, ow) that contain (ih, iw) of g[f, oh, ow, c] * [idx[f, oh, ow, c] == position]
// (fp32 sum of up to four bf16 terms, rounded once).  The byte compare and the masking are SIMD-in-register; the masked
// bf16 pairs are accumulated by the mixed-precision FMA (x * 1.0 + acc).
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
  uint2 kk[4];
  uint4 vv[4];
  int pos[4];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int oh = a == 0 ? oh0 : oh1;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ow = b == 0 ? ow0 : ow1;
      const int w = a * 2 + b;
      const bool on = !(a == 1 && oh1 == oh0) && oh < Ho && !(b == 1 && ow1 == ow0) && ow < Wo;
      pos[w] = on ? (ih - (2 * oh - 1)) * 3 + (iw - (2 * ow - 1)) : 0xff;     // 0xff matches no recorded position
      kk[w] = make_uint2(0u, 0u);
      vv[w] = make_uint4(0u, 0u, 0u, 0u);
      if (on) {
        const long long o = ((f * Ho + oh) * Wo + ow) * C8 + c8;
        kk[w] = __ldg(idx + o);
        vv[w] = __ldg(g + o);
      }
    }
  }
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const uint32_t p4 = (uint32_t)pos[w] * 0x01010101u;
    const uint32_t e0 = __vcmpeq4(kk[w].x, p4), e1 = __vcmpeq4(kk[w].y, p4);   // 0xff per matching byte
    const uint32_t w4[4] = {vv[w].x & __byte_perm(e0, 0u, 0x1100), vv[w].y & __byte_perm(e0, 0u, 0x3322),
                            vv[w].z & __byte_perm(e1, 0u, 0x1100), vv[w].w & __byte_perm(e1, 0u, 0x3322)};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      asm("{\n\t.reg .b16 lo, hi, one;\n\t"
          "mov.b32 {lo, hi}, %2;\n\t"
          "mov.b16 one, 0x3f80;\n\t"
          "fma.rn.f32.bf16 %0, lo, one, %0;\n\t"
          "fma.rn.f32.bf16 %1, hi, one, %1;\n\t}"
          : "+f"(acc[2 * q]), "+f"(acc[2 * q + 1])
          : "r"(w4[q]));
    }
  }
  uint4 o;
  o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]);
  o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
  dx[i] = o;
}

// Even H and W (the stem: 112 x 112): one thread per 2 x 2 block of input pixels and 8 channels.  The block (2a.., 2b..) is
// covered by the four windows (a | a+1, b | b+1) only, and every (pixel, window) pair has a FIXED tap position:
//   (2a, 2b): window (a, b) tap 4                       (2a, 2b+1): (a, b) tap 5, (a, b+1) tap 3
//   (2a+1, 2b): (a, b) tap 7, (a+1, b) tap 1            (2a+1, 2b+1): (a, b) 8, (a, b+1) 6, (a+1, b) 2, (a+1, b+1) 0
// so four (g, idx) loads serve four pixels (the per-pixel kernel above loads nine) and the position compares are against
// constants.  byte == tap  <=>  ((byte ^ tap) + 0x7f) has its top bit clear (bytes are < 16); PRMT's sign-replicate mode
// spreads that bit over the bf16 half it guards.
// PRMT in its generic PTX form: selector nibbles 8..11 replicate the SIGN of source byte 0..3 over the target byte
// (CUDA's __byte_perm masks the selector to three bits and cannot express this)
__device__ __forceinline__ uint32_t prmt_sign(uint32_t a, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(d) : "r"(a), "r"(sel));
  return d;
}
__device__ __forceinline__ void pool_take(float (&acc)[8], const uint4& gv, const uint2& kv, uint32_t tap) {
  const uint32_t p4 = tap * 0x01010101u;
  const uint32_t n0 = (kv.x ^ p4) + 0x7f7f7f7fu, n1 = (kv.y ^ p4) + 0x7f7f7f7fu;      // top bit of a byte set <=> not this tap
  const uint32_t w4[4] = {gv.x & ~prmt_sign(n0, 0x9988u), gv.y & ~prmt_sign(n0, 0xbbaau),
                          gv.z & ~prmt_sign(n1, 0x9988u), gv.w & ~prmt_sign(n1, 0xbbaau)};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    asm("{\n\t.reg .b16 lo, hi, one;\n\t"
        "mov.b32 {lo, hi}, %2;\n\t"
        "mov.b16 one, 0x3f80;\n\t"
        "fma.rn.f32.bf16 %0, lo, one, %0;\n\t"
        "fma.rn.f32.bf16 %1, hi, one, %1;\n\t}"
        : "+f"(acc[2 * q]), "+f"(acc[2 * q + 1])
        : "r"(w4[q]));
  }
}
__device__ __forceinline__ uint4 pool_pack(const float (&acc)[8]) {
  uint4 o;
  o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]);
  o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
  return o;
}

template <typename I>
__global__ void __launch_bounds__(256)
maxpool_bwd_block_kernel(const uint4* __restrict__ g, const uint2* __restrict__ idx, uint4* __restrict__ dx, int W, int Ho, int Wo,
                         int C8, long long total) {
  const I i = (I)blockIdx.x * (I)blockDim.x + threadIdx.x;       // (f, a, b, c8): the layout of g itself
  if ((long long)i >= total) return;
  const int c8 = (int)(i % (I)C8);
  I t = i / (I)C8;
  const int b = (int)(t % (I)Wo); t /= (I)Wo;
  const int a = (int)(t % (I)Ho);
  const long long f = (long long)(t / (I)Ho);
  const bool right = b + 1 < Wo, down = a + 1 < Ho;
  const int rowo = Wo * C8;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  const uint2 k9 = make_uint2(0x09090909u, 0x09090909u);          // tap 9 does not exist: matches nothing
  const uint4 g00 = __ldg(g + i);
  const uint2 k00 = __ldg(idx + i);
  const uint4 g01 = right ? __ldg(g + i + C8) : z4;
  const uint2 k01 = right ? __ldg(idx + i + C8) : k9;
  const uint4 g10 = down ? __ldg(g + i + rowo) : z4;
  const uint2 k10 = down ? __ldg(idx + i + rowo) : k9;
  const uint4 g11 = (right && down) ? __ldg(g + i + rowo + C8) : z4;
  const uint2 k11 = (right && down) ? __ldg(idx + i + rowo + C8) : k9;
  uint4* out = dx + ((f * (2 * Ho) + 2 * a) * (long long)W + 2 * b) * C8 + c8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  pool_take(acc, g00, k00, 4u);
  out[0] = pool_pack(acc);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  pool_take(acc, g00, k00, 5u); pool_take(acc, g01, k01, 3u);
  out[C8] = pool_pack(acc);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  pool_take(acc, g00, k00, 7u); pool_take(acc, g10, k10, 1u);
  out[(long long)W * C8] = pool_pack(acc);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  pool_take(acc, g00, k00, 8u); pool_take(acc, g01, k01, 6u); pool_take(acc, g10, k10, 2u); pool_take(acc, g11, k11, 0u);
  out[(long long)W * C8 + C8] = pool_pack(acc);
}


// ---------------------------------------------------------------- norm1 + ReLU + max-pool in one pass (and its backward)
// ResNet.forward's `maxpool(relu(norm1(conv1(x))))` (backbones/resnet.py:481-484).  As three launches the stem's tail
// moves the largest activation of the network five times (bn_apply: read + write, max-pool: read; backward: the routed
// gradient written by maxpool_bwd and read twice by bn_bwd).  y = relu(scale * x + shift) rounded to bf16 is a monotone
// function of x per channel -- non-decreasing for scale >= 0, non-increasing below -- so the window maximum of y is that
// function of the window's largest (smallest) x: the forward takes the packed-bf16 maximum of the RAW convolution
// outputs, with the sign bit of the channels whose scale is negative flipped, and applies the affine + ReLU + rounding to
// the winner only.  The result equals max-pooling the rounded activations bit for bit; the recorded position is the
// first tap holding the extreme x (where two different x round to the same activation ATen would record the earlier
// one: both carry the same forward value, and the fp32 reference agrees with THIS choice).  A window whose maximum is
// <= 0 records position 9, which matches no tap: the ReLU's backward needs nothing else.  The backward gathers the
// pooled gradient per 2 x 2 input block (maxpool_bwd_block_kernel's scheme) inside BOTH BatchNorm backward passes, so
// neither the activation nor its gradient ever exists in memory.  Work split as in bn.cu: a thread owns 8 channels.
struct PoolBnArgs {
  const uint4* x;                        // (F, H, W, C) convolution output
  uint4* y;                              // (F, Ho, Wo, C)
  uint2* idx;                            // (F, Ho, Wo, C) bytes
  long long F, M;                        // M = F * H * W (the BatchNorm's population)
  int H, W, Ho, Wo, C8, training;
  float eps, momentum;
  const float *sums, *gamma, *beta;
  float *running_mean, *running_var, *save_mean, *save_rstd;
};

#ifndef POOLBN_FWD_MINB
#define POOLBN_FWD_MINB 4
#endif
#ifndef POOLBN_RED_MINB
#define POOLBN_RED_MINB 2
#endif
#ifndef POOLBN_APPLY_MINB
#define POOLBN_APPLY_MINB 2
#endif
template <typename I>
__global__ void __launch_bounds__(256, POOLBN_FWD_MINB)
bn_relu_maxpool_fwd_kernel(const PoolBnArgs a) {
  extern __shared__ float4 s_coef[];                             // [2 * C8] scale, then [2 * C8] shift: read once per window
  const int C8 = a.C8, C = 8 * C8;
  const int vec = threadIdx.x % C8, rowlane = threadIdx.x / C8, rows_par = 256 / C8;
  for (int c = threadIdx.x; c < C; c += 256) {                   // bn_apply_kernel's prologue (bn.cu), once per CTA
    float mean, rstd;
    if (a.training) {
      const double m = (double)a.M;
      const double mu = (double)a.sums[c] / m;
      double var = (double)a.sums[C + c] / m - mu * mu;
      if (var < 0) var = 0;
      mean = (float)mu;
      rstd = (float)(1.0 / sqrt(var + (double)a.eps));
      if (blockIdx.x == 0) {
        a.save_mean[c] = mean;
        a.save_rstd[c] = rstd;
        if (a.running_mean) {
          const double unb = m > 1 ? var * m / (m - 1) : var;
          a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mean;
          a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unb;
        }
      }
    } else {
      mean = a.running_mean[c];
      rstd = 1.f / sqrtf(a.running_var[c] + a.eps);
      if (blockIdx.x == 0 && a.save_mean) { a.save_mean[c] = mean; a.save_rstd[c] = rstd; }
    }
    const float sc = a.gamma[c] * rstd;
    reinterpret_cast<float*>(s_coef)[c] = sc;
    reinterpret_cast<float*>(s_coef)[C + c] = a.beta[c] - mean * sc;
  }
  __syncthreads();
  uint32_t flip[4];
  {
    const float* sc = reinterpret_cast<const float*>(s_coef) + vec * 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) flip[q] = (sc[2 * q] < 0.f ? 0x00008000u : 0u) | (sc[2 * q + 1] < 0.f ? 0x80000000u : 0u);
  }
  const int H = a.H, W = a.W, Ho = a.Ho, Wo = a.Wo;
  const int rowp = W * C8;
  const I npix = (I)(a.F * Ho * Wo);
  for (I p = (I)blockIdx.x * (I)rows_par + (I)rowlane; p < npix; p += (I)gridDim.x * (I)rows_par) {
    const int ow = (int)(p % (I)Wo);
    const I t = p / (I)Wo;
    const int oh = (int)(t % (I)Ho);
    const long long f = (long long)(t / (I)Ho);
    const int ih0 = 2 * oh - 1, iw0 = 2 * ow - 1;
    const uint4* px = a.x + ((f * H + ih0) * W + iw0) * C8 + vec;
    uint32_t okmask = 0;
    uint32_t v[9][4];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tp = kh * 3 + kw;
        const bool ok = (unsigned)(ih0 + kh) < (unsigned)H && (unsigned)(iw0 + kw) < (unsigned)W;
        uint4 q = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // -inf in the flipped domain: never wins
        if (ok) {
          q = __ldg(px + kh * rowp + kw * C8);
          q.x ^= flip[0]; q.y ^= flip[1]; q.z ^= flip[2]; q.w ^= flip[3];
          okmask |= 1u << tp;
        }
        v[tp][0] = q.x; v[tp][1] = q.y; v[tp][2] = q.z; v[tp][3] = q.w;
      }
    }
    uint32_t m[4], a2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      m[q] = v[0][q];
#pragma unroll
      for (int tp = 1; tp < 9; ++tp) m[q] = hmax2_nan(m[q], v[tp][q]);
      a2[q] = 0x00040004u;
    }
#pragma unroll
    for (int tp = 8; tp >= 0; --tp) {                            // the earliest tap equal to the extreme is written last
      const uint32_t on = (okmask >> tp) & 1u ? 0xffffffffu : 0u;
      const uint32_t code = (uint32_t)tp * 0x00010001u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t e = heq2(v[tp][q], m[q]) & on;
        a2[q] = (e & code) | (~e & a2[q]);
      }
    }
    uint32_t o[4];
    const float4 sc0 = s_coef[2 * vec], sc1 = s_coef[2 * vec + 1], sh0 = s_coef[2 * C8 + 2 * vec], sh1 = s_coef[2 * C8 + 2 * vec + 1];
    const float scale[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
    const float shift[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t raw = m[q] ^ flip[q];
      const float lo = fmaxf(fmaf(bf16_lo(raw), scale[2 * q], shift[2 * q]), 0.f);
      const float hi = fmaxf(fmaf(bf16_hi(raw), scale[2 * q + 1], shift[2 * q + 1]), 0.f);
      o[q] = pack_bf16(lo, hi);
      if (!(lo > 0.f)) a2[q] = (a2[q] & 0xffff0000u) | 0x00000009u;     // dead ReLU: no tap receives the gradient
      if (!(hi > 0.f)) a2[q] = (a2[q] & 0x0000ffffu) | 0x00090000u;
    }
    const long long oi = (long long)p * C8 + vec;
    a.y[oi] = make_uint4(o[0], o[1], o[2], o[3]);
    uint2 k;
    k.x = __byte_perm(a2[0], a2[1], 0x6420);
    k.y = __byte_perm(a2[2], a2[3], 0x6420);
    a.idx[oi] = k;
  }
}

struct PoolBnBwdArgs {
  const uint4* g;                        // (F, Ho, Wo, C) gradient of the pooled output
  const uint2* idx;
  const uint4* x;                        // (F, H, W, C) convolution output
  uint4* dx;                             // (F, H, W, C) gradient of the convolution output
  long long F, M;
  int H, W, Ho, Wo, C8, training;
  const float *gamma, *mean, *rstd;
  float *sums, *dgamma, *dbeta;          // sums [2][C]: sum g', sum g' xhat
};

// APPLY = false: sums += (g', g' * xhat) per channel;  APPLY = true: dx = gamma rstd (g' - sum g'/M - xhat sum g' xhat/M).
// g' of a pixel = the pooled gradients of the <= 4 windows that recorded it (fp32 sum, never rounded to bf16 in between).
// Per-channel coefficients live in shared memory (three float4 pairs per use) so that the gathers keep the registers:
//   reduce: xhat = x * k0 + k1                 (k0 = rstd, k1 = -mean rstd)
//   apply : dx = g' * k0 + (x * k1 + k2)       (k0 = gamma rstd, k1 = -k0 rstd c, k2 = -k0 (b + c k1'),  b = sum g'/M, c = sum g' xhat/M)
template <bool APPLY, typename I>
__global__ void __launch_bounds__(256, APPLY ? POOLBN_APPLY_MINB : POOLBN_RED_MINB)
bn_relu_maxpool_bwd_kernel(const PoolBnBwdArgs a) {
  extern __shared__ float4 s_dyn[];                              // [3][2 * C8] coefficients, then the reduction scratch
  const int C8 = a.C8, C = 8 * C8;
  const int vec = threadIdx.x % C8, rowlane = threadIdx.x / C8, rows_par = 256 / C8;
  float* s_k = reinterpret_cast<float*>(s_dyn);
  float* scratch = s_k + 3 * C;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float rs = a.rstd[c], mr = -a.mean[c] * rs;
    if (APPLY) {
      const float inv_m = 1.f / (float)a.M;
      const float ga = a.gamma[c] * rs;
      const float t1 = a.sums[c], t2 = a.sums[C + c];
      const float b = a.training ? t1 * inv_m : 0.f, cc = a.training ? t2 * inv_m : 0.f;
      s_k[c] = ga;
      s_k[C + c] = -ga * cc * rs;
      s_k[2 * C + c] = -ga * (b + cc * mr);
      if (blockIdx.x == 0) { a.dbeta[c] = t1; a.dgamma[c] = t2; }
    } else {
      s_k[c] = rs;
      s_k[C + c] = mr;
    }
  }
  __syncthreads();
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  const int W = a.W, Ho = a.Ho, Wo = a.Wo;
  const int rowo = Wo * C8;
  const long long rowx = (long long)W * C8;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  const uint2 k9 = make_uint2(0x09090909u, 0x09090909u);
  const I nblk = (I)(a.F * Ho * Wo);
  auto use = [&](const float (&gv)[8], const uint4& xq, uint4* out) {
    const uint32_t xw[4] = {xq.x, xq.y, xq.z, xq.w};
    const float4 p0 = s_dyn[2 * vec], p1 = s_dyn[2 * vec + 1], q0 = s_dyn[2 * C8 + 2 * vec], q1 = s_dyn[2 * C8 + 2 * vec + 1];
    const float k0[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    const float k1[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    float o[8];
    if (APPLY) {
      const float4 r0 = s_dyn[4 * C8 + 2 * vec], r1 = s_dyn[4 * C8 + 2 * vec + 1];
      const float k2[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xv = (j & 1) ? bf16_hi(xw[j >> 1]) : bf16_lo(xw[j >> 1]);
        o[j] = fmaf(gv[j], k0[j], fmaf(xv, k1[j], k2[j]));
      }
      *out = pool_pack(o);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xv = (j & 1) ? bf16_hi(xw[j >> 1]) : bf16_lo(xw[j >> 1]);
        s1[j] += gv[j];
        s2[j] = fmaf(gv[j], fmaf(xv, k0[j], k1[j]), s2[j]);
      }
    }
  };
  for (I p = (I)blockIdx.x * (I)rows_par + (I)rowlane; p < nblk; p += (I)gridDim.x * (I)rows_par) {
    const int bb = (int)(p % (I)Wo);
    const I t = p / (I)Wo;
    const int aa = (int)(t % (I)Ho);
    const long long f = (long long)(t / (I)Ho);
    const bool right = bb + 1 < Wo, down = aa + 1 < Ho;
    const long long i = (long long)p * C8 + vec;
    const uint4 g00 = __ldg(a.g + i);
    const uint2 k00 = __ldg(a.idx + i);
    const uint4 g01 = right ? __ldg(a.g + i + C8) : z4;
    const uint2 k01 = right ? __ldg(a.idx + i + C8) : k9;
    const uint4 g10 = down ? __ldg(a.g + i + rowo) : z4;
    const uint2 k10 = down ? __ldg(a.idx + i + rowo) : k9;
    const uint4 g11 = (right && down) ? __ldg(a.g + i + rowo + C8) : z4;
    const uint2 k11 = (right && down) ? __ldg(a.idx + i + rowo + C8) : k9;
    const long long xo = ((f * (2 * Ho) + 2 * aa) * (long long)W + 2 * bb) * C8 + vec;
    const uint4 x00 = __ldg(a.x + xo), x01 = __ldg(a.x + xo + C8), x10 = __ldg(a.x + xo + rowx), x11 = __ldg(a.x + xo + rowx + C8);
    uint4* out = a.dx + xo;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    pool_take(acc, g00, k00, 4u);
    use(acc, x00, out);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    pool_take(acc, g00, k00, 5u); pool_take(acc, g01, k01, 3u);
    use(acc, x01, out + C8);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    pool_take(acc, g00, k00, 7u); pool_take(acc, g10, k10, 1u);
    use(acc, x10, out + rowx);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    pool_take(acc, g00, k00, 8u); pool_take(acc, g01, k01, 6u); pool_take(acc, g10, k10, 2u); pool_take(acc, g11, k11, 0u);
    use(acc, x11, out + rowx + C8);
  }
  if (!APPLY) {                                                  // registers -> shared memory -> one atomic per channel and CTA
    float* mine = scratch + (size_t)threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { mine[j] = s1[j]; mine[8 + j] = s2[j]; }
    __syncthreads();
    for (int i = threadIdx.x; i < C8 * 16; i += 256) {
      const int v = i / 16, q = i - v * 16;
      float s = 0.f;
      for (int r = 0; r < rows_par; ++r) s += scratch[(size_t)(r * C8 + v) * 16 + q];
      atomicAdd(&a.sums[(q >> 3) * C + v * 8 + (q & 7)], s);
    }
  }
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
  if (W % 8 == 0 && F * Ho < (1LL << 31) && !(reinterpret_cast<uintptr_t>(x) & 15)) {   // one CTA per output row
    // right padding: the last chunk of a row ends at element 6*(Wo-1) - 9 + 24 <= 3*W + 11
    const int pitch = ((kRowLead + W * 6 + 24 + 15) / 16) * 16 + 16;
    stem_im2col_rows_kernel<<<(unsigned)(F * Ho), 256, 7 * pitch, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)a, H, W, Ho, Wo, pitch);
    count_launch();
    MVFB_LAUNCH_CHECK();
    return MVFB_OK;
  }
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
  if (H % 2 == 0 && W % 2 == 0) {                                // 2 x 2 input block per thread
    const long long blocks4 = F * Ho * Wo * (C / 8);
    if (blocks4 < (1LL << 31))
      maxpool_bwd_block_kernel<uint32_t><<<(unsigned)ceil_div_ll(blocks4, 256), 256, 0, (cudaStream_t)stream>>>(
          (const uint4*)g, (const uint2*)idx, (uint4*)dx, W, Ho, Wo, C / 8, blocks4);
    else
      maxpool_bwd_block_kernel<long long><<<(unsigned)ceil_div_ll(blocks4, 256), 256, 0, (cudaStream_t)stream>>>(
          (const uint4*)g, (const uint2*)idx, (uint4*)dx, W, Ho, Wo, C / 8, blocks4);
    count_launch();
    MVFB_LAUNCH_CHECK();
    return MVFB_OK;
  }
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

int bn_relu_maxpool_fwd(const mvfb_bn_desc* d, const void* x, long long F, int H, int W, const float* sums, const float* gamma,
                        const float* beta, float* running_mean, float* running_var, float* save_mean, float* save_rstd, void* y,
                        void* idx, mvfb_stream_t stream) {
  MVFB_CHECK(d && x && y && idx && gamma && beta && F > 0 && H > 0 && W > 0, MVFB_ERR_ARG, "bn_relu_maxpool_fwd: bad arguments");
  MVFB_CHECK(d->C > 0 && d->C % 8 == 0 && 256 % (d->C / 8) == 0 && d->relu && d->M == F * H * W, MVFB_ERR_UNSUPPORTED,
             "bn_relu_maxpool_fwd: C / 8 must divide 256, relu = 1, M = F*H*W");
  MVFB_CHECK(d->training ? (sums && save_mean && save_rstd) : (running_mean && running_var), MVFB_ERR_ARG,
             "bn_relu_maxpool_fwd: training needs sums / save_*, inference the running statistics");
  MVFB_CHECK(!((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) && !(reinterpret_cast<uintptr_t>(idx) & 7),
             MVFB_ERR_ARG, "bn_relu_maxpool_fwd: misaligned tensors");
  PoolBnArgs a;
  a.x = (const uint4*)x; a.y = (uint4*)y; a.idx = (uint2*)idx;
  a.F = F; a.M = d->M; a.H = H; a.W = W; a.Ho = (H - 1) / 2 + 1; a.Wo = (W - 1) / 2 + 1; a.C8 = d->C / 8;
  a.training = d->training; a.eps = d->eps; a.momentum = d->momentum;
  a.sums = sums; a.gamma = gamma; a.beta = beta;
  a.running_mean = running_mean; a.running_var = running_var; a.save_mean = save_mean; a.save_rstd = save_rstd;
  const long long npix = F * a.Ho * a.Wo;
  const int rows_par = 256 / a.C8;
  long long blocks = ceil_div_ll(npix, rows_par);
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  if (npix < (1LL << 31) - (1LL << 24))
    bn_relu_maxpool_fwd_kernel<uint32_t><<<(unsigned)blocks, 256, 2 * d->C * sizeof(float), (cudaStream_t)stream>>>(a);
  else
    bn_relu_maxpool_fwd_kernel<long long><<<(unsigned)blocks, 256, 2 * d->C * sizeof(float), (cudaStream_t)stream>>>(a);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int bn_relu_maxpool_bwd(const mvfb_bn_desc* d, const void* g, const void* idx, const void* x, long long F, int H, int W,
                        const float* gamma, const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                        float* sums, mvfb_stream_t stream) {
  MVFB_CHECK(d && g && idx && x && dx && gamma && mean && rstd && dgamma && dbeta && sums && F > 0 && H > 0 && W > 0, MVFB_ERR_ARG,
             "bn_relu_maxpool_bwd: bad arguments");
  MVFB_CHECK(d->C > 0 && d->C % 8 == 0 && 256 % (d->C / 8) == 0 && d->relu && d->M == F * H * W && H % 2 == 0 && W % 2 == 0,
             MVFB_ERR_UNSUPPORTED, "bn_relu_maxpool_bwd: C / 8 must divide 256, relu = 1, M = F*H*W, even H and W");
  MVFB_CHECK(!((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx)) & 15) &&
                 !(reinterpret_cast<uintptr_t>(idx) & 7),
             MVFB_ERR_ARG, "bn_relu_maxpool_bwd: misaligned tensors");
  cudaStream_t st = (cudaStream_t)stream;
  PoolBnBwdArgs a;
  a.g = (const uint4*)g; a.idx = (const uint2*)idx; a.x = (const uint4*)x; a.dx = (uint4*)dx;
  a.F = F; a.M = d->M; a.H = H; a.W = W; a.Ho = H / 2; a.Wo = W / 2; a.C8 = d->C / 8; a.training = d->training;
  a.gamma = gamma; a.mean = mean; a.rstd = rstd; a.sums = sums; a.dgamma = dgamma; a.dbeta = dbeta;
  MVFB_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * d->C, st));
  const long long nblk = F * a.Ho * a.Wo;
  const int rows_par = 256 / a.C8;
  long long blocks = ceil_div_ll(nblk, rows_par);
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  const size_t coef = 3 * (size_t)d->C * sizeof(float), red = coef + 256 * 16 * sizeof(float);   // C <= 2048: <= 40 KB
  if (nblk < (1LL << 31) - (1LL << 24)) {
    bn_relu_maxpool_bwd_kernel<false, uint32_t><<<(unsigned)blocks, 256, red, st>>>(a);
    count_launch();
    MVFB_LAUNCH_CHECK();
    bn_relu_maxpool_bwd_kernel<true, uint32_t><<<(unsigned)blocks, 256, coef, st>>>(a);
  } else {
    bn_relu_maxpool_bwd_kernel<false, long long><<<(unsigned)blocks, 256, red, st>>>(a);
    count_launch();
    MVFB_LAUNCH_CHECK();
    bn_relu_maxpool_bwd_kernel<true, long long><<<(unsigned)blocks, 256, coef, st>>>(a);
  }
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

}  // extern "C"

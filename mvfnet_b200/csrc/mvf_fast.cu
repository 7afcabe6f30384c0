// MVF module, B200 fast path: bf16 NHWC activations, TMA-staged (T,H,W) tiles in shared memory.
//
// Work decomposition.  One CTA owns ONE clip n and ONE group of Cg slab channels for the whole (T,H,W)
// volume, so no halo is ever re-read from HBM:
//   * frames are streamed through a ring of R shared-memory slots by TMA (cp.async.bulk.tensor.4d over the
//     (c, w, h, frame) view of the NHWC tensor, box = Cg x (W+2) x (H+2) x 1 starting at (c0,-1,-1,f)); the
//     TMA unit zero-fills the out-of-bounds ring of the box, which IS the convolution's zero padding in
//     H and W; a permanently-zero slot plays frames t=-1 and t=T (the clip boundary, MVF.py:109);
//   * a thread handles V consecutive channels (16 B / 8 B vectors) of one pixel; its channel vector is
//     fixed for the CTA's lifetime, so the 7 distinct stencil coefficients per channel live in registers;
//   * z = 7-point cross stencil (the three centre taps are pre-summed), then BN scale/shift + hard-swish.
// Train-mode BatchNorm needs a grid-wide reduction between the stencil and the activation: pass 1 writes
// per-(clip, channel) partial sums (no atomics, no memset), pass 2 re-streams the slab (L2-resident) and each
// CTA reduces the N partials of its own channels in its prologue, overlapped with its first TMA loads.
// Backward mirrors this: kernel A reduces (sum du, sum du*z) partials, kernel B rebuilds dz frame by frame in
// a 3-slot fp32 ring (zero halo), accumulates the 7 tap-gradient sums against the staged x neighbours and
// emits dx(t-1) = transposed stencil of dz(t-2..t).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "mvf_internal.cuh"
#include "ptx.cuh"

namespace mvfb {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxRing = 16;

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct Geo {
  int N, T, Cs, H, W;
  int Cg, R;      // channels per CTA, x ring slots
  int Hp, Wp;     // padded frame extents (H+2, W+2)
  int slot_x;     // bytes of one padded bf16 frame slot
  int slot_g;     // bytes of one unpadded bf16 frame slot (backward: dL/dy)
  int slot_dz;    // bytes of one padded fp32 frame slot (backward: dz ring)
};

// ------------------------------------------------------------------------------------------ vectors
template <int V>
struct Vec;
template <>
struct Vec<8> {
  uint4 v;
  __device__ __forceinline__ void load_shared(const uint8_t* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
  }
  static __device__ __forceinline__ void store_global(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = o;
  }
};
template <>
struct Vec<4> {
  uint2 v;
  __device__ __forceinline__ void load_shared(const uint8_t* p) { v = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void unpack(float (&f)[4]) const {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  }
  static __device__ __forceinline__ void store_global(__nv_bfloat16* p, const float (&f)[4]) {
    uint2 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = o;
  }
};

// The 7 distinct coefficients of the cross stencil for V channels: centre (sum of the views' middle taps),
// t-1, t+1, h-1, h+1, w-1, w+1.   z[p] = kc x[p] + kt0 x[t-1] + kt2 x[t+1] + kh0 x[h-1] + ...  (MVF.py:118-120)
template <int V>
struct Coef {
  float c[V], t0[V], t2[V], h0[V], h2[V], w0[V], w2[V];
  __device__ __forceinline__ void load(const float* wt, const float* wh, const float* ww, int ch0) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int c3 = (ch0 + j) * 3;
      t0[j] = round_bf16(wt[c3]); c[j] = round_bf16(wt[c3 + 1]); t2[j] = round_bf16(wt[c3 + 2]);
      h0[j] = h2[j] = w0[j] = w2[j] = 0.f;
      if (wh) { h0[j] = round_bf16(wh[c3]); c[j] += round_bf16(wh[c3 + 1]); h2[j] = round_bf16(wh[c3 + 2]); }
      if (ww) { w0[j] = round_bf16(ww[c3]); c[j] += round_bf16(ww[c3 + 1]); w2[j] = round_bf16(ww[c3 + 2]); }
    }
  }
};

__device__ __forceinline__ float hswish_f(float u) { return u * __saturatef(fmaf(u, 1.f / 6.f, 0.5f)); }
__device__ __forceinline__ float hswish_grad_f(float u) {
  return __saturatef(fmaf(u, 1.f / 6.f, 0.5f)) + ((u > -3.f && u < 3.f) ? u * (1.f / 6.f) : 0.f);
}

// per-thread walk over the items (pixel, channel-vector) of one frame: item i = tid + k*kThreads,
// pixel = i / G, vector = i % G (G = Cg / V, a power of two that divides kThreads -> constant per thread)
struct Walk {
  int h0, w0, dh, dw, items;
  __device__ __forceinline__ void init(int tid, int G, int H, int W) {
    const int lg = 31 - __clz(G);
    const int pix0 = tid >> lg, step = kThreads >> lg;
    h0 = pix0 / W; w0 = pix0 - h0 * W;
    dh = step / W; dw = step - dh * W;
    items = H * W * G;
  }
};

struct Ring {
  uint64_t* bars;      // [R] x-frame barriers (+ [2] g-frame barriers in the backward kernels)
  uint8_t* zero;       // permanently-zero padded frame (t = -1 and t = T)
  uint8_t* slots;      // R slots of slot_x bytes
};

__device__ __forceinline__ void issue_x(const CUtensorMap* tm, const Ring& r, const Geo& g, int n, int c0, int t) {
  const int s = t % g.R;
  mbar_arrive_expect_tx(&r.bars[s], (uint32_t)(g.Hp * g.Wp * g.Cg * 2));
  tma_load_4d(r.slots + (size_t)s * g.slot_x, tm, &r.bars[s], c0, -1, -1, n * g.T + t);
}

template <int V>
__device__ __forceinline__ void stencil(const uint8_t* sc, const uint8_t* sm, const uint8_t* sp, int off, int rowb,
                                        int pixb, const Coef<V>& k, float (&z)[V], float (&xc)[V]) {
  Vec<V> v;
  float f[V];
  v.load_shared(sc + off); v.unpack(xc);
#pragma unroll
  for (int j = 0; j < V; ++j) z[j] = k.c[j] * xc[j];
  v.load_shared(sm + off); v.unpack(f);
#pragma unroll
  for (int j = 0; j < V; ++j) z[j] = fmaf(k.t0[j], f[j], z[j]);
  v.load_shared(sp + off); v.unpack(f);
#pragma unroll
  for (int j = 0; j < V; ++j) z[j] = fmaf(k.t2[j], f[j], z[j]);
  v.load_shared(sc + off - rowb); v.unpack(f);
#pragma unroll
  for (int j = 0; j < V; ++j) z[j] = fmaf(k.h0[j], f[j], z[j]);
  v.load_shared(sc + off + rowb); v.unpack(f);
#pragma unroll
  for (int j = 0; j < V; ++j) z[j] = fmaf(k.h2[j], f[j], z[j]);
  v.load_shared(sc + off - pixb); v.unpack(f);
#pragma unroll
  for (int j = 0; j < V; ++j) z[j] = fmaf(k.w0[j], f[j], z[j]);
  v.load_shared(sc + off + pixb); v.unpack(f);
#pragma unroll
  for (int j = 0; j < V; ++j) z[j] = fmaf(k.w2[j], f[j], z[j]);
}

// sum acc[0..K) over all threads of the CTA that own the same channel vector (tid % G); result for vector g,
// slot q lands in out[g * K + q] (shared memory, valid after the trailing __syncthreads()).
template <int K>
__device__ __forceinline__ void cta_reduce_by_vector(float (&acc)[K], int G, float* scratch /*[kWarps][G][K]*/,
                                                     float* out /*[G][K]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < K; ++q) {
    float a = acc[q];
    for (int o = 16; o >= G; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    acc[q] = a;
  }
  if (lane < G) {
#pragma unroll
    for (int q = 0; q < K; ++q) scratch[(warp * G + lane) * K + q] = acc[q];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * K; i += kThreads) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) a += scratch[w * G * K + i];
    out[i] = a;
  }
  __syncthreads();
}

// sum the per-clip partials [N][Cs][2] of this CTA's Cg channels: dsum[ch*2 + kind] (double, shared)
__device__ __forceinline__ void reduce_partials(const float* partials, int N, int Cs, int c0, int Cg, double* dpart,
                                                double* dsum) {
  const int per = 2 * Cg;                 // values per clip for this CTA (contiguous in memory)
  const int parts = kThreads / per;       // >= 2 for Cg <= 64
  const int k = threadIdx.x % per, part = threadIdx.x / per;
  if (part < parts) {
    double a = 0.0;
    for (int n = part; n < N; n += parts) a += (double)partials[((size_t)n * Cs + c0) * 2 + k];
    dpart[part * per + k] = a;
  }
  __syncthreads();
  if (threadIdx.x < per) {
    double a = 0.0;
    for (int p = 0; p < parts; ++p) a += dpart[p * per + threadIdx.x];
    dsum[threadIdx.x] = a;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------ forward
struct FwdArgs {
  Geo g;
  int use_hs, training;
  float eps, momentum;
  const float *wt, *wh, *ww, *gamma, *beta;
  float *running_mean, *running_var, *save_mean, *save_rstd;
  float* partials;           // [N][Cs][2]  (sum z, sum z^2) per clip
  __nv_bfloat16* y;
  long long y_pix;           // elements between consecutive pixels of y
};

constexpr int PASS_APPLY = 0;   // eval-mode BN (running stats) or no BN at all
constexpr int PASS_STATS = 1;   // train: per-clip partial sums only
constexpr int PASS_TRAIN = 2;   // train: batch statistics from the partials, then apply

__device__ __forceinline__ Ring carve_ring(uint8_t* smem, const Geo& g, uint8_t*& rest) {
  Ring r;
  r.bars = reinterpret_cast<uint64_t*>(smem);
  r.zero = smem + 256;
  r.slots = r.zero + g.slot_x;
  rest = r.slots + (size_t)g.R * g.slot_x;
  return r;
}

template <int PASS>
__global__ void __launch_bounds__(kThreads, 2)
mvf_fast_fwd_kernel(const __grid_constant__ CUtensorMap tmx, const FwdArgs a) {
  constexpr int V = 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  const Geo& g = a.g;
  const int tid = threadIdx.x;
  const int G = g.Cg / V;
  const int ngroups = g.Cs / g.Cg;
  const int n = blockIdx.x / ngroups, cg = blockIdx.x - n * ngroups;
  const int c0 = cg * g.Cg;
  uint8_t* rest;
  Ring ring = carve_ring(smem, g, rest);
  float* s_scale = reinterpret_cast<float*>(rest);            // [Cg]
  float* s_shift = s_scale + g.Cg;                            // [Cg]
  float* s_red = s_shift + g.Cg;                              // [kWarps][G][16] floats, or doubles for partials
  float* s_out = s_red + kWarps * G * 16;                     // [G][16]

  if (tid == 0) {
    tma_prefetch_desc(&tmx);
    for (int s = 0; s < g.R; ++s) mbar_init(&ring.bars[s], 1);
    fence_barrier_init();
  }
  for (int i = tid * 16; i < g.slot_x; i += kThreads * 16) *reinterpret_cast<uint4*>(ring.zero + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (tid == 0) {
    const int pre = g.R < g.T ? g.R : g.T;
    for (int t = 0; t < pre; ++t) issue_x(&tmx, ring, g, n, c0, t);
  }

  // ---- per-channel affine (overlaps with the loads in flight)
  if (PASS == PASS_TRAIN) {
    double* dpart = reinterpret_cast<double*>(s_red);
    double* dsum = dpart + kThreads;
    reduce_partials(a.partials, g.N, g.Cs, c0, g.Cg, dpart, dsum);
    if (tid < g.Cg) {
      const int c = c0 + tid;
      const double m = (double)g.N * g.T * g.H * g.W;
      const double mu = dsum[tid * 2] / m;
      double var = dsum[tid * 2 + 1] / m - mu * mu;
      if (var < 0) var = 0;
      const float mean = (float)mu, rstd = (float)(1.0 / sqrt(var + (double)a.eps));
      const float sc = a.gamma[c] * rstd;
      s_scale[tid] = sc;
      s_shift[tid] = a.beta[c] - mean * sc;
      if (n == 0) {
        a.save_mean[c] = mean;
        a.save_rstd[c] = rstd;
        if (a.running_mean) {
          const double unb = m > 1 ? var * m / (m - 1) : var;
          a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mean;
          a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unb;
        }
      }
    }
    __syncthreads();
  } else if (PASS == PASS_APPLY) {
    if (tid < g.Cg) {
      const int c = c0 + tid;
      float sc = 1.f, sh = 0.f;
      if (a.use_hs) {
        const float mean = a.running_mean[c], rstd = 1.f / sqrtf(a.running_var[c] + a.eps);
        sc = a.gamma[c] * rstd;
        sh = a.beta[c] - mean * sc;
        if (n == 0 && a.save_mean) { a.save_mean[c] = mean; a.save_rstd[c] = rstd; }
      }
      s_scale[tid] = sc;
      s_shift[tid] = sh;
    }
    __syncthreads();
  }

  const int vec = tid % G, ch0 = c0 + vec * V;
  Coef<V> k;
  k.load(a.wt, a.wh, a.ww, ch0);
  float scale[V], shift[V];
  if (PASS != PASS_STATS) {
#pragma unroll
    for (int j = 0; j < V; ++j) { scale[j] = s_scale[vec * V + j]; shift[j] = s_shift[vec * V + j]; }
  }
  float acc[2 * V];
#pragma unroll
  for (int j = 0; j < 2 * V; ++j) acc[j] = 0.f;

  Walk wk;
  wk.init(tid, G, g.H, g.W);
  const int pixb = g.Cg * 2, rowb = g.Wp * pixb;
  const bool reload = g.T > g.R;

  for (int t = 0; t < g.T; ++t) {
    if (t == 0) mbar_wait(&ring.bars[0], 0);
    if (t + 1 < g.T) mbar_wait(&ring.bars[(t + 1) % g.R], ((t + 1) / g.R) & 1);
    const uint8_t* sc = ring.slots + (size_t)(t % g.R) * g.slot_x;
    const uint8_t* sm = t > 0 ? ring.slots + (size_t)((t - 1) % g.R) * g.slot_x : ring.zero;
    const uint8_t* sp = t + 1 < g.T ? ring.slots + (size_t)((t + 1) % g.R) * g.slot_x : ring.zero;
    __nv_bfloat16* yf = a.y + (size_t)(n * g.T + t) * g.H * g.W * a.y_pix + ch0;
    int h = wk.h0, w = wk.w0;
    for (int i = tid; i < wk.items; i += kThreads) {
      const int off = ((h + 1) * g.Wp + (w + 1)) * pixb + vec * (V * 2);
      float z[V], xc[V];
      stencil<V>(sc, sm, sp, off, rowb, pixb, k, z, xc);
      if (PASS == PASS_STATS) {
#pragma unroll
        for (int j = 0; j < V; ++j) { acc[j] += z[j]; acc[V + j] = fmaf(z[j], z[j], acc[V + j]); }
      } else {
        if (a.use_hs) {
#pragma unroll
          for (int j = 0; j < V; ++j) z[j] = hswish_f(fmaf(z[j], scale[j], shift[j]));
        }
        Vec<V>::store_global(yf + (size_t)(h * g.W + w) * a.y_pix, z);
      }
      w += wk.dw; h += wk.dh;
      if (w >= g.W) { w -= g.W; h += 1; }
    }
    if (reload) {
      __syncthreads();                       // every thread is done with frame t-1's slot
      if (tid == 0 && t >= 1 && t - 1 + g.R < g.T) issue_x(&tmx, ring, g, n, c0, t - 1 + g.R);
    }
  }

  if (PASS == PASS_STATS) {
    cta_reduce_by_vector<2 * V>(acc, G, s_red, s_out);
    // s_out[vec*16 + kind*8 + j]  ->  partials[n][c0 + vec*8 + j][kind]
    for (int i = tid; i < 2 * g.Cg; i += kThreads) {
      const int ch = i >> 1, kind = i & 1;
      a.partials[((size_t)n * g.Cs + c0) * 2 + i] = s_out[(ch / V) * (2 * V) + kind * V + (ch % V)];
    }
  }
}

// ------------------------------------------------------------------------------------------ forward, resident clip
// Second-generation forward kernel for T in {4, 8, 16} when the whole clip volume of the channel group fits
// in shared memory.  The first kernel above is bound by instruction issue (40 lane-instructions per element,
// ncu r1d), not by HBM, so this one minimises instructions per element:
//   * every thread owns fixed (pixel, 8-channel vector) items and walks t = 0..T-1 with the unpacked centre
//     values of frames t-1, t, t+1 rolling through registers (fully unrolled: no moves), so only 5 instead of
//     7 shared-memory vectors are loaded and unpacked per output vector;
//   * all arithmetic is packed fp32x2 (FFMA2 on sm_100): 28 instead of 56 FMAs per vector;
//   * no per-frame barriers: the T frame loads are all in flight from the first instruction of the CTA and
//     are waited for once; the temporal zero padding lives in registers, so no zero slot is needed.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }

struct F8 {
  float2 p[4];
};
__device__ __forceinline__ F8 lds_unpack8(const uint8_t* ptr) {
  const uint4 v = *reinterpret_cast<const uint4*>(ptr);
  F8 r;
  r.p[0] = make_float2(bf16_lo(v.x), bf16_hi(v.x));
  r.p[1] = make_float2(bf16_lo(v.y), bf16_hi(v.y));
  r.p[2] = make_float2(bf16_lo(v.z), bf16_hi(v.z));
  r.p[3] = make_float2(bf16_lo(v.w), bf16_hi(v.w));
  return r;
}
__device__ __forceinline__ F8 zero8() {
  F8 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) r.p[j] = make_float2(0.f, 0.f);
  return r;
}
struct Coef8 {
  F8 c, t0, t2, h0, h2, w0, w2;
  __device__ __forceinline__ void load(const float* wt, const float* wh, const float* ww, int ch0) {
    float a[7][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c3 = (ch0 + j) * 3;
      a[1][j] = round_bf16(wt[c3]); a[0][j] = round_bf16(wt[c3 + 1]); a[2][j] = round_bf16(wt[c3 + 2]);
      a[3][j] = a[4][j] = a[5][j] = a[6][j] = 0.f;
      if (wh) { a[3][j] = round_bf16(wh[c3]); a[0][j] += round_bf16(wh[c3 + 1]); a[4][j] = round_bf16(wh[c3 + 2]); }
      if (ww) { a[5][j] = round_bf16(ww[c3]); a[0][j] += round_bf16(ww[c3 + 1]); a[6][j] = round_bf16(ww[c3 + 2]); }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      c.p[j] = make_float2(a[0][2 * j], a[0][2 * j + 1]);
      t0.p[j] = make_float2(a[1][2 * j], a[1][2 * j + 1]);
      t2.p[j] = make_float2(a[2][2 * j], a[2][2 * j + 1]);
      h0.p[j] = make_float2(a[3][2 * j], a[3][2 * j + 1]);
      h2.p[j] = make_float2(a[4][2 * j], a[4][2 * j + 1]);
      w0.p[j] = make_float2(a[5][2 * j], a[5][2 * j + 1]);
      w2.p[j] = make_float2(a[6][2 * j], a[6][2 * j + 1]);
    }
  }
};

template <int PASS, int T>
__global__ void __launch_bounds__(kThreads, 2)
mvf_fwd_resident_kernel(const __grid_constant__ CUtensorMap tmx, const FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Geo& g = a.g;
  const int tid = threadIdx.x;
  const int G = g.Cg / 8;
  const int ngroups = g.Cs / g.Cg;
  const int n = blockIdx.x / ngroups, cg = blockIdx.x - n * ngroups;
  const int c0 = cg * g.Cg;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* slots = smem + 256;
  uint8_t* rest = slots + (size_t)T * g.slot_x;
  float* s_scale = reinterpret_cast<float*>(rest);            // [Cg]
  float* s_shift = s_scale + g.Cg;                            // [Cg]
  float* s_red = s_shift + g.Cg;                              // [kWarps][G][16] floats / doubles for partials
  float* s_out = s_red + kWarps * G * 16;                     // [G][16]

  if (tid == 0) {
    tma_prefetch_desc(&tmx);
#pragma unroll
    for (int t = 0; t < T; ++t) mbar_init(&bars[t], 1);
    fence_barrier_init();
#pragma unroll
    for (int t = 0; t < T; ++t) {
      mbar_arrive_expect_tx(&bars[t], (uint32_t)(g.Hp * g.Wp * g.Cg * 2));
      tma_load_4d(slots + (size_t)t * g.slot_x, &tmx, &bars[t], c0, -1, -1, n * T + t);
    }
  }
  __syncthreads();                                            // barrier inits visible before anyone waits

  if (PASS == PASS_TRAIN) {
    double* dpart = reinterpret_cast<double*>(s_red);
    double* dsum = dpart + kThreads;
    reduce_partials(a.partials, g.N, g.Cs, c0, g.Cg, dpart, dsum);
    if (tid < g.Cg) {
      const int c = c0 + tid;
      const double m = (double)g.N * T * g.H * g.W;
      const double mu = dsum[tid * 2] / m;
      double var = dsum[tid * 2 + 1] / m - mu * mu;
      if (var < 0) var = 0;
      const float mean = (float)mu, rstd = (float)(1.0 / sqrt(var + (double)a.eps));
      const float sc = a.gamma[c] * rstd;
      s_scale[tid] = sc;
      s_shift[tid] = a.beta[c] - mean * sc;
      if (n == 0) {
        a.save_mean[c] = mean;
        a.save_rstd[c] = rstd;
        if (a.running_mean) {
          const double unb = m > 1 ? var * m / (m - 1) : var;
          a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mean;
          a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unb;
        }
      }
    }
    __syncthreads();
  } else if (PASS == PASS_APPLY) {
    if (tid < g.Cg) {
      const int c = c0 + tid;
      float sc = 1.f, sh = 0.f;
      if (a.use_hs) {
        const float mean = a.running_mean[c], rstd = 1.f / sqrtf(a.running_var[c] + a.eps);
        sc = a.gamma[c] * rstd;
        sh = a.beta[c] - mean * sc;
        if (n == 0 && a.save_mean) { a.save_mean[c] = mean; a.save_rstd[c] = rstd; }
      }
      s_scale[tid] = sc;
      s_shift[tid] = sh;
    }
    __syncthreads();
  }

  const int vec = tid % G, ch0 = c0 + vec * 8;
  Coef8 k;
  k.load(a.wt, a.wh, a.ww, ch0);
  F8 scale, shift;
  if (PASS != PASS_STATS) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      scale.p[j] = make_float2(s_scale[vec * 8 + 2 * j], s_scale[vec * 8 + 2 * j + 1]);
      shift.p[j] = make_float2(s_shift[vec * 8 + 2 * j], s_shift[vec * 8 + 2 * j + 1]);
    }
  }
  F8 sum = zero8(), sq = zero8();
  Walk wk;
  wk.init(tid, G, g.H, g.W);
  const int pixb = g.Cg * 2, rowb = g.Wp * pixb;
  const size_t frame_elems = (size_t)g.H * g.W * a.y_pix;
  const bool hs = a.use_hs != 0;

#pragma unroll 1
  for (int t = 0; t < T; ++t) mbar_wait(&bars[t], 0);

  int h = wk.h0, w = wk.w0;
#pragma unroll 1
  for (int i = tid; i < wk.items; i += kThreads) {
    const uint8_t* p0 = slots + ((h + 1) * g.Wp + (w + 1)) * pixb + vec * 16;
    __nv_bfloat16* yp = a.y + ((size_t)n * T * g.H * g.W + (size_t)(h * g.W + w)) * a.y_pix + ch0;
    F8 xm = zero8(), xc = lds_unpack8(p0), xp;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const uint8_t* pt = p0 + (size_t)t * g.slot_x;
      xp = (t + 1 < T) ? lds_unpack8(pt + g.slot_x) : zero8();
      F8 z;
#pragma unroll
      for (int j = 0; j < 4; ++j) z.p[j] = ffma2(k.t2.p[j], xp.p[j], ffma2(k.t0.p[j], xm.p[j], fmul2(k.c.p[j], xc.p[j])));
      F8 f = lds_unpack8(pt - rowb);
#pragma unroll
      for (int j = 0; j < 4; ++j) z.p[j] = ffma2(k.h0.p[j], f.p[j], z.p[j]);
      f = lds_unpack8(pt + rowb);
#pragma unroll
      for (int j = 0; j < 4; ++j) z.p[j] = ffma2(k.h2.p[j], f.p[j], z.p[j]);
      f = lds_unpack8(pt - pixb);
#pragma unroll
      for (int j = 0; j < 4; ++j) z.p[j] = ffma2(k.w0.p[j], f.p[j], z.p[j]);
      f = lds_unpack8(pt + pixb);
#pragma unroll
      for (int j = 0; j < 4; ++j) z.p[j] = ffma2(k.w2.p[j], f.p[j], z.p[j]);
      if (PASS == PASS_STATS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sum.p[j] = fadd2(sum.p[j], z.p[j]);
          sq.p[j] = ffma2(z.p[j], z.p[j], sq.p[j]);
        }
      } else {
        if (hs) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 u = ffma2(z.p[j], scale.p[j], shift.p[j]);
            float2 s;
            s.x = __saturatef(fmaf(u.x, 1.f / 6.f, 0.5f));
            s.y = __saturatef(fmaf(u.y, 1.f / 6.f, 0.5f));
            z.p[j] = fmul2(u, s);
          }
        }
        uint4 o;
        o.x = pack_bf16(z.p[0].x, z.p[0].y); o.y = pack_bf16(z.p[1].x, z.p[1].y);
        o.z = pack_bf16(z.p[2].x, z.p[2].y); o.w = pack_bf16(z.p[3].x, z.p[3].y);
        *reinterpret_cast<uint4*>(yp + (size_t)t * frame_elems) = o;
      }
      xm = xc;
      xc = xp;
    }
    w += wk.dw; h += wk.dh;
    if (w >= g.W) { w -= g.W; h += 1; }
  }

  if (PASS == PASS_STATS) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[2 * j] = sum.p[j].x; acc[2 * j + 1] = sum.p[j].y;
      acc[8 + 2 * j] = sq.p[j].x; acc[8 + 2 * j + 1] = sq.p[j].y;
    }
    cta_reduce_by_vector<16>(acc, G, s_red, s_out);
    for (int i = tid; i < 2 * g.Cg; i += kThreads) {
      const int ch = i >> 1, kind = i & 1;
      a.partials[((size_t)n * g.Cs + c0) * 2 + i] = s_out[(ch / 8) * 16 + kind * 8 + (ch % 8)];
    }
  }
}

// ------------------------------------------------------------------------------------------ backward
struct BwdArgs {
  Geo g;
  int use_hs, training, share_h, share_w;
  const float *wt, *wh, *ww, *gamma, *beta, *mean, *rstd;
  float* partials;           // [N][Cs][2]  (sum du, sum du*z) per clip
  float *dwt, *dwh, *dww, *dgamma, *dbeta;
  __nv_bfloat16* dx;
  long long dx_pix;
};

__device__ __forceinline__ void issue_g(const CUtensorMap* tm, uint64_t* bars, uint8_t* gslots, const Geo& g, int n,
                                        int c0, int t) {
  const int s = t & 1;
  mbar_arrive_expect_tx(&bars[s], (uint32_t)(g.H * g.W * g.Cg * 2));
  tma_load_4d(gslots + (size_t)s * g.slot_g, tm, &bars[s], c0, 0, 0, n * g.T + t);
}

// Kernel A: partial sums of du and du*z per (clip, channel); also zeroes the tap-gradient outputs that
// kernel B accumulates into with atomics.
__global__ void __launch_bounds__(kThreads, 2)
mvf_fast_bwd_reduce_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmg,
                           const BwdArgs a) {
  constexpr int V = 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  const Geo& g = a.g;
  const int tid = threadIdx.x;
  const int G = g.Cg / V;
  const int ngroups = g.Cs / g.Cg;
  const int n = blockIdx.x / ngroups, cg = blockIdx.x - n * ngroups;
  const int c0 = cg * g.Cg;
  uint8_t* rest;
  Ring ring = carve_ring(smem, g, rest);
  uint64_t* gbars = ring.bars + kMaxRing;
  uint8_t* gslots = rest;
  float* s_red = reinterpret_cast<float*>(gslots + 2 * (size_t)g.slot_g);
  float* s_out = s_red + kWarps * G * 16;

  if (tid == 0) {
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmg);
    for (int s = 0; s < g.R; ++s) mbar_init(&ring.bars[s], 1);
    mbar_init(&gbars[0], 1);
    mbar_init(&gbars[1], 1);
    fence_barrier_init();
  }
  for (int i = tid * 16; i < g.slot_x; i += kThreads * 16) *reinterpret_cast<uint4*>(ring.zero + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (tid == 0) {
    const int pre = g.R < g.T ? g.R : g.T;
    for (int t = 0; t < pre; ++t) issue_x(&tmx, ring, g, n, c0, t);
    issue_g(&tmg, gbars, gslots, g, n, c0, 0);
    if (g.T > 1) issue_g(&tmg, gbars, gslots, g, n, c0, 1);
  }
  if (n == 0) {
    for (int i = tid; i < g.Cg * 3; i += kThreads) {
      a.dwt[c0 * 3 + i] = 0.f;
      if (a.dwh) a.dwh[c0 * 3 + i] = 0.f;
      if (a.dww) a.dww[c0 * 3 + i] = 0.f;
    }
  }

  const int vec = tid % G, ch0 = c0 + vec * V;
  Coef<V> k;
  k.load(a.wt, a.wh, a.ww, ch0);
  float scale[V], shift[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    scale[j] = a.gamma[ch0 + j] * a.rstd[ch0 + j];
    shift[j] = a.beta[ch0 + j] - a.mean[ch0 + j] * scale[j];
  }
  float acc[2 * V];
#pragma unroll
  for (int j = 0; j < 2 * V; ++j) acc[j] = 0.f;
  Walk wk;
  wk.init(tid, G, g.H, g.W);
  const int pixb = g.Cg * 2, rowb = g.Wp * pixb;

  for (int t = 0; t < g.T; ++t) {
    if (t == 0) mbar_wait(&ring.bars[0], 0);
    if (t + 1 < g.T) mbar_wait(&ring.bars[(t + 1) % g.R], ((t + 1) / g.R) & 1);
    mbar_wait(&gbars[t & 1], (t >> 1) & 1);
    const uint8_t* sc = ring.slots + (size_t)(t % g.R) * g.slot_x;
    const uint8_t* sm = t > 0 ? ring.slots + (size_t)((t - 1) % g.R) * g.slot_x : ring.zero;
    const uint8_t* sp = t + 1 < g.T ? ring.slots + (size_t)((t + 1) % g.R) * g.slot_x : ring.zero;
    const uint8_t* sg = gslots + (size_t)(t & 1) * g.slot_g;
    int h = wk.h0, w = wk.w0;
    for (int i = tid; i < wk.items; i += kThreads) {
      const int off = ((h + 1) * g.Wp + (w + 1)) * pixb + vec * (V * 2);
      float z[V], xc[V], gv[V];
      stencil<V>(sc, sm, sp, off, rowb, pixb, k, z, xc);
      Vec<V> v;
      v.load_shared(sg + (h * g.W + w) * pixb + vec * (V * 2));
      v.unpack(gv);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float du = gv[j] * hswish_grad_f(fmaf(z[j], scale[j], shift[j]));
        acc[j] += du;
        acc[V + j] = fmaf(du, z[j], acc[V + j]);
      }
      w += wk.dw; h += wk.dh;
      if (w >= g.W) { w -= g.W; h += 1; }
    }
    __syncthreads();
    if (tid == 0) {
      if (t >= 1 && t - 1 + g.R < g.T) issue_x(&tmx, ring, g, n, c0, t - 1 + g.R);
      if (t + 2 < g.T) issue_g(&tmg, gbars, gslots, g, n, c0, t + 2);
    }
  }
  cta_reduce_by_vector<2 * V>(acc, G, s_red, s_out);
  for (int i = tid; i < 2 * g.Cg; i += kThreads) {
    const int ch = i >> 1, kind = i & 1;
    a.partials[((size_t)n * g.Cs + c0) * 2 + i] = s_out[(ch / V) * (2 * V) + kind * V + (ch % V)];
  }
}

// Kernel B: dz (BatchNorm + hard-swish backward), tap-gradient sums and dx.
__global__ void __launch_bounds__(kThreads, 1)
mvf_fast_bwd_dx_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmg,
                       const BwdArgs a) {
  constexpr int V = 4;
  extern __shared__ __align__(1024) uint8_t smem[];
  const Geo& g = a.g;
  const int tid = threadIdx.x;
  const int G = g.Cg / V;
  const int ngroups = g.Cs / g.Cg;
  const int n = blockIdx.x / ngroups, cg = blockIdx.x - n * ngroups;
  const int c0 = cg * g.Cg;
  uint8_t* rest;
  Ring ring = carve_ring(smem, g, rest);
  uint64_t* gbars = ring.bars + kMaxRing;
  uint8_t* gslots = rest;
  uint8_t* dzs = gslots + 2 * (size_t)g.slot_g;               // [4] padded fp32 frames; slot 3 stays zero
  float* s_coef = reinterpret_cast<float*>(dzs + 4 * (size_t)g.slot_dz);   // scale, shift, c1, c0 : [4][Cg]
  float* s_red = s_coef + 4 * g.Cg;                           // [kWarps][G][7*V] floats / doubles for partials
  float* s_out = s_red + kWarps * G * 7 * V;                  // [G][7*V]

  if (tid == 0) {
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmg);
    for (int s = 0; s < g.R; ++s) mbar_init(&ring.bars[s], 1);
    mbar_init(&gbars[0], 1);
    mbar_init(&gbars[1], 1);
    fence_barrier_init();
  }
  for (int i = tid * 16; i < g.slot_x; i += kThreads * 16) *reinterpret_cast<uint4*>(ring.zero + i) = make_uint4(0, 0, 0, 0);
  for (int i = tid * 16; i < 4 * g.slot_dz; i += kThreads * 16) *reinterpret_cast<uint4*>(dzs + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (tid == 0) {
    const int pre = g.R < g.T ? g.R : g.T;
    for (int t = 0; t < pre; ++t) issue_x(&tmx, ring, g, n, c0, t);
    issue_g(&tmg, gbars, gslots, g, n, c0, 0);
    if (g.T > 1) issue_g(&tmg, gbars, gslots, g, n, c0, 1);
  }

  // ---- per-channel constants:  u = z*scale + shift ;  dz = scale*du + u*c1 + c0   (c1 = c0 = 0 in eval mode)
  if (a.use_hs) {
    double* dpart = reinterpret_cast<double*>(s_red);
    double* dsum = dpart + kThreads;
    reduce_partials(a.partials, g.N, g.Cs, c0, g.Cg, dpart, dsum);
    if (tid < g.Cg) {
      const int c = c0 + tid;
      const double s1 = dsum[tid * 2], s2 = dsum[tid * 2 + 1];
      const double mean = a.mean[c], rstd = a.rstd[c], gam = a.gamma[c], bet = a.beta[c];
      const double dgamma = rstd * (s2 - mean * s1);
      const double m = (double)g.N * g.T * g.H * g.W;
      const double sc = gam * rstd;
      double c1 = 0.0, cc0 = 0.0;
      if (a.training) {
        c1 = -rstd * dgamma / m;
        cc0 = -sc * s1 / m + bet * rstd * dgamma / m;
      }
      s_coef[tid] = (float)sc;
      s_coef[g.Cg + tid] = (float)(bet - mean * sc);
      s_coef[2 * g.Cg + tid] = (float)c1;
      s_coef[3 * g.Cg + tid] = (float)cc0;
      if (n == 0) { a.dgamma[c] = (float)dgamma; a.dbeta[c] = (float)s1; }
    }
    __syncthreads();
  }

  const int vec = tid % G, ch0 = c0 + vec * V;
  Coef<V> k;
  k.load(a.wt, a.wh, a.ww, ch0);
  float scale[V], shift[V], c1[V], cc0[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    if (a.use_hs) {
      scale[j] = s_coef[vec * V + j]; shift[j] = s_coef[g.Cg + vec * V + j];
      c1[j] = s_coef[2 * g.Cg + vec * V + j]; cc0[j] = s_coef[3 * g.Cg + vec * V + j];
    } else {
      scale[j] = 1.f; shift[j] = 0.f; c1[j] = 0.f; cc0[j] = 0.f;
    }
  }
  // tap-gradient accumulators: centre, t-1, t+1, h-1, h+1, w-1, w+1   (dw[k] = sum dz[p] * x[p + (k-1)])
  float acc[7 * V];
#pragma unroll
  for (int j = 0; j < 7 * V; ++j) acc[j] = 0.f;

  Walk wk;
  wk.init(tid, G, g.H, g.W);
  const int pixb = g.Cg * 2, rowb = g.Wp * pixb;         // bf16 frames
  const int dpixb = g.Cg * 4, drowb = g.Wp * dpixb;      // fp32 dz frames
  const uint8_t* dzero = dzs + 3 * (size_t)g.slot_dz;

  for (int t = 0; t <= g.T; ++t) {
    if (t < g.T) {
      if (t == 0) mbar_wait(&ring.bars[0], 0);
      if (t + 1 < g.T) mbar_wait(&ring.bars[(t + 1) % g.R], ((t + 1) / g.R) & 1);
      mbar_wait(&gbars[t & 1], (t >> 1) & 1);
      const uint8_t* sc = ring.slots + (size_t)(t % g.R) * g.slot_x;
      const uint8_t* sm = t > 0 ? ring.slots + (size_t)((t - 1) % g.R) * g.slot_x : ring.zero;
      const uint8_t* sp = t + 1 < g.T ? ring.slots + (size_t)((t + 1) % g.R) * g.slot_x : ring.zero;
      const uint8_t* sg = gslots + (size_t)(t & 1) * g.slot_g;
      uint8_t* dzt = dzs + (size_t)(t % 3) * g.slot_dz;
      int h = wk.h0, w = wk.w0;
      for (int i = tid; i < wk.items; i += kThreads) {
        const int pp = (h + 1) * g.Wp + (w + 1);
        const int off = pp * pixb + vec * (V * 2);
        Vec<V> v;
        float xc[V], f[V], z[V], dz[V];
        v.load_shared(sc + off); v.unpack(xc);
        v.load_shared(sg + (h * g.W + w) * pixb + vec * (V * 2)); v.unpack(dz);   // dz <- g for now
        if (a.use_hs) {
          // z needs all seven neighbours first; the tap sums need dz, so stage the neighbours twice (shared
          // memory reads are cheap next to keeping 7*V more registers alive)
          float dummy[V];
          stencil<V>(sc, sm, sp, off, rowb, pixb, k, z, dummy);
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const float u = fmaf(z[j], scale[j], shift[j]);
            const float du = dz[j] * hswish_grad_f(u);
            dz[j] = fmaf(scale[j], du, fmaf(u, c1[j], cc0[j]));
          }
        }
        *reinterpret_cast<float4*>(dzt + pp * dpixb + vec * (V * 4)) = make_float4(dz[0], dz[1], dz[2], dz[3]);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = fmaf(dz[j], xc[j], acc[j]);
        v.load_shared(sm + off); v.unpack(f);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[V + j] = fmaf(dz[j], f[j], acc[V + j]);
        v.load_shared(sp + off); v.unpack(f);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[2 * V + j] = fmaf(dz[j], f[j], acc[2 * V + j]);
        v.load_shared(sc + off - rowb); v.unpack(f);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[3 * V + j] = fmaf(dz[j], f[j], acc[3 * V + j]);
        v.load_shared(sc + off + rowb); v.unpack(f);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[4 * V + j] = fmaf(dz[j], f[j], acc[4 * V + j]);
        v.load_shared(sc + off - pixb); v.unpack(f);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[5 * V + j] = fmaf(dz[j], f[j], acc[5 * V + j]);
        v.load_shared(sc + off + pixb); v.unpack(f);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[6 * V + j] = fmaf(dz[j], f[j], acc[6 * V + j]);
        w += wk.dw; h += wk.dh;
        if (w >= g.W) { w -= g.W; h += 1; }
      }
    }
    __syncthreads();                       // dz(t) complete; x(t-1) and g(t) slots are free
    if (tid == 0 && t < g.T) {
      if (t >= 1 && t - 1 + g.R < g.T) issue_x(&tmx, ring, g, n, c0, t - 1 + g.R);
      if (t + 2 < g.T) issue_g(&tmg, gbars, gslots, g, n, c0, t + 2);
    }
    if (t >= 1) {
      // dx(t-1) = kc dz(t-1) + kt0 dz(t) + kt2 dz(t-2) + kh0 dz[h+1] + kh2 dz[h-1] + kw0 dz[w+1] + kw2 dz[w-1]
      const int tt = t - 1;
      const uint8_t* d0 = dzs + (size_t)(tt % 3) * g.slot_dz;
      const uint8_t* dn = t < g.T ? dzs + (size_t)(t % 3) * g.slot_dz : dzero;
      const uint8_t* dp = tt >= 1 ? dzs + (size_t)((tt - 1) % 3) * g.slot_dz : dzero;
      __nv_bfloat16* dxf = a.dx + (size_t)(n * g.T + tt) * g.H * g.W * a.dx_pix + ch0;
      int h = wk.h0, w = wk.w0;
      for (int i = tid; i < wk.items; i += kThreads) {
        const int off = ((h + 1) * g.Wp + (w + 1)) * dpixb + vec * (V * 4);
        float4 q;
        float o[V];
        q = *reinterpret_cast<const float4*>(d0 + off);
        o[0] = k.c[0] * q.x; o[1] = k.c[1] * q.y; o[2] = k.c[2] * q.z; o[3] = k.c[3] * q.w;
        q = *reinterpret_cast<const float4*>(dn + off);
        o[0] = fmaf(k.t0[0], q.x, o[0]); o[1] = fmaf(k.t0[1], q.y, o[1]); o[2] = fmaf(k.t0[2], q.z, o[2]); o[3] = fmaf(k.t0[3], q.w, o[3]);
        q = *reinterpret_cast<const float4*>(dp + off);
        o[0] = fmaf(k.t2[0], q.x, o[0]); o[1] = fmaf(k.t2[1], q.y, o[1]); o[2] = fmaf(k.t2[2], q.z, o[2]); o[3] = fmaf(k.t2[3], q.w, o[3]);
        q = *reinterpret_cast<const float4*>(d0 + off + drowb);
        o[0] = fmaf(k.h0[0], q.x, o[0]); o[1] = fmaf(k.h0[1], q.y, o[1]); o[2] = fmaf(k.h0[2], q.z, o[2]); o[3] = fmaf(k.h0[3], q.w, o[3]);
        q = *reinterpret_cast<const float4*>(d0 + off - drowb);
        o[0] = fmaf(k.h2[0], q.x, o[0]); o[1] = fmaf(k.h2[1], q.y, o[1]); o[2] = fmaf(k.h2[2], q.z, o[2]); o[3] = fmaf(k.h2[3], q.w, o[3]);
        q = *reinterpret_cast<const float4*>(d0 + off + dpixb);
        o[0] = fmaf(k.w0[0], q.x, o[0]); o[1] = fmaf(k.w0[1], q.y, o[1]); o[2] = fmaf(k.w0[2], q.z, o[2]); o[3] = fmaf(k.w0[3], q.w, o[3]);
        q = *reinterpret_cast<const float4*>(d0 + off - dpixb);
        o[0] = fmaf(k.w2[0], q.x, o[0]); o[1] = fmaf(k.w2[1], q.y, o[1]); o[2] = fmaf(k.w2[2], q.z, o[2]); o[3] = fmaf(k.w2[3], q.w, o[3]);
        Vec<V>::store_global(dxf + (size_t)(h * g.W + w) * a.dx_pix, o);
        w += wk.dw; h += wk.dh;
        if (w >= g.W) { w -= g.W; h += 1; }
      }
      __syncthreads();                     // dz(t-2)'s slot may be overwritten by dz(t+1)
    }
  }

  // ---- tap gradients: CTA reduction, then one atomic per (channel, tap) per CTA
  cta_reduce_by_vector<7 * V>(acc, G, s_red, s_out);
  float* dst_h = a.share_h ? a.dwt : a.dwh;
  float* dst_w = a.share_w ? a.dwt : a.dww;
  for (int i = tid; i < g.Cg * 7; i += kThreads) {
    const int ch = i / 7, q = i - ch * 7;
    const float val = s_out[(ch / V) * (7 * V) + q * V + (ch % V)];
    const int c3 = (c0 + ch) * 3;
    switch (q) {
      case 0:
        atomicAdd(&a.dwt[c3 + 1], val);
        if (dst_h) atomicAdd(&dst_h[c3 + 1], val);
        if (dst_w) atomicAdd(&dst_w[c3 + 1], val);
        break;
      case 1: atomicAdd(&a.dwt[c3 + 0], val); break;
      case 2: atomicAdd(&a.dwt[c3 + 2], val); break;
      case 3: if (dst_h) atomicAdd(&dst_h[c3 + 0], val); break;
      case 4: if (dst_h) atomicAdd(&dst_h[c3 + 2], val); break;
      case 5: if (dst_w) atomicAdd(&dst_w[c3 + 0], val); break;
      default: if (dst_w) atomicAdd(&dst_w[c3 + 2], val); break;
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
size_t fwd_smem(const Geo& g) {
  const int G = g.Cg / 8;
  return 256 + (size_t)(g.R + 1) * g.slot_x + 2 * g.Cg * 4 + (size_t)kWarps * G * 16 * 4 + G * 16 * 4 +
         3 * kThreads * 8 /* double scratch of reduce_partials aliases s_red; keep room */;
}
size_t bwdA_smem(const Geo& g) {
  const int G = g.Cg / 8;
  return 256 + (size_t)(g.R + 1) * g.slot_x + 2 * (size_t)g.slot_g + (size_t)kWarps * G * 16 * 4 + G * 16 * 4 + 64;
}
size_t bwdB_smem(const Geo& g) {
  const int G = g.Cg / 4;
  return 256 + (size_t)(g.R + 1) * g.slot_x + 2 * (size_t)g.slot_g + 4 * (size_t)g.slot_dz + 4 * g.Cg * 4 +
         (size_t)kWarps * G * 28 * 4 + G * 28 * 4 + 3 * kThreads * 8;
}

void fill_geo(Geo& g, const mvfb_mvf_desc* d, int Cg, int R) {
  g.N = d->N; g.T = d->T; g.Cs = d->Cs; g.H = d->H; g.W = d->W;
  g.Cg = Cg; g.R = R; g.Hp = d->H + 2; g.Wp = d->W + 2;
  g.slot_x = round_up(g.Hp * g.Wp * Cg * 2, 128);
  g.slot_g = round_up(d->H * d->W * Cg * 2, 128);
  g.slot_dz = round_up(g.Hp * g.Wp * Cg * 4, 128);
}

// Channel-group / ring selection.  Wider groups give longer contiguous runs per pixel (Cg*2 bytes); more
// groups give more CTAs.  Prefer >= 2 CTAs per SM worth of work, rows of >= 32 B, and a ring that holds the
// whole clip (all loads in flight at once) when that still leaves room for 2 CTAs per SM.
bool choose_geo(const mvfb_mvf_desc* d, bool backward, Geo& out) {
  if (d->dtype != MVFB_BF16 || d->layout != MVFB_NHWC) return false;
  if (d->Cs % 8 != 0 || d->C % 8 != 0 || d->H + 2 > 256 || d->W + 2 > 256) return false;
  const int want = 2 * num_sms();
  int best = 0;
  const int cands[4] = {64, 32, 16, 8};
  for (int i = 0; i < 4; ++i) {
    const int Cg = cands[i];
    if (d->Cs % Cg) continue;
    if (Cg == 8 && best) break;                                // 16-byte rows only when nothing wider fits
    Geo g;
    fill_geo(g, d, Cg, d->T < 4 ? d->T : 4);
    const size_t need = backward ? bwdB_smem(g) : fwd_smem(g);
    if (need > (size_t)kSmemLimit) continue;
    if (!backward && need > 110 * 1024 && Cg > 16) continue;   // keep two forward CTAs per SM when possible
    best = Cg;
    if ((long long)d->N * (d->Cs / Cg) >= want || Cg == 8) break;
  }
  if (!best) return false;
  int R = d->T < 4 ? d->T : 4;
  if (d->T <= kMaxRing) {                                      // resident clip if it is cheap enough
    Geo g;
    fill_geo(g, d, best, d->T);
    const size_t need = backward ? bwdB_smem(g) : fwd_smem(g);
    if (need <= (size_t)(backward ? kSmemLimit : 110 * 1024)) R = d->T;
  }
  fill_geo(out, d, best, R);
  return true;
}

int make_tmap(CUtensorMap* tm, const void* base, long long pix_stride_elems, const Geo& g, bool padded) {
  const uint64_t dims[4] = {(uint64_t)g.Cs, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N * g.T};
  const uint64_t strides[3] = {(uint64_t)pix_stride_elems * 2, (uint64_t)g.W * pix_stride_elems * 2,
                               (uint64_t)g.H * g.W * pix_stride_elems * 2};
  const uint32_t box[4] = {(uint32_t)g.Cg, (uint32_t)(padded ? g.Wp : g.W), (uint32_t)(padded ? g.Hp : g.H), 1u};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, nullptr,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename K>
int set_smem(K kernel, size_t bytes) {
  MVFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return MVFB_OK;
}

}  // namespace

static bool choose_resident(const mvfb_mvf_desc* d, Geo& out, size_t& smem);

bool mvf_fast_supported(const mvfb_mvf_desc* d) {
  Geo g;
  size_t smem;
  return choose_resident(d, g, smem) || choose_geo(d, false, g);
}

bool mvf_fast_bwd_supported(const mvfb_mvf_desc* d) {
  Geo g;
  return choose_geo(d, true, g);
}

size_t mvf_fast_ws(const mvfb_mvf_desc* d) { return (size_t)d->N * d->Cs * 2 * sizeof(float) + 256; }

// resident-clip geometry: widest channel group whose whole (T, H+2, W+2) volume leaves room for two CTAs per SM
// and still yields >= 2 CTAs per SM of work (or the narrowest that fits)
static bool choose_resident(const mvfb_mvf_desc* d, Geo& out, size_t& smem) {
  if (d->dtype != MVFB_BF16 || d->layout != MVFB_NHWC) return false;
  if (d->T != 4 && d->T != 8 && d->T != 16) return false;
  if (d->Cs % 8 != 0 || d->C % 8 != 0 || d->H + 2 > 256 || d->W + 2 > 256) return false;
  const int want = 2 * num_sms();
  const int cands[4] = {64, 32, 16, 8};
  bool found = false;
  for (int i = 0; i < 4; ++i) {
    const int Cg = cands[i];
    if (d->Cs % Cg) continue;
    if (Cg == 8 && found) break;
    Geo g;
    fill_geo(g, d, Cg, d->T);
    const int G = Cg / 8;
    const size_t need = 256 + (size_t)d->T * g.slot_x + 2 * Cg * 4 + (size_t)kWarps * G * 16 * 4 + G * 16 * 4 + 3 * kThreads * 8;
    if (need > (size_t)kSmemLimit) continue;
    if (need > 112 * 1024 && Cg > 16) continue;
    out = g;
    smem = need;
    found = true;
    if ((long long)d->N * (d->Cs / Cg) >= want) break;
  }
  return found;
}

template <int T>
static int launch_resident(const mvfb_mvf_desc* d, const CUtensorMap& tmx, const FwdArgs& a, size_t smem,
                           cudaStream_t st) {
  static DevOnce once;
  int rc;
  if (once.pending()) {
    if ((rc = set_smem(mvf_fwd_resident_kernel<PASS_STATS, T>, kSmemLimit))) return rc;
    if ((rc = set_smem(mvf_fwd_resident_kernel<PASS_TRAIN, T>, kSmemLimit))) return rc;
    if ((rc = set_smem(mvf_fwd_resident_kernel<PASS_APPLY, T>, kSmemLimit))) return rc;
    once.done();
  }
  const dim3 grid(d->N * (d->Cs / a.g.Cg));
  if (d->use_hs && d->training) {
    mvf_fwd_resident_kernel<PASS_STATS, T><<<grid, kThreads, smem, st>>>(tmx, a);
    count_launch();
    MVFB_LAUNCH_CHECK();
    mvf_fwd_resident_kernel<PASS_TRAIN, T><<<grid, kThreads, smem, st>>>(tmx, a);
  } else {
    mvf_fwd_resident_kernel<PASS_APPLY, T><<<grid, kThreads, smem, st>>>(tmx, a);
  }
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int mvf_fast_fwd(const mvfb_mvf_desc* d, const void* x, void* y, long long y_stride, const float* wt,
                 const float* wh, const float* ww, const float* gamma, const float* beta, float* rm, float* rv,
                 float* save_mean, float* save_rstd, void* ws, cudaStream_t st) {
  Geo g;
  size_t smem_res = 0;
  const bool resident = choose_resident(d, g, smem_res);
  if (!resident && !choose_geo(d, false, g)) return MVFB_ERR_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(y) || y_stride % 8 != 0) return MVFB_ERR_UNSUPPORTED;
  CUtensorMap tmx;
  int rc = make_tmap(&tmx, x, d->C, g, true);
  if (rc) return rc;
  FwdArgs a;
  a.g = g;
  a.use_hs = d->use_hs; a.training = d->training; a.eps = d->eps; a.momentum = d->momentum;
  a.wt = wt; a.wh = d->mode != MVFB_MODE_T ? wh : nullptr; a.ww = d->mode == MVFB_MODE_THW ? ww : nullptr;
  a.gamma = gamma; a.beta = beta; a.running_mean = rm; a.running_var = rv;
  a.save_mean = save_mean; a.save_rstd = save_rstd;
  a.partials = (float*)ws;
  a.y = (__nv_bfloat16*)y; a.y_pix = y_stride;
  if (resident) {
    if (d->T == 4) return launch_resident<4>(d, tmx, a, smem_res, st);
    if (d->T == 8) return launch_resident<8>(d, tmx, a, smem_res, st);
    return launch_resident<16>(d, tmx, a, smem_res, st);
  }
  const size_t smem = fwd_smem(g);
  const dim3 grid(d->N * (d->Cs / g.Cg));
  if (d->use_hs && d->training) {
    static DevOnce once;
    if (once.pending()) {
      if ((rc = set_smem(mvf_fast_fwd_kernel<PASS_STATS>, kSmemLimit))) return rc;
      if ((rc = set_smem(mvf_fast_fwd_kernel<PASS_TRAIN>, kSmemLimit))) return rc;
      once.done();
    }
    mvf_fast_fwd_kernel<PASS_STATS><<<grid, kThreads, smem, st>>>(tmx, a);
    count_launch();
    MVFB_LAUNCH_CHECK();
    mvf_fast_fwd_kernel<PASS_TRAIN><<<grid, kThreads, smem, st>>>(tmx, a);
    count_launch();
    MVFB_LAUNCH_CHECK();
  } else {
    static DevOnce once;
    if (once.pending()) {
      if ((rc = set_smem(mvf_fast_fwd_kernel<PASS_APPLY>, kSmemLimit))) return rc;
      once.done();
    }
    mvf_fast_fwd_kernel<PASS_APPLY><<<grid, kThreads, smem, st>>>(tmx, a);
    count_launch();
    MVFB_LAUNCH_CHECK();
  }
  return MVFB_OK;
}

int mvf_fast_bwd(const mvfb_mvf_desc* d, const void* gp, long long g_stride, const void* x, void* dx,
                 long long dx_stride, const float* wt, const float* wh, const float* ww, const float* gamma,
                 const float* beta, const float* mean, const float* rstd, float* dwt, float* dwh, float* dww,
                 float* dgamma, float* dbeta, void* ws, cudaStream_t st) {
  Geo g;
  if (!choose_geo(d, true, g) || !aligned16(x) || !aligned16(gp) || (reinterpret_cast<uintptr_t>(dx) & 7) ||
      g_stride % 8 != 0 || dx_stride % 4 != 0)
    return MVFB_ERR_UNSUPPORTED;
  CUtensorMap tmx, tmg;
  int rc = make_tmap(&tmx, x, d->C, g, true);
  if (rc) return rc;
  if ((rc = make_tmap(&tmg, gp, g_stride, g, false))) return rc;
  const bool has_h = d->mode != MVFB_MODE_T, has_w = d->mode == MVFB_MODE_THW;
  BwdArgs a;
  a.g = g;
  a.use_hs = d->use_hs; a.training = d->training;
  a.share_h = has_h && wh == wt; a.share_w = has_w && ww == wt;
  a.wt = wt; a.wh = has_h ? wh : nullptr; a.ww = has_w ? ww : nullptr;
  a.gamma = gamma; a.beta = beta; a.mean = mean; a.rstd = rstd;
  a.partials = (float*)ws;
  a.dwt = dwt; a.dwh = (has_h && !a.share_h) ? dwh : nullptr; a.dww = (has_w && !a.share_w) ? dww : nullptr;
  a.dgamma = dgamma; a.dbeta = dbeta;
  a.dx = (__nv_bfloat16*)dx; a.dx_pix = dx_stride;
  static DevOnce once;
  if (once.pending()) {
    if ((rc = set_smem(mvf_fast_bwd_reduce_kernel, kSmemLimit))) return rc;
    if ((rc = set_smem(mvf_fast_bwd_dx_kernel, kSmemLimit))) return rc;
    once.done();
  }
  const dim3 grid(d->N * (d->Cs / g.Cg));
  if (d->use_hs) {
    mvf_fast_bwd_reduce_kernel<<<grid, kThreads, bwdA_smem(g), st>>>(tmx, tmg, a);
    count_launch();
    MVFB_LAUNCH_CHECK();
  } else {
    MVFB_CUDA(cudaMemsetAsync(dwt, 0, sizeof(float) * 3 * d->Cs, st));
    if (a.dwh) MVFB_CUDA(cudaMemsetAsync(a.dwh, 0, sizeof(float) * 3 * d->Cs, st));
    if (a.dww) MVFB_CUDA(cudaMemsetAsync(a.dww, 0, sizeof(float) * 3 * d->Cs, st));
  }
  mvf_fast_bwd_dx_kernel<<<grid, kThreads, bwdB_smem(g), st>>>(tmx, tmg, a);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

}  // namespace mvfb

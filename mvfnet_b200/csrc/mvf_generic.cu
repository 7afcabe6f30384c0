// MVF module, layout/dtype-generic kernels (any shape, NCHW or NHWC, fp32 or bf16 storage, fp32 math).
// This is the exact-layout drop-in path used for fp32 parity with the reference (MVF.py:104-137) and as
// the path for shapes the TMA kernels do not cover.  One CTA per (slab channel c, clip n); threads walk
// (t,h,w); per-channel reductions are CTA tree reductions + one fp64 atomic per CTA.
#include <cuda_bf16.h>

#include "common.cuh"
#include "mvf_internal.cuh"
#include "ptx.cuh"

namespace mvfb {

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) {
  return __ldg(p);
}
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) {
  *p = v;
}
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

struct Strides {
  long long f, c, h, w;
};

__host__ __device__ inline Strides make_strides(int layout, long long frame_stride_or_C, int Cch, int H, int W) {
  // NCHW: caller passes the frame stride (elements); NHWC: caller passes the per-pixel channel count.
  Strides s;
  if (layout == MVFB_NCHW) {
    s.w = 1; s.h = W; s.c = (long long)H * W; s.f = frame_stride_or_C;
  } else {
    s.c = 1; s.w = frame_stride_or_C; s.h = (long long)W * frame_stride_or_C; s.f = (long long)H * W * frame_stride_or_C;
  }
  (void)Cch;
  return s;
}

struct Taps {
  float t0, t1, t2, h0, h1, h2, w0, w1, w2;
};

// bf16 activations imply bf16 taps (see round_bf16 in ptx.cuh); fp32 activations use the taps as given.
template <typename T>
__device__ __forceinline__ Taps load_taps(const float* wt, const float* wh, const float* ww, int c) {
  Taps k;
  k.t0 = wt[c * 3 + 0]; k.t1 = wt[c * 3 + 1]; k.t2 = wt[c * 3 + 2];
  k.h0 = k.h1 = k.h2 = k.w0 = k.w1 = k.w2 = 0.f;
  if (wh) { k.h0 = wh[c * 3 + 0]; k.h1 = wh[c * 3 + 1]; k.h2 = wh[c * 3 + 2]; }
  if (ww) { k.w0 = ww[c * 3 + 0]; k.w1 = ww[c * 3 + 1]; k.w2 = ww[c * 3 + 2]; }
  if (sizeof(T) == 2) {
    k.t0 = round_bf16(k.t0); k.t1 = round_bf16(k.t1); k.t2 = round_bf16(k.t2);
    k.h0 = round_bf16(k.h0); k.h1 = round_bf16(k.h1); k.h2 = round_bf16(k.h2);
    k.w0 = round_bf16(k.w0); k.w1 = round_bf16(k.w1); k.w2 = round_bf16(k.w2);
  }
  return k;
}

// z at (t,h,w) of the (n,c) volume whose element (0,0,0) is at p; add order (t + h) + w as MVF.py:120.
template <typename T>
__device__ __forceinline__ float stencil_at(const T* p, const Strides& s, const Taps& k, int t, int h, int w, int Tn,
                                            int H, int W, bool has_h, bool has_w) {
  const T* q = p + t * s.f + h * s.h + w * s.w;
  float xc = ldf(q);
  float zt = k.t1 * xc;
  if (t > 0) zt = fmaf(k.t0, ldf(q - s.f), zt);
  if (t + 1 < Tn) zt = fmaf(k.t2, ldf(q + s.f), zt);
  float z = zt;
  if (has_h) {
    float zh = k.h1 * xc;
    if (h > 0) zh = fmaf(k.h0, ldf(q - s.h), zh);
    if (h + 1 < H) zh = fmaf(k.h2, ldf(q + s.h), zh);
    z += zh;
  }
  if (has_w) {
    float zw = k.w1 * xc;
    if (w > 0) zw = fmaf(k.w0, ldf(q - s.w), zw);
    if (w + 1 < W) zw = fmaf(k.w2, ldf(q + s.w), zw);
    z += zw;
  }
  return z;
}

template <int K>
__device__ __forceinline__ void block_reduce_atomic(float (&v)[K], double* dst, int dst_stride) {
  __shared__ float red[K][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    float a = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[i][warp] = a;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < K; ++i) {
      float a = lane < nw ? red[i][lane] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) atomicAdd(dst + (size_t)i * dst_stride, (double)a);
    }
  }
}

__device__ __forceinline__ float hswish(float u) { return u * __saturatef(fmaf(u, 1.f / 6.f, 0.5f)); }
__device__ __forceinline__ float hswish_grad(float u) {
  float s = __saturatef(fmaf(u, 1.f / 6.f, 0.5f));
  return s + ((u > -3.f && u < 3.f) ? u * (1.f / 6.f) : 0.f);
}

struct GenParams {
  int N, T, Cs, H, W;
  int has_h, has_w, use_hs, training;
  float eps, momentum;
  Strides sx, sy;           // x addressing; y / g / dx addressing
  const float *wt, *wh, *ww, *gamma, *beta;
  float *running_mean, *running_var, *save_mean, *save_rstd;
  double* sums;             // [2][Cs] forward; backward: [11][Cs]
};

// ---------------------------------------------------------------- forward
template <typename T>
__global__ void __launch_bounds__(256) gen_fwd_stats(const T* __restrict__ x, GenParams p) {
  const int c = blockIdx.x, n = blockIdx.y;
  const Taps k = load_taps<T>(p.wt, p.wh, p.ww, c);
  const T* base = x + (long long)n * p.T * p.sx.f + c * p.sx.c;
  const int vol = p.T * p.H * p.W;
  float acc[2] = {0.f, 0.f};
  for (int i = threadIdx.x; i < vol; i += blockDim.x) {
    int w = i % p.W, h = (i / p.W) % p.H, t = i / (p.W * p.H);
    float z = stencil_at(base, p.sx, k, t, h, w, p.T, p.H, p.W, p.has_h, p.has_w);
    acc[0] += z;
    acc[1] = fmaf(z, z, acc[1]);
  }
  block_reduce_atomic<2>(acc, p.sums + c, p.Cs);
}

// scale/shift of channel c for the apply passes; also emits saved / running statistics once.
__device__ __forceinline__ void bn_coeffs(const GenParams& p, int c, bool writer, float& mean, float& rstd) {
  if (p.training) {
    double m = (double)p.N * p.T * p.H * p.W;
    double mu = p.sums[c] / m;
    double var = p.sums[p.Cs + c] / m - mu * mu;
    if (var < 0) var = 0;
    mean = (float)mu;
    rstd = (float)(1.0 / sqrt(var + (double)p.eps));
    if (writer) {
      if (p.save_mean) p.save_mean[c] = mean;
      if (p.save_rstd) p.save_rstd[c] = rstd;
      if (p.running_mean) {
        double unb = m > 1 ? var * m / (m - 1) : var;
        p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * mean;
        p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)unb;
      }
    }
  } else {
    mean = p.running_mean[c];
    rstd = 1.f / sqrtf(p.running_var[c] + p.eps);
    if (writer) {
      if (p.save_mean) p.save_mean[c] = mean;
      if (p.save_rstd) p.save_rstd[c] = rstd;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gen_fwd_apply(const T* __restrict__ x, T* __restrict__ y, GenParams p) {
  const int c = blockIdx.x, n = blockIdx.y;
  const Taps k = load_taps<T>(p.wt, p.wh, p.ww, c);
  float scale = 1.f, shift = 0.f;
  if (p.use_hs) {
    float mean, rstd;
    bn_coeffs(p, c, n == 0 && threadIdx.x == 0, mean, rstd);
    scale = p.gamma[c] * rstd;
    shift = p.beta[c] - mean * scale;
  }
  const T* base = x + (long long)n * p.T * p.sx.f + c * p.sx.c;
  T* out = y + (long long)n * p.T * p.sy.f + c * p.sy.c;
  const int vol = p.T * p.H * p.W;
  for (int i = threadIdx.x; i < vol; i += blockDim.x) {
    int w = i % p.W, h = (i / p.W) % p.H, t = i / (p.W * p.H);
    float z = stencil_at(base, p.sx, k, t, h, w, p.T, p.H, p.W, p.has_h, p.has_w);
    float v = p.use_hs ? hswish(fmaf(z, scale, shift)) : z;
    stf(out + t * p.sy.f + h * p.sy.h + w * p.sy.w, v);
  }
}

// ---------------------------------------------------------------- backward
struct GenBwdParams {
  GenParams f;              // shapes, taps, x strides (sx), g/dx strides (sy)
  const float *mean, *rstd; // per-channel statistics used in the forward
  float* dz;                // fp32 scratch (N, Cs, T, H, W)
  double* sums;             // [0]=sum du, [1]=sum du*zhat, [2..10]=tap sums (t0,t1,t2,h0,h1,h2,w0,w1,w2)
  Strides sdx;
};

template <typename T>
__global__ void __launch_bounds__(256) gen_bwd_reduce(const T* __restrict__ g, const T* __restrict__ x, GenBwdParams q) {
  const GenParams& p = q.f;
  const int c = blockIdx.x, n = blockIdx.y;
  const Taps k = load_taps<T>(p.wt, p.wh, p.ww, c);
  const float mean = q.mean[c], rstd = q.rstd[c], gam = p.gamma[c], bet = p.beta[c];
  const T* base = x + (long long)n * p.T * p.sx.f + c * p.sx.c;
  const T* gb = g + (long long)n * p.T * p.sy.f + c * p.sy.c;
  const int vol = p.T * p.H * p.W;
  float acc[2] = {0.f, 0.f};
  for (int i = threadIdx.x; i < vol; i += blockDim.x) {
    int w = i % p.W, h = (i / p.W) % p.H, t = i / (p.W * p.H);
    float z = stencil_at(base, p.sx, k, t, h, w, p.T, p.H, p.W, p.has_h, p.has_w);
    float zhat = (z - mean) * rstd;
    float u = fmaf(zhat, gam, bet);
    float du = ldf(gb + t * p.sy.f + h * p.sy.h + w * p.sy.w) * hswish_grad(u);
    acc[0] += du;
    acc[1] = fmaf(du, zhat, acc[1]);
  }
  block_reduce_atomic<2>(acc, q.sums + c, p.Cs);
}

template <typename T>
__global__ void __launch_bounds__(256) gen_bwd_dz(const T* __restrict__ g, const T* __restrict__ x, GenBwdParams q) {
  const GenParams& p = q.f;
  const int c = blockIdx.x, n = blockIdx.y;
  const int vol = p.T * p.H * p.W;
  const T* gb = g + (long long)n * p.T * p.sy.f + c * p.sy.c;
  float* dz = q.dz + ((long long)n * p.Cs + c) * vol;
  if (!p.use_hs) {
    for (int i = threadIdx.x; i < vol; i += blockDim.x) {
      int w = i % p.W, h = (i / p.W) % p.H, t = i / (p.W * p.H);
      dz[i] = ldf(gb + t * p.sy.f + h * p.sy.h + w * p.sy.w);
    }
    return;
  }
  const Taps k = load_taps<T>(p.wt, p.wh, p.ww, c);
  const float mean = q.mean[c], rstd = q.rstd[c], gam = p.gamma[c], bet = p.beta[c];
  const double m = (double)p.N * vol;
  const float mdb = p.training ? (float)(q.sums[c] / m) : 0.f;
  const float mdg = p.training ? (float)(q.sums[p.Cs + c] / m) : 0.f;
  const float gr = gam * rstd;
  const T* base = x + (long long)n * p.T * p.sx.f + c * p.sx.c;
  for (int i = threadIdx.x; i < vol; i += blockDim.x) {
    int w = i % p.W, h = (i / p.W) % p.H, t = i / (p.W * p.H);
    float z = stencil_at(base, p.sx, k, t, h, w, p.T, p.H, p.W, p.has_h, p.has_w);
    float zhat = (z - mean) * rstd;
    float u = fmaf(zhat, gam, bet);
    float du = ldf(gb + t * p.sy.f + h * p.sy.h + w * p.sy.w) * hswish_grad(u);
    dz[i] = gr * (du - mdb - zhat * mdg);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gen_bwd_dx(const T* __restrict__ x, T* __restrict__ dx, GenBwdParams q) {
  const GenParams& p = q.f;
  const int c = blockIdx.x, n = blockIdx.y;
  const Taps k = load_taps<T>(p.wt, p.wh, p.ww, c);
  const int vol = p.T * p.H * p.W, HW = p.H * p.W;
  const float* dz = q.dz + ((long long)n * p.Cs + c) * vol;
  const T* base = x + (long long)n * p.T * p.sx.f + c * p.sx.c;
  T* out = dx + (long long)n * p.T * q.sdx.f + c * q.sdx.c;
  float acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.f;
  for (int i = threadIdx.x; i < vol; i += blockDim.x) {
    int w = i % p.W, h = (i / p.W) % p.H, t = i / HW;
    const float d = dz[i];
    const T* xq = base + t * p.sx.f + h * p.sx.h + w * p.sx.w;
    const float xc = ldf(xq);
    // transposed stencil: dx[p] = sum_k w[k] * dz[p - (k-1)]
    float v = k.t1 * d;
    if (t + 1 < p.T) v = fmaf(k.t0, dz[i + HW], v);
    if (t > 0) v = fmaf(k.t2, dz[i - HW], v);
    acc[1] = fmaf(d, xc, acc[1]);
    if (t > 0) acc[0] = fmaf(d, ldf(xq - p.sx.f), acc[0]);
    if (t + 1 < p.T) acc[2] = fmaf(d, ldf(xq + p.sx.f), acc[2]);
    if (p.has_h) {
      v = fmaf(k.h1, d, v);
      if (h + 1 < p.H) v = fmaf(k.h0, dz[i + p.W], v);
      if (h > 0) v = fmaf(k.h2, dz[i - p.W], v);
      acc[4] = fmaf(d, xc, acc[4]);
      if (h > 0) acc[3] = fmaf(d, ldf(xq - p.sx.h), acc[3]);
      if (h + 1 < p.H) acc[5] = fmaf(d, ldf(xq + p.sx.h), acc[5]);
    }
    if (p.has_w) {
      v = fmaf(k.w1, d, v);
      if (w + 1 < p.W) v = fmaf(k.w0, dz[i + 1], v);
      if (w > 0) v = fmaf(k.w2, dz[i - 1], v);
      acc[7] = fmaf(d, xc, acc[7]);
      if (w > 0) acc[6] = fmaf(d, ldf(xq - p.sx.w), acc[6]);
      if (w + 1 < p.W) acc[8] = fmaf(d, ldf(xq + p.sx.w), acc[8]);
    }
    stf(out + t * q.sdx.f + h * q.sdx.h + w * q.sdx.w, v);
  }
  block_reduce_atomic<9>(acc, q.sums + 2 * p.Cs + c, p.Cs);
}

// sums (fp64) -> fp32 parameter gradients.  share: wh/ww alias wt -> their tap sums fold into dwt.
__global__ void mvf_bwd_finalize(const double* sums, int Cs, int h_shares, int w_shares, int has_h, int has_w,
                                 int use_hs, float* dwt, float* dwh, float* dww, float* dgamma, float* dbeta) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cs) return;
  if (use_hs) {
    if (dbeta) dbeta[c] = (float)sums[c];
    if (dgamma) dgamma[c] = (float)sums[Cs + c];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double t = sums[(2 + k) * Cs + c], h = sums[(5 + k) * Cs + c], w = sums[(8 + k) * Cs + c];
    if (has_h && h_shares) t += h;
    if (has_w && w_shares) t += w;
    dwt[c * 3 + k] = (float)t;
    if (has_h && !h_shares && dwh) dwh[c * 3 + k] = (float)h;
    if (has_w && !w_shares && dww) dww[c * 3 + k] = (float)w;
  }
}

// ---------------------------------------------------------------- host launchers
static GenParams make_params(const mvfb_mvf_desc* d, long long y_stride, const float* wt, const float* wh,
                             const float* ww, const float* gamma, const float* beta, float* rm, float* rv,
                             float* save_mean, float* save_rstd, double* sums) {
  GenParams p;
  p.N = d->N; p.T = d->T; p.Cs = d->Cs; p.H = d->H; p.W = d->W;
  p.has_h = d->mode != MVFB_MODE_T;
  p.has_w = d->mode == MVFB_MODE_THW;
  p.use_hs = d->use_hs; p.training = d->training; p.eps = d->eps; p.momentum = d->momentum;
  p.sx = make_strides(d->layout, d->layout == MVFB_NCHW ? (long long)d->C * d->H * d->W : d->C, d->C, d->H, d->W);
  p.sy = make_strides(d->layout, y_stride, d->Cs, d->H, d->W);
  p.wt = wt; p.wh = p.has_h ? wh : nullptr; p.ww = p.has_w ? ww : nullptr;
  p.gamma = gamma; p.beta = beta; p.running_mean = rm; p.running_var = rv;
  p.save_mean = save_mean; p.save_rstd = save_rstd; p.sums = sums;
  return p;
}

template <typename T>
static int gen_fwd_t(const mvfb_mvf_desc* d, const void* x, void* y, GenParams p, cudaStream_t st) {
  dim3 grid(d->Cs, d->N);
  if (d->use_hs && d->training) {
    MVFB_CUDA(cudaMemsetAsync(p.sums, 0, sizeof(double) * 2 * d->Cs, st));
    gen_fwd_stats<T><<<grid, 256, 0, st>>>((const T*)x, p);
    count_launch();
    MVFB_LAUNCH_CHECK();
  }
  gen_fwd_apply<T><<<grid, 256, 0, st>>>((const T*)x, (T*)y, p);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int mvf_generic_fwd(const mvfb_mvf_desc* d, const void* x, void* y, long long y_stride, const float* wt,
                    const float* wh, const float* ww, const float* gamma, const float* beta, float* rm, float* rv,
                    float* save_mean, float* save_rstd, void* ws, cudaStream_t st) {
  GenParams p = make_params(d, y_stride, wt, wh, ww, gamma, beta, rm, rv, save_mean, save_rstd, (double*)ws);
  return d->dtype == MVFB_F32 ? gen_fwd_t<float>(d, x, y, p, st) : gen_fwd_t<__nv_bfloat16>(d, x, y, p, st);
}

template <typename T>
static int gen_bwd_t(const mvfb_mvf_desc* d, const void* g, const void* x, void* dx, GenBwdParams q,
                     cudaStream_t st) {
  dim3 grid(d->Cs, d->N);
  MVFB_CUDA(cudaMemsetAsync(q.sums, 0, sizeof(double) * 11 * d->Cs, st));
  if (d->use_hs) {
    gen_bwd_reduce<T><<<grid, 256, 0, st>>>((const T*)g, (const T*)x, q);
    count_launch();
    MVFB_LAUNCH_CHECK();
  }
  gen_bwd_dz<T><<<grid, 256, 0, st>>>((const T*)g, (const T*)x, q);
  count_launch();
  MVFB_LAUNCH_CHECK();
  gen_bwd_dx<T><<<grid, 256, 0, st>>>((const T*)x, (T*)dx, q);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

// header of 2*roundup(Cs,64) floats is reserved by mvf_bwd (api.cu) for eval-mode (mean, rstd)
size_t mvf_generic_bwd_ws(const mvfb_mvf_desc* d) {
  size_t e = (size_t)d->N * d->T * d->Cs * d->H * d->W;
  return 2 * sizeof(float) * (((size_t)d->Cs + 63) / 64 * 64) + 16 * 8 * (size_t)d->Cs + e * sizeof(float) + 256;
}

__global__ void eval_stats_kernel(const float* rm, const float* rv, float eps, int Cs, float* mean, float* rstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < Cs) {
    mean[c] = rm[c];
    rstd[c] = 1.f / sqrtf(rv[c] + eps);
  }
}

int mvf_eval_stats(const float* rm, const float* rv, float eps, int Cs, float* mean, float* rstd, cudaStream_t st) {
  eval_stats_kernel<<<ceil_div(Cs, 128), 128, 0, st>>>(rm, rv, eps, Cs, mean, rstd);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int mvf_generic_bwd(const mvfb_mvf_desc* d, const void* g, long long g_stride, const void* x, void* dx,
                    long long dx_stride, const float* wt, const float* wh, const float* ww, const float* gamma,
                    const float* beta, const float* mean, const float* rstd, float* dwt, float* dwh, float* dww,
                    float* dgamma, float* dbeta, void* ws, cudaStream_t st) {
  GenBwdParams q;
  q.sums = (double*)ws;
  q.dz = (float*)((char*)ws + 16 * 8 * (size_t)d->Cs);
  q.f = make_params(d, g_stride, wt, wh, ww, gamma, beta, nullptr, nullptr, nullptr, nullptr, q.sums);
  q.mean = mean; q.rstd = rstd;
  q.sdx = make_strides(d->layout, dx_stride, d->Cs, d->H, d->W);
  int rc = d->dtype == MVFB_F32 ? gen_bwd_t<float>(d, g, x, dx, q, st) : gen_bwd_t<__nv_bfloat16>(d, g, x, dx, q, st);
  if (rc) return rc;
  const bool has_h = d->mode != MVFB_MODE_T, has_w = d->mode == MVFB_MODE_THW;
  mvf_bwd_finalize<<<ceil_div(d->Cs, 128), 128, 0, st>>>(q.sums, d->Cs, has_h && wh == wt, has_w && ww == wt, has_h,
                                                          has_w, d->use_hs, dwt, dwh, dww, dgamma, dbeta);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

}  // namespace mvfb

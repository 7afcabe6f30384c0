// The ends of the training step around the bottleneck stack: all HBM-bound streaming kernels.
//
//   preprocess_u8      Normalize + FormatShape + ToTensor of the data pipeline (datasets/pipelines/augmentations.py:343-396,
//                      formating.py:134-185) on the GPU: uint8 HWC frames -> normalised bf16 NHWC frames, so the host ships
//                      1 byte per sample instead of the reference's float32 (B, T, 3, H, W) wire format.
//   head_pool_fwd/bwd  TSNClsHead: spatial average pool + dropout (heads/tsn_clshead.py:71-92), forward and backward.
//   head_ce_fwd/bwd    ... new_fc bias + SimpleConsensus mean over the T segments (tsn_clshead.py:93-98,
//                      segmental_consensuses/simple_consensus.py:41-61) + cross-entropy (heads/base.py:40-45); the 2048 -> 400
//                      projection itself runs on conv1x1_gemm / conv1x1_wgrad with the class axis padded to 448.
//   flat_sqnorm        sum of squares of the flat gradient buffer (the clip_grad norm, core/dist_utils.py:65-66)
//   sgd_nesterov_step  DistOptimizerHook.after_train_iter's tail in ONE pass over the flat buffers: / world (dist_utils.py:32),
//                      clip to max_norm (optimizer_config grad_clip, r50_dense.py:154), weight decay, momentum, Nesterov
//                      update (torch.optim.SGD as configured at r50_dense.py:152-153), and the bf16 copy of the updated
//                      parameters that the next forward's GEMMs read.
#include <cuda_bf16.h>

#include "common.cuh"
#include "mvf_internal.cuh"
#include "ptx.cuh"

namespace mvfb {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// all threads of the CTA get the result; `red` holds >= 32 floats
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = warp_sum(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
  if (threadIdx.x < 32) t = warp_max(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

// ------------------------------------------------------------------------------------------ input pre-processing
// 4 pixels (12 bytes in, 24 bytes out) per thread; channel of byte j is j % 3 whatever the pixel.
__global__ void __launch_bounds__(kThreads)
preprocess_u8_kernel(const uint32_t* __restrict__ x, uint2* __restrict__ y, long long groups, float m0, float m1,
                     float m2, float s0, float s1, float s2, int swap_rb) {
  const long long stride = (long long)gridDim.x * kThreads;
  // mean / inverse std per position of the 12-byte group, in OUTPUT channel order
  const float mean[3] = {m0, m1, m2}, inv[3] = {s0, s1, s2};
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < groups; i += stride) {
    const uint32_t a = __ldg(x + 3 * i), b = __ldg(x + 3 * i + 1), c = __ldg(x + 3 * i + 2);
    float v[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = (float)((a >> (8 * j)) & 0xffu);
      v[4 + j] = (float)((b >> (8 * j)) & 0xffu);
      v[8 + j] = (float)((c >> (8 * j)) & 0xffu);
    }
    float o[12];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const int src = swap_rb ? 2 - ch : ch;                       // BGR -> RGB: output channel ch reads input 2 - ch
        o[3 * p + ch] = (v[3 * p + src] - mean[ch]) * inv[ch];
      }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      uint2 w;
      w.x = pack_bf16(o[4 * q], o[4 * q + 1]);
      w.y = pack_bf16(o[4 * q + 2], o[4 * q + 3]);
      y[3 * i + q] = w;
    }
  }
}

// ------------------------------------------------------------------------------------------ head: pool + dropout
// counter-based keep decision of element (row f, channel c): identical in forward and backward, no mask is stored
__device__ __forceinline__ bool keep_elem(unsigned long long seed, unsigned long long idx, float p) {
  unsigned long long z = idx + seed * 0x9E3779B97F4A7C15ull + 0x9E3779B97F4A7C15ull;   // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.0f / 16777216.0f);         // 24 random bits -> [0, 1)
  return u >= p;
}

// x: (F, HW, C) bf16; feat: (F, C) bf16 = dropout(mean over HW).  One thread per (f, 8-channel vector).
__global__ void __launch_bounds__(kThreads)
head_pool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ feat, long long F, int HW, int C,
                     float p, unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev) seed ^= *seed_dev * 0xD6E8FEB86659FD93ull;          // per-replay part of the seed (CUDA-graph steps)
  const int vecs = C >> 3;
  const long long total = F * vecs;
  const float inv_hw = 1.f / (float)HW, keep_scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const long long f = i / vecs;
    const int vec = (int)(i - f * vecs);
    const __nv_bfloat16* px = x + (f * HW) * (long long)C + vec * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int q = 0; q < HW; ++q) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(px + (long long)q * C));
      acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
      acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // the pooled value is a bf16 tensor in the reference's autocast graph: round before the dropout scale
      const float m = __bfloat162float(__float2bfloat16_rn(acc[j] * inv_hw));
      o[j] = (p > 0.f) ? (keep_elem(seed, (unsigned long long)(f * C + vec * 8 + j), p) ? m * keep_scale : 0.f) : m;
    }
    uint4 w;
    w.x = pack_bf16(o[0], o[1]); w.y = pack_bf16(o[2], o[3]); w.z = pack_bf16(o[4], o[5]); w.w = pack_bf16(o[6], o[7]);
    *reinterpret_cast<uint4*>(feat + f * C + vec * 8) = w;
  }
}

// dfeat: (F, C) bf16 -> dx: (F, HW, C) bf16 = dfeat * keep / HW broadcast over the HW pixels
__global__ void __launch_bounds__(kThreads)
head_pool_bwd_kernel(const __nv_bfloat16* __restrict__ dfeat, __nv_bfloat16* __restrict__ dx, long long F, int HW, int C,
                     float p, unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev) seed ^= *seed_dev * 0xD6E8FEB86659FD93ull;
  const int vecs = C >> 3;
  const long long total = F * vecs;
  const float scale = (p > 0.f ? 1.f / (1.f - p) : 1.f) / (float)HW;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const long long f = i / vecs;
    const int vec = (int)(i - f * vecs);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(dfeat + f * C + vec * 8));
    float g[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y), bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool keep = p > 0.f ? keep_elem(seed, (unsigned long long)(f * C + vec * 8 + j), p) : true;
      g[j] = keep ? g[j] * scale : 0.f;
    }
    uint4 w;
    w.x = pack_bf16(g[0], g[1]); w.y = pack_bf16(g[2], g[3]); w.z = pack_bf16(g[4], g[5]); w.w = pack_bf16(g[6], g[7]);
    __nv_bfloat16* pd = dx + (f * HW) * (long long)C + vec * 8;
    for (int q = 0; q < HW; ++q) *reinterpret_cast<uint4*>(pd + (long long)q * C) = w;
  }
}

// ------------------------------------------------------------------------------------------ head: consensus + CE
// One CTA per clip.  s[c] = bias[c] + mean_t logits[b*T + t, c];  loss += (logsumexp(s) - s[label]) / B;
// ds[b, c] = (softmax(s)[c] - [c == label]) / B  (fp32, consumed by head_ce_bwd_kernel);  dbias[c] += ds[b, c].
__global__ void __launch_bounds__(kThreads)
head_ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, long long ldl, const float* __restrict__ bias,
                   const long long* __restrict__ labels, int B, int T, int NC, float* __restrict__ score,
                   float* __restrict__ ds, float* __restrict__ dbias, float* __restrict__ loss) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float inv_t = 1.f / (float)T, inv_b = 1.f / (float)B;
  const long long label = labels[b];
  // classes handled by this thread: c = tid, tid + 256 (NC <= 512)
  float s[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int c = threadIdx.x + k * kThreads;
    if (c < NC) {
      float a = 0.f;
      for (int t = 0; t < T; ++t) a += __bfloat162float(logits[((long long)b * T + t) * ldl + c]);
      s[k] = a * inv_t + (bias ? bias[c] : 0.f);
      if (score) score[(long long)b * NC + c] = s[k];
    }
  }
  const float mx = block_max(fmaxf(s[0], s[1]), red);
  float e[2] = {0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (threadIdx.x + k * kThreads < NC) e[k] = __expf(s[k] - mx);
  const float sum = block_sum(e[0] + e[1], red);
  const float lse = mx + __logf(sum);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int c = threadIdx.x + k * kThreads;
    if (c < NC) {
      const float g = (e[k] / sum - (c == label ? 1.f : 0.f)) * inv_b;
      ds[(long long)b * NC + c] = g;
      if (dbias) atomicAdd(&dbias[c], g);
      if (c == label) atomicAdd(loss, (lse - s[k]) * inv_b);
    }
  }
}

// dlogits[b*T + t, c] = gout * ds[b, c] / T for c < NC, 0 for the padding columns NC <= c < ldl
__global__ void __launch_bounds__(kThreads)
head_ce_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ gout, __nv_bfloat16* __restrict__ dlogits,
                   long long ldl, int B, int T, int NC) {
  const float g = (gout ? *gout : 1.f) / (float)T;
  const long long total = (long long)B * T * ldl;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const long long row = i / ldl;
    const int c = (int)(i - row * ldl);
    const long long b = row / T;
    dlogits[i] = __float2bfloat16_rn(c < NC ? ds[b * NC + c] * g : 0.f);
  }
}

// ------------------------------------------------------------------------------------------ optimizer tail
__global__ void __launch_bounds__(kThreads)
flat_sqnorm_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  __shared__ float red[32];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(g4 + i);
    a0 = fmaf(v.x, v.x, a0); a1 = fmaf(v.y, v.y, a1); a2 = fmaf(v.z, v.z, a2); a3 = fmaf(v.w, v.w, a3);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    a0 = fmaf(v, v, a0);
  }
  const float s = block_sum((a0 + a1) + (a2 + a3), red);
  if (threadIdx.x == 0) atomicAdd(out, (double)s);
}

struct SgdArgs {
  float* p;                 // flat fp32 parameters (updated in place)
  float* mom;               // flat momentum buffer (updated in place)
  const float* g;           // flat gradients (sum over ranks when world > 1)
  __nv_bfloat16* p16;       // optional bf16 copy of the updated parameters
  long long n;
  const double* sqnorm;     // device: sum of squares of g (before grad_scale)
  float* norm_out;          // optional device scalar: total gradient norm after grad_scale (what clip_grad_norm_ returns)
  float grad_scale;         // 1 / world
  float max_norm;           // <= 0: no clipping
  float lr, momentum, weight_decay;
  int nesterov;
};

__global__ void __launch_bounds__(kThreads)
sgd_nesterov_kernel(const SgdArgs a) {
  float coef = a.grad_scale;
  if (a.max_norm > 0.f) {
    const float total = sqrtf((float)*a.sqnorm) * a.grad_scale;
    const float c = a.max_norm / (total + 1e-6f);                    // torch.nn.utils.clip_grad_norm_
    coef *= fminf(c, 1.f);
    if (a.norm_out && blockIdx.x == 0 && threadIdx.x == 0) *a.norm_out = total;
  }
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * kThreads;
  auto upd = [&](float& p, float& m, float g) {
    g = fmaf(a.weight_decay, p, g * coef);
    m = fmaf(a.momentum, m, g);
    g = a.nesterov ? fmaf(a.momentum, m, g) : m;
    p = fmaf(-a.lr, g, p);
  };
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(a.p)[i], m = reinterpret_cast<float4*>(a.mom)[i];
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.g) + i);
    upd(p.x, m.x, g.x); upd(p.y, m.y, g.y); upd(p.z, m.z, g.z); upd(p.w, m.w, g.w);
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.mom)[i] = m;
    if (a.p16) {
      uint2 w;
      w.x = pack_bf16(p.x, p.y); w.y = pack_bf16(p.z, p.w);
      reinterpret_cast<uint2*>(a.p16)[i] = w;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float p = a.p[i], m = a.mom[i];
    upd(p, m, a.g[i]);
    a.p[i] = p; a.mom[i] = m;
    if (a.p16) a.p16[i] = __float2bfloat16_rn(p);
  }
}

int stream_grid(long long work_items) {
  long long blocks = (work_items + kThreads - 1) / kThreads;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace mvfb

using namespace mvfb;

extern "C" {

int preprocess_u8(const void* x, void* y, long long pixels, const float* mean, const float* std_, int to_rgb,
                  mvfb_stream_t stream) {
  MVFB_CHECK(x && y && mean && std_ && pixels > 0, MVFB_ERR_ARG, "null argument / no pixels");
  MVFB_CHECK(pixels % 4 == 0, MVFB_ERR_UNSUPPORTED, "pixel count %lld must be a multiple of 4", pixels);
  MVFB_CHECK(!((uintptr_t)x & 3) && !((uintptr_t)y & 7), MVFB_ERR_UNSUPPORTED, "x must be 4-byte, y 8-byte aligned");
  MVFB_CHECK(std_[0] != 0.f && std_[1] != 0.f && std_[2] != 0.f, MVFB_ERR_ARG, "zero std");
  const long long groups = pixels / 4;
  // (mean, 1/std) are indexed by the OUTPUT channel (after the optional BGR -> RGB swap), as mmcv.imnormalize applies them
  preprocess_u8_kernel<<<stream_grid(groups), kThreads, 0, (cudaStream_t)stream>>>(
      (const uint32_t*)x, (uint2*)y, groups, mean[0], mean[1], mean[2], (float)(1.0 / (double)std_[0]),
      (float)(1.0 / (double)std_[1]), (float)(1.0 / (double)std_[2]), to_rgb);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int head_pool_fwd(const void* x, void* feat, long long F, int HW, int C, float p, unsigned long long seed,
                  const unsigned long long* seed_dev, mvfb_stream_t stream) {
  MVFB_CHECK(x && feat && F > 0 && HW > 0 && C > 0, MVFB_ERR_ARG, "null argument / bad shape");
  MVFB_CHECK(C % 8 == 0 && !((uintptr_t)x & 15) && !((uintptr_t)feat & 15), MVFB_ERR_UNSUPPORTED,
             "C must be a multiple of 8, tensors 16-byte aligned");
  MVFB_CHECK(p >= 0.f && p < 1.f, MVFB_ERR_ARG, "dropout ratio %f outside [0, 1)", p);
  head_pool_fwd_kernel<<<stream_grid(F * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (__nv_bfloat16*)feat, F, HW, C, p, seed, seed_dev);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int head_pool_bwd(const void* dfeat, void* dx, long long F, int HW, int C, float p, unsigned long long seed,
                  const unsigned long long* seed_dev, mvfb_stream_t stream) {
  MVFB_CHECK(dfeat && dx && F > 0 && HW > 0 && C > 0, MVFB_ERR_ARG, "null argument / bad shape");
  MVFB_CHECK(C % 8 == 0 && !((uintptr_t)dfeat & 15) && !((uintptr_t)dx & 15), MVFB_ERR_UNSUPPORTED,
             "C must be a multiple of 8, tensors 16-byte aligned");
  MVFB_CHECK(p >= 0.f && p < 1.f, MVFB_ERR_ARG, "dropout ratio %f outside [0, 1)", p);
  head_pool_bwd_kernel<<<stream_grid(F * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dfeat, (__nv_bfloat16*)dx, F, HW, C, p, seed, seed_dev);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int head_ce_fwd(const void* logits, long long ldl, const float* bias, const long long* labels, int B, int T, int NC,
                float* score, float* ds, float* dbias, float* loss, mvfb_stream_t stream) {
  MVFB_CHECK(logits && labels && ds && loss && B > 0 && T > 0, MVFB_ERR_ARG, "null argument / bad shape");
  MVFB_CHECK(NC > 0 && NC <= 2 * kThreads && ldl >= NC, MVFB_ERR_UNSUPPORTED, "classes %d must be in (0, %d] and <= ldl", NC,
             2 * kThreads);
  cudaStream_t st = (cudaStream_t)stream;
  MVFB_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  if (dbias) MVFB_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * NC, st));
  head_ce_fwd_kernel<<<B, kThreads, 0, st>>>((const __nv_bfloat16*)logits, ldl, bias, labels, B, T, NC, score, ds, dbias,
                                             loss);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int head_ce_bwd(const float* ds, const float* gout, void* dlogits, long long ldl, int B, int T, int NC,
                mvfb_stream_t stream) {
  MVFB_CHECK(ds && dlogits && B > 0 && T > 0 && NC > 0 && ldl >= NC, MVFB_ERR_ARG, "null argument / bad shape");
  head_ce_bwd_kernel<<<stream_grid((long long)B * T * ldl), kThreads, 0, (cudaStream_t)stream>>>(
      ds, gout, (__nv_bfloat16*)dlogits, ldl, B, T, NC);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int flat_sqnorm(const float* g, long long n, double* out, mvfb_stream_t stream) {
  MVFB_CHECK(g && out && n > 0, MVFB_ERR_ARG, "null argument / empty buffer");
  MVFB_CHECK(!((uintptr_t)g & 15), MVFB_ERR_UNSUPPORTED, "the flat buffer must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  MVFB_CUDA(cudaMemsetAsync(out, 0, sizeof(double), st));
  flat_sqnorm_kernel<<<stream_grid(n / 4 + 1), kThreads, 0, st>>>(g, n, out);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int sgd_nesterov_step(float* p, float* mom, const float* g, void* p_bf16, long long n, const double* sqnorm,
                      float* norm_out, float grad_scale, float max_norm, float lr, float momentum, float weight_decay,
                      int nesterov, mvfb_stream_t stream) {
  MVFB_CHECK(p && mom && g && n > 0, MVFB_ERR_ARG, "null argument / empty buffer");
  MVFB_CHECK(max_norm <= 0.f || sqnorm, MVFB_ERR_ARG, "clipping needs the squared norm (flat_sqnorm)");
  MVFB_CHECK(!((uintptr_t)p & 15) && !((uintptr_t)mom & 15) && !((uintptr_t)g & 15) && !((uintptr_t)p_bf16 & 7),
             MVFB_ERR_UNSUPPORTED, "flat buffers must be 16-byte aligned");
  SgdArgs a;
  a.p = p; a.mom = mom; a.g = g; a.p16 = (__nv_bfloat16*)p_bf16; a.n = n; a.sqnorm = sqnorm; a.norm_out = norm_out;
  a.grad_scale = grad_scale; a.max_norm = max_norm; a.lr = lr; a.momentum = momentum; a.weight_decay = weight_decay;
  a.nesterov = nesterov;
  sgd_nesterov_kernel<<<stream_grid(n / 4 + 1), kThreads, 0, (cudaStream_t)stream>>>(a);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ weight operand forms
// The input-gradient GEMMs read W^T (1x1: (Cin, Cout); 3x3: (Cin, 3, 3, Cout) with the taps rotated by 180 degrees).
// All of them are rebuilt from the flat bf16 weight buffer in ONE launch per optimizer step: a table of 32 x 32 tiles
// {src, dst, leading dimensions, extents} built once by the host (the buffers never move), one CTA per tile, a
// shared-memory transpose with coalesced reads and writes.  (Per-layer torch transposes were 66 launches, 1.4 ms.)
namespace mvfb {
namespace {
struct TransposeTile {
  const __nv_bfloat16* src;   // element (r, c) at src[r * lds + c]
  __nv_bfloat16* dst;         // element (c, r) at dst[c * ldd + r]
  int lds, ldd, rows, cols;   // extents of the matrix this tile belongs to
  int r0, c0;                 // tile origin
};
__global__ void __launch_bounds__(256)
transpose_tiles_kernel(const TransposeTile* __restrict__ tiles) {
  __shared__ __nv_bfloat16 tile[32][34];
  const TransposeTile t = tiles[blockIdx.x];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = t.r0 + ty + 8 * k, c = t.c0 + tx;
    if (r < t.rows && c < t.cols) tile[ty + 8 * k][tx] = t.src[(size_t)r * t.lds + c];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = t.c0 + ty + 8 * k, r = t.r0 + tx;
    if (r < t.rows && c < t.cols) t.dst[(size_t)c * t.ldd + r] = tile[tx][ty + 8 * k];
  }
}
}  // namespace
}  // namespace mvfb

extern "C" int transpose_tiles(const void* tiles_dev, long long ntiles, mvfb_stream_t stream) {
  MVFB_CHECK(tiles_dev && ntiles > 0 && ntiles < (1LL << 31), MVFB_ERR_ARG, "bad tile table");
  mvfb::transpose_tiles_kernel<<<(unsigned)ntiles, 256, 0, (cudaStream_t)stream>>>((const mvfb::TransposeTile*)tiles_dev);
  mvfb::count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

// Train/eval BatchNorm2d (+ residual add) (+ ReLU) on NHWC bf16 activations, forward and backward.
//
// Replaces, per Bottleneck (backbones/resnet.py:208-244): norm1+relu, norm2+relu, norm3 (+downsample norm)
// + `out += identity` + relu -- in the reference each of those is a separate full-tensor HBM round trip
// (and ncu on the torch build shows the four ATen batch-norm kernels at 54 % of the training step,
// profiles/r01_step_launches_torch_bn.csv).  Here a layer costs:
//   forward :  [stats: 1 read -- or 0 when the producing GEMM already accumulated them]  + apply: 1 read, 1 write
//   backward:  reduce: reads g, y(mask), x   + apply: reads g, y, x, writes dx (+ d residual)
// All four kernels are HBM-bound streaming kernels: a thread owns 8 consecutive channels (16-byte vectors),
// a CTA covers whole rows so that the channel vector of a thread never changes (scale/shift live in registers),
// column reductions go registers -> shared memory -> one fp32 atomic per channel per CTA.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "mvf_internal.cuh"
#include "ptx.cuh"

namespace mvfb {

namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerIter = 4;     // rows a thread has in flight (measured: 1 -> 97.4 ms step, 4 -> 90.9 ms, 8 -> 94.5 ms)

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
  return o;
}

// Reduce K per-thread accumulators over the threads of the CTA that share a channel vector (tid % vecs) and add the
// result to dst[q * C + channel].   K = 16: (8 sums, 8 second sums).
template <int K>
__device__ __forceinline__ void column_reduce_atomic(float (&acc)[K], int vecs, int rows_par, int vec, int rowlane,
                                                     float* dst, int C, float* scratch) {
  // scratch: [rows_par][vecs][K]; the (kThreads % vecs) left-over threads hold nothing
  if (rowlane < rows_par) {
    float* mine = scratch + ((size_t)rowlane * vecs + vec) * K;
#pragma unroll
    for (int q = 0; q < K; ++q) mine[q] = acc[q];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < vecs * K; i += kThreads) {
    float s = 0.f;
    for (int r = 0; r < rows_par; ++r) s += scratch[(size_t)r * vecs * K + i];
    const int v = i / K, q = i - v * K;               // q = kind * 8 + j
    atomicAdd(&dst[(q >> 3) * C + v * 8 + (q & 7)], s);
  }
}

// ---------------------------------------------------------------- forward statistics
__global__ void __launch_bounds__(kThreads)
bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, long long M, int C, float* __restrict__ sums) {
  extern __shared__ float scratch[];
  const int vecs = C >> 3;
  const int rows_par = kThreads / vecs;                 // host guarantees vecs <= kThreads
  const int vec = threadIdx.x % vecs, rowlane = threadIdx.x / vecs;
  float acc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) acc[q] = 0.f;
  if (rowlane < rows_par) {
    const long long stride = (long long)gridDim.x * rows_par;
    long long r = (long long)blockIdx.x * rows_par + rowlane;
    const __nv_bfloat16* p = x + vec * 8;
    for (; r + 3 * stride < M; r += 4 * stride) {       // four independent 16-byte loads in flight per thread
      uint4 v0 = ldg16(p + r * ldx), v1 = ldg16(p + (r + stride) * ldx), v2 = ldg16(p + (r + 2 * stride) * ldx),
            v3 = ldg16(p + (r + 3 * stride) * ldx);
      float f[8];
      unpack8(v0, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j] += f[j]; acc[8 + j] = fmaf(f[j], f[j], acc[8 + j]); }
      unpack8(v1, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j] += f[j]; acc[8 + j] = fmaf(f[j], f[j], acc[8 + j]); }
      unpack8(v2, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j] += f[j]; acc[8 + j] = fmaf(f[j], f[j], acc[8 + j]); }
      unpack8(v3, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j] += f[j]; acc[8 + j] = fmaf(f[j], f[j], acc[8 + j]); }
    }
    for (; r < M; r += stride) {
      float f[8];
      unpack8(ldg16(p + r * ldx), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j] += f[j]; acc[8 + j] = fmaf(f[j], f[j], acc[8 + j]); }
    }
  }
  column_reduce_atomic<16>(acc, vecs, rows_par, vec, rowlane, sums, C, scratch);
}

// ---------------------------------------------------------------- forward apply
struct ApplyArgs {
  const __nv_bfloat16 *x, *res;
  __nv_bfloat16* y;
  uint8_t* mask;                         // optional: 1 bit per element, (y > 0); byte (row, 8-channel vector)
  long long ldx, ldr, ldy, M;
  int C, relu, training;
  float eps, momentum;
  const float *sums, *gamma, *beta;
  float *running_mean, *running_var, *save_mean, *save_rstd;
};

template <int U>
__global__ void __launch_bounds__(kThreads)
bn_apply_kernel(const ApplyArgs a) {
  const int vecs = a.C >> 3;
  const int rows_par = kThreads / vecs;
  const int vec = threadIdx.x % vecs, rowlane = threadIdx.x / vecs;
  if (rowlane >= rows_par) return;
  float scale[8], shift[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = vec * 8 + j;
    float mean, rstd;
    if (a.training) {
      const double m = (double)a.M;
      const double mu = (double)a.sums[c] / m;
      double var = (double)a.sums[a.C + c] / m - mu * mu;
      if (var < 0) var = 0;
      mean = (float)mu;
      rstd = (float)(1.0 / sqrt(var + (double)a.eps));
      if (blockIdx.x == 0 && rowlane == 0) {
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
      if (blockIdx.x == 0 && rowlane == 0 && a.save_mean) { a.save_mean[c] = mean; a.save_rstd[c] = rstd; }
    }
    scale[j] = a.gamma[c] * rstd;
    shift[j] = a.beta[c] - mean * scale[j];
  }
  const long long stride = (long long)gridDim.x * rows_par;
  const int co = vec * 8;
  // four rows per iteration, every load issued before the first use: ~100 KB in flight per SM (the kernel is
  // register-limited to a few CTAs per SM, one row at a time left HBM latency exposed: 64-69 % of the copy rate)
  auto body = [&](long long r, const uint4& xq, const uint4& rq) {
    float f[8];
    unpack8(xq, f);
    if (a.res) {
      float g[8];
      unpack8(rq, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], scale[j], shift[j]) + g[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], scale[j], shift[j]);
    }
    if (a.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      if (a.mask) {
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) bits |= (f[j] > 0.f ? 1u : 0u) << j;
        a.mask[r * vecs + vec] = (uint8_t)bits;
      }
    }
    *reinterpret_cast<uint4*>(a.y + r * a.ldy + co) = pack8(f);
  };
  long long r = (long long)blockIdx.x * rows_par + rowlane;
  for (; r + (U - 1) * stride < a.M; r += U * stride) {
    uint4 xq[U], rq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) xq[u] = ldg16(a.x + (r + u * stride) * a.ldx + co);
#pragma unroll
    for (int u = 0; u < U; ++u) rq[u] = a.res ? ldg16(a.res + (r + u * stride) * a.ldr + co) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < U; ++u) body(r + u * stride, xq[u], rq[u]);
  }
  for (; r < a.M; r += stride)
    body(r, ldg16(a.x + r * a.ldx + co), a.res ? ldg16(a.res + r * a.ldr + co) : make_uint4(0, 0, 0, 0));
}

// ---------------------------------------------------------------- backward
struct BwdArgs {
  const uint8_t* mask;                   // ReLU bit mask written by bn_apply (preferred over reading y back)
  const __nv_bfloat16 *g, *y, *x;        // dL/d(out), out (ReLU mask source; null = no ReLU), conv output
  __nv_bfloat16 *dx, *dres;              // dL/d(conv output), dL/d(residual) (null if none)
  long long ldg, ldy, ldx, lddx, lddr, M;
  int C, training;
  const float *gamma, *mean, *rstd;
  float* sums;                           // [2][C]: sum dy', sum dy' * xhat
  float *dgamma, *dbeta;
};

template <int U>
__global__ void __launch_bounds__(kThreads)
bn_bwd_reduce_kernel(const BwdArgs a) {
  extern __shared__ float scratch[];
  const int vecs = a.C >> 3;
  const int rows_par = kThreads / vecs;
  const int vec = threadIdx.x % vecs, rowlane = threadIdx.x / vecs;
  float acc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) acc[q] = 0.f;
  if (rowlane < rows_par) {
    float rs[8], mr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      rs[j] = a.rstd[vec * 8 + j];
      mr[j] = -a.mean[vec * 8 + j] * rs[j];
    }
    const long long stride = (long long)gridDim.x * rows_par;
    const int co = vec * 8;
    auto body = [&](const uint4& gq, const uint4& xq, uint32_t bits, const uint4& yq) {
      float gv[8], xv[8];
      unpack8(gq, gv);
      unpack8(xq, xv);
      if (a.mask) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gv[j] = (bits >> j) & 1u ? gv[j] : 0.f;
      } else if (a.y) {
        float yv[8];
        unpack8(yq, yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) gv[j] = yv[j] > 0.f ? gv[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] += gv[j];
        acc[8 + j] = fmaf(gv[j], fmaf(xv[j], rs[j], mr[j]), acc[8 + j]);
      }
    };
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    long long r = (long long)blockIdx.x * rows_par + rowlane;
    for (; r + (U - 1) * stride < a.M; r += U * stride) {             // four rows per iteration, loads first
      uint4 gq[U], xq[U], yq[U];
      uint32_t bits[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long ru = r + u * stride;
        gq[u] = ldg16(a.g + ru * a.ldg + co);
        xq[u] = ldg16(a.x + ru * a.ldx + co);
        bits[u] = a.mask ? __ldg(a.mask + ru * vecs + vec) : 0u;
        yq[u] = (!a.mask && a.y) ? ldg16(a.y + ru * a.ldy + co) : z4;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) body(gq[u], xq[u], bits[u], yq[u]);
    }
    for (; r < a.M; r += stride)
      body(ldg16(a.g + r * a.ldg + co), ldg16(a.x + r * a.ldx + co), a.mask ? __ldg(a.mask + r * vecs + vec) : 0u,
           (!a.mask && a.y) ? ldg16(a.y + r * a.ldy + co) : z4);
  }
  column_reduce_atomic<16>(acc, vecs, rows_par, vec, rowlane, a.sums, a.C, scratch);
}

template <int U>
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_kernel(const BwdArgs a) {
  const int vecs = a.C >> 3;
  const int rows_par = kThreads / vecs;
  const int vec = threadIdx.x % vecs, rowlane = threadIdx.x / vecs;
  if (rowlane >= rows_par) return;
  // dx = ga * (dy' - b - xhat * c)   with  ga = gamma * rstd,  b = sum dy' / M,  c = sum dy' xhat / M  (0 in eval)
  float rs[8], mr[8], ga[8], b[8], cc[8];
  const float inv_m = 1.f / (float)a.M;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = vec * 8 + j;
    rs[j] = a.rstd[c];
    mr[j] = -a.mean[c] * rs[j];
    ga[j] = a.gamma[c] * rs[j];
    const float s1 = a.sums[c], s2 = a.sums[a.C + c];
    b[j] = a.training ? s1 * inv_m : 0.f;
    cc[j] = a.training ? s2 * inv_m : 0.f;
    if (blockIdx.x == 0 && rowlane == 0) {
      a.dbeta[c] = s1;
      a.dgamma[c] = s2;
    }
  }
  const long long stride = (long long)gridDim.x * rows_par;
  const int co = vec * 8;
  auto body = [&](long long r, const uint4& gq, const uint4& xq, uint32_t bits, const uint4& yq) {
    float gv[8], xv[8];
    unpack8(gq, gv);
    unpack8(xq, xv);
    if (a.mask) {
#pragma unroll
      for (int j = 0; j < 8; ++j) gv[j] = (bits >> j) & 1u ? gv[j] : 0.f;
    } else if (a.y) {
      float yv[8];
      unpack8(yq, yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) gv[j] = yv[j] > 0.f ? gv[j] : 0.f;
    }
    if (a.dres) *reinterpret_cast<uint4*>(a.dres + r * a.lddr + co) = pack8(gv);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = fmaf(xv[j], rs[j], mr[j]);
      o[j] = ga[j] * (gv[j] - b[j] - xh * cc[j]);
    }
    *reinterpret_cast<uint4*>(a.dx + r * a.lddx + co) = pack8(o);
  };
  const uint4 z4 = make_uint4(0, 0, 0, 0);
  long long r = (long long)blockIdx.x * rows_par + rowlane;
  for (; r + (U - 1) * stride < a.M; r += U * stride) {               // four rows per iteration, loads first
    uint4 gq[U], xq[U], yq[U];
    uint32_t bits[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ru = r + u * stride;
      gq[u] = ldg16(a.g + ru * a.ldg + co);
      xq[u] = ldg16(a.x + ru * a.ldx + co);
      bits[u] = a.mask ? __ldg(a.mask + ru * vecs + vec) : 0u;
      yq[u] = (!a.mask && a.y) ? ldg16(a.y + ru * a.ldy + co) : z4;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) body(r + u * stride, gq[u], xq[u], bits[u], yq[u]);
  }
  for (; r < a.M; r += stride)
    body(r, ldg16(a.g + r * a.ldg + co), ldg16(a.x + r * a.ldx + co), a.mask ? __ldg(a.mask + r * vecs + vec) : 0u,
         (!a.mask && a.y) ? ldg16(a.y + r * a.ldy + co) : z4);
}

// rows x cols strided 2-D copy in 16-byte vectors (cols % 8 == 0): y[r, :cols] = x[r, :cols]
__global__ void __launch_bounds__(kThreads)
copy_cols_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy,
                 long long M, int vecs) {
  const long long total = M * vecs;
  const long long stride = (long long)gridDim.x * kThreads;
  long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  for (; i + 3 * stride < total; i += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long k = i + u * stride, r = k / vecs;
      v[u] = ldg16(x + r * ldx + (k - r * vecs) * 8);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long k = i + u * stride, r = k / vecs;
      *reinterpret_cast<uint4*>(y + r * ldy + (k - r * vecs) * 8) = v[u];
    }
  }
  for (; i < total; i += stride) {
    const long long r = i / vecs;
    *reinterpret_cast<uint4*>(y + r * ldy + (i - r * vecs) * 8) = ldg16(x + r * ldx + (i - r * vecs) * 8);
  }
}

int grid_for(long long M, int C) {
  const int vecs = C / 8;
  const int rows_par = kThreads / vecs;
  long long blocks = (M + rows_par - 1) / rows_par;
  const long long cap = (long long)num_sms() * 8;         // 8 resident 256-thread CTAs per SM, a few rows each
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

size_t scratch_bytes(int C) {
  const int vecs = C / 8;
  const int rows_par = kThreads / vecs;
  return (size_t)rows_par * vecs * 16 * sizeof(float);
}

int check_bn(const mvfb_bn_desc* d) {
  MVFB_CHECK(d != nullptr, MVFB_ERR_ARG, "null descriptor");
  MVFB_CHECK(d->M > 0 && d->C > 0, MVFB_ERR_ARG, "bad shape M=%lld C=%d", d->M, d->C);
  MVFB_CHECK(d->C % 8 == 0 && d->C / 8 <= kThreads, MVFB_ERR_UNSUPPORTED, "C=%d must be a multiple of 8 and <= %d", d->C,
             8 * kThreads);
  return MVFB_OK;
}

bool ok16(const void* p, long long ld) { return ((uintptr_t)p & 15) == 0 && ld % 8 == 0; }

}  // namespace

}  // namespace mvfb

using namespace mvfb;

extern "C" {

int copy_cols(const void* x, long long ldx, void* y, long long ldy, long long M, int cols, mvfb_stream_t stream) {
  MVFB_CHECK(x && y && M > 0 && cols > 0 && cols % 8 == 0, MVFB_ERR_ARG, "bad copy_cols arguments");
  MVFB_CHECK(ok16(x, ldx) && ok16(y, ldy), MVFB_ERR_UNSUPPORTED, "tensors must be 16-byte aligned with ld %% 8 == 0");
  const int vecs = cols / 8;
  long long blocks = (M * vecs + kThreads * 4 - 1) / (kThreads * 4);
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  copy_cols_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)y, ldy, M, vecs);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int bn_stats(const mvfb_bn_desc* d, const void* x, long long ldx, float* sums, mvfb_stream_t stream) {
  int rc = check_bn(d);
  if (rc) return rc;
  MVFB_CHECK(x && sums, MVFB_ERR_ARG, "null x / sums");
  MVFB_CHECK(ok16(x, ldx), MVFB_ERR_UNSUPPORTED, "x must be 16-byte aligned with ld %% 8 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  MVFB_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * d->C, st));
  const size_t sm = scratch_bytes(d->C);
  static DevOnce once;
  if (once.pending()) {
    MVFB_CUDA(cudaFuncSetAttribute(bn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    MVFB_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_kernel<kRowsPerIter>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    once.done();
  }
  bn_stats_kernel<<<grid_for(d->M, d->C), kThreads, sm, st>>>((const __nv_bfloat16*)x, ldx, d->M, d->C, sums);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int bn_apply(const mvfb_bn_desc* d, const void* x, long long ldx, const void* residual, long long ldr, void* y,
             long long ldy, const float* sums, const float* gamma, const float* beta, float* running_mean,
             float* running_var, float* save_mean, float* save_rstd, unsigned char* relu_mask, mvfb_stream_t stream) {
  int rc = check_bn(d);
  if (rc) return rc;
  MVFB_CHECK(x && y && gamma && beta, MVFB_ERR_ARG, "null x / y / gamma / beta");
  MVFB_CHECK(!d->training || (sums && save_mean && save_rstd), MVFB_ERR_ARG, "training needs sums and save_mean/save_rstd");
  MVFB_CHECK(d->training || (running_mean && running_var), MVFB_ERR_ARG, "eval mode needs running statistics");
  MVFB_CHECK(ok16(x, ldx) && ok16(y, ldy) && (!residual || ok16(residual, ldr)), MVFB_ERR_UNSUPPORTED,
             "tensors must be 16-byte aligned with ld %% 8 == 0");
  ApplyArgs a;
  a.x = (const __nv_bfloat16*)x; a.res = (const __nv_bfloat16*)residual; a.y = (__nv_bfloat16*)y;
  a.mask = d->relu ? relu_mask : nullptr;
  a.ldx = ldx; a.ldr = ldr; a.ldy = ldy; a.M = d->M; a.C = d->C; a.relu = d->relu; a.training = d->training;
  a.eps = d->eps; a.momentum = d->momentum; a.sums = sums; a.gamma = gamma; a.beta = beta;
  a.running_mean = running_mean; a.running_var = running_var; a.save_mean = save_mean; a.save_rstd = save_rstd;
  bn_apply_kernel<kRowsPerIter><<<grid_for(d->M, d->C), kThreads, 0, (cudaStream_t)stream>>>(a);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

int bn_bwd(const mvfb_bn_desc* d, const void* g, long long ldg, const void* y, long long ldy, const void* x,
           long long ldx, const float* gamma, const float* mean, const float* rstd, void* dx, long long lddx,
           void* dres, long long lddr, float* dgamma, float* dbeta, float* sums, const unsigned char* relu_mask,
           mvfb_stream_t stream) {
  int rc = check_bn(d);
  if (rc) return rc;
  MVFB_CHECK(g && x && gamma && mean && rstd && dx && dgamma && dbeta && sums, MVFB_ERR_ARG, "null argument");
  MVFB_CHECK(ok16(g, ldg) && ok16(x, ldx) && ok16(dx, lddx) && (!y || relu_mask || ok16(y, ldy)) && (!dres || ok16(dres, lddr)),
             MVFB_ERR_UNSUPPORTED, "tensors must be 16-byte aligned with ld %% 8 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  BwdArgs a;
  a.mask = d->relu ? relu_mask : nullptr;
  a.g = (const __nv_bfloat16*)g; a.y = d->relu ? (const __nv_bfloat16*)y : nullptr; a.x = (const __nv_bfloat16*)x;
  a.dx = (__nv_bfloat16*)dx; a.dres = (__nv_bfloat16*)dres;
  a.ldg = ldg; a.ldy = ldy; a.ldx = ldx; a.lddx = lddx; a.lddr = lddr; a.M = d->M; a.C = d->C; a.training = d->training;
  a.gamma = gamma; a.mean = mean; a.rstd = rstd; a.sums = sums; a.dgamma = dgamma; a.dbeta = dbeta;
  MVFB_CHECK(!d->relu || y || relu_mask, MVFB_ERR_ARG, "relu backward needs the forward output y or its bit mask");
  MVFB_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * d->C, st));
  static DevOnce once;
  if (once.pending()) {
    MVFB_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_kernel<kRowsPerIter>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    once.done();
  }
  const int grid = grid_for(d->M, d->C);
  bn_bwd_reduce_kernel<kRowsPerIter><<<grid, kThreads, scratch_bytes(d->C), st>>>(a);
  count_launch();
  MVFB_LAUNCH_CHECK();
  bn_bwd_apply_kernel<kRowsPerIter><<<grid, kThreads, 0, st>>>(a);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

}  // extern "C"

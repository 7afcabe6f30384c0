// MVF forward, third generation: persistent, warp-specialised frame-stream kernel (bf16 NHWC).
//
// ncu on the first two generations (profiles/r01_mvf_v1_*, r01_mvf_v2_*) showed the forward pass bound by
// instruction issue and per-CTA fixed cost (parameter loads, barrier setup, launch ramp, load/compute phases
// that do not overlap), not by HBM.  This kernel removes the fixed costs and overlaps everything:
//   * persistent CTAs: CTA b owns ONE channel group (and one H tile) for the whole launch and walks over its
//     share of the clips, so stencil coefficients / BN affine are set up once;
//   * the frames of those clips form ONE continuous stream through a ring of R shared-memory slots: a
//     producer warp issues cp.async.bulk.tensor.4d loads (box = Cg x (W+2) x (Hs+2) x 1, out-of-bounds
//     zero fill = the convolution's H/W zero padding) as slots are released; consumer warps wait on the
//     slot's `full` mbarrier and release it with one `empty` arrive per warp -- no CTA-wide barrier in the
//     steady state, loads run R-2 frames ahead of the arithmetic;
//   * every consumer thread owns exactly one (pixel, 8-channel vector) item; the unpacked centre values of
//     frames t-1, t, t+1 roll through registers (zeros at the clip boundary), so a frame's slot is touched
//     only for the centre of t+1 and the four H/W neighbours of t: 5 vector loads + unpacks per output vector;
//   * packed fp32x2 arithmetic (FFMA2), BN folded to one scale/shift, hard-swish via FFMA.SAT.
// Train-mode BatchNorm: PASS_STATS accumulates per-thread (sum z, sum z^2) over ALL the CTA's clips and
// writes one partial row per CTA; PASS_TRAIN reduces the rows of its channel group in its prologue while its
// first loads are in flight.
#include <cuda_bf16.h>

#include "common.cuh"
#include "mvf_internal.cuh"
#include "ptx.cuh"

namespace mvfb {

namespace {

constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxItems = 416;          // consumer threads per CTA (one item each); 14 warps = 4 per SM sub-partition (16K registers each) -> at most 128 registers per thread
constexpr int kMaxThreads = kMaxItems + 32;
constexpr int kMaxRing = 12;

constexpr int PASS_APPLY = 0;   // eval-mode BN (running stats) or no BN at all
constexpr int PASS_STATS = 1;   // train: partial sums only
constexpr int PASS_TRAIN = 2;   // train: batch statistics from the partials, then apply

struct StreamGeo {
  int N, T, Cs, H, W;
  int Cg, G, ngroups;     // channels per CTA, 8-channel vectors per pixel, channel groups
  int hsplit, Hs;         // H tiles and rows per tile
  int Hp, Wp;             // padded tile extents (Hs+2, W+2)
  int slot;               // bytes of one frame slot (multiple of 128)
  int R;                  // ring slots
  int items;              // Hs*W*G  (<= kMaxItems)
  int cwarps;             // consumer warps
  int P;                  // CTAs sharing one (channel group, H tile): clips are dealt round-robin
};

struct StreamArgs {
  StreamGeo g;
  int use_hs;
  float eps, momentum;
  const float *wt, *wh, *ww, *gamma, *beta;
  float *running_mean, *running_var, *save_mean, *save_rstd;
  float* partials;        // [grid][2*Cg]
  __nv_bfloat16* y;
  long long y_pix;
};

struct F8 {
  float2 p[4];
};
__device__ __forceinline__ F8 zero8() {
  F8 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) r.p[j] = make_float2(0.f, 0.f);
  return r;
}
__device__ __forceinline__ F8 lds_unpack8(const uint8_t* ptr) {
  const uint4 v = *reinterpret_cast<const uint4*>(ptr);
  F8 r;
  r.p[0] = make_float2(bf16_lo(v.x), bf16_hi(v.x));
  r.p[1] = make_float2(bf16_lo(v.y), bf16_hi(v.y));
  r.p[2] = make_float2(bf16_lo(v.z), bf16_hi(v.z));
  r.p[3] = make_float2(bf16_lo(v.w), bf16_hi(v.w));
  return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ F8 unpack8(const uint4& v) {
  F8 r;
  r.p[0] = make_float2(bf16_lo(v.x), bf16_hi(v.x));
  r.p[1] = make_float2(bf16_lo(v.y), bf16_hi(v.y));
  r.p[2] = make_float2(bf16_lo(v.z), bf16_hi(v.z));
  r.p[3] = make_float2(bf16_lo(v.w), bf16_hi(v.w));
  return r;
}
__device__ __forceinline__ bool try_wait_u32(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void slow_wait_u32(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!try_wait_u32(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void wait_u32(uint32_t bar, uint32_t parity) {
  if (!try_wait_u32(bar, parity)) slow_wait_u32(bar, parity);
}
__device__ __forceinline__ void arrive_u32(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ F8 lds_f8(const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  F8 r;
  r.p[0] = make_float2(a.x, a.y); r.p[1] = make_float2(a.z, a.w);
  r.p[2] = make_float2(b.x, b.y); r.p[3] = make_float2(b.z, b.w);
  return r;
}
__device__ __forceinline__ void fma8(F8& z, const F8& k, const F8& x) {
#pragma unroll
  for (int j = 0; j < 4; ++j) z.p[j] = __ffma2_rn(k.p[j], x.p[j], z.p[j]);
}

template <int PASS>
__global__ void __launch_bounds__(kMaxThreads, 1)
mvf_stream_fwd_kernel(const __grid_constant__ CUtensorMap tmx, const StreamArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const StreamGeo& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x;
  // CTA -> (channel group, H tile, clip lane)
  const int cg = blockIdx.x % g.ngroups;
  const int rest = blockIdx.x / g.ngroups;
  const int hs = rest % g.hsplit, p = rest / g.hsplit;
  const int c0 = cg * g.Cg, h0 = hs * g.Hs;
  const int nclips = p < g.N ? (g.N - p + g.P - 1) / g.P : 0;
  const int Q = nclips * g.T;                                  // frames in this CTA's stream

  uint64_t* full = reinterpret_cast<uint64_t*>(smem);         // [R]
  uint64_t* empty = full + 16;                                 // [R]
  uint8_t* slots = smem + 256;
  float* s_coef = reinterpret_cast<float*>(slots + (size_t)g.R * g.slot);   // [7][Cg]
  float* s_scale = s_coef + 7 * g.Cg;                          // [Cg]
  float* s_shift = s_scale + g.Cg;                             // [Cg]
  double* s_dpart = reinterpret_cast<double*>(s_shift + g.Cg); // [kMaxThreads]
  double* s_dsum = s_dpart + kMaxThreads;                      // [2*Cg]
  float* s_red = reinterpret_cast<float*>(s_dsum + 128);      // [cwarps][G][16]
  float* s_out = s_red + 16 * 8 * 16;                          // [G][16]

  const uint32_t frame_bytes = (uint32_t)(g.Hp * g.Wp * g.Cg * 2);
  if (tid == 0) {
    tma_prefetch_desc(&tmx);
    for (int s = 0; s < g.R; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], g.cwarps);
    }
    fence_barrier_init();
  }
  __syncthreads();
  // ---- the first R frames can be requested before anything else happens
  const bool producer = warp == g.cwarps;
  if (producer && lane == 0) {
    const int pre = Q < g.R ? Q : g.R;
    int n = p, t = 0;
    for (int q = 0; q < pre; ++q) {
      mbar_arrive_expect_tx(&full[q], frame_bytes);
      tma_load_4d(slots + (size_t)q * g.slot, &tmx, &full[q], c0, -1, h0 - 1, n * g.T + t);
      if (++t == g.T) { t = 0; n += g.P; }
    }
  }

  // ---- per-channel constants, computed once per CTA
  for (int i = tid; i < 7 * g.Cg; i += nthreads) {
    const int q = i / g.Cg, ch = i - q * g.Cg, c3 = (c0 + ch) * 3;
    float v;
    switch (q) {
      case 0: v = a.wt[c3 + 1] + (a.wh ? a.wh[c3 + 1] : 0.f) + (a.ww ? a.ww[c3 + 1] : 0.f); break;   // centre
      case 1: v = a.wt[c3]; break;                                                                 // t-1
      case 2: v = a.wt[c3 + 2]; break;                                                             // t+1
      case 3: v = a.wh ? a.wh[c3] : 0.f; break;                                                    // h-1
      case 4: v = a.wh ? a.wh[c3 + 2] : 0.f; break;                                                // h+1
      case 5: v = a.ww ? a.ww[c3] : 0.f; break;                                                    // w-1
      default: v = a.ww ? a.ww[c3 + 2] : 0.f; break;                                               // w+1
    }
    s_coef[i] = v;
  }
  if (PASS == PASS_TRAIN) {
    const int per = 2 * g.Cg, rows = g.hsplit * g.P;          // partial rows of this channel group
    const int parts = nthreads / per;
    const int k = tid % per, part = tid / per;
    if (part < parts) {
      double acc = 0.0;
      for (int r = part; r < rows; r += parts) acc += (double)a.partials[((size_t)r * g.ngroups + cg) * per + k];
      s_dpart[part * per + k] = acc;
    }
    __syncthreads();
    if (tid < per) {
      double acc = 0.0;
      for (int q = 0; q < parts; ++q) acc += s_dpart[q * per + tid];
      s_dsum[tid] = acc;
    }
    __syncthreads();
    if (tid < g.Cg) {
      const int c = c0 + tid;
      const double m = (double)g.N * g.T * g.H * g.W;
      const double mu = s_dsum[tid * 2] / m;
      double var = s_dsum[tid * 2 + 1] / m - mu * mu;
      if (var < 0) var = 0;
      const float mean = (float)mu, rstd = (float)(1.0 / sqrt(var + (double)a.eps));
      const float sc = a.gamma[c] * rstd;
      s_scale[tid] = sc;
      s_shift[tid] = a.beta[c] - mean * sc;
      if (rest == 0) {
        a.save_mean[c] = mean;
        a.save_rstd[c] = rstd;
        if (a.running_mean) {
          const double unb = m > 1 ? var * m / (m - 1) : var;
          a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mean;
          a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unb;
        }
      }
    }
  } else if (PASS == PASS_APPLY) {
    if (tid < g.Cg) {
      const int c = c0 + tid;
      float sc = 1.f, sh = 0.f;
      if (a.use_hs) {
        const float mean = a.running_mean[c], rstd = 1.f / sqrtf(a.running_var[c] + a.eps);
        sc = a.gamma[c] * rstd;
        sh = a.beta[c] - mean * sc;
        if (rest == 0 && a.save_mean) { a.save_mean[c] = mean; a.save_rstd[c] = rstd; }
      }
      s_scale[tid] = sc;
      s_shift[tid] = sh;
    }
  }
  __syncthreads();

  F8 sum = zero8(), sq = zero8();
  const int vec = tid % g.G;

  if (producer) {
    // ===== TMA producer: refill slots as the consumers release them
    if (lane == 0) {
      int s = 0, use = 1;                                       // frame q = R goes to slot 0, second use
      int n = p, t = 0;
      for (int q = 0; q < g.R && q < Q; ++q) { if (++t == g.T) { t = 0; n += g.P; } }
      for (int q = g.R; q < Q; ++q) {
        mbar_wait(&empty[s], (use - 1) & 1);
        mbar_arrive_expect_tx(&full[s], frame_bytes);
        tma_load_4d(slots + (size_t)s * g.slot, &tmx, &full[s], c0, -1, h0 - 1, n * g.T + t);
        if (++t == g.T) { t = 0; n += g.P; }
        if (++s == g.R) { s = 0; ++use; }
      }
    }
  } else {
    // ===== consumers: one (pixel, 8-channel vector) item per thread
    const bool active = tid < g.items;
    const int pix = tid / g.G;
    const int hl = pix / g.W, w = pix - hl * g.W;             // row within the H tile, column
    const int pixb = g.Cg * 2, rowb = g.Wp * pixb;
    const int off = ((hl + 1) * g.Wp + (w + 1)) * pixb + vec * 16;
    F8 kc = lds_f8(s_coef + vec * 8), kt0 = lds_f8(s_coef + g.Cg + vec * 8), kt2 = lds_f8(s_coef + 2 * g.Cg + vec * 8);
    F8 kh0 = lds_f8(s_coef + 3 * g.Cg + vec * 8), kh2 = lds_f8(s_coef + 4 * g.Cg + vec * 8);
    F8 kw0 = lds_f8(s_coef + 5 * g.Cg + vec * 8), kw2 = lds_f8(s_coef + 6 * g.Cg + vec * 8);
    F8 scale = zero8(), shift = zero8();
    if (PASS != PASS_STATS) { scale = lds_f8(s_scale + vec * 8); shift = lds_f8(s_shift + vec * 8); }
    const bool hs_on = a.use_hs != 0;
    const size_t frame_elems = (size_t)g.H * g.W * a.y_pix;
    const size_t ypix = ((size_t)(h0 + hl) * g.W + w) * a.y_pix + c0 + vec * 8;
    // 32-bit shared-space addresses keep the per-frame bookkeeping to a handful of integer instructions
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
    const uint32_t base = smem_u32(slots) + (uint32_t)off;
    const uint32_t slotb = (uint32_t)g.slot, ringb = (uint32_t)g.R * slotb;

    uint32_t cur = base, fb = full0, ph = 0;                   // current frame: slot address, full barrier, parity
    for (int kclip = 0; kclip < nclips; ++kclip) {
      const int n = p + kclip * g.P;
      __nv_bfloat16* yp = a.y + (size_t)n * g.T * frame_elems + ypix;
      F8 xm = zero8(), xc = zero8(), xp;
      wait_u32(fb, ph);
      if (active) xc = unpack8(lds128(cur));
#pragma unroll 1
      for (int t = 0; t < g.T; ++t) {
        uint32_t nxt = cur + slotb, fb1 = fb + 8, ph1 = ph;
        if (nxt == base + ringb) { nxt = base; fb1 = full0; ph1 ^= 1; }
        xp = zero8();
        if (t + 1 < g.T) {
          wait_u32(fb1, ph1);
          if (active) xp = unpack8(lds128(nxt));
        }
        if (active) {
          F8 z;
#pragma unroll
          for (int j = 0; j < 4; ++j) z.p[j] = __fmul2_rn(kc.p[j], xc.p[j]);
          fma8(z, kt0, xm);
          fma8(z, kt2, xp);
          fma8(z, kh0, unpack8(lds128(cur - rowb)));
          fma8(z, kh2, unpack8(lds128(cur + rowb)));
          fma8(z, kw0, unpack8(lds128(cur - pixb)));
          fma8(z, kw2, unpack8(lds128(cur + pixb)));
          if (PASS == PASS_STATS) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              sum.p[j] = __fadd2_rn(sum.p[j], z.p[j]);
              sq.p[j] = __ffma2_rn(z.p[j], z.p[j], sq.p[j]);
            }
          } else {
            if (hs_on) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 u = __ffma2_rn(z.p[j], scale.p[j], shift.p[j]);
                float2 sg;
                sg.x = __saturatef(fmaf(u.x, 1.f / 6.f, 0.5f));
                sg.y = __saturatef(fmaf(u.y, 1.f / 6.f, 0.5f));
                z.p[j] = __fmul2_rn(u, sg);
              }
            }
            uint4 o;
            o.x = pack_bf16(z.p[0].x, z.p[0].y); o.y = pack_bf16(z.p[1].x, z.p[1].y);
            o.z = pack_bf16(z.p[2].x, z.p[2].y); o.w = pack_bf16(z.p[3].x, z.p[3].y);
            *reinterpret_cast<uint4*>(yp) = o;
          }
        }
        yp += frame_elems;
        __syncwarp();
        if (lane == 0) arrive_u32(empty0 + (fb - full0));       // this warp is done with frame t's slot
        xm = xc;
        xc = xp;
        cur = nxt;
        fb = fb1;
        ph = ph1;
      }
    }
  }

  if (PASS == PASS_STATS) {
    // CTA reduction of (sum, sumsq) by channel vector: lanes with equal lane % G own the same channels
    float acc[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[2 * j] = sum.p[j].x; acc[2 * j + 1] = sum.p[j].y;
      acc[8 + 2 * j] = sq.p[j].x; acc[8 + 2 * j + 1] = sq.p[j].y;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      float v = acc[q];
      for (int o = 16; o >= g.G; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[q] = v;
    }
    if (!producer && lane < g.G) {
#pragma unroll
      for (int q = 0; q < 16; ++q) s_red[(warp * g.G + lane) * 16 + q] = acc[q];
    }
    __syncthreads();
    for (int i = tid; i < 2 * g.Cg; i += nthreads) {
      const int ch = i >> 1, kind = i & 1;
      const int idx = (ch / 8) * 16 + kind * 8 + (ch % 8);
      float v = 0.f;
      for (int wv = 0; wv < g.cwarps; ++wv) v += s_red[wv * g.G * 16 + idx];
      a.partials[(size_t)blockIdx.x * 2 * g.Cg + i] = v;
    }
  }
  (void)s_out;
}

size_t stream_smem(const StreamGeo& g) {
  return 256 + (size_t)g.R * g.slot + (size_t)(7 + 2) * g.Cg * 4 + (size_t)kMaxThreads * 8 + 128 * 8 +
         (size_t)16 * 8 * 16 * 4 + 8 * 16 * 4 + 64;
}

bool choose_stream(const mvfb_mvf_desc* d, StreamGeo& g) {
  if (d->dtype != MVFB_BF16 || d->layout != MVFB_NHWC) return false;
  if (d->Cs % 8 != 0 || d->C % 8 != 0 || d->W + 2 > 256) return false;
  const int splits[4] = {1, 2, 4, 7};
  const int cands[4] = {64, 32, 16, 8};
  for (int si = 0; si < 4; ++si) {
    const int hsplit = splits[si];
    if (d->H % hsplit) continue;
    const int Hs = d->H / hsplit;
    if (Hs + 2 > 256) continue;
    for (int ci = 0; ci < 4; ++ci) {
      const int Cg = cands[ci];
      if (d->Cs % Cg) continue;
      const int items = Hs * d->W * (Cg / 8);
      if (items > kMaxItems) continue;
      g.N = d->N; g.T = d->T; g.Cs = d->Cs; g.H = d->H; g.W = d->W;
      g.Cg = Cg; g.G = Cg / 8; g.ngroups = d->Cs / Cg;
      g.hsplit = hsplit; g.Hs = Hs; g.Hp = Hs + 2; g.Wp = d->W + 2;
      g.slot = (g.Hp * g.Wp * Cg * 2 + 127) / 128 * 128;
      g.items = items;
      g.cwarps = (items + 31) / 32;
      int R = (int)((150 * 1024) / g.slot);
      if (R > kMaxRing) R = kMaxRing;
      if (R < 3) continue;
      g.R = R;
      const int lanes = g.ngroups * hsplit;
      int P = num_sms() / lanes;
      if (P < 1) P = 1;
      if (P > d->N) P = d->N;
      g.P = P;
      if (stream_smem(g) > (size_t)kSmemLimit) continue;
      return true;
    }
  }
  return false;
}

}  // namespace

bool mvf_stream_supported(const mvfb_mvf_desc* d) {
  StreamGeo g;
  return choose_stream(d, g);
}

size_t mvf_stream_ws(const mvfb_mvf_desc* d) {
  StreamGeo g;
  if (!choose_stream(d, g)) return 0;
  return (size_t)g.ngroups * g.hsplit * g.P * 2 * g.Cg * sizeof(float) + 256;
}

int mvf_stream_fwd(const mvfb_mvf_desc* d, const void* x, void* y, long long y_stride, const float* wt,
                   const float* wh, const float* ww, const float* gamma, const float* beta, float* rm, float* rv,
                   float* save_mean, float* save_rstd, void* ws, cudaStream_t st) {
  StreamGeo g;
  if (!choose_stream(d, g)) return MVFB_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) || y_stride % 8 != 0)
    return MVFB_ERR_UNSUPPORTED;
  CUtensorMap tmx;
  const uint64_t dims[4] = {(uint64_t)g.Cs, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N * g.T};
  const uint64_t strides[3] = {(uint64_t)d->C * 2, (uint64_t)g.W * d->C * 2, (uint64_t)g.H * g.W * d->C * 2};
  const uint32_t box[4] = {(uint32_t)g.Cg, (uint32_t)g.Wp, (uint32_t)g.Hp, 1u};
  int rc = encode_tmap(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, nullptr,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (rc) return rc;
  StreamArgs a;
  a.g = g;
  a.use_hs = d->use_hs; a.eps = d->eps; a.momentum = d->momentum;
  a.wt = wt; a.wh = d->mode != MVFB_MODE_T ? wh : nullptr; a.ww = d->mode == MVFB_MODE_THW ? ww : nullptr;
  a.gamma = gamma; a.beta = beta; a.running_mean = rm; a.running_var = rv;
  a.save_mean = save_mean; a.save_rstd = save_rstd;
  a.partials = (float*)ws;
  a.y = (__nv_bfloat16*)y; a.y_pix = y_stride;
  static bool once = false;
  if (!once) {
    MVFB_CUDA(cudaFuncSetAttribute(mvf_stream_fwd_kernel<PASS_APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    MVFB_CUDA(cudaFuncSetAttribute(mvf_stream_fwd_kernel<PASS_STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    MVFB_CUDA(cudaFuncSetAttribute(mvf_stream_fwd_kernel<PASS_TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    once = true;
  }
  const dim3 grid(g.ngroups * g.hsplit * g.P), block(32 * (g.cwarps + 1));
  const size_t smem = stream_smem(g);
  if (d->use_hs && d->training) {
    mvf_stream_fwd_kernel<PASS_STATS><<<grid, block, smem, st>>>(tmx, a);
    count_launch();
    MVFB_LAUNCH_CHECK();
    mvf_stream_fwd_kernel<PASS_TRAIN><<<grid, block, smem, st>>>(tmx, a);
  } else {
    mvf_stream_fwd_kernel<PASS_APPLY><<<grid, block, smem, st>>>(tmx, a);
  }
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

}  // namespace mvfb

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
#include <stdlib.h>

#include "common.cuh"
#include "mvf_internal.cuh"
#include "ptx.cuh"
#include "mvf_stream.cuh"

namespace mvfb {

using namespace stream;

namespace {

constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxRing = 12;

// Consumer threads per CTA (one item each).  V = channels per item.  V = 8: 13 + 1 warps = 4 per SM sub-partition
// (16K registers each) -> at most 128 registers per thread (136 fails to launch: measured).  V = 4: half the
// per-thread state, 25 + 1 warps at <= 72 registers -- twice the warps to hide latency with.
template <int V>
struct Lim {
  static constexpr int kMaxItems = V == 8 ? 416 : 800;
  static constexpr int kMaxThreads = kMaxItems + 32;
};
constexpr int kMaxThreadsAny = 832;
constexpr int kMaxCWarps = 25;

constexpr int PASS_APPLY = 0;   // eval-mode BN (running stats) or no BN at all
constexpr int PASS_STATS = 1;   // train: partial sums only
constexpr int PASS_TRAIN = 2;   // train: batch statistics from the partials, then apply

struct StreamGeo {
  int N, T, Cs, H, W;
  int V;                  // channels per item (8 or 4)
  int Cg, G, ngroups;     // channels per CTA, V-channel vectors per pixel, channel groups
  int hsplit, Hs;         // H tiles and rows per tile
  int Hp, Wp;             // padded tile extents (Hs+2, W+2)
  int slot;               // bytes of one frame slot (multiple of 128)
  int R;                  // ring slots
  int items;              // Hs*W*G  (<= Lim<V>::kMaxItems)
  int cwarps;             // consumer warps
  int P;                  // CTAs sharing one (channel group, H tile): clips are dealt round-robin
};

struct StreamArgs {
  StreamGeo g;
  int use_hs;
  int debug;              // tuning experiments (MVFB_STREAM_DEBUG): 1 = no stores, 2 = no arithmetic / smem reads
  float eps, momentum;
  const float *wt, *wh, *ww, *gamma, *beta;
  float *running_mean, *running_var, *save_mean, *save_rstd;
  float* partials;        // [grid][2*Cg]
  __nv_bfloat16* y;
  long long y_pix;
};

template <int PASS, int V>
__global__ void __launch_bounds__(Lim<V>::kMaxThreads, 1)
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
  double* s_dpart = reinterpret_cast<double*>(s_shift + g.Cg); // [kMaxThreadsAny]
  double* s_dsum = s_dpart + kMaxThreadsAny;                   // [2*Cg]
  float* s_red = reinterpret_cast<float*>(s_dsum + 128);      // [cwarps][2*Cg]

  const uint32_t frame_bytes = (uint32_t)(g.Hp * g.Wp * g.Cg * 2);
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(a.partials) + (size_t)blockIdx.x * 4;
  if ((a.debug & 4) && tid == 0) stamps[0] = gtimer();
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

  // ---- per-channel constants, computed once per CTA.  All global loads are issued before the first use so the
  // prologue costs one memory round trip, not three.
  float bn_m = 0.f, bn_v = 1.f, bn_g = 1.f, bn_b = 0.f;
  if (PASS == PASS_APPLY && a.use_hs && tid < g.Cg) {
    bn_m = a.running_mean[c0 + tid]; bn_v = a.running_var[c0 + tid];
    bn_g = a.gamma[c0 + tid]; bn_b = a.beta[c0 + tid];
  }
  if (PASS == PASS_TRAIN && tid < g.Cg) { bn_g = a.gamma[c0 + tid]; bn_b = a.beta[c0 + tid]; }
  for (int i = tid; i < 3 * g.Cg; i += nthreads) {
    // taps of view i / Cg: (w[0], w[1], w[2]) -> s_coef rows: centre accumulates all views' middle taps later
    const int view = i / g.Cg, ch = i - view * g.Cg, c3 = (c0 + ch) * 3;
    const float* wv = view == 0 ? a.wt : (view == 1 ? a.wh : a.ww);
    float w0 = 0.f, w1 = 0.f, w2 = 0.f;
    if (wv) { w0 = round_bf16(wv[c3]); w1 = round_bf16(wv[c3 + 1]); w2 = round_bf16(wv[c3 + 2]); }
    s_coef[(1 + 2 * view) * g.Cg + ch] = w0;                   // rows 1,3,5: t-1, h-1, w-1
    s_coef[(2 + 2 * view) * g.Cg + ch] = w2;                   // rows 2,4,6: t+1, h+1, w+1
    s_red[view * g.Cg + ch] = w1;                              // middle taps, summed below
  }
  if (PASS == PASS_TRAIN) {
    // Programmatic dependent launch: this grid may start while the statistics grid is still draining (its barrier
    // setup, first TMA loads of x and coefficient loads above overlap that tail); the partial sums are only valid
    // once the primary grid has completed and flushed.
    asm volatile("griddepcontrol.wait;" ::: "memory");
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
      const float sc = bn_g * rstd;
      s_scale[tid] = sc;
      s_shift[tid] = bn_b - mean * sc;
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
        const float mean = bn_m, rstd = 1.f / sqrtf(bn_v + a.eps);
        sc = bn_g * rstd;
        sh = bn_b - mean * sc;
        if (rest == 0 && a.save_mean) { a.save_mean[c] = mean; a.save_rstd[c] = rstd; }
      }
      s_scale[tid] = sc;
      s_shift[tid] = sh;
    }
  }
  __syncthreads();
  if (tid < g.Cg) s_coef[tid] = s_red[tid] + s_red[g.Cg + tid] + s_red[2 * g.Cg + tid];   // centre coefficient
  __syncthreads();

  FV<V> sum = zerov<V>(), sq = zerov<V>();
  const int vec = tid % g.G;
  if (PASS == PASS_STATS) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the apply grid queue up
  if ((a.debug & 4) && tid == 0) stamps[1] = gtimer();

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
    // ===== consumers: one (pixel, V-channel vector) item per thread
    const bool active = tid < g.items && !(a.debug & 2);
    const int pix = tid / g.G;
    const int hl = pix / g.W, w = pix - hl * g.W;             // row within the H tile, column
    const int pixb = g.Cg * 2, rowb = g.Wp * pixb;
    const int off = ((hl + 1) * g.Wp + (w + 1)) * pixb + vec * (2 * V);
    const FV<V> kc = lds_f32<V>(s_coef + vec * V), kt0 = lds_f32<V>(s_coef + g.Cg + vec * V),
                kt2 = lds_f32<V>(s_coef + 2 * g.Cg + vec * V), kh0 = lds_f32<V>(s_coef + 3 * g.Cg + vec * V),
                kh2 = lds_f32<V>(s_coef + 4 * g.Cg + vec * V), kw0 = lds_f32<V>(s_coef + 5 * g.Cg + vec * V),
                kw2 = lds_f32<V>(s_coef + 6 * g.Cg + vec * V);
    FV<V> scale = zerov<V>(), shift = zerov<V>();
    if (PASS != PASS_STATS) { scale = lds_f32<V>(s_scale + vec * V); shift = lds_f32<V>(s_shift + vec * V); }
    const bool hs_on = a.use_hs != 0;
    const size_t frame_elems = (size_t)g.H * g.W * a.y_pix;
    const size_t ypix = ((size_t)(h0 + hl) * g.W + w) * a.y_pix + c0 + vec * V;
    // 32-bit shared-space addresses keep the per-frame bookkeeping to a handful of integer instructions
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
    const uint32_t base = smem_u32(slots) + (uint32_t)off;
    const uint32_t slotb = (uint32_t)g.slot, ringb = (uint32_t)g.R * slotb;

    uint32_t cur = base, fb = full0, ph = 0;                   // current frame: slot address, full barrier, parity
    __nv_bfloat16* yp = a.y;
    // one frame step; (xm, xc, xp) = centre values of frames t-1, t, t+1.  The caller rotates the three register
    // sets instead of moving them (3-way unrolled frame loop).
    auto step = [&](const FV<V>& xm, const FV<V>& xc, FV<V>& xp, int t) {
      uint32_t nxt = cur + slotb, fb1 = fb + 8, ph1 = ph;
      if (nxt == base + ringb) { nxt = base; fb1 = full0; ph1 ^= 1; }
      xp = zerov<V>();
      if (t + 1 < g.T) {
        wait_u32(fb1, ph1);
        if (active) xp = lds_bf16<V>(nxt);
      }
      if (active) {
        FV<V> z;
#pragma unroll
        for (int j = 0; j < V / 2; ++j) z.p[j] = __fmul2_rn(kc.p[j], xc.p[j]);
        fmav<V>(z, kt0, xm);
        fmav<V>(z, kt2, xp);
        fmav<V>(z, kh0, lds_bf16<V>(cur - rowb));
        fmav<V>(z, kh2, lds_bf16<V>(cur + rowb));
        fmav<V>(z, kw0, lds_bf16<V>(cur - pixb));
        fmav<V>(z, kw2, lds_bf16<V>(cur + pixb));
        if (PASS == PASS_STATS) {
#pragma unroll
          for (int j = 0; j < V / 2; ++j) {
            sum.p[j] = __fadd2_rn(sum.p[j], z.p[j]);
            sq.p[j] = __ffma2_rn(z.p[j], z.p[j], sq.p[j]);
          }
        } else {
          if (hs_on) {
#pragma unroll
            for (int j = 0; j < V / 2; ++j) {
              const float2 u = __ffma2_rn(z.p[j], scale.p[j], shift.p[j]);
              float2 sg;
              sg.x = __saturatef(fmaf(u.x, 1.f / 6.f, 0.5f));
              sg.y = __saturatef(fmaf(u.y, 1.f / 6.f, 0.5f));
              z.p[j] = __fmul2_rn(u, sg);
            }
          }
          if (!(a.debug & 1)) store_bf16<V>(yp, z);
        }
      }
      yp += frame_elems;
      __syncwarp();
      if (lane == 0) arrive_u32(empty0 + (fb - full0));         // this warp is done with frame t's slot
      cur = nxt;
      fb = fb1;
      ph = ph1;
    };
    for (int kclip = 0; kclip < nclips; ++kclip) {
      const int n = p + kclip * g.P;
      yp = a.y + (size_t)n * g.T * frame_elems + ypix;
      FV<V> ra = zerov<V>(), rb = zerov<V>(), rc;
      wait_u32(fb, ph);
      if ((a.debug & 4) && tid == 0 && kclip == 0) stamps[2] = gtimer();
      if (active) rb = lds_bf16<V>(cur);
#pragma unroll 1
      for (int t = 0; t < g.T; t += 3) {
        step(ra, rb, rc, t);
        if (t + 1 < g.T) step(rb, rc, ra, t + 1);
        if (t + 2 < g.T) step(rc, ra, rb, t + 2);
      }
    }
  }

  if ((a.debug & 4) && tid == 0) stamps[3] = gtimer();
  if (PASS == PASS_STATS) {
    // CTA reduction of (sum, sumsq) by channel vector: lanes with equal lane % G own the same channels
    float acc[2 * V];
#pragma unroll
    for (int j = 0; j < V / 2; ++j) {
      acc[2 * j] = sum.p[j].x; acc[2 * j + 1] = sum.p[j].y;
      acc[V + 2 * j] = sq.p[j].x; acc[V + 2 * j + 1] = sq.p[j].y;
    }
#pragma unroll
    for (int q = 0; q < 2 * V; ++q) {
      float v = acc[q];
      for (int o = 16; o >= g.G; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[q] = v;
    }
    if (!producer && lane < g.G) {
#pragma unroll
      for (int q = 0; q < 2 * V; ++q) s_red[(warp * g.G + lane) * (2 * V) + q] = acc[q];
    }
    __syncthreads();
    for (int i = tid; i < 2 * g.Cg; i += nthreads) {
      const int ch = i >> 1, kind = i & 1;
      const int idx = (ch / V) * (2 * V) + kind * V + (ch % V);
      float v = 0.f;
      for (int wv = 0; wv < g.cwarps; ++wv) v += s_red[wv * 2 * g.Cg + idx];
      a.partials[(size_t)blockIdx.x * 2 * g.Cg + i] = v;
    }
  }
}

size_t stream_smem(const StreamGeo& g) {
  return 256 + (size_t)g.R * g.slot + (size_t)(7 + 2) * g.Cg * 4 + (size_t)kMaxThreadsAny * 8 + 128 * 8 +
         (size_t)kMaxCWarps * 2 * g.Cg * 4 + 64;
}

bool choose_stream(const mvfb_mvf_desc* d, StreamGeo& g) {
  if (d->dtype != MVFB_BF16 || d->layout != MVFB_NHWC) return false;
  if (d->Cs % 8 != 0 || d->C % 8 != 0 || d->W + 2 > 256) return false;
  const int V = 8;                      // 8-channel items (4-channel items at twice the warps measured the same)
  const int max_items = V == 8 ? Lim<8>::kMaxItems : Lim<4>::kMaxItems;
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
      const int G = Cg / V;
      if (G > 16) continue;                                  // lanes sharing a vector must tile a warp
      const int items = Hs * d->W * G;
      if (items > max_items) continue;
      g.N = d->N; g.T = d->T; g.Cs = d->Cs; g.H = d->H; g.W = d->W;
      g.V = V; g.Cg = Cg; g.G = G; g.ngroups = d->Cs / Cg;
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

template <int V>
int launch_stream(const mvfb_mvf_desc* d, const CUtensorMap& tmx, const StreamArgs& a, cudaStream_t st) {
  static DevOnce once;
  if (once.pending()) {
    MVFB_CUDA(cudaFuncSetAttribute(mvf_stream_fwd_kernel<PASS_APPLY, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    MVFB_CUDA(cudaFuncSetAttribute(mvf_stream_fwd_kernel<PASS_STATS, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    MVFB_CUDA(cudaFuncSetAttribute(mvf_stream_fwd_kernel<PASS_TRAIN, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    once.done();
  }
  const StreamGeo& g = a.g;
  const dim3 grid(g.ngroups * g.hsplit * g.P), block(32 * (g.cwarps + 1));
  const size_t smem = stream_smem(g);
  if (d->use_hs && d->training) {
    mvf_stream_fwd_kernel<PASS_STATS, V><<<grid, block, smem, st>>>(tmx, a);
    count_launch();
    MVFB_LAUNCH_CHECK();
    // second pass as a programmatic dependent launch (PDL): no host-visible gap between the two grids
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    MVFB_CUDA(cudaLaunchKernelEx(&cfg, mvf_stream_fwd_kernel<PASS_TRAIN, V>, tmx, a));
  } else {
    mvf_stream_fwd_kernel<PASS_APPLY, V><<<grid, block, smem, st>>>(tmx, a);
  }
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
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
  a.debug = 0;
  a.wt = wt; a.wh = d->mode != MVFB_MODE_T ? wh : nullptr; a.ww = d->mode == MVFB_MODE_THW ? ww : nullptr;
  a.gamma = gamma; a.beta = beta; a.running_mean = rm; a.running_var = rv;
  a.save_mean = save_mean; a.save_rstd = save_rstd;
  a.partials = (float*)ws;
  a.y = (__nv_bfloat16*)y; a.y_pix = y_stride;
  return g.V == 8 ? launch_stream<8>(d, tmx, a, st) : launch_stream<4>(d, tmx, a, st);
}

}  // namespace mvfb

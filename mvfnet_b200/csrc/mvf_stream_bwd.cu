// MVF backward, persistent warp-specialised frame-stream kernels (bf16 NHWC) -- the backward counterpart of
// mvf_stream.cu (same ring / producer-warp / rolling-register design; see that file's header).
//
//   kernel A  mvf_stream_bwd_reduce  : streams x (padded frames) and g = dL/dy (unpadded frames) together, recomputes
//             z = stencil(x), u = z*scale + shift, du = g * hardswish'(u) and accumulates per-thread
//             (sum du, sum du*z) over all of the CTA's clips; one partial row per CTA.  Also zeroes dwt/dwh/dww.
//   kernel B  mvf_stream_bwd_dx      : reduces the partial rows of its channel group (dgamma, dbeta and the two
//             BatchNorm-backward correction terms), then per frame
//                 dz(t)   = scale*du + u*c1 + c0                      (BN + hardswish backward, in registers)
//                 a_q    += dz(t)[p] * x[p + q]                       (7 tap-gradient sums, x neighbours already staged)
//                 dz(t) -> double-buffered fp32 frame in smem (zero halo), one consumer barrier
//                 P(t)    = kc dz(t) + kt2 dz(t-1) + kh0 dz(t)[h+1] + kh2 dz(t)[h-1] + kw0 dz(t)[w+1] + kw2 dz(t)[w-1]
//                 dx(t-1) = P(t-1) + kt0 dz(t)                        (transposed stencil; the temporal terms roll
//                                                                      through registers exactly like the forward)
//             dx may alias g: a frame of g is in shared memory long before that frame's dx is written.
// Needs the whole H x W frame in one CTA (the dz halo exchange), i.e. no H tiling: 28x28 / 32x32 slabs stay on the
// ring kernels of mvf_fast.cu.
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
constexpr int kMaxRing = 10;
constexpr int kMaxItems = 416;                 // 13 consumer warps + 1 producer warp, <= 128 registers per thread
constexpr int kMaxThreads = kMaxItems + 32;
constexpr int VB = 4;                          // channels per item in kernel B (register budget: 7 tap accumulators)

struct BGeo {
  int N, T, Cs, H, W;
  int Cg, ngroups;        // channels per CTA, channel groups
  int Hp, Wp;             // padded frame extents
  int slot_x, slot_g;     // bytes: padded bf16 x frame, unpadded bf16 g frame
  int stage;              // slot_x + slot_g (multiple of 128)
  int slot_dz;            // padded fp32 dz frame
  int R;
  int P;                  // CTAs per channel group (clips dealt round-robin)
};

struct BArgs {
  BGeo g;
  int use_hs, training, share_h, share_w;
  const float *wt, *wh, *ww, *gamma, *beta, *mean, *rstd;
  float* partials;        // [PA][Cs][2]: (sum du, sum du*z) per clip lane of kernel A
  int PA;                 // clip lanes of kernel A (rows of `partials`)
  float *dwt, *dwh, *dww, *dgamma, *dbeta;
  __nv_bfloat16* dx;
  long long dx_pix;
};

__device__ __forceinline__ float hswish_grad_f(float u) {
  return __saturatef(fmaf(u, 1.f / 6.f, 0.5f)) + ((u > -3.f && u < 3.f) ? u * (1.f / 6.f) : 0.f);
}

struct Carve {
  uint64_t *full, *empty;
  uint8_t* stages;
  uint8_t* dz;            // 2 padded fp32 frames (kernel B only)
  float *s_coef, *s_const, *s_red;
  double *s_dpart, *s_dsum;
};

__device__ __forceinline__ Carve carve(uint8_t* smem, const BGeo& g, bool with_dz) {
  Carve c;
  c.full = reinterpret_cast<uint64_t*>(smem);
  c.empty = c.full + 16;
  c.stages = smem + 256;
  c.dz = c.stages + (size_t)g.R * g.stage;
  uint8_t* rest = c.dz + (with_dz ? 2 * (size_t)g.slot_dz : 0);
  c.s_coef = reinterpret_cast<float*>(rest);                  // [7][Cg]
  c.s_const = c.s_coef + 7 * g.Cg;                            // [4][Cg]: scale, shift, c1, c0
  c.s_dpart = reinterpret_cast<double*>(c.s_const + 4 * g.Cg + (g.Cg & 1 ? 1 : 0));   // [kMaxThreads]
  c.s_dsum = c.s_dpart + kMaxThreads;                         // [2*Cg]
  c.s_red = reinterpret_cast<float*>(c.s_dsum + 128);        // [cwarps][G][K]
  return c;
}

size_t smem_bytes(const BGeo& g, bool with_dz) {
  return 256 + (size_t)g.R * g.stage + (with_dz ? 2 * (size_t)g.slot_dz : 0) + (size_t)12 * g.Cg * 4 +
         (size_t)kMaxThreads * 8 + 128 * 8 + (size_t)13 * 7 * g.Cg * 4 + 256;
}

// stencil coefficient table [7][Cg]: centre, t-1, t+1, h-1, h+1, w-1, w+1  (all threads; caller syncs)
__device__ __forceinline__ void load_coef(float* s_coef, float* scratch, const BArgs& a, int c0, int nthreads) {
  const BGeo& g = a.g;
  for (int i = threadIdx.x; i < 3 * g.Cg; i += nthreads) {
    const int view = i / g.Cg, ch = i - view * g.Cg, c3 = (c0 + ch) * 3;
    const float* wv = view == 0 ? a.wt : (view == 1 ? a.wh : a.ww);
    float w0 = 0.f, w1 = 0.f, w2 = 0.f;
    if (wv) { w0 = round_bf16(wv[c3]); w1 = round_bf16(wv[c3 + 1]); w2 = round_bf16(wv[c3 + 2]); }
    s_coef[(1 + 2 * view) * g.Cg + ch] = w0;
    s_coef[(2 + 2 * view) * g.Cg + ch] = w2;
    scratch[view * g.Cg + ch] = w1;
  }
  __syncthreads();
  if (threadIdx.x < g.Cg) s_coef[threadIdx.x] = scratch[threadIdx.x] + scratch[g.Cg + threadIdx.x] + scratch[2 * g.Cg + threadIdx.x];
}

__device__ __forceinline__ void issue_stage(const CUtensorMap* tmx, const CUtensorMap* tmg, const Carve& c, const BGeo& g,
                                            int s, int c0, int frame) {
  mbar_arrive_expect_tx(&c.full[s], (uint32_t)(g.Hp * g.Wp * g.Cg * 2 + g.H * g.W * g.Cg * 2));
  uint8_t* dst = c.stages + (size_t)s * g.stage;
  tma_load_4d(dst, tmx, &c.full[s], c0, -1, -1, frame);
  tma_load_4d(dst + g.slot_x, tmg, &c.full[s], c0, 0, 0, frame);
}

// ------------------------------------------------------------------------------------------ kernel A
__global__ void __launch_bounds__(kMaxThreads, 1)
mvf_stream_bwd_reduce(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmg, const BArgs a) {
  constexpr int V = 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  const BGeo& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
  const int G = g.Cg / V, items = g.H * g.W * G, cwarps = (items + 31) / 32;
  const int cg = blockIdx.x % g.ngroups, p = blockIdx.x / g.ngroups;
  const int c0 = cg * g.Cg;
  const int nclips = p < g.N ? (g.N - p + g.P - 1) / g.P : 0;
  const int Q = nclips * g.T;
  Carve c = carve(smem, g, false);

  if (tid == 0) {
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmg);
    for (int s = 0; s < g.R; ++s) { mbar_init(&c.full[s], 1); mbar_init(&c.empty[s], cwarps); }
    fence_barrier_init();
  }
  __syncthreads();
  const bool producer = warp == cwarps;
  if (producer && lane == 0) {
    const int pre = Q < g.R ? Q : g.R;
    int n = p, t = 0;
    for (int q = 0; q < pre; ++q) {
      issue_stage(&tmx, &tmg, c, g, q, c0, n * g.T + t);
      if (++t == g.T) { t = 0; n += g.P; }
    }
  }
  if (p == 0) {                                     // zero the tap-gradient outputs kernel B accumulates into
    for (int i = tid; i < g.Cg * 3; i += nthreads) {
      a.dwt[c0 * 3 + i] = 0.f;
      if (a.dwh) a.dwh[c0 * 3 + i] = 0.f;
      if (a.dww) a.dww[c0 * 3 + i] = 0.f;
    }
  }
  float bn_sc = 1.f, bn_sh = 0.f;
  if (tid < g.Cg) {
    const int ch = c0 + tid;
    bn_sc = a.gamma[ch] * a.rstd[ch];
    bn_sh = a.beta[ch] - a.mean[ch] * bn_sc;
  }
  load_coef(c.s_coef, c.s_red, a, c0, nthreads);
  if (tid < g.Cg) { c.s_const[tid] = bn_sc; c.s_const[g.Cg + tid] = bn_sh; }
  __syncthreads();

  FV<V> sum = zerov<V>(), sq = zerov<V>();
  const int vec = tid % G;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // the dx kernel may queue up behind this grid
  if (producer) {
    if (lane == 0) {
      int s = 0, use = 1, n = p, t = 0;
      for (int q = 0; q < g.R && q < Q; ++q) { if (++t == g.T) { t = 0; n += g.P; } }
      for (int q = g.R; q < Q; ++q) {
        mbar_wait(&c.empty[s], (use - 1) & 1);
        issue_stage(&tmx, &tmg, c, g, s, c0, n * g.T + t);
        if (++t == g.T) { t = 0; n += g.P; }
        if (++s == g.R) { s = 0; ++use; }
      }
    }
  } else {
    const bool active = tid < items;
    const int pix = tid / G, hl = pix / g.W, w = pix - hl * g.W;
    const int pixb = g.Cg * 2, rowb = g.Wp * pixb;
    const uint32_t off = (uint32_t)(((hl + 1) * g.Wp + (w + 1)) * pixb + vec * (2 * V));
    const uint32_t goff = (uint32_t)(g.slot_x + (hl * g.W + w) * pixb + vec * (2 * V));
    const FV<V> kc = lds_f32<V>(c.s_coef + vec * V), kt0 = lds_f32<V>(c.s_coef + g.Cg + vec * V),
                kt2 = lds_f32<V>(c.s_coef + 2 * g.Cg + vec * V), kh0 = lds_f32<V>(c.s_coef + 3 * g.Cg + vec * V),
                kh2 = lds_f32<V>(c.s_coef + 4 * g.Cg + vec * V), kw0 = lds_f32<V>(c.s_coef + 5 * g.Cg + vec * V),
                kw2 = lds_f32<V>(c.s_coef + 6 * g.Cg + vec * V);
    const FV<V> scale = lds_f32<V>(c.s_const + vec * V), shift = lds_f32<V>(c.s_const + g.Cg + vec * V);
    const uint32_t full0 = smem_u32(c.full), empty0 = smem_u32(c.empty);
    const uint32_t base = smem_u32(c.stages), stageb = (uint32_t)g.stage, ringb = (uint32_t)g.R * stageb;
    const bool hs_on = a.use_hs != 0;
    uint32_t cur = base, fb = full0, ph = 0;
    auto step = [&](const FV<V>& xm, const FV<V>& xc, FV<V>& xp, int t) {
      uint32_t nxt = cur + stageb, fb1 = fb + 8, ph1 = ph;
      if (nxt == base + ringb) { nxt = base; fb1 = full0; ph1 ^= 1; }
      xp = zerov<V>();
      if (t + 1 < g.T) {
        wait_u32(fb1, ph1);
        if (active) xp = lds_bf16<V>(nxt + off);
      }
      if (active) {
        FV<V> z;
#pragma unroll
        for (int j = 0; j < V / 2; ++j) z.p[j] = __fmul2_rn(kc.p[j], xc.p[j]);
        fmav<V>(z, kt0, xm);
        fmav<V>(z, kt2, xp);
        fmav<V>(z, kh0, lds_bf16<V>(cur + off - rowb));
        fmav<V>(z, kh2, lds_bf16<V>(cur + off + rowb));
        fmav<V>(z, kw0, lds_bf16<V>(cur + off - pixb));
        fmav<V>(z, kw2, lds_bf16<V>(cur + off + pixb));
        const FV<V> gv = lds_bf16<V>(cur + goff);
#pragma unroll
        for (int j = 0; j < V / 2; ++j) {
          float2 du = gv.p[j];
          if (hs_on) {
            const float2 u = __ffma2_rn(z.p[j], scale.p[j], shift.p[j]);
            du.x *= hswish_grad_f(u.x);
            du.y *= hswish_grad_f(u.y);
          }
          sum.p[j] = __fadd2_rn(sum.p[j], du);
          sq.p[j] = __ffma2_rn(du, z.p[j], sq.p[j]);
        }
      }
      __syncwarp();
      if (lane == 0) arrive_u32(empty0 + (fb - full0));
      cur = nxt; fb = fb1; ph = ph1;
    };
    for (int kclip = 0; kclip < nclips; ++kclip) {
      FV<V> ra = zerov<V>(), rb = zerov<V>(), rc;
      wait_u32(fb, ph);
      if (active) rb = lds_bf16<V>(cur + off);
#pragma unroll 1
      for (int t = 0; t < g.T; t += 3) {
        step(ra, rb, rc, t);
        if (t + 1 < g.T) step(rb, rc, ra, t + 1);
        if (t + 2 < g.T) step(rc, ra, rb, t + 2);
      }
    }
  }
  // CTA reduction by channel vector -> one partial row per CTA
  float acc[2 * V];
#pragma unroll
  for (int j = 0; j < V / 2; ++j) {
    acc[2 * j] = sum.p[j].x; acc[2 * j + 1] = sum.p[j].y;
    acc[V + 2 * j] = sq.p[j].x; acc[V + 2 * j + 1] = sq.p[j].y;
  }
#pragma unroll
  for (int q = 0; q < 2 * V; ++q) {
    float v = acc[q];
    for (int o = 16; o >= G; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    acc[q] = v;
  }
  __syncthreads();                                  // s_red doubled as load_coef scratch
  if (!producer && lane < G) {
#pragma unroll
    for (int q = 0; q < 2 * V; ++q) c.s_red[(warp * G + lane) * (2 * V) + q] = acc[q];
  }
  __syncthreads();
  for (int i = tid; i < 2 * g.Cg; i += nthreads) {
    const int ch = i >> 1, kind = i & 1;
    const int idx = (ch / V) * (2 * V) + kind * V + (ch % V);
    float v = 0.f;
    for (int wv = 0; wv < cwarps; ++wv) v += c.s_red[wv * 2 * g.Cg + idx];
    a.partials[((size_t)p * g.Cs + c0) * 2 + i] = v;
  }
}

// ------------------------------------------------------------------------------------------ kernel B
__global__ void __launch_bounds__(kMaxThreads, 1)
mvf_stream_bwd_dx(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmg, const BArgs a) {
  constexpr int V = VB;
  extern __shared__ __align__(1024) uint8_t smem[];
  const BGeo& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
  const int G = g.Cg / V, items = g.H * g.W * G, cwarps = (items + 31) / 32;
  const int cg = blockIdx.x % g.ngroups, p = blockIdx.x / g.ngroups;
  const int c0 = cg * g.Cg;
  const int nclips = p < g.N ? (g.N - p + g.P - 1) / g.P : 0;
  const int Q = nclips * g.T;
  Carve c = carve(smem, g, true);

  if (tid == 0) {
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmg);
    for (int s = 0; s < g.R; ++s) { mbar_init(&c.full[s], 1); mbar_init(&c.empty[s], cwarps); }
    fence_barrier_init();
  }
  for (int i = tid * 16; i < 2 * g.slot_dz; i += nthreads * 16) *reinterpret_cast<uint4*>(c.dz + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const bool producer = warp == cwarps;
  if (producer && lane == 0) {
    const int pre = Q < g.R ? Q : g.R;
    int n = p, t = 0;
    for (int q = 0; q < pre; ++q) {
      issue_stage(&tmx, &tmg, c, g, q, c0, n * g.T + t);
      if (++t == g.T) { t = 0; n += g.P; }
    }
  }
  // ---- per-channel constants:  u = z*scale + shift ;  dz = scale*du + u*c1 + c0   (c1 = c0 = 0 in eval mode)
  float bn_g = 1.f, bn_b = 0.f, bn_m = 0.f, bn_r = 1.f;
  if (a.use_hs && tid < g.Cg) {
    bn_g = a.gamma[c0 + tid]; bn_b = a.beta[c0 + tid]; bn_m = a.mean[c0 + tid]; bn_r = a.rstd[c0 + tid];
  }
  load_coef(c.s_coef, c.s_red, a, c0, nthreads);
  if (a.use_hs) {
    // launched as a programmatic dependent of the reduce kernel: everything above overlapped its tail; its partial
    // rows, its zeroing of dw*, and the last reads of g (which dx may alias) are complete after this wait
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int per = 2 * g.Cg, rows = a.PA;
    const int parts = nthreads / per;
    const int k = tid % per, part = tid / per;
    if (part < parts) {
      double acc = 0.0;
      for (int r = part; r < rows; r += parts) acc += (double)a.partials[((size_t)r * g.Cs + c0) * 2 + k];
      c.s_dpart[part * per + k] = acc;
    }
    __syncthreads();
    if (tid < per) {
      double acc = 0.0;
      for (int q = 0; q < parts; ++q) acc += c.s_dpart[q * per + tid];
      c.s_dsum[tid] = acc;
    }
    __syncthreads();
    if (tid < g.Cg) {
      const double s1 = c.s_dsum[tid * 2], s2 = c.s_dsum[tid * 2 + 1];
      const double dgamma = (double)bn_r * (s2 - (double)bn_m * s1);
      const double m = (double)g.N * g.T * g.H * g.W;
      const double sc = (double)bn_g * bn_r;
      double c1 = 0.0, cc0 = 0.0;
      if (a.training) {
        c1 = -(double)bn_r * dgamma / m;
        cc0 = -sc * s1 / m + (double)bn_b * bn_r * dgamma / m;
      }
      c.s_const[tid] = (float)sc;
      c.s_const[g.Cg + tid] = (float)((double)bn_b - (double)bn_m * sc);
      c.s_const[2 * g.Cg + tid] = (float)c1;
      c.s_const[3 * g.Cg + tid] = (float)cc0;
      if (p == 0) { a.dgamma[c0 + tid] = (float)dgamma; a.dbeta[c0 + tid] = (float)s1; }
    }
  } else if (tid < g.Cg) {
    c.s_const[tid] = 1.f; c.s_const[g.Cg + tid] = 0.f; c.s_const[2 * g.Cg + tid] = 0.f; c.s_const[3 * g.Cg + tid] = 0.f;
  }
  __syncthreads();

  FV<V> acc[7];                                     // tap sums: centre, t-1, t+1, h-1, h+1, w-1, w+1
#pragma unroll
  for (int q = 0; q < 7; ++q) acc[q] = zerov<V>();
  const int vec = tid % G;

  if (producer) {
    if (lane == 0) {
      int s = 0, use = 1, n = p, t = 0;
      for (int q = 0; q < g.R && q < Q; ++q) { if (++t == g.T) { t = 0; n += g.P; } }
      for (int q = g.R; q < Q; ++q) {
        mbar_wait(&c.empty[s], (use - 1) & 1);
        issue_stage(&tmx, &tmg, c, g, s, c0, n * g.T + t);
        if (++t == g.T) { t = 0; n += g.P; }
        if (++s == g.R) { s = 0; ++use; }
      }
    }
  } else {
    const bool active = tid < items;
    const int pix = tid / G, hl = pix / g.W, w = pix - hl * g.W;
    const int pixb = g.Cg * 2, rowb = g.Wp * pixb;            // bf16 frames
    const int dpixb = g.Cg * 4, drowb = g.Wp * dpixb;          // fp32 dz frames
    const uint32_t off = (uint32_t)(((hl + 1) * g.Wp + (w + 1)) * pixb + vec * (2 * V));
    const uint32_t goff = (uint32_t)(g.slot_x + (hl * g.W + w) * pixb + vec * (2 * V));
    const uint32_t doff = (uint32_t)(((hl + 1) * g.Wp + (w + 1)) * dpixb + vec * (4 * V));
    const FV<V> kc = lds_f32<V>(c.s_coef + vec * V), kt0 = lds_f32<V>(c.s_coef + g.Cg + vec * V),
                kt2 = lds_f32<V>(c.s_coef + 2 * g.Cg + vec * V), kh0 = lds_f32<V>(c.s_coef + 3 * g.Cg + vec * V),
                kh2 = lds_f32<V>(c.s_coef + 4 * g.Cg + vec * V), kw0 = lds_f32<V>(c.s_coef + 5 * g.Cg + vec * V),
                kw2 = lds_f32<V>(c.s_coef + 6 * g.Cg + vec * V);
    const FV<V> scale = lds_f32<V>(c.s_const + vec * V), shift = lds_f32<V>(c.s_const + g.Cg + vec * V),
                c1 = lds_f32<V>(c.s_const + 2 * g.Cg + vec * V), cc0 = lds_f32<V>(c.s_const + 3 * g.Cg + vec * V);
    const uint32_t full0 = smem_u32(c.full), empty0 = smem_u32(c.empty);
    const uint32_t base = smem_u32(c.stages), stageb = (uint32_t)g.stage, ringb = (uint32_t)g.R * stageb;
    const uint32_t dz0 = smem_u32(c.dz) + doff, dzb = (uint32_t)g.slot_dz;
    const bool hs_on = a.use_hs != 0;
    const int nconsumer = 32 * cwarps;
    const size_t frame_elems = (size_t)g.H * g.W * a.dx_pix;
    const size_t pix_elems = ((size_t)hl * g.W + w) * a.dx_pix + c0 + vec * V;
    uint32_t cur = base, fb = full0, ph = 0, par = 0;            // par: which dz buffer this frame writes
    __nv_bfloat16* dxp = a.dx;
    FV<V> pprev = zerov<V>(), dzprev = zerov<V>();

    auto lds_dz = [&](uint32_t addr) {
      FV<V> r;
      float4 q;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(addr));
      r.p[0] = make_float2(q.x, q.y); r.p[1] = make_float2(q.z, q.w);
      return r;
    };
    auto step = [&](const FV<V>& xm, const FV<V>& xc, FV<V>& xp, int t) {
      uint32_t nxt = cur + stageb, fb1 = fb + 8, ph1 = ph;
      if (nxt == base + ringb) { nxt = base; fb1 = full0; ph1 ^= 1; }
      xp = zerov<V>();
      if (t + 1 < g.T) {
        wait_u32(fb1, ph1);
        if (active) xp = lds_bf16<V>(nxt + off);
      }
      FV<V> dz = zerov<V>();
      const uint32_t dzw = dz0 + par * dzb;
      if (active) {
        const FV<V> xhm = lds_bf16<V>(cur + off - rowb), xhp = lds_bf16<V>(cur + off + rowb),
                    xwm = lds_bf16<V>(cur + off - pixb), xwp = lds_bf16<V>(cur + off + pixb);
        dz = lds_bf16<V>(cur + goff);                            // g for now
        if (hs_on) {
          FV<V> z;
#pragma unroll
          for (int j = 0; j < V / 2; ++j) z.p[j] = __fmul2_rn(kc.p[j], xc.p[j]);
          fmav<V>(z, kt0, xm);
          fmav<V>(z, kt2, xp);
          fmav<V>(z, kh0, xhm);
          fmav<V>(z, kh2, xhp);
          fmav<V>(z, kw0, xwm);
          fmav<V>(z, kw2, xwp);
#pragma unroll
          for (int j = 0; j < V / 2; ++j) {
            const float2 u = __ffma2_rn(z.p[j], scale.p[j], shift.p[j]);
            float2 du;
            du.x = dz.p[j].x * hswish_grad_f(u.x);
            du.y = dz.p[j].y * hswish_grad_f(u.y);
            dz.p[j] = __ffma2_rn(scale.p[j], du, __ffma2_rn(u, c1.p[j], cc0.p[j]));
          }
        }
        fmav<V>(acc[0], dz, xc);
        fmav<V>(acc[1], dz, xm);
        fmav<V>(acc[2], dz, xp);
        fmav<V>(acc[3], dz, xhm);
        fmav<V>(acc[4], dz, xhp);
        fmav<V>(acc[5], dz, xwm);
        fmav<V>(acc[6], dz, xwp);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dzw), "f"(dz.p[0].x), "f"(dz.p[0].y), "f"(dz.p[1].x),
                     "f"(dz.p[1].y) : "memory");
      }
      __syncwarp();
      if (lane == 0) arrive_u32(empty0 + (fb - full0));          // x(t) neighbours and g(t) are consumed
      consumer_bar_sync(nconsumer);                              // dz(t) of every pixel is in shared memory
      if (active) {
        // dx(t-1) = P(t-1) + kt0 dz(t)
        if (t > 0) {
          FV<V> o = pprev;
          fmav<V>(o, kt0, dz);
          store_bf16<V>(dxp, o);
        }
        FV<V> pn;
#pragma unroll
        for (int j = 0; j < V / 2; ++j) pn.p[j] = __fmul2_rn(kc.p[j], dz.p[j]);
        fmav<V>(pn, kt2, dzprev);
        fmav<V>(pn, kh0, lds_dz(dzw + drowb));
        fmav<V>(pn, kh2, lds_dz(dzw - drowb));
        fmav<V>(pn, kw0, lds_dz(dzw + dpixb));
        fmav<V>(pn, kw2, lds_dz(dzw - dpixb));
        pprev = pn;
        dzprev = dz;
      }
      if (t > 0) dxp += frame_elems;
      par ^= 1;
      cur = nxt; fb = fb1; ph = ph1;
    };
    for (int kclip = 0; kclip < nclips; ++kclip) {
      const int n = p + kclip * g.P;
      dxp = a.dx + (size_t)n * g.T * frame_elems + pix_elems;    // frame 0 of this clip
      pprev = zerov<V>();
      dzprev = zerov<V>();
      FV<V> ra = zerov<V>(), rb = zerov<V>(), rc;
      wait_u32(fb, ph);
      if (active) rb = lds_bf16<V>(cur + off);
#pragma unroll 1
      for (int t = 0; t < g.T; t += 3) {
        step(ra, rb, rc, t);
        if (t + 1 < g.T) step(rb, rc, ra, t + 1);
        if (t + 2 < g.T) step(rc, ra, rb, t + 2);
      }
      if (active) store_bf16<V>(dxp, pprev);                     // dx(T-1): no frame T
    }
  }

  // ---- tap gradients: reduce over the CTA, then one atomic per (channel, tap)
  float flat[7 * V];
#pragma unroll
  for (int q = 0; q < 7; ++q)
#pragma unroll
    for (int j = 0; j < V / 2; ++j) { flat[q * V + 2 * j] = acc[q].p[j].x; flat[q * V + 2 * j + 1] = acc[q].p[j].y; }
#pragma unroll
  for (int q = 0; q < 7 * V; ++q) {
    float v = flat[q];
    for (int o = 16; o >= G; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    flat[q] = v;
  }
  __syncthreads();
  if (!producer && lane < G) {
#pragma unroll
    for (int q = 0; q < 7 * V; ++q) c.s_red[(warp * G + lane) * (7 * V) + q] = flat[q];
  }
  __syncthreads();
  float* dst_h = a.share_h ? a.dwt : a.dwh;
  float* dst_w = a.share_w ? a.dwt : a.dww;
  for (int i = tid; i < g.Cg * 7; i += nthreads) {
    const int ch = i / 7, q = i - ch * 7;
    const int idx = (ch / V) * (7 * V) + q * V + (ch % V);
    float val = 0.f;
    for (int wv = 0; wv < cwarps; ++wv) val += c.s_red[wv * 7 * g.Cg + idx];
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

// geometry for items of V channels (kernel A: V = 8, kernel B: V = 4); the two kernels tile the channels independently
bool choose_bwd(const mvfb_mvf_desc* d, int V, bool with_dz, BGeo& g) {
  if (d->dtype != MVFB_BF16 || d->layout != MVFB_NHWC) return false;
  if (d->Cs % 8 != 0 || d->C % 8 != 0 || d->W + 2 > 256 || d->H + 2 > 256) return false;
  const int cands[4] = {64, 32, 16, 8};
  for (int ci = 0; ci < 4; ++ci) {
    const int Cg = cands[ci];
    if (d->Cs % Cg) continue;
    const int GB = Cg / V;
    if (GB > 16 || GB < 1) continue;
    if (d->H * d->W * GB > kMaxItems) continue;
    g.N = d->N; g.T = d->T; g.Cs = d->Cs; g.H = d->H; g.W = d->W;
    g.Cg = Cg; g.ngroups = d->Cs / Cg;
    g.Hp = d->H + 2; g.Wp = d->W + 2;
    g.slot_x = (g.Hp * g.Wp * Cg * 2 + 127) / 128 * 128;
    g.slot_g = (d->H * d->W * Cg * 2 + 127) / 128 * 128;
    g.stage = g.slot_x + g.slot_g;
    g.slot_dz = (g.Hp * g.Wp * Cg * 4 + 127) / 128 * 128;
    int R = (int)((120 * 1024) / g.stage);
    if (R > kMaxRing) R = kMaxRing;
    if (R < 3) continue;
    g.R = R;
    int P = num_sms() / g.ngroups;
    if (P < 1) P = 1;
    if (P > d->N) P = d->N;
    g.P = P;
    if (smem_bytes(g, with_dz) > (size_t)kSmemLimit) continue;
    return true;
  }
  return false;
}

int make_map(CUtensorMap* tm, const void* base, long long pix_stride, const BGeo& g, bool padded) {
  const uint64_t dims[4] = {(uint64_t)g.Cs, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N * g.T};
  const uint64_t strides[3] = {(uint64_t)pix_stride * 2, (uint64_t)g.W * pix_stride * 2, (uint64_t)g.H * g.W * pix_stride * 2};
  const uint32_t box[4] = {(uint32_t)g.Cg, (uint32_t)(padded ? g.Wp : g.W), (uint32_t)(padded ? g.Hp : g.H), 1u};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, nullptr,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

}  // namespace

bool mvf_stream_bwd_supported(const mvfb_mvf_desc* d) {
  BGeo ga, gb;
  return choose_bwd(d, 8, false, ga) && choose_bwd(d, VB, true, gb);
}

size_t mvf_stream_bwd_ws(const mvfb_mvf_desc* d) {
  BGeo ga;
  if (!choose_bwd(d, 8, false, ga)) return 0;
  return (size_t)ga.P * d->Cs * 2 * sizeof(float) + 256;
}

int mvf_stream_bwd(const mvfb_mvf_desc* d, const void* gp, long long g_stride, const void* x, void* dx,
                   long long dx_stride, const float* wt, const float* wh, const float* ww, const float* gamma,
                   const float* beta, const float* mean, const float* rstd, float* dwt, float* dwh, float* dww,
                   float* dgamma, float* dbeta, void* ws, cudaStream_t st) {
  BGeo ga, g;
  if (!choose_bwd(d, 8, false, ga) || !choose_bwd(d, VB, true, g)) return MVFB_ERR_UNSUPPORTED;
  if (((uintptr_t)x & 15) || ((uintptr_t)gp & 15) || ((uintptr_t)dx & 7) || g_stride % 8 != 0 || dx_stride % 4 != 0)
    return MVFB_ERR_UNSUPPORTED;
  CUtensorMap tmx, tmg, tmxa, tmga;
  int rc;
  if ((rc = make_map(&tmx, x, d->C, g, true))) return rc;
  if ((rc = make_map(&tmg, gp, g_stride, g, false))) return rc;
  if ((rc = make_map(&tmxa, x, d->C, ga, true))) return rc;
  if ((rc = make_map(&tmga, gp, g_stride, ga, false))) return rc;
  const bool has_h = d->mode != MVFB_MODE_T, has_w = d->mode == MVFB_MODE_THW;
  BArgs a;
  a.g = g;
  a.use_hs = d->use_hs; a.training = d->training;
  a.share_h = has_h && wh == wt; a.share_w = has_w && ww == wt;
  a.wt = wt; a.wh = has_h ? wh : nullptr; a.ww = has_w ? ww : nullptr;
  a.gamma = gamma; a.beta = beta; a.mean = mean; a.rstd = rstd;
  a.partials = (float*)ws;
  a.PA = ga.P;
  a.dwt = dwt; a.dwh = (has_h && !a.share_h) ? dwh : nullptr; a.dww = (has_w && !a.share_w) ? dww : nullptr;
  a.dgamma = dgamma; a.dbeta = dbeta;
  a.dx = (__nv_bfloat16*)dx; a.dx_pix = dx_stride;
  static DevOnce once;
  if (once.pending()) {
    MVFB_CUDA(cudaFuncSetAttribute(mvf_stream_bwd_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    MVFB_CUDA(cudaFuncSetAttribute(mvf_stream_bwd_dx, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    once.done();
  }
  const dim3 grid(g.ngroups * g.P);
  if (d->use_hs) {
    BArgs aa = a;
    aa.g = ga;
    const int itemsA = ga.H * ga.W * (ga.Cg / 8);
    mvf_stream_bwd_reduce<<<ga.ngroups * ga.P, 32 * ((itemsA + 31) / 32 + 1), smem_bytes(ga, false), st>>>(tmxa, tmga, aa);
    count_launch();
    MVFB_LAUNCH_CHECK();
  } else {
    MVFB_CUDA(cudaMemsetAsync(dwt, 0, sizeof(float) * 3 * d->Cs, st));
    if (a.dwh) MVFB_CUDA(cudaMemsetAsync(a.dwh, 0, sizeof(float) * 3 * d->Cs, st));
    if (a.dww) MVFB_CUDA(cudaMemsetAsync(a.dww, 0, sizeof(float) * 3 * d->Cs, st));
  }
  const int itemsB = g.H * g.W * (g.Cg / VB);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(32 * ((itemsB + 31) / 32 + 1)); cfg.dynamicSmemBytes = smem_bytes(g, true);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = d->use_hs ? 1 : 0;   // PDL behind the reduce kernel only
  cfg.attrs = attr; cfg.numAttrs = 1;
  MVFB_CUDA(cudaLaunchKernelEx(&cfg, mvf_stream_bwd_dx, tmx, tmg, a));
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

}  // namespace mvfb

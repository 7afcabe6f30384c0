// MVF backward, sweep generation: ONE cooperative launch on the bf16-operand FMA -- the backward counterpart of
// mvf_sweep.cu.  Replaces autograd of MVF.py:109-137 (three depthwise Conv3d, BatchNorm3d, HardSwish) for bf16 NHWC slabs.
//
// What the stream generation (mvf_stream_bwd.cu) was bound by (profiles/r01_own_kernels_in_step_ncu_full.csv: 6.8 % DRAM
// throughput, 0.12 of the 3*E*s roofline in the step): two launches with two prologues and two pipeline ramps, fp32
// FFMA2 arithmetic with a bf16 -> fp32 unpack in front of every operand, an fp32 dz exchange (16 bytes per 4 channels
// written, 5 x 16 bytes read), 16-byte TMA rows, and no kernel at all for 28 x 28 / 16 x 16 frames (ring / generic tiers).
// This kernel:
//   * streams every frame of (x, g) of the CTA's clips TWICE through one TMA ring inside one launch: sweep 0 recomputes
//     z, u and du = g * hardswish'(u) and accumulates (sum du, sum du*z) per channel; the CTAs of a channel group
//     exchange their partial rows as {value, tag} words (same protocol as mvf_sweep.cu: no flag, no atomic, tag = host
//     call number + device replay counter); sweep 1 recomputes du, forms dz = scale*du + u*c1 + c0 (BatchNorm backward
//     folded into three per-channel constants), accumulates the seven tap-gradient sums and applies the transposed
//     stencil.  Cooperative launch: all CTAs resident.  use_hs = False needs no statistics: sweep 1 only.
//   * all stencil arithmetic is `fma.rn.f32.bf16` (FHFMA): x, g, dz and the taps stay PACKED; products are exact, sums fp32.
//     dz is rounded to bf16 once -- for the shared-memory exchange (8 bytes per 4 channels instead of 16), the
//     tap-gradient products dz * x and the transposed stencil alike; dx is a bf16 tensor anyway.
//   * items of 4 channels, 1 / 2 / 4 items per thread (all of the SAME channel vector, so taps, BatchNorm constants and
//     the 28 tap-gradient accumulators are shared); whole frames per CTA, channel group as wide as the thread budget
//     allows: 64 channels at 7 x 7 (128-byte TMA rows), 16 at 14 x 14, 8 at 28 x 28 -- every R50 / R101 slab at 224 px
//     and the 16 x 16 / 8 x 8 slabs at 256 px.
//   * temporal terms roll through registers like the forward's: a thread keeps its own dz(t-1), dz(t-2) packed, and
//     dx(t-1) is finished in step t, in the same barrier interval as the computation of dz(t) (one consumer barrier
//     per frame, two independent dependency chains per item between barriers).
// dx may alias g (include/mvf_b200.h): a frame of g is in shared memory before any dx of that frame is written, and
// the frames a CTA re-reads in sweep 1 are never frames whose dx has been written (its own clips, in order).
#include <cuda_bf16.h>

#include <atomic>

#include "common.cuh"
#include "mvf_internal.cuh"
#include "ptx.cuh"
#include "mvf_stream.cuh"

namespace mvfb {

using namespace stream;

namespace {

constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxRing = 10;
constexpr int kMaxCWarps = 13;                 // 13 consumer warps + 1 producer warp = 448 threads: <= 144 registers each
constexpr int kMaxThreads = 32 * (kMaxCWarps + 1);
constexpr int V = 4;                           // channels per item

struct WGeo {
  int N, T, Cs, H, W;
  int Cg, G, ngroups;     // channels per CTA, 4-channel vectors per pixel, channel groups
  int Hp, Wp;             // padded frame extents
  int slot_x, slot_g;     // bytes: padded bf16 x frame, unpadded bf16 g frame (multiples of 128)
  int stage;              // slot_x + slot_g
  int R;                  // ring slots
  int P;                  // CTAs per channel group (clips are dealt round-robin)
  int IT;                 // items per thread
  int PP;                 // pixels per pass: item i of a thread is pixel (tid / G) + i * PP
  int cwarps;             // consumer warps
};

struct WArgs {
  WGeo g;
  int use_hs, training, share_h, share_w;
  const float *wt, *wh, *ww, *gamma, *beta, *mean, *rstd;
  uint2* partials;        // [grid][2*Cg] {fp32 partial sum, tag}
  unsigned int epoch_hi;  // host call number (24 bits)
  unsigned int* epochs;   // [ngroups] device replay counters
  float *dwt, *dwh, *dww, *dgamma, *dbeta;
  __nv_bfloat16* dx;
  long long dx_pix;
  const __nv_bfloat16* dx_add;   // optional: dx += dx_add (addressed like dx): the gradient of the block's identity path
};

struct P4 {               // 4 bf16 channels, packed as loaded
  uint32_t v[2];
};
__device__ __forceinline__ P4 lds_p4(uint32_t addr) {
  P4 r;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.v[0]), "=r"(r.v[1]) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts_p4(uint32_t addr, const P4& p) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(p.v[0]), "r"(p.v[1]) : "memory");
}
__device__ __forceinline__ P4 zero_p4() {
  P4 r;
  r.v[0] = 0u; r.v[1] = 0u;
  return r;
}
// (lo, hi) += a.(lo, hi) * b.(lo, hi): two FHFMA.BF16, operands read from the register halves directly
__device__ __forceinline__ void fh2(float2& z, uint32_t a, uint32_t b) {
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
      "mov.b32 {al, ah}, %2;\n\t"
      "mov.b32 {bl, bh}, %3;\n\t"
      "fma.rn.f32.bf16 %0, al, bl, %0;\n\t"
      "fma.rn.f32.bf16 %1, ah, bh, %1;\n\t}"
      : "+f"(z.x), "+f"(z.y)
      : "r"(a), "r"(b));
}
__device__ __forceinline__ void fh4(float2 (&z)[2], const P4& a, const P4& b) {
  fh2(z[0], a.v[0], b.v[0]);
  fh2(z[1], a.v[1], b.v[1]);
}
__device__ __forceinline__ float hswish_grad_f(float u) {
  return __saturatef(fmaf(u, 1.f / 6.f, 0.5f)) + ((u > -3.f && u < 3.f) ? u * (1.f / 6.f) : 0.f);
}
__device__ __forceinline__ uint2 ld_tagged(const uint2* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return make_uint2((uint32_t)w, (uint32_t)(w >> 32));
}
__device__ __forceinline__ void st_tagged(uint2* p, uint2 v) {
  const unsigned long long w = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}

struct Carve {
  uint64_t *full, *empty;
  uint8_t* stages;
  uint8_t* dz;            // 2 padded bf16 frames
  float* s_const;         // [4][Cg]: scale, shift, c1, c0
  double *s_dpart, *s_dsum;
  float* s_red;           // [cwarps][G][28]
};

__device__ __forceinline__ Carve carve(uint8_t* smem, const WGeo& g) {
  Carve c;
  c.full = reinterpret_cast<uint64_t*>(smem);
  c.empty = c.full + 16;
  c.stages = smem + 256;
  c.dz = c.stages + (size_t)g.R * g.stage;
  c.s_const = reinterpret_cast<float*>(c.dz + 2 * (size_t)g.slot_x);
  c.s_dpart = reinterpret_cast<double*>(c.s_const + 4 * g.Cg);
  c.s_dsum = c.s_dpart + kMaxThreads;
  c.s_red = reinterpret_cast<float*>(c.s_dsum + 128);
  return c;
}

size_t smem_bytes(const WGeo& g) {
  return 256 + (size_t)g.R * g.stage + 2 * (size_t)g.slot_x + (size_t)4 * g.Cg * 4 + (size_t)kMaxThreads * 8 + 128 * 8 +
         (size_t)kMaxCWarps * 7 * g.Cg * 4 + 256;
}

template <int IT>
__global__ void __launch_bounds__(kMaxThreads, 1)
mvf_sweep_bwd_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmg, const WArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const WGeo& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cwarps = g.cwarps, nconsumer = 32 * cwarps, G = g.G;
  const int cg = blockIdx.x % g.ngroups, p = blockIdx.x / g.ngroups;
  const int c0 = cg * g.Cg;
  const int nclips = p < g.N ? (g.N - p + g.P - 1) / g.P : 0;
  const int Q = nclips * g.T;                                   // frames per sweep
  const bool two = a.use_hs != 0;                               // statistics sweep + grid exchange first
  const int QQ = two ? 2 * Q : Q;
  unsigned int tag = 0;
  if (two) tag = (a.epoch_hi << 8) | ((__ldcv(a.epochs + cg) + 1u) & 0xffu);
  Carve c = carve(smem, g);

  if (tid == 0) {
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmg);
    for (int s = 0; s < g.R; ++s) { mbar_init(&c.full[s], 1); mbar_init(&c.empty[s], cwarps); }
    fence_barrier_init();
  }
  // zero both dz frames: their halo (and pixels no item owns) must read as 0 for the transposed stencil
  for (int i = tid * 16; i < 2 * g.slot_x; i += blockDim.x * 16) *reinterpret_cast<uint4*>(c.dz + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const bool producer = warp == cwarps;
  const uint32_t frame_tx = (uint32_t)(g.Hp * g.Wp * g.Cg * 2 + g.H * g.W * g.Cg * 2);
  auto issue = [&](int s, int q) {                              // frame q of this CTA's stream (both sweeps)
    const int qq = q < Q ? q : q - Q;
    const int kclip = qq / g.T, t = qq - kclip * g.T;
    const int frame = (p + kclip * g.P) * g.T + t;
    mbar_arrive_expect_tx(&c.full[s], frame_tx);
    uint8_t* dst = c.stages + (size_t)s * g.stage;
    tma_load_4d(dst, &tmx, &c.full[s], c0, -1, -1, frame);
    tma_load_4d(dst + g.slot_x, &tmg, &c.full[s], c0, 0, 0, frame);
  };

  if (producer) {
    if (lane == 0) {
      int s = 0, use = 0;
      for (int q = 0; q < QQ; ++q) {
        if (use > 0) mbar_wait(&c.empty[s], (use - 1) & 1);
        issue(s, q);
        if (++s == g.R) { s = 0; ++use; }
      }
    }
    return;
  }

  // ===== consumers =====
  const int vec = tid % G, pbase = tid / G;
  const int HW = g.H * g.W;
  const int pixb = g.Cg * 2, rowb = g.Wp * pixb;
  bool act[IT];
  uint32_t off[IT], goff[IT], pixoff[IT];
#pragma unroll
  for (int i = 0; i < IT; ++i) {
    const int pix = pbase + i * g.PP;
    act[i] = pbase < g.PP && pix < HW;
    const int pp = act[i] ? pix : 0;
    const int h = pp / g.W, w = pp - h * g.W;
    off[i] = (uint32_t)(((h + 1) * g.Wp + (w + 1)) * pixb + vec * (2 * V));
    goff[i] = (uint32_t)(g.slot_x + (h * g.W + w) * pixb + vec * (2 * V));
    pixoff[i] = (uint32_t)((h * g.W + w) * (int)a.dx_pix + c0 + vec * V);
  }
  // stencil taps of this thread's 4 channels, rounded to bf16 (the forward's taps), packed; merged centre as (hi, lo)
  P4 kc, kcl, kt0, kt2, kh0, kh2, kw0, kw2;
  {
    float w[3][12];
#pragma unroll
    for (int view = 0; view < 3; ++view) {
      const float* wv = view == 0 ? a.wt : (view == 1 ? a.wh : a.ww);
#pragma unroll
      for (int q = 0; q < 12; ++q) w[view][q] = wv ? __ldg(wv + (size_t)(c0 + vec * V) * 3 + q) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int e = 6 * j, o = 6 * j + 3;                       // taps of channels 2j and 2j+1: [ch][3]
      const float ce = round_bf16(w[0][e + 1]) + round_bf16(w[1][e + 1]) + round_bf16(w[2][e + 1]);
      const float co = round_bf16(w[0][o + 1]) + round_bf16(w[1][o + 1]) + round_bf16(w[2][o + 1]);
      kc.v[j] = pack_bf16(ce, co);
      kcl.v[j] = pack_bf16(ce - round_bf16(ce), co - round_bf16(co));
      kt0.v[j] = pack_bf16(w[0][e], w[0][o]); kt2.v[j] = pack_bf16(w[0][e + 2], w[0][o + 2]);
      kh0.v[j] = pack_bf16(w[1][e], w[1][o]); kh2.v[j] = pack_bf16(w[1][e + 2], w[1][o + 2]);
      kw0.v[j] = pack_bf16(w[2][e], w[2][o]); kw2.v[j] = pack_bf16(w[2][e + 2], w[2][o + 2]);
    }
  }
  float2 scale[2], shift[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) { scale[j] = make_float2(1.f, 1.f); shift[j] = make_float2(0.f, 0.f); }
  if (two) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ch = c0 + vec * V + 2 * j;
      const float s0 = a.gamma[ch] * a.rstd[ch], s1 = a.gamma[ch + 1] * a.rstd[ch + 1];
      scale[j] = make_float2(s0, s1);
      shift[j] = make_float2(a.beta[ch] - a.mean[ch] * s0, a.beta[ch + 1] - a.mean[ch + 1] * s1);
    }
  }

  const uint32_t full0 = smem_u32(c.full), empty0 = smem_u32(c.empty);
  const uint32_t base = smem_u32(c.stages), stageb = (uint32_t)g.stage, ringb = (uint32_t)g.R * stageb;
  uint32_t cur = base, fb = full0, ph = 0;                      // current slot, its full barrier, its parity
  auto advance = [&](uint32_t& nxt, uint32_t& fb1, uint32_t& ph1) {
    nxt = cur + stageb; fb1 = fb + 8; ph1 = ph;
    if (nxt == base + ringb) { nxt = base; fb1 = full0; ph1 ^= 1; }
  };
  // z of item i at the current frame from its rolled centres and the staged neighbours
  auto stencil = [&](const P4& xm, const P4& xc, const P4& xp, uint32_t ctr, float2 (&z)[2], P4& xhm, P4& xhp, P4& xwm, P4& xwp) {
    xhm = lds_p4(ctr - rowb); xhp = lds_p4(ctr + rowb); xwm = lds_p4(ctr - pixb); xwp = lds_p4(ctr + pixb);
    z[0] = make_float2(0.f, 0.f); z[1] = make_float2(0.f, 0.f);
    fh4(z, xc, kc); fh4(z, xc, kcl);
    fh4(z, xm, kt0); fh4(z, xp, kt2);
    fh4(z, xhm, kh0); fh4(z, xhp, kh2);
    fh4(z, xwm, kw0); fh4(z, xwp, kw2);
  };

  // ---------------------------------------------------------------- sweep 0: (sum du, sum du*z)
  if (two) {
    float2 sum[2], sq[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) { sum[j] = make_float2(0.f, 0.f); sq[j] = make_float2(0.f, 0.f); }
    for (int kclip = 0; kclip < nclips; ++kclip) {
      P4 xm[IT], xc[IT];
      wait_u32(fb, ph);
#pragma unroll
      for (int i = 0; i < IT; ++i) { xm[i] = zero_p4(); xc[i] = act[i] ? lds_p4(cur + off[i]) : zero_p4(); }
#pragma unroll 1
      for (int t = 0; t < g.T; ++t) {
        uint32_t nxt, fb1, ph1;
        advance(nxt, fb1, ph1);
        P4 xp[IT];
        if (t + 1 < g.T) wait_u32(fb1, ph1);
#pragma unroll
        for (int i = 0; i < IT; ++i) xp[i] = (t + 1 < g.T && act[i]) ? lds_p4(nxt + off[i]) : zero_p4();
#pragma unroll
        for (int i = 0; i < IT; ++i) {
          if (act[i]) {
            float2 z[2];
            P4 n0, n1, n2, n3;
            stencil(xm[i], xc[i], xp[i], cur + off[i], z, n0, n1, n2, n3);
            const P4 gq = lds_p4(cur + goff[i]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float2 u = __ffma2_rn(z[j], scale[j], shift[j]);
              float2 du;
              du.x = bf16_lo(gq.v[j]) * hswish_grad_f(u.x);
              du.y = bf16_hi(gq.v[j]) * hswish_grad_f(u.y);
              sum[j] = __fadd2_rn(sum[j], du);
              sq[j] = __ffma2_rn(du, z[j], sq[j]);
            }
          }
          xm[i] = xc[i]; xc[i] = xp[i];
        }
        __syncwarp();
        if (lane == 0) arrive_u32(empty0 + (fb - full0));
        cur = nxt; fb = fb1; ph = ph1;
      }
    }
    // ---- CTA partial row -> grid exchange -> BatchNorm-backward constants of this channel group
    float acc8[8] = {sum[0].x, sum[0].y, sum[1].x, sum[1].y, sq[0].x, sq[0].y, sq[1].x, sq[1].y};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float v = acc8[q];
      for (int o = 16; o >= G; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc8[q] = v;
    }
    if (lane < G) {                                             // lane < G: vec == lane (32 % G == 0, tid % 32 == lane)
#pragma unroll
      for (int q = 0; q < 8; ++q) c.s_red[(warp * G + lane) * 8 + q] = acc8[q];
    }
    consumer_bar_sync(nconsumer);
    const int per = 2 * g.Cg, rows = g.P;                       // row layout [vec][sum 0..3 | sumsq 0..3]
    if (tid < per) {
      float v = 0.f;
      for (int wv = 0; wv < cwarps; ++wv) v += c.s_red[wv * per + tid];
      st_tagged(a.partials + (size_t)blockIdx.x * per + tid, make_uint2(__float_as_uint(v), tag));
    }
    float bn_g = 1.f, bn_b = 0.f, bn_m = 0.f, bn_r = 1.f;
    if (tid < g.Cg) { bn_g = a.gamma[c0 + tid]; bn_b = a.beta[c0 + tid]; bn_m = a.mean[c0 + tid]; bn_r = a.rstd[c0 + tid]; }
    const int parts = nconsumer / per;
    {
      const int k = tid % per, part = tid / per;
      if (part < parts) {
        double accd = 0.0;
        const long long t0 = clock64();
        constexpr int kBatch = 8;
        for (int r0 = part; r0 < rows; r0 += kBatch * parts) {
          uint2 v[kBatch];
          bool ok;
          do {
            ok = true;
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
              const int r = r0 + u * parts;
              v[u] = make_uint2(0u, tag);
              if (r < rows) v[u] = ld_tagged(a.partials + ((size_t)r * g.ngroups + cg) * per + k);
              ok = ok && v[u].y == tag;
            }
            if (!ok && clock64() - t0 > 4000000000LL) __trap();  // a CTA that never arrives must not hang the GPU
          } while (!ok);
#pragma unroll
          for (int u = 0; u < kBatch; ++u) accd += (double)__uint_as_float(v[u].x);
        }
        c.s_dpart[part * per + k] = accd;
      }
    }
    consumer_bar_sync(nconsumer);
    if (tid < g.Cg) {
      const int i1 = (tid / V) * 8 + (tid % V), i2 = i1 + 4;
      double s1 = 0.0, s2 = 0.0;
      for (int q = 0; q < parts; ++q) { s1 += c.s_dpart[q * per + i1]; s2 += c.s_dpart[q * per + i2]; }
      // s2 = sum du*z;  dgamma = sum du*zhat = rstd * (s2 - mean*s1)
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
    if (p == 0 && tid == 0) a.epochs[cg] = tag & 0xffu;          // every CTA of the group has read the counter
    consumer_bar_sync(nconsumer);
  }

  // ---------------------------------------------------------------- sweep 1: dz, tap gradients, transposed stencil
  float2 c1v[2], c0v[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) { c1v[j] = make_float2(0.f, 0.f); c0v[j] = make_float2(0.f, 0.f); }
  if (two) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ch = vec * V + 2 * j;
      c1v[j] = make_float2(c.s_const[2 * g.Cg + ch], c.s_const[2 * g.Cg + ch + 1]);
      c0v[j] = make_float2(c.s_const[3 * g.Cg + ch], c.s_const[3 * g.Cg + ch + 1]);
    }
  }
  float2 acc[7][2];                                             // tap sums: centre, t-1, t+1, h-1, h+1, w-1, w+1
#pragma unroll
  for (int q = 0; q < 7; ++q) { acc[q][0] = make_float2(0.f, 0.f); acc[q][1] = make_float2(0.f, 0.f); }
  const uint32_t dz0 = smem_u32(c.dz), dzb = (uint32_t)g.slot_x;
  const size_t frame_elems = (size_t)HW * a.dx_pix;
  uint32_t par = 0;                                             // which dz buffer this frame writes
  // dx(s) = kc dz(s) + kt2 dz(s-1) + kt0 dz(s+1) + kh0 dz(s)[h+1] + kh2 dz(s)[h-1] + kw0 dz(s)[w+1] + kw2 dz(s)[w-1]
  // (the transposed stencil).  The H / W neighbours of dz(s) come from the shared-memory frame that step s filled and
  // the barrier at the end of step s published; the three temporal terms are this thread's own registers.  It runs in
  // step s + 1, in the same barrier interval as the computation of dz(s+1): two independent dependency chains per item.
  // `addq`: the identity path's gradient of this item (mvf_bwd_add), requested at the TOP of the step -- a global load
  // inside the dependency chain of finish() costs a DRAM round trip per frame (measured: +100 us per launch)
  auto finish = [&](int i, uint32_t dzbuf, const P4& d0, const P4& dm, const P4& dp, __nv_bfloat16* dst, const uint2& addq) {
    float2 o[2] = {make_float2(bf16_lo(addq.x), bf16_hi(addq.x)), make_float2(bf16_lo(addq.y), bf16_hi(addq.y))};
    const uint32_t dc = dzbuf + off[i];
    const P4 n0 = lds_p4(dc + rowb), n1 = lds_p4(dc - rowb), n2 = lds_p4(dc + pixb), n3 = lds_p4(dc - pixb);
    fh4(o, d0, kc); fh4(o, d0, kcl);
    fh4(o, dm, kt2); fh4(o, dp, kt0);
    fh4(o, n0, kh0); fh4(o, n1, kh2);
    fh4(o, n2, kw0); fh4(o, n3, kw2);
    uint2 w2;
    w2.x = pack_bf16(o[0].x, o[0].y); w2.y = pack_bf16(o[1].x, o[1].y);
    *reinterpret_cast<uint2*>(dst + pixoff[i]) = w2;
  };
  for (int kclip = 0; kclip < nclips; ++kclip) {
    __nv_bfloat16* dxf = a.dx + (size_t)(p + kclip * g.P) * g.T * frame_elems;   // frame 0 of this clip
    P4 xm[IT], xc[IT], dz1[IT], dz2[IT];                        // dz(t-1), dz(t-2)
    wait_u32(fb, ph);
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      xm[i] = zero_p4(); dz1[i] = zero_p4(); dz2[i] = zero_p4();
      xc[i] = act[i] ? lds_p4(cur + off[i]) : zero_p4();
    }
#pragma unroll 1
    for (int t = 0; t < g.T; ++t) {
      uint32_t nxt, fb1, ph1;
      advance(nxt, fb1, ph1);
      P4 xp[IT];
      uint2 addq[IT];
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        addq[i] = make_uint2(0u, 0u);
        if (a.dx_add && t > 0 && act[i])
          addq[i] = __ldg(reinterpret_cast<const uint2*>(a.dx_add + (dxf - a.dx) + (size_t)(t - 1) * frame_elems + pixoff[i]));
      }
      if (t + 1 < g.T) wait_u32(fb1, ph1);
#pragma unroll
      for (int i = 0; i < IT; ++i) xp[i] = (t + 1 < g.T && act[i]) ? lds_p4(nxt + off[i]) : zero_p4();
      const uint32_t dzw = dz0 + par * dzb, dzr = dz0 + (par ^ 1u) * dzb;
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        P4 dzp = zero_p4();
        if (act[i]) {
          float2 z[2];
          P4 xhm, xhp, xwm, xwp;
          stencil(xm[i], xc[i], xp[i], cur + off[i], z, xhm, xhp, xwm, xwp);
          const P4 gq = lds_p4(cur + goff[i]);
          if (two) {
            float2 dzf[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float2 u = __ffma2_rn(z[j], scale[j], shift[j]);
              float2 du;
              du.x = bf16_lo(gq.v[j]) * hswish_grad_f(u.x);
              du.y = bf16_hi(gq.v[j]) * hswish_grad_f(u.y);
              dzf[j] = __ffma2_rn(scale[j], du, __ffma2_rn(u, c1v[j], c0v[j]));
            }
            dzp.v[0] = pack_bf16(dzf[0].x, dzf[0].y);
            dzp.v[1] = pack_bf16(dzf[1].x, dzf[1].y);
          } else {
            dzp = gq;                                            // y = z: dz is the incoming gradient itself
          }
          fh4(acc[0], dzp, xc[i]);
          fh4(acc[1], dzp, xm[i]);
          fh4(acc[2], dzp, xp[i]);
          fh4(acc[3], dzp, xhm);
          fh4(acc[4], dzp, xhp);
          fh4(acc[5], dzp, xwm);
          fh4(acc[6], dzp, xwp);
          sts_p4(dzw + off[i], dzp);
          if (t > 0) finish(i, dzr, dz1[i], dz2[i], dzp, dxf + (size_t)(t - 1) * frame_elems, addq[i]);
        }
        dz2[i] = dz1[i]; dz1[i] = dzp;
        xm[i] = xc[i]; xc[i] = xp[i];
      }
      __syncwarp();
      if (lane == 0) arrive_u32(empty0 + (fb - full0));          // x(t) neighbours and g(t) are consumed
      consumer_bar_sync(nconsumer);                              // dz(t) of every pixel is in shared memory
      par ^= 1;
      cur = nxt; fb = fb1; ph = ph1;
    }
    // dx(T-1): there is no frame T.  dz(T-1) sits in the buffer the last step wrote (par was flipped after it).
    uint2 addl[IT];
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      addl[i] = make_uint2(0u, 0u);
      if (a.dx_add && act[i])
        addl[i] = __ldg(reinterpret_cast<const uint2*>(a.dx_add + (dxf - a.dx) + (size_t)(g.T - 1) * frame_elems + pixoff[i]));
    }
#pragma unroll
    for (int i = 0; i < IT; ++i)
      if (act[i]) finish(i, dz0 + (par ^ 1u) * dzb, dz1[i], dz2[i], zero_p4(), dxf + (size_t)(g.T - 1) * frame_elems, addl[i]);
  }

  // ---- tap gradients: reduce over the CTA, then one atomic per (channel, tap) (the outputs were zeroed before the launch)
  float flat[7 * V];
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    flat[q * V + 0] = acc[q][0].x; flat[q * V + 1] = acc[q][0].y;
    flat[q * V + 2] = acc[q][1].x; flat[q * V + 3] = acc[q][1].y;
  }
#pragma unroll
  for (int q = 0; q < 7 * V; ++q) {
    float v = flat[q];
    for (int o = 16; o >= G; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    flat[q] = v;
  }
  consumer_bar_sync(nconsumer);                                  // s_red doubled as the statistics scratch
  if (lane < G) {
#pragma unroll
    for (int q = 0; q < 7 * V; ++q) c.s_red[(warp * G + lane) * (7 * V) + q] = flat[q];
  }
  consumer_bar_sync(nconsumer);
  float* dst_h = a.share_h ? a.dwt : a.dwh;
  float* dst_w = a.share_w ? a.dwt : a.dww;
  for (int i = tid; i < g.Cg * 7; i += nconsumer) {
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

// Whole frames per CTA.  Candidates are (channel group, items per thread) pairs whose items fit the thread budget; the
// score prefers wide TMA rows (a 16-byte row costs the TMA unit as much as a 64-byte one) and <= 2 items per thread
// (4 items need ~145 registers and spill at the 128 the thread count allows).
bool fill(const mvfb_mvf_desc* d, int Cg, int IT, WGeo& g) {
  const int HW = d->H * d->W;
  const int G = Cg / V;                                         // 16, 8, 4, 2: divides 32
  const int PP = (HW + IT - 1) / IT;
  if (d->Cs % Cg || PP * G > 32 * kMaxCWarps) return false;
  g.N = d->N; g.T = d->T; g.Cs = d->Cs; g.H = d->H; g.W = d->W;
  g.Cg = Cg; g.G = G; g.ngroups = d->Cs / Cg;
  g.Hp = d->H + 2; g.Wp = d->W + 2;
  g.slot_x = (g.Hp * g.Wp * Cg * 2 + 127) / 128 * 128;
  g.slot_g = (HW * Cg * 2 + 127) / 128 * 128;
  g.stage = g.slot_x + g.slot_g;
  g.IT = IT; g.PP = PP;
  g.cwarps = (PP * G + 31) / 32;
  if (2 * Cg > 32 * g.cwarps) return false;                     // the statistics row needs 2*Cg consumer threads
  int R = (int)((150 * 1024 - 2 * g.slot_x) / g.stage);
  if (R > kMaxRing) R = kMaxRing;
  if (R < 3) return false;
  g.R = R;
  int P = num_sms() / g.ngroups;
  if (P < 1) P = 1;
  if (P > d->N) P = d->N;
  g.P = P;
  return smem_bytes(g) <= (size_t)kSmemLimit;
}

bool choose(const mvfb_mvf_desc* d, WGeo& g) {
  if (d->dtype != MVFB_BF16 || d->layout != MVFB_NHWC) return false;
  if (d->Cs % 8 != 0 || d->C % 8 != 0 || d->W + 2 > 256 || d->H + 2 > 256) return false;
  const int cands[4] = {64, 32, 16, 8};
  const double row_score[4] = {1.0, 1.0, 0.8, 0.5};
  double best = 0.0;
  for (int ci = 0; ci < 4; ++ci) {
    for (int IT = 1; IT <= 4; IT *= 2) {
      WGeo c;
      if (!fill(d, cands[ci], IT, c)) continue;
      const double score = row_score[ci] * (IT == 4 ? 0.7 : 1.0) + 1e-3 * (4 - ci) - 1e-4 * IT;
      if (score > best) { best = score; g = c; }
    }
  }
  return best > 0.0;
}

size_t partial_bytes(const WGeo& g) { return ((size_t)g.ngroups * g.P * 2 * g.Cg * sizeof(uint2) + 255) / 256 * 256; }

int make_map(CUtensorMap* tm, const void* base, long long pix_stride, const WGeo& g, bool padded) {
  const uint64_t dims[4] = {(uint64_t)g.Cs, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N * g.T};
  const uint64_t strides[3] = {(uint64_t)pix_stride * 2, (uint64_t)g.W * pix_stride * 2, (uint64_t)g.H * g.W * pix_stride * 2};
  const uint32_t box[4] = {(uint32_t)g.Cg, (uint32_t)(padded ? g.Wp : g.W), (uint32_t)(padded ? g.Hp : g.H), 1u};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, nullptr,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

unsigned int next_call() {
  static std::atomic<unsigned int> e{0xb0d001u};
  return e.fetch_add(1u, std::memory_order_relaxed) & 0xffffffu;
}

template <int IT>
int launch(const CUtensorMap& tmx, const CUtensorMap& tmg, WArgs& a, cudaStream_t st) {
  static DevOnce once;
  if (once.pending()) {
    MVFB_CUDA(cudaFuncSetAttribute(mvf_sweep_bwd_kernel<IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    once.done();
  }
  const WGeo& g = a.g;
  const dim3 grid(g.ngroups * g.P), block(32 * (g.cwarps + 1));
  const size_t smem = smem_bytes(g);
  if (a.use_hs) {
    a.epoch_hi = next_call();
    void* params[3] = {(void*)&tmx, (void*)&tmg, (void*)&a};
    // cooperative launch: the driver guarantees that all CTAs are resident, which the grid exchange relies on
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)mvf_sweep_bwd_kernel<IT>, grid, block, params, smem, st);
    if (e == cudaErrorCooperativeLaunchTooLarge) {
      (void)cudaGetLastError();
      return MVFB_ERR_UNSUPPORTED;                               // device shared with another context: two-launch tier
    }
    MVFB_CUDA(e);
  } else {
    mvf_sweep_bwd_kernel<IT><<<grid, block, smem, st>>>(tmx, tmg, a);
  }
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

}  // namespace

bool mvf_sweep_bwd_supported(const mvfb_mvf_desc* d) {
  WGeo g;
  return choose(d, g);
}

size_t mvf_sweep_bwd_ws(const mvfb_mvf_desc* d) {
  WGeo g;
  if (!choose(d, g)) return 0;
  return partial_bytes(g) + (size_t)g.ngroups * sizeof(unsigned int) + 256;
}

int mvf_sweep_bwd(const mvfb_mvf_desc* d, const void* gp, long long g_stride, const void* x, void* dx,
                  long long dx_stride, const float* wt, const float* wh, const float* ww, const float* gamma,
                  const float* beta, const float* mean, const float* rstd, float* dwt, float* dwh, float* dww,
                  float* dgamma, float* dbeta, void* ws, const void* dx_add, cudaStream_t st) {
  WGeo g;
  if (!choose(d, g)) return MVFB_ERR_UNSUPPORTED;
  if ((uintptr_t)dx_add & 7) return MVFB_ERR_UNSUPPORTED;
  if (((uintptr_t)x & 15) || ((uintptr_t)gp & 15) || ((uintptr_t)dx & 7) || g_stride % 8 != 0 || dx_stride % 4 != 0 ||
      (long long)d->H * d->W * dx_stride + d->Cs >= (1LL << 32))
    return MVFB_ERR_UNSUPPORTED;
  CUtensorMap tmx, tmg;
  int rc;
  if ((rc = make_map(&tmx, x, d->C, g, true))) return rc;
  if ((rc = make_map(&tmg, gp, g_stride, g, false))) return rc;
  const bool has_h = d->mode != MVFB_MODE_T, has_w = d->mode == MVFB_MODE_THW;
  WArgs a;
  a.g = g;
  a.use_hs = d->use_hs; a.training = d->training;
  a.share_h = has_h && wh == wt; a.share_w = has_w && ww == wt;
  a.wt = wt; a.wh = has_h ? wh : nullptr; a.ww = has_w ? ww : nullptr;
  a.gamma = gamma; a.beta = beta; a.mean = mean; a.rstd = rstd;
  a.partials = (uint2*)ws;
  a.epoch_hi = 0;
  a.epochs = reinterpret_cast<unsigned int*>((char*)ws + partial_bytes(g));
  a.dwt = dwt; a.dwh = (has_h && !a.share_h) ? dwh : nullptr; a.dww = (has_w && !a.share_w) ? dww : nullptr;
  a.dgamma = dgamma; a.dbeta = dbeta;
  a.dx = (__nv_bfloat16*)dx; a.dx_pix = dx_stride;
  a.dx_add = (const __nv_bfloat16*)dx_add;
  // the tap-gradient outputs are accumulated with atomics: zero them first (one memset when the caller allocated the
  // three of them back to back, as mvfnet_b200/ops.py does)
  const size_t tapb = sizeof(float) * 3 * d->Cs;
  if (a.dwh == dwt + 3 * d->Cs && a.dww == a.dwh + 3 * d->Cs) {
    MVFB_CUDA(cudaMemsetAsync(dwt, 0, 3 * tapb, st));
  } else {
    MVFB_CUDA(cudaMemsetAsync(dwt, 0, tapb, st));
    if (a.dwh) MVFB_CUDA(cudaMemsetAsync(a.dwh, 0, tapb, st));
    if (a.dww) MVFB_CUDA(cudaMemsetAsync(a.dww, 0, tapb, st));
  }
  switch (g.IT) {
    case 1: return launch<1>(tmx, tmg, a, st);
    case 2: return launch<2>(tmx, tmg, a, st);
    default: return launch<4>(tmx, tmg, a, st);
  }
}

}  // namespace mvfb

// MVF forward, fourth generation: persistent frame-stream kernel on the sm_100 mixed-precision FMA, with train-mode
// BatchNorm in ONE launch (two sweeps over the CTA's frame stream separated by a grid barrier).
//
// What the third generation (mvf_stream.cu) was bound by: 5 bf16 -> fp32 unpacks per output element on
// the half-rate integer pipe, 72 registers of fp32 coefficients per thread, 13 consumer warps that split 4-3-3-3 over
// the four SM sub-partitions, ring bookkeeping in every frame step and, in train mode, two launches with two
// prologues.  This kernel changes the arithmetic, the thread mapping and the launch structure, not the data movement:
//   * `fma.rn.f32.bf16` (SASS FHFMA.BF16, full FFMA rate -- tools/ubench/fhfma.cu) multiplies two bf16 values taken
//     from either half of a 32-bit register and accumulates in fp32: activations stay PACKED in registers, there is
//     no unpack instruction at all, and the 7 stencil taps of 8 channels are 28 packed registers instead of 56.
//     Each of the nine taps is rounded to bf16 (what the reference's own autocast path does to its Conv3d weights, and
//     what every bf16 MVF kernel of this library does, forward and backward); the merged centre coefficient is kept
//     as a bf16 (hi, lo) pair so it is not rounded twice.  Products are exact in fp32, accumulation is fp32.
//     BatchNorm scale/shift and hard-swish stay fp32.
//   * 7 consumer warps + 1 producer warp = 2 warps per SM sub-partition (balanced), every consumer thread owns TWO
//     (pixel, 8-channel vector) items of the same channel vector (16 independent accumulation chains, shared
//     coefficient registers, up to 255 registers per thread);
//   * the ring has R = m*T slots of a compile-time stride and the frame loop is unrolled over T, so a frame's slot,
//     barrier and parity are compile-time offsets from per-clip bases: no ring bookkeeping in the frame step;
//   * the four H/W neighbour loads of frame t need no barrier (frame t is already resident), so they are issued
//     before the wait on frame t+1 and their latency overlaps it;
//   * train mode: sweep 0 accumulates (sum z, sum z^2) over all the CTA's clips, one partial row per CTA, grid
//     barrier (cooperative launch, all CTAs resident), every CTA reduces the rows of its channel group, sweep 1
//     re-streams the same frames (L2 hits; the producer warp never stops, so the ring is already full of sweep-1
//     frames when the barrier opens), normalises, applies hard-swish and stores.  HBM sees x once and y once.
// CTA = (channel group, H tile, clip lane) as in mvf_stream.cu; frames arrive by cp.async.bulk.tensor.4d with the
// H/W zero padding supplied by TMA out-of-bounds fill.  T must be 4, 8 or 16 and a padded frame tile at most
// kSlotB bytes; everything else stays on mvf_stream.cu.
#include <cuda_bf16.h>
#include <stdlib.h>

#include <atomic>
#include <type_traits>
#include <utility>

#include "common.cuh"
#include "mvf_internal.cuh"
#include "ptx.cuh"
#include "mvf_stream.cuh"

namespace mvfb {

using namespace stream;

namespace {

constexpr int kSmemLimit = 227 * 1024;
constexpr int kRing = 16;                      // slots (a multiple of every supported T)
constexpr int kSlotB = 10496;                  // compile-time slot stride: 14x14/16ch = 8192, 7x28/16ch = 8640, 7x7/64ch = 10368
constexpr int kCWarps = 7;                     // consumer warps (+1 producer warp = 2 warps per SM sub-partition)
constexpr int kMaxItems = 2 * 32 * kCWarps;    // two items per consumer thread
constexpr int kThreads = 32 * (kCWarps + 1);
constexpr int kTailB = 8192;                   // BN affine, reduction scratch (<= 64 channels per CTA)

constexpr int MODE_PLAIN = 0;                  // use_hs = False: the stencil only
constexpr int MODE_EVAL = 1;                   // BN from running statistics + hard-swish
constexpr int MODE_TRAIN = 2;                  // batch statistics: two sweeps, grid barrier in between

struct SwGeo {
  int N, T, Cs, H, W;
  int Cg, G, ngroups;     // channels per CTA, 8-channel vectors per pixel, channel groups
  int hsplit, wsplit;     // H x W tiles per frame
  int Hs, Ws;             // rows / columns per tile
  int Hp, Wp;             // padded tile extents (Hs+2, Ws+2)
  int rowpair;            // 1: warp i owns rows (2i, 2i+1) of the tile, lanes = (column, vector); results leave through
                          //    one TMA store per warp and frame.  0: items dealt linearly, results stored from registers
  int pixels, pixhalf;    // rowpair = 0: Hs*Ws; pixels owned as "item B" start at pixhalf = ceil(pixels/2)
  int cthreads;           // rowpair = 0: pixhalf*G consumer threads carry items
  int cwarps;             // consumer warps (<= kCWarps)
  int P;                  // CTAs sharing one (channel group, tile): clips are dealt round-robin
};

struct SwArgs {
  SwGeo g;
  float eps, momentum;
  const float *wt, *wh, *ww, *gamma, *beta;
  float *running_mean, *running_var, *save_mean, *save_rstd;
  uint2* partials;        // [grid][2*Cg] {fp32 partial sum, epoch}
  unsigned int epoch_hi;  // host side: unique per launch CALL (24 bits) -- distinguishes this launch's partials from
                          // anything an earlier call (any geometry, any workspace reuse) left in the workspace
  unsigned int* epochs;   // [ngroups] device-resident replay counters (low 8 bits of the tag), one per channel group:
                          // every CTA of the group reads the word, and the group's first CTA advances it once the
                          // exchange has completed -- so REPLAYS of one captured launch (same epoch_hi, same
                          // workspace) carry different tags too.  tag = epoch_hi << 8 | counter.
  int pre_frames;         // sweep-1 frames requested before the grid exchange has completed
  __nv_bfloat16* y;
  long long y_pix;
  unsigned long long* stamps;   // MVFB_SWEEP_DEBUG: [grid][8] %globaltimer stamps (tools/stream_timeline.py), else null
};

struct P8 {               // 8 bf16 channels, packed as loaded
  uint32_t v[4];
};
__device__ __forceinline__ P8 lds_p8(uint32_t addr) {
  P8 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "r"(addr));
  return r;
}
// (lo, hi) += x.(lo, hi) * k.(lo, hi): two FHFMA.BF16, operands read from the register halves directly
__device__ __forceinline__ void fh2(float2& z, uint32_t x, uint32_t k) {
  asm("{\n\t.reg .b16 xl, xh, kl, kh;\n\t"
      "mov.b32 {xl, xh}, %2;\n\t"
      "mov.b32 {kl, kh}, %3;\n\t"
      "fma.rn.f32.bf16 %0, xl, kl, %0;\n\t"
      "fma.rn.f32.bf16 %1, xh, kh, %1;\n\t}"
      : "+f"(z.x), "+f"(z.y)
      : "r"(x), "r"(k));
}
__device__ __forceinline__ void fh2_first(float2& z, uint32_t x, uint32_t k) {
  asm("{\n\t.reg .b16 xl, xh, kl, kh;\n\t"
      "mov.b32 {xl, xh}, %2;\n\t"
      "mov.b32 {kl, kh}, %3;\n\t"
      "fma.rn.f32.bf16 %0, xl, kl, 0f00000000;\n\t"
      "fma.rn.f32.bf16 %1, xh, kh, 0f00000000;\n\t}"
      : "=f"(z.x), "=f"(z.y)
      : "r"(x), "r"(k));
}
__device__ __forceinline__ void fh8(float2 (&z)[4], const P8& x, const P8& k) {
#pragma unroll
  for (int j = 0; j < 4; ++j) fh2(z[j], x.v[j], k.v[j]);
}
// {fp32 value, epoch} travels as ONE 64-bit scalar access: single-copy atomic in the PTX memory model (a .v2.u32
// vector access is two independent 32-bit accesses as far as the model is concerned, so a reader could pair the new
// epoch with a stale value half).
__device__ __forceinline__ uint2 ld_relaxed_v2(const uint2* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return make_uint2((uint32_t)w, (uint32_t)(w >> 32));
}
__device__ __forceinline__ void st_relaxed_v2(uint2* p, uint2 v) {
  const unsigned long long w = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
// Wait on an mbarrier phase.  The fast path is try_wait + one branch; the spin (labels are local to the PTX block)
// gives up after ~4 s of %globaltimer so that a transfer that never lands traps instead of hanging the GPU.
__device__ __forceinline__ void wait_phase(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u64 t0, t1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "mov.u64 t0, %%globaltimer;\n"
      "SPIN:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "mov.u64 t1, %%globaltimer;\n\t"
      "sub.u64 t1, t1, t0;\n\t"
      "setp.lt.u64 p, t1, 4000000000;\n\t"
      "@p bra SPIN;\n\t"
      "trap;\n"
      "DONE:\n\t"
      "}"
      ::"r"(bar), "r"(parity)
      : "memory");
}

// Non-blocking probe of an mbarrier phase; the result is consumed a whole frame step later.
__device__ __forceinline__ uint32_t test_phase(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}

template <int... I, class F>
__device__ __forceinline__ void static_for(std::integer_sequence<int, I...>, F&& f) {
  (f(std::integral_constant<int, I>{}), ...);
}

template <int MODE, int TT, bool RP>
__global__ void __launch_bounds__(kThreads, 1)
mvf_sweep_kernel(const __grid_constant__ CUtensorMap tmx, const SwArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int T = TT;
  constexpr int M = kRing / TT;                                 // clips resident in the ring
  const SwGeo& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nconsumer = 32 * g.cwarps;
  // CTA -> (channel group, tile, clip lane)
  const int cg = blockIdx.x % g.ngroups;
  const int rest = blockIdx.x / g.ngroups;
  const int ntiles = g.hsplit * g.wsplit;
  const int tile = rest % ntiles, p = rest / ntiles;
  const int c0 = cg * g.Cg, h0 = (tile / g.wsplit) * g.Hs, w0 = (tile % g.wsplit) * g.Ws;
  const int nclips = p < g.N ? (g.N - p + g.P - 1) / g.P : 0;
  const int KK = MODE == MODE_TRAIN ? 2 * nclips : nclips;     // clips in this CTA's stream (both sweeps)
  unsigned int epoch = 0;                                      // train mode: this launch's tag (requested first of all)
  // low byte = counter + 1 (the counter is set to it when the exchange is over), high bits = the host's call number:
  // no row left behind by an earlier call or an earlier replay of this call carries the tag
  if (MODE == MODE_TRAIN) epoch = (a.epoch_hi << 8) | ((__ldcv(a.epochs + cg) + 1u) & 0xffu);

  uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [kRing]
  uint64_t* empty = full + kRing;                               // [kRing]
  uint64_t* gate = empty + kRing;                               // train mode: opened when this CTA has its batch statistics
  uint8_t* slots = smem + 384;
  float* s_scale = reinterpret_cast<float*>(slots + (size_t)kRing * kSlotB);   // [Cg]
  float* s_shift = s_scale + g.Cg;                              // [Cg]
  double* s_dpart = reinterpret_cast<double*>(s_shift + g.Cg); // [kThreads]
  float* s_red = reinterpret_cast<float*>(s_dpart + kThreads); // [cwarps][2*Cg]

  const uint32_t frame_bytes = (uint32_t)(g.Hp * g.Wp * g.Cg * 2);
  // item mapping of a consumer thread (the producer warp computes it too and ignores it)
  const bool actA = RP ? lane < g.Ws * g.G : tid < g.cthreads;
  // spare lanes shadow an active lane of their own quarter warp (same address: a broadcast, not a bank conflict)
  int it = RP ? lane : tid;
  if (!actA) {
    if (RP) { while (it >= g.Ws * g.G) it -= g.G; } else { it = 0; }
  }
  const int vec = it % g.G, pixA = it / g.G;                    // rowpair: pixA is the column
  // Stencil taps (and, in eval mode, BatchNorm parameters) of this thread's 8 channels straight from global memory:
  // threads with equal `vec` read the same 96 B per view (L1 broadcast).  Issued FIRST, ahead of the TMA burst that
  // fills the ring, so that they are one uncongested memory round trip hidden behind the first frame.
  float4 wraw[3][6];
  float4 bnraw[4][2];
  {
#pragma unroll
    for (int view = 0; view < 3; ++view) {
      const float* wv = view == 0 ? a.wt : (view == 1 ? a.wh : a.ww);
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        wraw[view][q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (wv) wraw[view][q] = __ldg(reinterpret_cast<const float4*>(wv + (size_t)(c0 + vec * 8) * 3) + q);
      }
    }
    if (MODE == MODE_EVAL) {
      const int cb8 = c0 + vec * 8;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        bnraw[0][q] = __ldg(reinterpret_cast<const float4*>(a.gamma + cb8) + q);
        bnraw[1][q] = __ldg(reinterpret_cast<const float4*>(a.beta + cb8) + q);
        bnraw[2][q] = __ldg(reinterpret_cast<const float4*>(a.running_mean + cb8) + q);
        bnraw[3][q] = __ldg(reinterpret_cast<const float4*>(a.running_var + cb8) + q);
      }
    }
  }
  unsigned long long* stamps = a.stamps ? a.stamps + (size_t)blockIdx.x * 8 : nullptr;
  if (stamps && tid == 0) stamps[0] = gtimer();
  if (tid == 0) {
    tma_prefetch_desc(&tmx);
    for (int s = 0; s < kRing; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], g.cwarps);
    }
    mbar_init(gate, 1);
    fence_barrier_init();
  }
  __syncthreads();
  const bool producer = warp == g.cwarps;
  // clips requested before anything else happens: they need nothing but the barriers.  Train mode starts with sweep-0
  // clips only; sweep 1 is gated (below).
  const int kk0 = MODE == MODE_TRAIN ? (nclips < M ? nclips : M) : (KK < M ? KK : M);
  if (producer && lane == 0) {
    for (int kk = 0; kk < kk0; ++kk) {
      const int n = p + kk * g.P;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int s = kk * T + t;
        mbar_arrive_expect_tx(&full[s], frame_bytes);
        tma_load_4d(slots + (size_t)s * kSlotB, &tmx, &full[s], c0, w0 - 1, h0 - 1, n * T + t);
      }
    }
  }

  if (producer) {
    // ===== TMA producer: refill a clip's T slots as the consumers release them, straight through both sweeps.
    // Train mode: only `pre_frames` frames of sweep 1 are requested before this CTA has passed the grid exchange --
    // 144 CTAs refilling their whole rings at once keep L2 busy for microseconds, exactly when the exchange's few
    // latency-critical words are in flight.
    if (lane == 0) {
      int cm = kk0 % M, use = kk0 / M;                          // slot group of clip kk and how often it was used before
      for (int kk = kk0; kk < KK; ++kk) {
        const int n = p + (kk < nclips ? kk : kk - nclips) * g.P;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int s = cm * T + t;
          if (MODE == MODE_TRAIN && kk * T + t == nclips * T + a.pre_frames) mbar_wait(gate, 0);
          if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
          mbar_arrive_expect_tx(&full[s], frame_bytes);
          tma_load_4d(slots + (size_t)s * kSlotB, &tmx, &full[s], c0, w0 - 1, h0 - 1, n * T + t);
        }
        if (++cm == M) { cm = 0; ++use; }
      }
    }
  } else if (nclips > 0) {
    // ===== consumers: two (pixel, 8-channel vector) items per thread, same channel vector.
    // rowpair: item A = (row 2*warp, column), item B = the pixel below it, so A's lower and B's upper neighbour are
    // the other item's centre (8 shared-memory loads per frame step instead of 10, all conflict-free: a quarter warp
    // reads 128 contiguous bytes) and the warp's results are the rectangle rows (2*warp, 2*warp+1) x Ws x Cg.
    const bool actB = RP ? actA : (actA && pixA + g.pixhalf < g.pixels);
    const int pixB = actB ? pixA + g.pixhalf : pixA;
    const int hA = RP ? 2 * warp : pixA / g.Ws, wA = RP ? pixA : pixA - hA * g.Ws;   // row / column within the tile
    const int hB = RP ? hA + 1 : pixB / g.Ws, wB = RP ? wA : pixB - hB * g.Ws;
    const int pixb = g.Cg * 2, rowb = g.Wp * pixb;
    const uint32_t offA = (uint32_t)(((hA + 1) * g.Wp + (wA + 1)) * pixb + vec * 16);
    const uint32_t offB = (uint32_t)(((hB + 1) * g.Wp + (wB + 1)) * pixb + vec * 16);
    P8 kc, kcl, kt0, kt2, kh0, kh2, kw0, kw2;
    {
      float w[3][24];
#pragma unroll
      for (int view = 0; view < 3; ++view)
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          w[view][4 * q] = wraw[view][q].x; w[view][4 * q + 1] = wraw[view][q].y;
          w[view][4 * q + 2] = wraw[view][q].z; w[view][4 * q + 3] = wraw[view][q].w;
        }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = 6 * j, o = 6 * j + 3;                     // taps of channels 2j and 2j+1: [ch][3]
        // centre = sum of the three views' middle taps, each rounded to bf16 like every other tap; the sum itself is
        // carried as a bf16 (hi, lo) pair so that it is not rounded a second time
        const float ce = round_bf16(w[0][e + 1]) + round_bf16(w[1][e + 1]) + round_bf16(w[2][e + 1]);
        const float co = round_bf16(w[0][o + 1]) + round_bf16(w[1][o + 1]) + round_bf16(w[2][o + 1]);
        kc.v[j] = pack_bf16(ce, co);
        kcl.v[j] = pack_bf16(ce - round_bf16(ce), co - round_bf16(co));
        kt0.v[j] = pack_bf16(w[0][e], w[0][o]); kt2.v[j] = pack_bf16(w[0][e + 2], w[0][o + 2]);
        kh0.v[j] = pack_bf16(w[1][e], w[1][o]); kh2.v[j] = pack_bf16(w[1][e + 2], w[1][o + 2]);
        kw0.v[j] = pack_bf16(w[2][e], w[2][o]); kw2.v[j] = pack_bf16(w[2][e + 2], w[2][o + 2]);
      }
    }
    float2 scale[4], shift[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { scale[j] = make_float2(1.f, 1.f); shift[j] = make_float2(0.f, 0.f); }
    if (MODE == MODE_EVAL) {                                    // BN affine from the running statistics, in registers
      const int cb8 = c0 + vec * 8;
      float gm[8], bt[8], rm[8], rv[8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        gm[4 * q] = bnraw[0][q].x; gm[4 * q + 1] = bnraw[0][q].y; gm[4 * q + 2] = bnraw[0][q].z; gm[4 * q + 3] = bnraw[0][q].w;
        bt[4 * q] = bnraw[1][q].x; bt[4 * q + 1] = bnraw[1][q].y; bt[4 * q + 2] = bnraw[1][q].z; bt[4 * q + 3] = bnraw[1][q].w;
        rm[4 * q] = bnraw[2][q].x; rm[4 * q + 1] = bnraw[2][q].y; rm[4 * q + 2] = bnraw[2][q].z; rm[4 * q + 3] = bnraw[2][q].w;
        rv[4 * q] = bnraw[3][q].x; rv[4 * q + 1] = bnraw[3][q].y; rv[4 * q + 2] = bnraw[3][q].z; rv[4 * q + 3] = bnraw[3][q].w;
      }
      float sc[8], sh[8], rs[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        rs[q] = rsqrtf(rv[q] + a.eps);
        sc[q] = gm[q] * rs[q];
        sh[q] = bt[q] - rm[q] * sc[q];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { scale[j] = make_float2(sc[2 * j], sc[2 * j + 1]); shift[j] = make_float2(sh[2 * j], sh[2 * j + 1]); }
      if (rest == 0 && a.save_mean && tid < g.G) {              // pixel 0's threads cover every channel vector once
#pragma unroll
        for (int q = 0; q < 8; ++q) { a.save_mean[cb8 + q] = rm[q]; a.save_rstd[cb8 + q] = rs[q]; }
      }
    }
    auto load_affine = [&]() {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        scale[j] = *reinterpret_cast<const float2*>(s_scale + vec * 8 + 2 * j);
        shift[j] = *reinterpret_cast<const float2*>(s_shift + vec * 8 + 2 * j);
      }
    };
    const size_t frame_elems = (size_t)g.H * g.W * a.y_pix;
    const size_t ypixA = ((size_t)(h0 + hA) * g.W + w0 + wA) * a.y_pix + c0 + vec * 8;
    const size_t ypixB = ((size_t)(h0 + hB) * g.W + w0 + wB) * a.y_pix + c0 + vec * 8;

    const uint32_t full0 = smem_u32(full), slots0 = smem_u32(slots);

    float2 sum[4], sq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { sum[j] = make_float2(0.f, 0.f); sq[j] = make_float2(0.f, 0.f); }

    int cm = 0;                                                 // slot group of the current clip
    uint32_t par = 0;                                           // its full-barrier parity
    int left = KK;                                              // clips of the stream not yet finished
    if (stamps && tid == 0) stamps[1] = gtimer();               // prologue done
    wait_phase(full0, 0);
    if (stamps && tid == 0) stamps[2] = gtimer();               // first frame landed

    // Everything a frame step reads from shared memory is loaded ONE STEP AHEAD: step t computes frame t from
    // registers (centres of t-1, t, t+1 and the H/W neighbours of t) while the loads of slot t+1 (centre of t+1, used
    // by this step's last tap, and the neighbours of t+1, used by the next step) are in flight, and the barrier of
    // slot t+2 is TESTED (non-blocking) so that the next step branches on a predicate that is already there.  Neither
    // the barrier round trip nor the load latency is on the critical path of a step.
    struct Nb { P8 hm, hp, wm, wp; };                            // rowpair: A uses hm/wm/wp, B uses hp/wm/wp
    auto load_nb = [&](uint32_t cA, uint32_t cB, Nb& nA, Nb& nB) {
      nA.hm = lds_p8(cA - rowb); nA.wm = lds_p8(cA - pixb); nA.wp = lds_p8(cA + pixb);
      nB.hp = lds_p8(cB + rowb); nB.wm = lds_p8(cB - pixb); nB.wp = lds_p8(cB + pixb);
      if (!RP) { nA.hp = lds_p8(cA + rowb); nB.hm = lds_p8(cB - rowb); }
    };
    P8 xcA = lds_p8(slots0 + offA), xcB = lds_p8(slots0 + offB), xmA, xmB;   // centre of the first frame
    Nb nA, nB;
    load_nb(slots0 + offA, slots0 + offB, nA, nB);
#pragma unroll
    for (int j = 0; j < 4; ++j) { xmA.v[j] = 0u; xmB.v[j] = 0u; }
    uint32_t ready = test_phase(full0 + 8u, 0);                 // slot 1 (T >= 4: same slot group, same parity)

    // one sweep over the CTA's clips; `stats` is a compile-time tag so that the statistics accumulators (sweep 0)
    // and the BatchNorm affine (sweep 1) never hold registers at the same time
    auto run_sweep = [&](auto stats_tag) {
      constexpr bool stats = decltype(stats_tag)::value;
#pragma unroll 1
      for (int kclip = 0; kclip < nclips; ++kclip) {
        const uint32_t cb = slots0 + (uint32_t)cm * (uint32_t)(T * kSlotB);
        const uint32_t fullc = full0 + (uint32_t)cm * (uint32_t)(T * 8);
        int cmn = cm + 1;
        uint32_t parn = par;
        if (cmn == M) { cmn = 0; parn ^= 1u; }
        const uint32_t cbn = slots0 + (uint32_t)cmn * (uint32_t)(T * kSlotB);
        const uint32_t fulln = full0 + (uint32_t)cmn * (uint32_t)(T * 8);
        const bool has_next = --left > 0;
        const uint32_t aC = cb + offA, bC = cb + offB;          // centres in slot 0 of this clip; frame t adds t*kSlotB
        const uint32_t aN = cbn + offA, bN = cbn + offB;        // ... and of the next clip
        __nv_bfloat16* ypA = a.y + (size_t)(p + kclip * g.P) * T * frame_elems + ypixA;
        __nv_bfloat16* ypB = a.y + (size_t)(p + kclip * g.P) * T * frame_elems + ypixB;
        static_for(std::make_integer_sequence<int, T>{}, [&](auto tc) {
          constexpr int t = decltype(tc)::value;
          // (a) slot of frame t+1: its barrier was tested a step ago; load its centre and neighbours
          P8 xnA, xnB;
          Nb mA = nA, mB = nB;
#pragma unroll
          for (int j = 0; j < 4; ++j) { xnA.v[j] = 0u; xnB.v[j] = 0u; }
          if (t + 1 < T) {
            if (!ready) wait_phase(fullc + 8u * (t + 1), par);
            xnA = lds_p8(aC + (uint32_t)((t + 1) * kSlotB));
            xnB = lds_p8(bC + (uint32_t)((t + 1) * kSlotB));
            load_nb(aC + (uint32_t)((t + 1) * kSlotB), bC + (uint32_t)((t + 1) * kSlotB), mA, mB);
          } else if (has_next) {                                // the next clip's first frame
            if (!ready) wait_phase(fulln, parn);
            xnA = lds_p8(aN);
            xnB = lds_p8(bN);
            load_nb(aN, bN, mA, mB);
          }
          // (b) test the barrier of frame t+2's slot for the next step
          if (t + 2 < T) ready = test_phase(fullc + 8u * (t + 2), par);
          else if (has_next) ready = test_phase(fulln + 8u * (t + 2 - T), parn);
          // (c) frame t from registers
          const P8& hpA = RP ? xcB : nA.hp;
          const P8& hmB = RP ? xcA : nB.hm;
          float2 zA[4], zB[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) { fh2_first(zA[j], xcA.v[j], kc.v[j]); fh2_first(zB[j], xcB.v[j], kc.v[j]); }
          fh8(zA, xcA, kcl); fh8(zB, xcB, kcl);
          fh8(zA, nA.hm, kh0); fh8(zB, hmB, kh0);
          fh8(zA, hpA, kh2); fh8(zB, nB.hp, kh2);
          fh8(zA, nA.wm, kw0); fh8(zB, nB.wm, kw0);
          fh8(zA, nA.wp, kw2); fh8(zB, nB.wp, kw2);
          if (t > 0) { fh8(zA, xmA, kt0); fh8(zB, xmB, kt0); }  // zero padding in T: the clip's first / last frame
          if (t + 1 < T) { fh8(zA, xnA, kt2); fh8(zB, xnB, kt2); }
          // (d) frame t's slot was last read a step ago (its values have just been consumed): release it
          __syncwarp();
          if (lane == 0) arrive_u32(fullc + 8u * kRing + 8u * t);
          if (stats) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (actA) { sum[j] = __fadd2_rn(sum[j], zA[j]); sq[j] = __ffma2_rn(zA[j], zA[j], sq[j]); }
              if (actB) { sum[j] = __fadd2_rn(sum[j], zB[j]); sq[j] = __ffma2_rn(zB[j], zB[j], sq[j]); }
            }
          } else {
            if (MODE != MODE_PLAIN) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 uA = __ffma2_rn(zA[j], scale[j], shift[j]), uB = __ffma2_rn(zB[j], scale[j], shift[j]);
                float2 sA, sB;
                sA.x = __saturatef(fmaf(uA.x, 1.f / 6.f, 0.5f));
                sA.y = __saturatef(fmaf(uA.y, 1.f / 6.f, 0.5f));
                sB.x = __saturatef(fmaf(uB.x, 1.f / 6.f, 0.5f));
                sB.y = __saturatef(fmaf(uB.y, 1.f / 6.f, 0.5f));
                zA[j] = __fmul2_rn(uA, sA);
                zB[j] = __fmul2_rn(uB, sB);
              }
            }
            uint4 oA, oB;
            oA.x = pack_bf16(zA[0].x, zA[0].y); oA.y = pack_bf16(zA[1].x, zA[1].y);
            oA.z = pack_bf16(zA[2].x, zA[2].y); oA.w = pack_bf16(zA[3].x, zA[3].y);
            oB.x = pack_bf16(zB[0].x, zB[0].y); oB.y = pack_bf16(zB[1].x, zB[1].y);
            oB.z = pack_bf16(zB[2].x, zB[2].y); oB.w = pack_bf16(zB[3].x, zB[3].y);
            if (actA) *reinterpret_cast<uint4*>(ypA + (size_t)t * frame_elems) = oA;
            if (actB) *reinterpret_cast<uint4*>(ypB + (size_t)t * frame_elems) = oB;
          }
          xmA = xcA; xcA = xnA; nA = mA;
          xmB = xcB; xcB = xnB; nB = mB;
        });
        cm = cmn;
        par = parn;
      }
    };
    if (MODE == MODE_TRAIN) {
      run_sweep(std::true_type{});
      if (stamps && tid == 0) stamps[3] = gtimer();             // statistics sweep done
      // ---- CTA partial sums -> one row per CTA -> grid barrier -> batch statistics of this channel group
      float acc[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] = sum[j].x; acc[2 * j + 1] = sum[j].y;
        acc[8 + 2 * j] = sq[j].x; acc[8 + 2 * j + 1] = sq[j].y;
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        float v = acc[q];
        for (int o = 16; o >= g.G; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[q] = v;
      }
      if (lane < g.G) {                                         // lane < G: vec == lane (32 % G == 0)
#pragma unroll
        for (int q = 0; q < 16; ++q) s_red[(warp * g.G + lane) * 16 + q] = acc[q];
      }
      consumer_bar_sync(nconsumer);
      if (stamps && tid == 0) stamps[6] = gtimer();             // every warp of this CTA has finished the statistics sweep
      // Grid-wide exchange without a flag: every partial travels as an 8-byte {value, epoch} word (single-copy
      // atomic), readers poll the words themselves until they carry this launch's epoch.  One store propagation plus
      // one load round trip after the last CTA arrives -- no fence, no atomic counter, no second round of loads.
      const int per = 2 * g.Cg, rows = ntiles * g.P;               // partial rows of this channel group
      if (tid < per) {                                          // row layout [vec][sum 0..7 | sumsq 0..7]
        float v = 0.f;
        for (int wv = 0; wv < g.cwarps; ++wv) v += s_red[wv * per + tid];
        st_relaxed_v2(a.partials + (size_t)blockIdx.x * per + tid, make_uint2(__float_as_uint(v), epoch));
      }
      float bn_g = 1.f, bn_b = 0.f, old_rm = 0.f, old_rv = 0.f;  // gamma / beta / running statistics travel with the wait
      if (tid < g.Cg) {
        bn_g = a.gamma[c0 + tid]; bn_b = a.beta[c0 + tid];
        if (rest == 0 && a.running_mean) { old_rm = a.running_mean[c0 + tid]; old_rv = a.running_var[c0 + tid]; }
      }
      const int parts = nconsumer / per;
      {
        const int k = tid % per, part = tid / per;
        if (part < parts) {
          double accd = 0.0;
          const long long t0 = clock64();
          constexpr int kBatch = 16;                            // rows polled per round trip (wide groups have parts = 1)
          for (int r0 = part; r0 < rows; r0 += kBatch * parts) {
            uint2 v[kBatch];
            bool ok;
            do {
              ok = true;
#pragma unroll
              for (int u = 0; u < kBatch; ++u) {
                const int r = r0 + u * parts;
                v[u] = make_uint2(0u, epoch);
                if (r < rows) v[u] = ld_relaxed_v2(a.partials + ((size_t)r * g.ngroups + cg) * per + k);
                ok = ok && v[u].y == epoch;
              }
              if (!ok && clock64() - t0 > 4000000000LL) __trap();  // a CTA that never arrives must not hang the GPU
            } while (!ok);
#pragma unroll
            for (int u = 0; u < kBatch; ++u) accd += (double)__uint_as_float(v[u].x);
          }
          s_dpart[part * per + k] = accd;
        }
      }
      if (stamps && tid == 0) stamps[7] = gtimer();             // this thread's partial rows have arrived
      consumer_bar_sync(nconsumer);
      if (tid < g.Cg) {
        const int c = c0 + tid;
        const int i1 = (tid / 8) * 16 + (tid % 8), i2 = i1 + 8;
        double s1 = 0.0, s2 = 0.0;
        for (int q = 0; q < parts; ++q) { s1 += s_dpart[q * per + i1]; s2 += s_dpart[q * per + i2]; }
        const double m = (double)g.N * T * g.H * g.W;
        const double mu = s1 / m;
        double var = s2 / m - mu * mu;
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
            a.running_mean[c] = (1.f - a.momentum) * old_rm + a.momentum * mean;
            a.running_var[c] = (1.f - a.momentum) * old_rv + a.momentum * (float)unb;
          }
        }
      }
      // every CTA of this channel group has published with `epoch`, i.e. has read the word: the next launch (or the
      // next replay of a captured graph) gets a different tag
      if (rest == 0 && tid == 0) a.epochs[cg] = epoch & 0xffu;
      if (tid == 0) mbar_arrive(gate);                            // the producer may now request the rest of sweep 1
      consumer_bar_sync(nconsumer);
      load_affine();
      if (stamps && tid == 0) stamps[4] = gtimer();             // grid barrier passed, batch statistics ready
    }
    run_sweep(std::false_type{});
    if (stamps && tid == 0) stamps[5] = gtimer();               // last frame stored
  }
}

size_t sweep_smem(const SwGeo& g) {
  return 384 + (size_t)kRing * kSlotB + kTailB + 64;
}

void finish_geo(const mvfb_mvf_desc* d, SwGeo& g) {
  g.N = d->N; g.T = d->T; g.Cs = d->Cs; g.H = d->H; g.W = d->W;
  g.G = g.Cg / 8; g.ngroups = d->Cs / g.Cg;
  g.Hs = d->H / g.hsplit; g.Ws = d->W / g.wsplit; g.Hp = g.Hs + 2; g.Wp = g.Ws + 2;
  g.pixels = g.Hs * g.Ws; g.pixhalf = (g.pixels + 1) / 2;
  if (g.rowpair) {
    g.cwarps = g.Hs / 2;
    g.cthreads = 32 * g.cwarps;
  } else {
    g.cthreads = g.pixhalf * g.G;
    g.cwarps = (g.cthreads + 31) / 32;
  }
  const int lanes = g.ngroups * g.hsplit * g.wsplit;
  int P = num_sms() / lanes;
  if (P < 1) P = 1;
  if (P > d->N) P = d->N;
  g.P = P;
}

bool choose_sweep(const mvfb_mvf_desc* d, SwGeo& g) {
  if (d->dtype != MVFB_BF16 || d->layout != MVFB_NHWC) return false;
  if (d->T != 4 && d->T != 8 && d->T != 16) return false;
  if (d->Cs % 8 != 0 || d->C % 8 != 0) return false;
  // 1) row-pair tiles: an even number of rows (one warp per pair, <= kCWarps) and a row of (column, vector) lanes
  //    that fits a warp.  Wide channel groups first: the TMA unit retires ~1 box row per 1.5 clocks whatever the
  //    row's size, so 64-byte rows (32 channels) halve its load per element against 32-byte rows (measured: a
  //    14x14x16-channel frame takes 0.205 us of TMA time, as long as its arithmetic).  Within a width take the
  //    tiling that loads the fewest halo pixels and idles the fewest lanes.
  const int cands_rp[2] = {32, 16};
  for (int ci = 0; ci < 2; ++ci) {
    const int Cg = cands_rp[ci];
    if (d->Cs % Cg) continue;
    double best = 1e30;
    SwGeo bg;
    for (int hsplit = 1; hsplit <= d->H; ++hsplit) {
      if (d->H % hsplit) continue;
      const int Hs = d->H / hsplit;
      if (Hs % 2 || Hs / 2 > kCWarps || 2 * Cg > 32 * (Hs / 2)) continue;
      for (int wsplit = 1; wsplit <= d->W; ++wsplit) {
        if (d->W % wsplit) continue;
        const int Ws = d->W / wsplit;
        if (Ws * (Cg / 8) > 32 || Ws < 4) continue;
        if ((Hs + 2) * (Ws + 2) * Cg * 2 > kSlotB) continue;
        const double halo = (double)((Hs + (hsplit > 1 ? 2 : 0)) * (Ws + (wsplit > 1 ? 2 : 0))) / (Hs * Ws);
        const double idle = 32.0 / (Ws * (Cg / 8)) * ((double)kCWarps / (Hs / 2) > 2.0 ? 1.5 : 1.0);
        const double cost = halo * idle;
        if (cost < best) {
          best = cost;
          bg.Cg = Cg; bg.hsplit = hsplit; bg.wsplit = wsplit; bg.rowpair = 1;
        }
      }
    }
    if (best < 1.8) {
      g = bg;
      finish_geo(d, g);
      if (sweep_smem(g) <= (size_t)kSmemLimit) return true;
    }
  }
  // 2) linear items (odd extents, wide channel groups): H tiles only
  if (d->W + 2 > 256) return false;
  const int splits[4] = {1, 2, 4, 7};
  const int cands[4] = {64, 32, 16, 8};
  for (int pass = 0; pass < 2; ++pass) {                     // first pass: >= 32-byte rows per pixel only
    for (int si = 0; si < 4; ++si) {
      const int hsplit = splits[si];
      if (d->H % hsplit) continue;
      const int Hs = d->H / hsplit;
      if (Hs + 2 > 256) continue;
      for (int ci = 0; ci < 4; ++ci) {
        const int Cg = cands[ci];
        if ((pass == 0) != (Cg >= 16)) continue;
        if (d->Cs % Cg) continue;
        const int pixels = Hs * d->W;
        if (pixels * (Cg / 8) > kMaxItems) continue;
        if (((pixels + 1) / 2) * (Cg / 8) > 32 * kCWarps) continue;
        if ((Hs + 2) * (d->W + 2) * Cg * 2 > kSlotB) continue;
        g.Cg = Cg; g.hsplit = hsplit; g.wsplit = 1; g.rowpair = 0;
        finish_geo(d, g);
        if (2 * Cg > 32 * g.cwarps) continue;                // the statistics reduction needs 2*Cg consumer threads
        if (sweep_smem(g) > (size_t)kSmemLimit) continue;
        return true;
      }
    }
  }
  return false;
}

// Call number of a train-mode launch (24 bits used): tags its partial sums so that words left in the workspace by any
// earlier call -- another geometry whose rows overlap this one's, recycled allocator blocks, uninitialised memory
// (probability 2^-32 per word) -- are never mistaken for this launch's.
unsigned int next_epoch() {
  static std::atomic<unsigned int> e{0x5eed01u};
  return e.fetch_add(1u, std::memory_order_relaxed) & 0xffffffu;
}

template <int MODE, int TT, bool RP>
int launch_mode(const CUtensorMap& tmx, SwArgs& a, cudaStream_t st) {
  static DevOnce once;
  if (once.pending()) {
    MVFB_CUDA(cudaFuncSetAttribute(mvf_sweep_kernel<MODE, TT, RP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    once.done();
  }
  const SwGeo& g = a.g;
  const dim3 grid(g.ngroups * g.hsplit * g.wsplit * g.P), block(32 * (g.cwarps + 1));
  const size_t smem = sweep_smem(g);
  if (MODE == MODE_TRAIN) {
    a.epoch_hi = next_epoch();
    void* params[2] = {(void*)&tmx, (void*)&a};
    // cooperative launch: the driver guarantees that all CTAs are resident, which the grid exchange relies on
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)mvf_sweep_kernel<MODE, TT, RP>, grid, block, params, smem, st);
    if (e == cudaErrorCooperativeLaunchTooLarge) {
      (void)cudaGetLastError();
      return MVFB_ERR_UNSUPPORTED;                           // device shared with another context: two-launch path
    }
    MVFB_CUDA(e);
  } else {
    mvf_sweep_kernel<MODE, TT, RP><<<grid, block, smem, st>>>(tmx, a);
  }
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

template <int TT, bool RP>
int launch_sweep(const mvfb_mvf_desc* d, const CUtensorMap& tmx, SwArgs& a, cudaStream_t st) {
  if (!d->use_hs) return launch_mode<MODE_PLAIN, TT, RP>(tmx, a, st);
  if (d->training) return launch_mode<MODE_TRAIN, TT, RP>(tmx, a, st);
  return launch_mode<MODE_EVAL, TT, RP>(tmx, a, st);
}

template <bool RP>
int launch_T(const mvfb_mvf_desc* d, const CUtensorMap& tmx, SwArgs& a, cudaStream_t st) {
  switch (a.g.T) {
    case 4: return launch_sweep<4, RP>(d, tmx, a, st);
    case 8: return launch_sweep<8, RP>(d, tmx, a, st);
    default: return launch_sweep<16, RP>(d, tmx, a, st);
  }
}

}  // namespace

static size_t sweep_partial_bytes(const SwGeo& g) {
  return ((size_t)g.ngroups * g.hsplit * g.wsplit * g.P * 2 * g.Cg * sizeof(uint2) + 255) / 256 * 256;
}

bool mvf_sweep_supported(const mvfb_mvf_desc* d) {
  SwGeo g;
  return choose_sweep(d, g);
}

size_t mvf_sweep_ws(const mvfb_mvf_desc* d) {
  SwGeo g;
  if (!choose_sweep(d, g)) return 0;
  return sweep_partial_bytes(g) + (size_t)g.ngroups * sizeof(unsigned int) + 256;
}

int mvf_sweep_fwd(const mvfb_mvf_desc* d, const void* x, void* y, long long y_stride, const float* wt,
                  const float* wh, const float* ww, const float* gamma, const float* beta, float* rm, float* rv,
                  float* save_mean, float* save_rstd, void* ws, size_t ws_bytes, cudaStream_t st) {
  SwGeo g;
  if (!choose_sweep(d, g)) return MVFB_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) || y_stride % 8 != 0)
    return MVFB_ERR_UNSUPPORTED;
  const void* vec_loaded[7] = {wt, wh, ww, gamma, beta, rm, rv};          // read as float4 by the kernel
  for (const void* q : vec_loaded)
    if (reinterpret_cast<uintptr_t>(q) & 15) return MVFB_ERR_UNSUPPORTED;
  CUtensorMap tmx;
  const uint64_t dims[4] = {(uint64_t)g.Cs, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N * g.T};
  const uint64_t strides[3] = {(uint64_t)d->C * 2, (uint64_t)g.W * d->C * 2, (uint64_t)g.H * g.W * d->C * 2};
  const uint32_t box[4] = {(uint32_t)g.Cg, (uint32_t)g.Wp, (uint32_t)g.Hp, 1u};
  int rc = encode_tmap(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, nullptr,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (rc) return rc;
  SwArgs a;
  a.g = g;
  a.eps = d->eps; a.momentum = d->momentum;
  a.wt = wt; a.wh = d->mode != MVFB_MODE_T ? wh : nullptr; a.ww = d->mode == MVFB_MODE_THW ? ww : nullptr;
  a.gamma = gamma; a.beta = beta; a.running_mean = rm; a.running_var = rv;
  a.save_mean = save_mean; a.save_rstd = save_rstd;
  a.partials = (uint2*)ws;
  a.epoch_hi = 0;
  a.epochs = reinterpret_cast<unsigned int*>((char*)ws + sweep_partial_bytes(g));
  a.pre_frames = 4;   // measured best of 1..16; must be >= 1: the last step of sweep 0 reads the first frame of sweep 1
  a.y = (__nv_bfloat16*)y; a.y_pix = y_stride;
  // tools/stream_timeline.py (mvf_b200_set_option(MVFB_OPT_SWEEP_DEBUG, 1)): %globaltimer stamps behind the partials
  const bool debug = option(OPT_SWEEP_DEBUG) != 0;
  if (debug && ws_bytes < (size_t)(1 << 20)) {
    set_error("MVFB_OPT_SWEEP_DEBUG needs a workspace of >= 1 MiB, got %zu bytes", ws_bytes);
    return MVFB_ERR_WORKSPACE;
  }
  a.stamps = debug ? reinterpret_cast<unsigned long long*>((char*)ws + (512 << 10)) : nullptr;
  return g.rowpair ? launch_T<true>(d, tmx, a, st) : launch_T<false>(d, tmx, a, st);
}

}  // namespace mvfb

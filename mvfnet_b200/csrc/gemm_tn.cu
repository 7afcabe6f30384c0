// 1x1 convolution on NHWC activations == "TN" GEMM on the 5th-generation tensor cores:
//
//     D[m, n] = sum_k A[m, k] * B[n, k]        A: (M, K) pixels x input channels  (K contiguous, bf16)
//                                               B: (N, K) output x input channels  (K contiguous, bf16)
//                                               D: (M, N) pixels x output channels (N contiguous, bf16)
//
// Replaces nn.Conv2d(k=1, bias=False) of Bottleneck.conv1 / conv3 / downsample (backbones/resnet.py:157-180,
// 299-303) and, called with (dY, W^T), its input-gradient.  The A operand may come from TWO tensors split
// along K: columns [0, K0) from A0 (the compact MVF slab written by mvf_fwd) and [K0, K) from A1 (the
// untouched channels of x), which is how MVF.forward's `torch.cat` + `.contiguous()` (MVF.py:135-137) are
// eliminated: the concatenation happens in TMEM, never in HBM.
//
// Kernel: persistent, one CTA per SM, warp-specialised:
//   warp 0   TMA producer   cp.async.bulk.tensor.2d, 128-byte swizzle, kStages-deep full/empty mbarrier ring
//   warp 1   MMA issuer     one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//                           fp32 accumulators in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i
//                           overlaps the main loop of tile i+1; tcgen05.commit releases smem slots / signals TMEM
//   warps 2-9 epilogue      tcgen05.ld 32x32b -> bf16 -> 128B-swizzled staging tile in smem -> per-column
//                           (sum, sum of squares) of the ROUNDED outputs for the following train-mode BatchNorm
//                           (fp32 atomicAdd per column and row half per tile) -> TMA store (clips the M tail).
//                           A warp reads the TMEM lanes of its quarter (warp % 4); the two warps of a quarter split
//                           the tile's columns.  Eight warps because the memory-bound layers (K = 64 .. 256) are paced
//                           by the epilogue: a 128 x 256 tile is ~500 instructions per thread (TMEM -> bf16 -> smem,
//                           2 FMAs per element for the statistics) against ~3400 clocks of HBM time per tile.
#include <cuda_bf16.h>

#include "common.cuh"
#include "mvf_internal.cuh"
#include "ptx.cuh"

namespace mvfb {

namespace {

constexpr int BM = 128;            // tile rows (UMMA M)
constexpr int BK = 64;             // K per stage: 64 bf16 = 128 B = one swizzle atom
constexpr int kEpiWarps = 8;       // two epilogue warps per TMEM lane quarter: each takes half of the tile's columns
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kThreads = 64 + kEpiThreads;   // + TMA producer warp + MMA issuer warp

struct GemmArgs {
  long long M;
  int N, K, K0;                    // K0: first K0 columns of A come from tensor map A0, the rest from A1
  int m_tiles, n_tiles;
  float* colsum;                   // [N] or null
  float* colsq;                    // [N] or null
  const __nv_bfloat16* res;        // optional (M, N) bf16 addend, leading dimension ldr: out = A B^T + res
  long long ldr;
  int res_col0;                    // the addend applies to columns >= res_col0 only (a multiple of 32)
  // inference epilogue (eval-mode BatchNorm folded into the convolution): out = [relu]((A B^T) * ep_scale[n] + ep_shift[n] [+ res])
  const float* ep_scale;           // [N] or null
  const float* ep_shift;           // [N] (with ep_scale)
  int ep_relu;
  // 3x3 convolution as an implicit GEMM (IM2COL kernels): K = 9 * Cin ordered (r, s, c); A tiles are gathered by
  // TMA im2col loads from the NHWC input, M = F * Ho * Wo output pixels
  int Cin, Ho, Wo, stride, ks, pad;   // ks = filter width: 3 (pad 1) or 1 (pad 0); the stride-2 dgrad's windows: 1 or 2 (pad 0)
  // scattered store (stride-2 input gradient): tile row m = (f, i, j) of the (Ho, Wo) grid goes to pixel (2i, 2j) of the
  // (2Ho, 2Wo) image `scat` points into (the parity's (ph, pw) offset is already in the pointer); null = TMA store to tmD
  __nv_bfloat16* scat;
  int scat_fill;                   // 1: also zero the three other pixels of the row's 2 x 2 cell
};

template <int BN, bool RES = false, int CL = 1>
struct Cfg {
  // RES (addend tile prefetched by TMA into two buffers of its own): fewer stages make room for them.  CL = 2 (CTA pair):
  // a CTA holds half of every B tile, which buys a fourth stage at BN = 256
  static constexpr int kStages = RES ? (BN >= 128 ? 3 : 4) : (BN >= 256 ? (CL == 2 ? 4 : 3) : (BN >= 128 ? 5 : 6));
  static constexpr int kABytes = BM * BK * 2;                 // 16 KB
  static constexpr int kBBytes = BN * BK * 2 / CL;            // this CTA's share of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kPanels = BN / 64;                     // 64-column (128 B) panels of the output tile
  static constexpr int kStagingBytes = BM * BN * 2;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN; // power of two >= 32: BN in {64,128,256} -> 128,256,512
  static constexpr int kStatParts = kEpiThreads / (BN / 2);   // row parts of the statistics pass (BN = 256: 2, 128: 4, 64: 8)
  static constexpr int kStatBytes = kStatParts * BN * 2 * 4;  // [part][sum | sumsq][BN] fp32 = 4 KB
  static constexpr size_t kSmem = 1024 /*align slack*/ + (size_t)kStages * kStageBytes + kStagingBytes + 256 + kStatBytes +
                                  (RES ? 2 * kStagingBytes : 0);
  static_assert(kSmem <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// RES: the addend (a.res) arrives by TMA (map tmR: the output's geometry over the addend) into two shared-memory buffers,
// requested two tiles ahead by the epilogue -- its latency never meets the epilogue.  (The first version copied the tile
// with __ldg inside the epilogue: on the K <= 128 layers, where the epilogue IS the critical path, the fused add cost
// as much as the separate add kernel it replaced -- profiles/r02_step_by_shape.txt.)  RES = false with a.res set is
// that older in-epilogue copy, kept for the im2col kernels (inference only).
// CL = 2: CTA pairs (clusters of two, tcgen05 cta_group::2).  The pair computes a 256-row tile of ONE N tile: each CTA loads
// its own 128 rows of A and HALF of the B tile, the leader (cluster rank 0) issues M = 256 MMAs over both CTAs' shared
// memory, each CTA's epilogue drains its own 128 accumulator rows.  A CTA then pulls A + B/2 instead of A + B through its
// L2 -> SM port, which is what paces the K >= 256 layers (profiles/r02_conv_halo.txt, GEMM ablations).  Barriers: all TMA
// loads of both CTAs complete on the LEADER's full[s]; the leader's MMA commits arrive on empty[s] / tmem_full[acc] of
// BOTH CTAs; both CTAs' epilogue warps arrive on the LEADER's tmem_empty[acc].
template <int BN, bool IM2COL, bool RES, int CL = 1>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
               const __grid_constant__ CUtensorMap tmR, const GemmArgs a) {
  using C = Cfg<BN, RES, CL>;
  static_assert(CL == 1 || (CL == 2 && !RES), "the CTA-pair mode has no addend path");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;                                        // kStages x [A tile | B tile]
  uint8_t* staging = smem + (size_t)C::kStages * C::kStageBytes;     // BM x BN bf16, 64-column swizzled panels
  uint8_t* addend = staging + C::kStagingBytes;                      // RES: two more tiles of the same layout
  uint8_t* tail = addend + (RES ? 2 * C::kStagingBytes : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full = bars;                      // [kStages]
  uint64_t* empty = bars + C::kStages;        // [kStages]
  uint64_t* tmem_full = empty + C::kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint64_t* res_full = tmem_empty + 2;        // [2] (RES)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 2);
  float* s_stat = reinterpret_cast<float*>(tail + 256);             // [kStatParts][2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = a.K / BK;
  const int total_tiles = a.m_tiles * a.n_tiles;
  // tile loop of this CTA: plain = tiles blockIdx.x, +gridDim.x, ...; pair = PAIR tiles (two M tiles, one N tile), the
  // CTA's own M tile being 2 * pair + rank
  const int crank = CL == 2 ? (int)cluster_ctarank() : 0;
  const int t_first = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int t_limit = CL == 2 ? ((a.m_tiles + 1) / 2) * a.n_tiles : total_tiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], (kEpiThreads / 32) * CL);     // pair: the leader's barrier counts both CTAs' epilogue warps
      mbar_init(&res_full[s], 1);
    }
    if (RES) tma_prefetch_desc(&tmR);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CL == 2) {
      tmem_alloc_pair(tmem_slot, C::kTmemCols);
    } else {
      tmem_alloc(tmem_slot, C::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();             // the peer's barriers exist before anything is sent to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = t_first; tile < t_limit; tile += t_step) {
        const int mt = (CL == 2 ? 2 : 1) * (tile / a.n_tiles) + crank, nt = tile % a.n_tiles;
        // IM2COL: first output pixel of the tile -> base input pixel of its filter window (pad 1)
        int bw = 0, bh = 0, bn = 0, cblocks = 1;
        if (IM2COL) {
          const long long m0 = (long long)mt * BM;
          const int q = (int)(m0 % a.Wo), pq = (int)(m0 / a.Wo);
          bw = q * a.stride - a.pad;
          bh = (pq % a.Ho) * a.stride - a.pad;
          bn = pq / a.Ho;
          cblocks = a.Cin / BK;
        }
        int rs = 0, cb = 0;                                    // filter tap (r*3+s) and channel block of this k-block
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = stage_base + (size_t)s * C::kStageBytes;
          const int k = kb * BK;
          if (CL == 2) {
            // both CTAs' boxes complete on the leader's barrier, which expects the bytes of both
            if (crank == 0) mbar_arrive_expect_tx(&full[s], 2 * C::kStageBytes);
            const uint32_t lead = leader_addr(&full[s]);
            if (IM2COL) {
              tma_load_im2col_4d_pair(sa, &tmA1, lead, cb * BK, bw, bh, bn, (uint16_t)(rs % a.ks), (uint16_t)(rs / a.ks));
              if (++cb == cblocks) { cb = 0; ++rs; }
            } else {
              tma_load_2d_pair(sa, k < a.K0 ? &tmA0 : &tmA1, lead, k, mt * BM);
            }
            tma_load_2d_pair(sa + C::kABytes, &tmB, lead, k, nt * BN + crank * (BN / 2));   // this CTA's half of the B tile
          } else {
            mbar_arrive_expect_tx(&full[s], C::kStageBytes);
            if (IM2COL) {
              tma_load_im2col_4d(sa, &tmA1, &full[s], cb * BK, bw, bh, bn, (uint16_t)(rs % a.ks), (uint16_t)(rs / a.ks));
              if (++cb == cblocks) { cb = 0; ++rs; }
            } else {
              tma_load_2d(sa, k < a.K0 ? &tmA0 : &tmA1, &full[s], k, mt * BM);
            }
            tma_load_2d(sa + C::kABytes, &tmB, &full[s], k, nt * BN);
          }
          if (++s == C::kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (all lanes run the loop, one elected lane issues: ptx.cuh) =====================
    if (CL == 1 || crank == 0) {                               // pair: the leader issues for both CTAs
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CL, BN, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = t_first; tile < t_limit; tile += t_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);               // epilogue(s) have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)s * C::kStageBytes);
          // one descriptor per operand tile; a K step of 16 elements = 32 bytes = 2 units of the start-address field
          const uint64_t adesc = umma_smem_desc_sw128(sa, 0, 1024), bdesc = umma_smem_desc_sw128(sa + C::kABytes, 0, 1024);
          if (CL == 2) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) umma_f16_elect_pair(d_tmem, adesc + 2u * kk, bdesc + 2u * kk, idesc, (kb | kk) != 0);
            umma_commit_elect_pair(&empty[s]);                 // both CTAs' producers may refill this stage
          } else {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) umma_f16_elect(d_tmem, adesc + 2u * kk, bdesc + 2u * kk, idesc, (kb | kk) != 0);
            umma_commit_elect(&empty[s]);                      // slot free once these MMAs have read it
          }
          if (++s == C::kStages) { s = 0; ph ^= 1; }
        }
        if (CL == 2) umma_commit_elect_pair(&tmem_full[acc]);  // accumulator complete, in both CTAs
        else umma_commit_elect(&tmem_full[acc]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int et = threadIdx.x - 64;                           // 0..255
    const int lane_grp = warp & 3;                             // TMEM lanes [32*lane_grp, +32) belong to this warp
    const int half = (warp - 2) >> 2;                          // which half of the tile's columns this warp converts
    const int row = lane_grp * 32 + lane;                      // output row within the tile
    constexpr int kHalfN = BN / 2;
    auto request_addend = [&](int tile_r, int buf) {           // one thread: the addend tile of `tile_r` -> buffer `buf`
      const int mr = tile_r / a.n_tiles, nr = tile_r - mr * a.n_tiles;
      mbar_arrive_expect_tx(&res_full[buf], C::kStagingBytes);
#pragma unroll
      for (int p = 0; p < C::kPanels; ++p)
        tma_load_2d(addend + (size_t)buf * C::kStagingBytes + (size_t)p * (BM * 128), &tmR, &res_full[buf], nr * BN + p * 64, mr * BM);
    };
    if (RES && et == 0) {
      if (blockIdx.x < total_tiles) request_addend(blockIdx.x, 0);
      if (blockIdx.x + gridDim.x < total_tiles) request_addend(blockIdx.x + gridDim.x, 1);
    }
    int it = 0;
    for (int tile = t_first; tile < t_limit; tile += t_step, ++it) {
      const int mt = (CL == 2 ? 2 : 1) * (tile / a.n_tiles) + crank, nt = tile % a.n_tiles;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const uint8_t* add_tile = RES ? addend + (size_t)acc * C::kStagingBytes : staging;
      // Two phases where the tile has at least two 64-column panels and no statistics are wanted: convert the first half
      // of the columns, hand its panels to the TMA engine, convert the second half while they are being read.  With one
      // phase the store of tile i and the conversion of tile i+1 take turns on the single staging buffer (163 -> 148 us,
      // 243 -> 228 us on the wide-N / small-K input-gradient GEMMs); WITH statistics the one-phase order is the faster one,
      // because there the statistics pass is what overlaps the store (profiles/r02_conv_halo.txt, GEMM ablations).
      const bool two_phase = C::kPanels >= 2 && !a.scat && !a.colsum && !(a.res && !RES);
      const int nphase = two_phase ? 2 : 1;
      for (int phase = 0; phase < nphase; ++phase) {
        // the panels this phase writes must have been read by the TMA engine (two-phase: one younger store group may still
        // be pending -- the other half's)
        if (et == 0) { if (two_phase) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
        named_bar_sync(1, kEpiThreads);
        if (phase == 0) {
          // optional addend (out = A B^T + res): its 128 x BN tile is copied into the staging buffer FIRST, coalesced (a
          // warp reads whole 128- .. 512-byte rows), in the swizzled layout the results will have -- while this tile's MMAs
          // are still running.  Each thread later adds its own row chunks from shared memory (the first version read the
          // addend row-wise from global memory per thread: 32 cache lines per load instruction, 3.5 ms per step).
          if (!RES && a.res) {
            constexpr int kChunksPerRow = BN / 8;                    // 16-byte chunks per tile row
            const long long m0 = (long long)mt * BM;
#pragma unroll 4
            for (int idx = et; idx < BM * kChunksPerRow; idx += kEpiThreads) {
              const int r = idx / kChunksPerRow, ch = idx - r * kChunksPerRow;
              uint4 v4 = make_uint4(0u, 0u, 0u, 0u);
              if (m0 + r < a.M) v4 = __ldg(reinterpret_cast<const uint4*>(a.res + (m0 + r) * a.ldr + nt * BN + ch * 8));
              *reinterpret_cast<uint4*>(staging + (size_t)(ch >> 3) * (BM * 128) + (size_t)r * 128 + (((ch & 7) ^ (r & 7)) << 4)) = v4;
            }
          }
          mbar_wait(&tmem_full[acc], acc_ph);
          tc_fence_after();
          if (!RES && a.res) named_bar_sync(1, kEpiThreads);       // every addend chunk is in place
          if (RES) mbar_wait(&res_full[acc], acc_ph);              // this tile's addend has landed (requested two tiles ago)
        }
        // columns of this warp in this phase: one phase = its half of the tile; two phases = a quarter each
        const int colbase = two_phase ? phase * kHalfN + half * (kHalfN / 2) : half * kHalfN;
        const int ncols = two_phase ? kHalfN / 2 : kHalfN;
        const uint32_t taddr = tmem_base + (uint32_t)(acc * BN + colbase) + ((uint32_t)(lane_grp * 32) << 16);
#pragma unroll 1
        for (int cc = 0; cc < ncols; cc += 32) {
          const int c0 = colbase + cc;                           // first column of this chunk within the tile
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + cc, v);
          tmem_ld_wait();
          if (a.ep_scale) {                                      // every lane reads the same 32 floats: L1 broadcast
            const float4* sc = reinterpret_cast<const float4*>(a.ep_scale + nt * BN + c0);
            const float4* sh = reinterpret_cast<const float4*>(a.ep_shift + nt * BN + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 s4 = __ldg(sc + q), h4 = __ldg(sh + q);
              v[4 * q + 0] = __float_as_uint(fmaf(__uint_as_float(v[4 * q + 0]), s4.x, h4.x));
              v[4 * q + 1] = __float_as_uint(fmaf(__uint_as_float(v[4 * q + 1]), s4.y, h4.y));
              v[4 * q + 2] = __float_as_uint(fmaf(__uint_as_float(v[4 * q + 2]), s4.z, h4.z));
              v[4 * q + 3] = __float_as_uint(fmaf(__uint_as_float(v[4 * q + 3]), s4.w, h4.w));
            }
          }
          uint8_t* panel = staging + (size_t)(c0 >> 6) * (BM * 128) + (size_t)row * 128;
          const uint8_t* apanel = add_tile + (size_t)(c0 >> 6) * (BM * 128) + (size_t)row * 128;
          const int chunk0 = (c0 & 63) >> 3;
          if (a.res && nt * BN + c0 >= a.res_col0) {             // this thread's own row chunks of the staged addend
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 r4 = *reinterpret_cast<const uint4*>(apanel + (((chunk0 + q) ^ (row & 7)) << 4));
              const uint32_t w4[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[q * 8 + 2 * e] = __float_as_uint(__uint_as_float(v[q * 8 + 2 * e]) + bf16_lo(w4[e]));
                v[q * 8 + 2 * e + 1] = __float_as_uint(__uint_as_float(v[q * 8 + 2 * e + 1]) + bf16_hi(w4[e]));
              }
            }
          }
          if (a.ep_relu) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(fmaxf(__uint_as_float(v[q]), 0.f));
          }
          // 32 fp32 -> 32 bf16 = 64 B = four 16-byte chunks of this row in panel c0/64
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
            o.y = pack_bf16(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
            o.z = pack_bf16(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
            o.w = pack_bf16(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
            *reinterpret_cast<uint4*>(panel + (((chunk0 + q) ^ (row & 7)) << 4)) = o;
          }
        }
        if (two_phase) {
          if (phase == 1) {                                      // TMEM accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CL == 2) mbar_arrive_cluster(leader_addr(&tmem_empty[acc])); else mbar_arrive(&tmem_empty[acc]); }
          }
          fence_proxy_async_smem();                              // staging writes -> visible to the TMA engine
          named_bar_sync(1, kEpiThreads);
          if (et == 0) {
#pragma unroll
            for (int p = 0; p < C::kPanels / 2; ++p) {
              const int pp = phase * (C::kPanels / 2) + p;
              tma_store_2d(&tmD, staging + (size_t)pp * (BM * 128), nt * BN + pp * 64, mt * BM);
            }
            tma_store_commit();
            // every epilogue thread has read this tile's addend (the barrier above): its buffer takes the tile after next
            if (RES && phase == 1 && tile + 2 * (int)gridDim.x < total_tiles) request_addend(tile + 2 * (int)gridDim.x, acc);
          }
        }
      }                                                        // phases
      if (!two_phase) {
        if (a.scat && et < BM) {                                 // destination of tile row `et` (M < 2^31: checked by the host)
          const unsigned m = (unsigned)mt * BM + et;
          long long off = -1;
          if (m < (unsigned)a.M) {
            const unsigned j = m % (unsigned)a.Wo, pq = m / (unsigned)a.Wo;
            const unsigned i = pq % (unsigned)a.Ho, f = pq / (unsigned)a.Ho;
            off = (((long long)f * (2 * a.Ho) + 2 * i) * (2 * a.Wo) + 2 * j) * (long long)a.N;
          }
          reinterpret_cast<long long*>(s_stat)[et] = off;
        }
        // TMEM accumulator fully read: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CL == 2) mbar_arrive_cluster(leader_addr(&tmem_empty[acc])); else mbar_arrive(&tmem_empty[acc]); }
        fence_proxy_async_smem();                                // staging writes -> visible to the TMA engine
        named_bar_sync(1, kEpiThreads);
        if (a.scat) {
          // every output row is a pixel of its own in the double-resolution image: 16-byte chunks, a warp covers whole rows
          // (row offsets were computed once per tile, before the barrier above: s_rowoff).  scat_fill: the row owns its
          // whole 2 x 2 cell and zeroes the other three pixels (stride-2 1x1 input gradient: dx needs no memset).
          constexpr int kChunksPerRow = BN / 8;
          const long long* s_rowoff = reinterpret_cast<const long long*>(s_stat);
          const long long right = a.N, down = 2ll * a.Wo * a.N;      // one pixel to the right / one image row down, in elements
#pragma unroll 4
          for (int idx = et; idx < BM * kChunksPerRow; idx += kEpiThreads) {
            const int r = idx / kChunksPerRow, ch = idx - r * kChunksPerRow;
            const long long off = s_rowoff[r];
            if (off >= 0) {
              const uint4 v4 = *reinterpret_cast<const uint4*>(staging + (size_t)(ch >> 3) * (BM * 128) + (size_t)r * 128 +
                                                               (((ch & 7) ^ (r & 7)) << 4));
              __nv_bfloat16* dst = a.scat + off + nt * BN + ch * 8;
              *reinterpret_cast<uint4*>(dst) = v4;
              if (a.scat_fill) {
                const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(dst + right) = z4;
                *reinterpret_cast<uint4*>(dst + down) = z4;
                *reinterpret_cast<uint4*>(dst + down + right) = z4;
              }
            }
          }
        } else if (et == 0) {
#pragma unroll
          for (int p = 0; p < C::kPanels; ++p) tma_store_2d(&tmD, staging + (size_t)p * (BM * 128), nt * BN + p * 64, mt * BM);
          tma_store_commit();
          // every epilogue thread has read this tile's addend (the barrier above): its buffer takes the tile after next
          if (RES && tile + 2 * (int)gridDim.x < total_tiles) request_addend(tile + 2 * (int)gridDim.x, acc);
        }
      }                                                        // !two_phase
      // per-column statistics of the rounded tile (rows beyond M were zero-filled by TMA: they add nothing).  A work
      // item is (column pair, row part): one 32-bit shared-memory load yields two columns of a row, consecutive threads
      // take consecutive column pairs (conflict-free whatever the swizzle: the 32 words of a warp lie in one 128-byte row)
      if (a.colsum) {
        constexpr int kPairs = BN / 2;                          // column pairs per tile
        constexpr int kParts = C::kStatParts;
        constexpr int kRows = BM / kParts;
        {
          const int pair = et % kPairs, part = et / kPairs;     // kPairs * kParts == kEpiThreads
          const int col = 2 * pair;
          const uint8_t* panel = staging + (size_t)(col >> 6) * (BM * 128);
          const int chunk = (col & 63) >> 3, within = (col & 7) * 2;
          float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll 8
          for (int rr = 0; rr < kRows; ++rr) {
            const int r = part * kRows + rr;
            const uint32_t h2 = *reinterpret_cast<const uint32_t*>(panel + r * 128 + ((chunk ^ (r & 7)) << 4) + within);
            bf16x2_sum_sq(h2, s1a, s1b, s2a, s2b);
          }
          float* mine = s_stat + (size_t)part * (2 * BN);
          *reinterpret_cast<float2*>(mine + col) = make_float2(s1a, s1b);
          *reinterpret_cast<float2*>(mine + BN + col) = make_float2(s2a, s2b);
        }
        // one atomic per column and statistic per tile: the partial sums of the row parts meet in shared memory first
        // (atomics on one address serialise in L2; with a part per atomic the K = 64 layers were bound by them)
        named_bar_sync(1, kEpiThreads);
        for (int i = et; i < 2 * BN; i += kEpiThreads) {
          float v = 0.f;
#pragma unroll
          for (int q = 0; q < kParts; ++q) v += s_stat[(size_t)q * (2 * BN) + i];
          atomicAdd(i < BN ? &a.colsum[nt * BN + i] : &a.colsq[nt * BN + i - BN], v);
        }
      }
    }
    if (et == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();             // the peer may still be reading this CTA's shared memory / TMEM
  if (warp == 2) {
    tc_fence_after();
    if (CL == 2) tmem_dealloc_pair(tmem_base, C::kTmemCols);
    else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

int make_2d_map(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_cols,
                uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {ld_elems * 2};
  const uint32_t box[2] = {box_cols, box_rows};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, nullptr,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

// The CTA-pair mode pays where a CTA's L2 -> SM traffic is the bound: 256-wide N tiles with at least eight K blocks (measured:
// 4-11 % faster from K = 512 on, 10-15 % SLOWER at K = 256, where the pair's extra synchronisation meets a short K loop --
// profiles/r02_conv_halo.txt) and enough tile pairs to fill the GPU.  The caller must encode tmB with a HALF-tile box.
inline bool pair_pays(long long M, int N, int K, int bn, bool im2col) {
  if (option(OPT_GEMM_PAIR_OFF) || bn != 256 || K < 8 * BK) return false;
  const long long m_tiles = (M + BM - 1) / BM;
  if (im2col && (m_tiles & 1)) return false;                   // a tile past the last pixel would gather out-of-range windows
  return ((m_tiles + 1) / 2) * (N / bn) >= num_sms() / 2;
}

template <int BN, bool IM2COL, bool RES = false, int CL = 1>
int launch_kernel(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmB, const CUtensorMap& tmD,
                  const CUtensorMap& tmR, GemmArgs a, cudaStream_t st) {
  using C = Cfg<BN, RES, CL>;
  a.m_tiles = (int)((a.M + BM - 1) / BM);
  a.n_tiles = a.N / BN;
  static DevOnce once;
  if (once.pending()) {
    MVFB_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN, IM2COL, RES, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
    once.done();
  }
  if (CL == 2) {
    int pairs = ((a.m_tiles + 1) / 2) * a.n_tiles;
    if (pairs > num_sms() / 2) pairs = num_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MVFB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tn_kernel<BN, IM2COL, RES, CL>, tmA0, tmA1, tmB, tmD, tmR, a));
    count_launch();
    return MVFB_OK;
  }
  int grid = a.m_tiles * a.n_tiles;
  if (grid > num_sms()) grid = num_sms();
  gemm_tn_kernel<BN, IM2COL, RES, CL><<<grid, kThreads, C::kSmem, st>>>(tmA0, tmA1, tmB, tmD, tmR, a);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

struct Epi {                       // optional inference epilogue
  const float* scale = nullptr;
  const float* shift = nullptr;
  int relu = 0;
  int res_col0 = 0;                // first column the addend applies to
};

template <int BN, bool RES = false>
int launch(const mvfb_gemm_desc* d, const void* a0, const void* a1, const void* b, const void* res, long long ldr,
           void* out, float* colsum, float* colsq, const Epi& ep, cudaStream_t st) {
  CUtensorMap tmA0, tmA1, tmB, tmD, tmR;
  int rc;
  // A1 is addressed with the GEMM's own k coordinate, so its map spans columns [0, K) of the source rows
  if ((rc = make_2d_map(&tmA1, a1, (uint64_t)d->K, (uint64_t)d->M, (uint64_t)d->lda1, BK, BM))) return rc;
  if (d->K0 > 0) {
    if ((rc = make_2d_map(&tmA0, a0, (uint64_t)d->K0, (uint64_t)d->M, (uint64_t)d->lda0, BK, BM))) return rc;
  } else {
    tmA0 = tmA1;
  }
  const bool pair = !RES && !res && pair_pays(d->M, d->N, d->K, BN, false);
  if ((rc = make_2d_map(&tmB, b, (uint64_t)d->K, (uint64_t)d->N, (uint64_t)d->ldb, BK, pair ? BN / 2 : BN))) return rc;
  if ((rc = make_2d_map(&tmD, out, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldd, 64, BM))) return rc;
  tmR = tmD;
  if (RES && (rc = make_2d_map(&tmR, res, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)ldr, 64, BM))) return rc;
  GemmArgs a;
  a.M = d->M; a.N = d->N; a.K = d->K; a.K0 = d->K0;
  a.colsum = colsum; a.colsq = colsq;
  a.res = (const __nv_bfloat16*)res; a.ldr = ldr; a.res_col0 = ep.res_col0;
  a.ep_scale = ep.scale; a.ep_shift = ep.shift; a.ep_relu = ep.relu;
  a.Cin = a.Ho = a.Wo = a.stride = a.pad = 0;
  a.ks = 1;
  a.scat = nullptr; a.scat_fill = 0;
  if constexpr (BN == 256 && !RES) {
    if (pair) return launch_kernel<BN, false, false, 2>(tmA0, tmA1, tmB, tmD, tmR, a, st);
  }
  return launch_kernel<BN, false, RES>(tmA0, tmA1, tmB, tmD, tmR, a, st);
}

// 3x3 / pad 1 convolution: A gathered by TMA im2col from x (F, H, W, Cin); B = weights (Cout, 3, 3, Cin)
template <int BN>
int launch_conv3x3(const mvfb_conv_desc* d, const void* x, const void* w, void* out, float* colsum, float* colsq,
                   const void* res, const Epi& ep, cudaStream_t st) {
  const int Ho = (d->H - 1) / d->stride + 1, Wo = (d->W - 1) / d->stride + 1;
  CUtensorMap tmA, tmB, tmD;
  const uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->F};
  const uint64_t strides[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
  const int pad = d->ksize == 3 ? 1 : 0, taps = d->ksize * d->ksize;
  const int lower[2] = {-pad, -pad}, upper[2] = {-pad, -pad};  // upper corner = pad - (kernel - 1)
  const uint32_t estr[4] = {1, (uint32_t)d->stride, (uint32_t)d->stride, 1};
  int rc = encode_tmap_im2col(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, lower, upper, BK, BM, estr,
                              CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const long long M = (long long)d->F * Ho * Wo;
  const bool pair = !res && pair_pays(M, d->Cout, taps * d->Cin, BN, true);
  if ((rc = make_2d_map(&tmB, w, (uint64_t)taps * d->Cin, (uint64_t)d->Cout, (uint64_t)taps * d->Cin, BK, pair ? BN / 2 : BN))) return rc;
  if ((rc = make_2d_map(&tmD, out, (uint64_t)d->Cout, (uint64_t)M, (uint64_t)d->Cout, 64, BM))) return rc;
  GemmArgs a;
  a.M = M; a.N = d->Cout; a.K = taps * d->Cin; a.K0 = 0;
  a.colsum = colsum; a.colsq = colsq;
  a.res = (const __nv_bfloat16*)res; a.ldr = d->Cout; a.res_col0 = 0;
  a.ep_scale = ep.scale; a.ep_shift = ep.shift; a.ep_relu = ep.relu;
  a.Cin = d->Cin; a.Ho = Ho; a.Wo = Wo; a.stride = d->stride; a.ks = d->ksize; a.pad = pad;
  a.scat = nullptr; a.scat_fill = 0;
  if constexpr (BN == 256) {
    if (pair) return launch_kernel<BN, true, false, 2>(tmA, tmA, tmB, tmD, tmD, a, st);
  }
  return launch_kernel<BN, true>(tmA, tmA, tmB, tmD, tmD, a, st);
}

// ---- stride-2 3x3 input gradient (backbones/resnet.py:163-170 with stride 2: conv2 of the first block of layer2/3/4).
// dx[h, w, c] = sum_{r, s, n} g[i, j, n] w[n, r, s, c] over 2i + r - 1 = h, 2j + s - 1 = w.  Split by the parity of
// (h, w) = (2a + ph, 2b + pw): ph = 0 takes filter row r = 1 from g row a; ph = 1 takes r = 2 from row a and r = 0 from
// row a + 1 (the same along w).  Each parity is therefore a STRIDE-1 convolution of g with a (1|2) x (1|2) window, no
// padding on the low side, zero fill beyond the last row / column -- 1 + 2 + 2 + 4 = 9 taps in total, no wasted MAC --
// and runs on the implicit-GEMM kernel (TMA im2col on g, M = F*Ho*Wo, K = taps*Cout, N = Cin); the epilogue scatters its
// rows to the pixels of that parity in dx (a first version wrote compact quarters and interleaved them in a second pass:
// 3x the dx traffic, 1.8 ms per step against cuDNN's 1.2).
template <int BN>
int launch_window(const void* g, int F, int Ho, int Wo, int Cout, int Cin, int kh, int kw, const void* wq, void* out,
                  cudaStream_t st) {
  CUtensorMap tmA, tmB, tmD;
  const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)F};
  const uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)Wo * Cout * 2, (uint64_t)Ho * Wo * Cout * 2};
  const int lower[2] = {0, 0}, upper[2] = {0, 0};              // one base pixel per output pixel; offsets reach past the edge
  const uint32_t estr[4] = {1, 1, 1, 1};
  int rc = encode_tmap_im2col(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, g, dims, strides, lower, upper, BK, BM, estr,
                              CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const long long M = (long long)F * Ho * Wo;
  const int taps = kh * kw;
  if ((rc = make_2d_map(&tmB, wq, (uint64_t)taps * Cout, (uint64_t)Cin, (uint64_t)taps * Cout, BK, BN))) return rc;
  tmD = tmB;                                                   // unused: the epilogue scatters the rows itself
  GemmArgs a;
  a.scat = (__nv_bfloat16*)out; a.scat_fill = 0;
  a.M = M; a.N = Cin; a.K = taps * Cout; a.K0 = 0;
  a.colsum = nullptr; a.colsq = nullptr;
  a.res = nullptr; a.ldr = 0; a.res_col0 = 0;
  a.ep_scale = nullptr; a.ep_shift = nullptr; a.ep_relu = 0;
  a.Cin = Cout; a.Ho = Ho; a.Wo = Wo; a.stride = 1; a.ks = kw; a.pad = 0;
  return launch_kernel<BN, true>(tmA, tmA, tmB, tmD, tmD, a, st);
}

}  // namespace

}  // namespace mvfb

using namespace mvfb;

// ---- stride-2 1x1 input gradient (the down-sampling convolution, make_res_layer resnet.py:299-303): the plain GEMM
// dY W on the compact gradient whose rows land on the even pixels of dx; the epilogue zeroes the rest of each 2 x 2 cell.
template <int BN>
static int launch_s2_1x1(const mvfb_conv_desc* d, const void* g, const void* wT, void* dx, cudaStream_t st) {
  const int Ho = d->H / 2, Wo = d->W / 2;
  const long long M = (long long)d->F * Ho * Wo;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_2d_map(&tmA, g, (uint64_t)d->Cout, (uint64_t)M, (uint64_t)d->Cout, BK, BM))) return rc;
  if ((rc = make_2d_map(&tmB, wT, (uint64_t)d->Cout, (uint64_t)d->Cin, (uint64_t)d->Cout, BK, BN))) return rc;
  GemmArgs a;
  a.M = M; a.N = d->Cin; a.K = d->Cout; a.K0 = 0;
  a.colsum = nullptr; a.colsq = nullptr;
  a.res = nullptr; a.ldr = 0; a.res_col0 = 0;
  a.ep_scale = nullptr; a.ep_shift = nullptr; a.ep_relu = 0;
  a.Cin = 0; a.Ho = Ho; a.Wo = Wo; a.stride = 1; a.ks = 1; a.pad = 0;
  a.scat = (__nv_bfloat16*)dx; a.scat_fill = 1;
  return launch_kernel<BN, false>(tmA, tmA, tmB, tmB, tmB, a, st);
}

extern "C" int conv1x1s2_dgrad(const mvfb_conv_desc* d, const void* g, const void* wT, void* dx, mvfb_stream_t stream) {
  MVFB_CHECK(d && g && wT && dx, MVFB_ERR_ARG, "null descriptor / operand");
  MVFB_CHECK(d->F > 0 && d->H > 0 && d->W > 0 && d->stride == 2 && d->ksize == 1 && d->H % 2 == 0 && d->W % 2 == 0,
             MVFB_ERR_UNSUPPORTED, "conv1x1s2_dgrad takes the 1x1 stride-2 layers with even H, W (F=%d H=%d W=%d)", d->F, d->H, d->W);
  MVFB_CHECK(d->Cin % 64 == 0 && d->Cout % BK == 0, MVFB_ERR_UNSUPPORTED, "Cin=%d and Cout=%d must be multiples of 64",
             d->Cin, d->Cout);
  MVFB_CHECK(!((uintptr_t)g & 15) && !((uintptr_t)wT & 15) && !((uintptr_t)dx & 15), MVFB_ERR_UNSUPPORTED,
             "operands must be 16-byte aligned");
  MVFB_CHECK((long long)d->F * d->H * d->W < (1ll << 31), MVFB_ERR_UNSUPPORTED, "F*H*W must stay below 2^31 pixels");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->Cin % 256 == 0) return launch_s2_1x1<256>(d, g, wT, dx, st);
  if (d->Cin % 128 == 0) return launch_s2_1x1<128>(d, g, wT, dx, st);
  return launch_s2_1x1<64>(d, g, wT, dx, st);
}

extern "C" int conv3x3s2_dgrad(const mvfb_conv_desc* d, const void* g, const void* wq, void* dx, mvfb_stream_t stream) {
  MVFB_CHECK(d && g && wq && dx, MVFB_ERR_ARG, "null descriptor / operand");
  MVFB_CHECK(d->F > 0 && d->H > 0 && d->W > 0 && d->stride == 2 && d->ksize == 3 && d->H % 2 == 0 && d->W % 2 == 0,
             MVFB_ERR_UNSUPPORTED, "conv3x3s2_dgrad takes the 3x3 stride-2 layers with even H, W (F=%d H=%d W=%d)", d->F, d->H, d->W);
  MVFB_CHECK(d->Cin % 64 == 0 && d->Cout % BK == 0, MVFB_ERR_UNSUPPORTED, "Cin=%d and Cout=%d must be multiples of 64",
             d->Cin, d->Cout);
  MVFB_CHECK(!((uintptr_t)g & 15) && !((uintptr_t)wq & 15) && !((uintptr_t)dx & 15), MVFB_ERR_UNSUPPORTED,
             "operands must be 16-byte aligned");
  MVFB_CHECK((long long)d->F * d->H * d->W < (1ll << 31), MVFB_ERR_UNSUPPORTED, "F*H*W must stay below 2^31 pixels");
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = d->H / 2, Wo = d->W / 2;
  const __nv_bfloat16* w = (const __nv_bfloat16*)wq;
  __nv_bfloat16* out = (__nv_bfloat16*)dx;
  const int khs[4] = {1, 1, 2, 2}, kws[4] = {1, 2, 1, 2};       // parity (ph, pw) = (p / 2, p % 2): window kh x kw
  for (int p = 0; p < 4; ++p) {
    __nv_bfloat16* o = out + ((size_t)(p >> 1) * d->W + (p & 1)) * d->Cin;   // pixel (ph, pw) of every 2 x 2 cell
    int rc;
    if (d->Cin % 256 == 0) rc = launch_window<256>(g, d->F, Ho, Wo, d->Cout, d->Cin, khs[p], kws[p], w, o, st);
    else if (d->Cin % 128 == 0) rc = launch_window<128>(g, d->F, Ho, Wo, d->Cout, d->Cin, khs[p], kws[p], w, o, st);
    else rc = launch_window<64>(g, d->F, Ho, Wo, d->Cout, d->Cin, khs[p], kws[p], w, o, st);
    if (rc) return rc;
    w += (size_t)khs[p] * kws[p] * d->Cout * d->Cin;
  }
  return MVFB_OK;
}

static int conv1x1_gemm_impl(const mvfb_gemm_desc* d, const void* a0, const void* a1, const void* b, const void* res,
                             long long ldr, void* out, float* colsum, float* colsq, const Epi& ep, mvfb_stream_t stream) {
  MVFB_CHECK(d && a1 && b && out, MVFB_ERR_ARG, "null descriptor / operand");
  MVFB_CHECK(d->M > 0 && d->N > 0 && d->K > 0, MVFB_ERR_ARG, "bad GEMM shape M=%lld N=%d K=%d", d->M, d->N, d->K);
  MVFB_CHECK(d->K % BK == 0 && d->K0 % BK == 0 && d->K0 >= 0 && d->K0 < d->K, MVFB_ERR_UNSUPPORTED,
             "K=%d and K0=%d must be multiples of %d with K0 < K", d->K, d->K0, BK);
  MVFB_CHECK(d->N % 64 == 0, MVFB_ERR_UNSUPPORTED, "N=%d must be a multiple of 64", d->N);
  MVFB_CHECK(d->K0 == 0 || a0, MVFB_ERR_ARG, "K0 > 0 needs the A0 operand");
  MVFB_CHECK((colsum == nullptr) == (colsq == nullptr), MVFB_ERR_ARG, "colsum and colsq go together");
  MVFB_CHECK(d->lda1 % 8 == 0 && d->ldb % 8 == 0 && d->ldd % 8 == 0 && (d->K0 == 0 || d->lda0 % 8 == 0), MVFB_ERR_UNSUPPORTED,
             "leading dimensions must be multiples of 8 elements (16 bytes)");
  MVFB_CHECK(!((uintptr_t)a1 & 15) && !((uintptr_t)b & 15) && !((uintptr_t)out & 15) && !((uintptr_t)a0 & 15),
             MVFB_ERR_UNSUPPORTED, "operands must be 16-byte aligned");
  MVFB_CHECK(!res || (!((uintptr_t)res & 15) && ldr % 8 == 0 && ldr >= d->N), MVFB_ERR_UNSUPPORTED,
             "the addend must be 16-byte aligned with a leading dimension that is a multiple of 8 and >= N");
  cudaStream_t st = (cudaStream_t)stream;
  MVFB_CHECK(!ep.scale || (ep.shift && !((uintptr_t)ep.scale & 15) && !((uintptr_t)ep.shift & 15)), MVFB_ERR_ARG,
             "the epilogue scale needs a shift, both 16-byte aligned");
  // addend by TMA (two extra tiles of shared memory, BN <= 128, 3-4 stages) where the epilogue paces the kernel; with a
  // long K loop the epilogue hides behind the MMAs anyway and the wider BN = 256 tile wins (layer4: K = 512)
  if (res && (d->K <= 256 || d->N % 256 != 0)) {
    if (d->N % 128 == 0) return launch<128, true>(d, a0, a1, b, res, ldr, out, colsum, colsq, ep, st);
    return launch<64, true>(d, a0, a1, b, res, ldr, out, colsum, colsq, ep, st);
  }
  if (d->N % 256 == 0) return launch<256>(d, a0, a1, b, res, ldr, out, colsum, colsq, ep, st);
  if (d->N % 128 == 0) return launch<128>(d, a0, a1, b, res, ldr, out, colsum, colsq, ep, st);
  return launch<64>(d, a0, a1, b, res, ldr, out, colsum, colsq, ep, st);
}

extern "C" int conv1x1_gemm(const mvfb_gemm_desc* d, const void* a0, const void* a1, const void* b, void* out,
                            float* colsum, float* colsq, mvfb_stream_t stream) {
  return conv1x1_gemm_impl(d, a0, a1, b, nullptr, 0, out, colsum, colsq, Epi{}, stream);
}

extern "C" int conv1x1_gemm_add(const mvfb_gemm_desc* d, const void* a0, const void* a1, const void* b, const void* res,
                                long long ldr, void* out, mvfb_stream_t stream) {
  MVFB_CHECK(res != nullptr, MVFB_ERR_ARG, "conv1x1_gemm_add needs the addend");
  return conv1x1_gemm_impl(d, a0, a1, b, res, ldr, out, nullptr, nullptr, Epi{}, stream);
}

extern "C" int conv1x1_gemm_add_cols(const mvfb_gemm_desc* d, const void* a0, const void* a1, const void* b, const void* res,
                                     long long ldr, int first_col, void* out, mvfb_stream_t stream) {
  MVFB_CHECK(res != nullptr, MVFB_ERR_ARG, "conv1x1_gemm_add_cols needs the addend");
  MVFB_CHECK(d && first_col >= 0 && first_col % 32 == 0 && first_col < d->N, MVFB_ERR_ARG,
             "first_col=%d must be a multiple of 32 inside [0, N)", first_col);
  Epi ep;
  ep.res_col0 = first_col;
  return conv1x1_gemm_impl(d, a0, a1, b, res, ldr, out, nullptr, nullptr, ep, stream);
}

extern "C" int conv1x1_gemm_bnact(const mvfb_gemm_desc* d, const void* a0, const void* a1, const void* b, const float* scale,
                                  const float* shift, const void* res, long long ldr, int relu, void* out,
                                  mvfb_stream_t stream) {
  MVFB_CHECK(scale && shift, MVFB_ERR_ARG, "conv1x1_gemm_bnact needs scale and shift");
  Epi ep;
  ep.scale = scale; ep.shift = shift; ep.relu = relu;
  return conv1x1_gemm_impl(d, a0, a1, b, res, ldr, out, nullptr, nullptr, ep, stream);
}

static int conv3x3_gemm_impl(const mvfb_conv_desc* d, const void* x, const void* w, void* out, float* colsum,
                             float* colsq, const void* res, const Epi& ep, mvfb_stream_t stream) {
  MVFB_CHECK(d && x && w && out, MVFB_ERR_ARG, "null descriptor / operand");
  MVFB_CHECK(d->F > 0 && d->H > 0 && d->W > 0 && (d->stride == 1 || d->stride == 2) && (d->ksize == 3 || d->ksize == 1),
             MVFB_ERR_ARG, "bad conv shape F=%d H=%d W=%d stride=%d ksize=%d", d->F, d->H, d->W, d->stride, d->ksize);
  MVFB_CHECK(d->Cin % BK == 0 && d->Cout % 64 == 0, MVFB_ERR_UNSUPPORTED, "Cin=%d and Cout=%d must be multiples of 64",
             d->Cin, d->Cout);
  MVFB_CHECK((colsum == nullptr) == (colsq == nullptr), MVFB_ERR_ARG, "colsum and colsq go together");
  MVFB_CHECK(!((uintptr_t)x & 15) && !((uintptr_t)w & 15) && !((uintptr_t)out & 15) && !((uintptr_t)res & 15),
             MVFB_ERR_UNSUPPORTED, "operands must be 16-byte aligned");
  MVFB_CHECK(!ep.scale || (ep.shift && !((uintptr_t)ep.scale & 15) && !((uintptr_t)ep.shift & 15)), MVFB_ERR_ARG,
             "the epilogue scale needs a shift, both 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  // the 56x56 / 28x28 layers: one halo band per tile instead of nine im2col gathers (conv_halo.cu)
  if (!res && !ep.scale && conv_halo_eligible(d)) return conv_halo(d, x, w, out, colsum, colsq, st);
  if (d->Cout % 256 == 0) return launch_conv3x3<256>(d, x, w, out, colsum, colsq, res, ep, st);
  if (d->Cout % 128 == 0) return launch_conv3x3<128>(d, x, w, out, colsum, colsq, res, ep, st);
  return launch_conv3x3<64>(d, x, w, out, colsum, colsq, res, ep, st);
}

extern "C" int conv3x3_gemm(const mvfb_conv_desc* d, const void* x, const void* w, void* out, float* colsum,
                            float* colsq, mvfb_stream_t stream) {
  return conv3x3_gemm_impl(d, x, w, out, colsum, colsq, nullptr, Epi{}, stream);
}

extern "C" int conv3x3_gemm_bnact(const mvfb_conv_desc* d, const void* x, const void* w, const float* scale,
                                  const float* shift, const void* res, int relu, void* out, mvfb_stream_t stream) {
  MVFB_CHECK(scale && shift, MVFB_ERR_ARG, "conv3x3_gemm_bnact needs scale and shift");
  Epi ep;
  ep.scale = scale; ep.shift = shift; ep.relu = relu;
  return conv3x3_gemm_impl(d, x, w, out, nullptr, nullptr, res, ep, stream);
}

// 3x3 / stride 1 / pad 1 convolution as an implicit GEMM whose A operand is ONE halo band in shared memory.
//
// gemm_tn.cu's im2col kernel gathers the A operand once per filter tap: nine TMA im2col loads of 128 pixels x 64 channels
// per tile and channel block.  Here a tile is R = 128 / (W + 2) whole image rows, addressed as CONSECUTIVE positions of the
// zero-padded frame (q = hp*(W+2) + wp, starting at the row's first real pixel), and the R + 2 padded rows it touches are
// loaded ONCE by a tiled 4-D TMA box (64 channels, W+2, R+2 rows, 1 frame; the padding is TMA's out-of-bounds zero fill).
// The box lands as consecutive 128-byte rows in the SWIZZLE_128B pattern, and tap (r, s) of the filter is the same band
// read from row r*(W+2) + s on: a K-major UMMA descriptor may start at ANY 128-byte row of such a tile, the hardware
// swizzle is a function of the absolute shared-memory address (tools/ubench/umma_shift.cu).  Per tile and channel block
// the kernel moves 4 x 58 = 232 rows (56 x 56) instead of 9 x 128.  Positions in the padding columns (and MMA rows past
// the tile's R rows) produce garbage; the epilogue zeroes them (so the BatchNorm sums stay exact) and skips them when it
// stores.  Weights: (Cout, 3, 3, Cin) bf16; with Cin = Cout = 64 they stay resident in shared memory.
#include <cuda_bf16.h>

#include "common.cuh"
#include "mvf_internal.cuh"
#include "ptx.cuh"

namespace mvfb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kMaxBands = 4;        // deeper rings (3 bands + 14 weight stages, 2 + 16) measured the same: r02_conv_halo.txt
constexpr int kMaxBStages = 8;
constexpr int kBarBytes = 512;      // barriers + the TMEM address slot
static_assert((2 * kMaxBands + 2 * kMaxBStages + 4) * 8 + 4 <= kBarBytes, "barrier area");

struct HaloArgs {
  int F, H, W, Cin, Cout;
  int Wp, R;                       // padded width W + 2; image rows per tile (R * Wp <= 130); a band holds R + 2 padded rows
  int band_bytes;                  // (2 * Wp + 2 + 128) * 128 (the furthest row any tap's 128-row view touches), rounded up to 1024
  int tiles_per_frame, m_tiles, n_tiles, cblocks;
  int bands, bstages;              // ring depths
  int bres;                        // weights resident in shared memory (n_tiles == 1)
  float* colsum;
  float* colsq;
  __nv_bfloat16* out;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// CL = 2: a CTA pair (cluster of two, tcgen05 cta_group::2) works on two consecutive tiles of the same output-channel
// tile: each CTA stages its own band and HALF of every weight tile, the leader (cluster rank 0) issues M = 256 MMAs over
// both CTAs' shared memory (gemm_tn.cu's pair mode).  For the streamed-weight case (C = 128) that halves the weight
// bytes each SM pulls per tile -- 288 of the 334 KB a tile needs.
template <int BN, int CL>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB, const HaloArgs a) {
  constexpr int kBBytes = BN * BK * 2 / CL;                    // this CTA's part of one (tap, channel block) weight tile
  constexpr int kStagingBytes = BM * BN * 2;
  constexpr int kPanels = BN / 64;
  constexpr int kTmemCols = 2 * BN;
  constexpr int kStatParts = kEpiThreads / (BN / 2);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* band_base = smem;
  uint8_t* b_base = band_base + (size_t)a.bands * a.band_bytes;
  const int b_tiles = a.bres ? 9 * a.cblocks : a.bstages;
  uint8_t* staging = b_base + (size_t)b_tiles * kBBytes;
  uint8_t* tail = staging + kStagingBytes;
  uint64_t* band_full = reinterpret_cast<uint64_t*>(tail);     // [kMaxBands]
  uint64_t* band_empty = band_full + kMaxBands;                // [kMaxBands]
  uint64_t* b_full = band_empty + kMaxBands;                   // [kMaxBStages] (slot 0 doubles as "weights resident")
  uint64_t* b_empty = b_full + kMaxBStages;                    // [kMaxBStages]
  uint64_t* tmem_full = b_empty + kMaxBStages;                 // [2]
  uint64_t* tmem_empty = tmem_full + 2;                        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  long long* s_rowoff2 = reinterpret_cast<long long*>(tail + kBarBytes);             // [2][BM]: tile i uses half i & 1
  float* s_stat = reinterpret_cast<float*>(tail + kBarBytes + 2 * BM * 8);           // [kStatParts][2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = a.m_tiles * a.n_tiles;
  const int crank = CL == 2 ? (int)cluster_ctarank() : 0;
  const int t_first = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int t_limit = CL == 2 ? ((a.m_tiles + 1) / 2) * a.n_tiles : total_tiles;   // pair: work items of two M tiles
  const uint32_t band_tx = (uint32_t)((a.R + 2) * a.Wp * 128);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kMaxBands; ++s) { mbar_init(&band_full[s], 1); mbar_init(&band_empty[s], 1); }
    for (int s = 0; s < kMaxBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], (kEpiThreads / 32) * CL); }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CL == 2) {
      tmem_alloc_pair(tmem_slot, kTmemCols);
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();             // the peer's barriers exist before anything is sent to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item -> (output-channel tile, frame, first image row); the odd CTA of a pair may get f == F (nothing to do: its
  // band is TMA's zero fill and its rows are all "padding")
  auto tile_geo = [&](int tile, int& nt, int& f, int& h0) {
    const int q = tile / a.n_tiles;
    const int mt = CL * q + crank;
    nt = tile - q * a.n_tiles;
    f = mt / a.tiles_per_frame;
    h0 = a.R * (mt - f * a.tiles_per_frame);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (CL == 1 && a.bres && (int)blockIdx.x < total_tiles) {   // the whole weight matrix, once
        mbar_arrive_expect_tx(&b_full[0], (uint32_t)(9 * a.cblocks * kBBytes));
        for (int kb = 0; kb < 9 * a.cblocks; ++kb) tma_load_2d(b_base + (size_t)kb * kBBytes, &tmB, &b_full[0], kb * BK, 0);
      }
      int bs = 0, ss = 0;
      uint32_t bph = 0, sph = 0;
      for (int tile = t_first; tile < t_limit; tile += t_step) {
        int nt, f, h0;
        tile_geo(tile, nt, f, h0);
        for (int cb = 0; cb < a.cblocks; ++cb) {
          mbar_wait(&band_empty[bs], bph ^ 1);
          if (CL == 2) {                                       // both CTAs' boxes complete on the leader's barrier
            if (crank == 0) mbar_arrive_expect_tx(&band_full[bs], 2 * band_tx);
            tma_load_4d_pair(band_base + (size_t)bs * a.band_bytes, &tmX, leader_addr(&band_full[bs]), cb * BK, -1, h0 - 1, f);
          } else {
            mbar_arrive_expect_tx(&band_full[bs], band_tx);
            tma_load_4d(band_base + (size_t)bs * a.band_bytes, &tmX, &band_full[bs], cb * BK, -1, h0 - 1, f);
          }
          if (++bs == a.bands) { bs = 0; bph ^= 1; }
          if (CL == 2 || !a.bres) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_empty[ss], sph ^ 1);
              if (CL == 2) {                                   // this CTA's half of the weight tile's output channels
                if (crank == 0) mbar_arrive_expect_tx(&b_full[ss], 2 * kBBytes);
                tma_load_2d_pair(b_base + (size_t)ss * kBBytes, &tmB, leader_addr(&b_full[ss]), tap * a.Cin + cb * BK,
                                 nt * BN + crank * (BN / 2));
              } else {
                mbar_arrive_expect_tx(&b_full[ss], kBBytes);
                tma_load_2d(b_base + (size_t)ss * kBBytes, &tmB, &b_full[ss], tap * a.Cin + cb * BK, nt * BN);
              }
              if (++ss == a.bstages) { ss = 0; sph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (CL == 1 || crank == 0) {                               // pair: the leader issues for both CTAs
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CL, BN, 0, 0);
      const bool bres = CL == 1 && a.bres;
      if (bres && (int)blockIdx.x < total_tiles) {
        mbar_wait(&b_full[0], 0);
        tc_fence_after();
      }
      int bs = 0, ss = 0, it = 0;
      uint32_t bph = 0, sph = 0;
      for (int tile = t_first; tile < t_limit; tile += t_step, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int cb = 0; cb < a.cblocks; ++cb) {
          mbar_wait(&band_full[bs], bph);
          tc_fence_after();
          // One descriptor per band and per weight tile, then 64-bit adds: the issuing thread's own instruction chain
          // paces N = 64 / 128 MMAs (54 / 64 clocks each at the hardware rate, tools/ubench/umma_rate.cu) -- building both
          // descriptors from scratch for every MMA cost ~150 clocks per MMA.  The start-address field counts 16-byte units
          // in the low 14 bits and never carries out of them for in-range shared-memory addresses.
          // tap (0, 0) of the tile's first position (padded row h0 + 1, column 1) is the band's very first row
          const uint64_t a_first = umma_smem_desc_sw128(smem_u32(band_base + (size_t)bs * a.band_bytes), 0, 1024);
          const uint64_t b_res = umma_smem_desc_sw128(smem_u32(b_base) + (uint32_t)cb * kBBytes, 0, 1024);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap - 3 * r;
            const uint64_t ad = a_first + (uint64_t)(uint32_t)((r * a.Wp + s) * 8);   // a row of 128 bytes = 8 descriptor units
            uint64_t bd;
            if (bres) {
              bd = b_res + (uint64_t)(uint32_t)(tap * a.cblocks * (kBBytes / 16));
            } else {
              mbar_wait(&b_full[ss], sph);
              tc_fence_after();
              bd = umma_smem_desc_sw128(smem_u32(b_base) + (uint32_t)ss * kBBytes, 0, 1024);
            }
            if (CL == 2) {
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk) umma_f16_elect_pair(d_tmem, ad + 2u * kk, bd + 2u * kk, idesc, (cb | tap | kk) != 0);
              umma_commit_elect_pair(&b_empty[ss]);                  // both CTAs' producers may refill this stage
              if (++ss == a.bstages) { ss = 0; sph ^= 1; }
            } else {
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk) umma_f16_elect(d_tmem, ad + 2u * kk, bd + 2u * kk, idesc, (cb | tap | kk) != 0);
              if (!bres) {
                umma_commit_elect(&b_empty[ss]);
                if (++ss == a.bstages) { ss = 0; sph ^= 1; }
              }
            }
          }
          // the band (both CTAs' in a pair) is free once these MMAs have read it
          if (CL == 2) umma_commit_elect_pair(&band_empty[bs]); else umma_commit_elect(&band_empty[bs]);
          if (++bs == a.bands) { bs = 0; bph ^= 1; }
        }
        if (CL == 2) umma_commit_elect_pair(&tmem_full[acc]); else umma_commit_elect(&tmem_full[acc]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int et = threadIdx.x - 64;
    const int lane_grp = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = lane_grp * 32 + lane;
    constexpr int kHalfN = BN / 2;
    // Per-column sums of the rounded tiles (gemm_tn.cu's scheme), kept in registers ACROSS the CTA's tiles: thread
    // (pair, part) owns two columns and every kStatParts-th block of rows; one shared-memory reduction and one atomic
    // per column when the output-channel tile changes (never, for the one-tile widths of layer1/2) and at the end.
    constexpr int kPairs = BN / 2;
    constexpr int kRows = BM / kStatParts;
    const int st_pair = et % kPairs, st_part = et / kPairs;
    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
    int stat_nt = -1;
    auto flush_stats = [&](int nt_of) {
      float* mine = s_stat + (size_t)st_part * (2 * BN);
      *reinterpret_cast<float2*>(mine + 2 * st_pair) = make_float2(s1a, s1b);
      *reinterpret_cast<float2*>(mine + BN + 2 * st_pair) = make_float2(s2a, s2b);
      named_bar_sync(2, kEpiThreads);
      for (int i = et; i < 2 * BN; i += kEpiThreads) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < kStatParts; ++q) v += s_stat[(size_t)q * (2 * BN) + i];
        atomicAdd(i < BN ? &a.colsum[nt_of * BN + i] : &a.colsq[nt_of * BN + i - BN], v);
      }
      named_bar_sync(2, kEpiThreads);
      s1a = s1b = s2a = s2b = 0.f;
    };
    int it = 0;
    for (int tile = t_first; tile < t_limit; tile += t_step, ++it) {
      int nt, f, h0;
      tile_geo(tile, nt, f, h0);
      const int acc = it & 1;
      // double-buffered: a fast warp may start tile i+1's offsets while a slow one still stores tile i
      long long* s_rowoff = s_rowoff2 + acc * BM;
      if (et < BM) {                                           // where tile row `et` goes, or -1 for a padding position
        const int j = et + 1;                                  // position within the tile's first padded row, from column 0
        const int dh = j / a.Wp, wp = j - dh * a.Wp;
        long long off = -1;
        if (wp >= 1 && wp <= a.W && dh < a.R && h0 + dh < a.H && f < a.F) off = (((long long)f * a.H + (h0 + dh)) * a.W + (wp - 1)) * a.Cout;
        s_rowoff[et] = off;
      }
      named_bar_sync(1, kEpiThreads);                          // offsets visible; the previous tile's staging reads are over
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const bool valid = s_rowoff[row] >= 0;
      const uint32_t taddr = tmem_base + (uint32_t)(acc * BN + half * kHalfN) + ((uint32_t)(lane_grp * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < kHalfN; cc += 32) {
        const int c0 = half * kHalfN + cc;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + cc, v);
        tmem_ld_wait();
        uint8_t* panel = staging + (size_t)(c0 >> 6) * (BM * 128) + (size_t)row * 128;
        const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o = make_uint4(0u, 0u, 0u, 0u);                // padding positions contribute nothing to the statistics
          if (valid) {
            o.x = pack_bf16(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
            o.y = pack_bf16(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
            o.z = pack_bf16(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
            o.w = pack_bf16(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
          }
          *reinterpret_cast<uint4*>(panel + (((chunk0 + q) ^ (row & 7)) << 4)) = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CL == 2) mbar_arrive_cluster(leader_addr(&tmem_empty[acc])); else mbar_arrive(&tmem_empty[acc]); }
      named_bar_sync(1, kEpiThreads);
      // store: 16-byte chunks, a warp covers whole rows
      {
        constexpr int kChunksPerRow = BN / 8;
#pragma unroll 4
        for (int idx = et; idx < BM * kChunksPerRow; idx += kEpiThreads) {
          const int r = idx / kChunksPerRow, ch = idx - r * kChunksPerRow;
          const long long off = s_rowoff[r];
          if (off >= 0) {
            const uint4 v4 = *reinterpret_cast<const uint4*>(staging + (size_t)(ch >> 3) * (BM * 128) + (size_t)r * 128 +
                                                             (((ch & 7) ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(a.out + off + nt * BN + ch * 8) = v4;
          }
        }
      }
      if (a.colsum) {
        if (nt != stat_nt) {
          if (stat_nt >= 0) flush_stats(stat_nt);
          stat_nt = nt;
        }
        const int col = 2 * st_pair;
        const uint8_t* panel = staging + (size_t)(col >> 6) * (BM * 128);
        const int chunk = (col & 63) >> 3, within = (col & 7) * 2;
#pragma unroll 8
        for (int rr = 0; rr < kRows; ++rr) {
          const int r = st_part * kRows + rr;
          const uint32_t h2 = *reinterpret_cast<const uint32_t*>(panel + r * 128 + ((chunk ^ (r & 7)) << 4) + within);
          bf16x2_sum_sq(h2, s1a, s1b, s2a, s2b);
        }
      }
    }
    if (a.colsum && stat_nt >= 0) flush_stats(stat_nt);
  }

  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();             // the peer may still be reading this CTA's shared memory / TMEM
  if (warp == 2) {
    tc_fence_after();
    if (CL == 2) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, int CL>
int launch_halo(const mvfb_conv_desc* d, const void* x, const void* w, void* out, float* colsum, float* colsq, cudaStream_t st) {
  HaloArgs a;
  a.F = d->F; a.H = d->H; a.W = d->W; a.Cin = d->Cin; a.Cout = d->Cout;
  a.Wp = d->W + 2;
  a.R = (BM + 2) / a.Wp;                                       // R * Wp - 2 positions per tile, at most 128
  a.tiles_per_frame = (d->H + a.R - 1) / a.R;
  a.m_tiles = d->F * a.tiles_per_frame;
  a.n_tiles = d->Cout / BN;
  a.cblocks = d->Cin / BK;
  a.band_bytes = (((2 * a.Wp + 2 + BM) * 128 + 1023) / 1024) * 1024;
  a.colsum = colsum; a.colsq = colsq;
  a.out = (__nv_bfloat16*)out;
  constexpr int kBBytes = BN * BK * 2 / CL;
  const size_t fixed = 1024 + (size_t)BM * BN * 2 + kBarBytes + 2 * BM * 8 + (size_t)(kEpiThreads / (BN / 2)) * BN * 2 * 4;
  const size_t budget = 227 * 1024;
  // weights resident when they fit beside two bands; otherwise a ring of (tap, channel block) tiles
  a.bres = 0;
  const size_t wbytes = (size_t)9 * a.cblocks * kBBytes;
  if (CL == 1 && a.n_tiles == 1 && fixed + wbytes + 2 * (size_t)a.band_bytes <= budget) a.bres = 1;
  const size_t bsm = a.bres ? wbytes : 0;
  a.bstages = a.bres ? 0 : 6;
  size_t left = budget - fixed - bsm - (a.bres ? 0 : (size_t)a.bstages * kBBytes);
  a.bands = (int)(left / a.band_bytes);
  if (a.bands > kMaxBands) a.bands = kMaxBands;
  if (a.bands < 2) return MVFB_ERR_UNSUPPORTED;
  if (!a.bres) {                                               // spend what is left on deeper weight staging
    left -= (size_t)a.bands * a.band_bytes;
    int extra = (int)(left / kBBytes);
    if (a.bstages + extra > kMaxBStages) extra = kMaxBStages - a.bstages;
    a.bstages += extra;
  }
  const size_t smem_bytes = fixed + (size_t)a.bands * a.band_bytes + (a.bres ? wbytes : (size_t)a.bstages * kBBytes);
  CUtensorMap tmX, tmB;
  const uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->F};
  const uint64_t strides[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
  const uint32_t box[4] = {(uint32_t)BK, (uint32_t)a.Wp, (uint32_t)(a.R + 2), 1u};
  int rc = encode_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (rc) return rc;
  const uint64_t bdims[2] = {(uint64_t)9 * d->Cin, (uint64_t)d->Cout};
  const uint64_t bstrides[1] = {(uint64_t)9 * d->Cin * 2};
  const uint32_t bbox[2] = {(uint32_t)BK, (uint32_t)(BN / CL)};
  rc = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, bdims, bstrides, bbox, nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (rc) return rc;
  static DevOnce once;
  if (once.pending()) {
    MVFB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    once.done();
  }
  if (CL == 2) {
    int pairs = ((a.m_tiles + 1) / 2) * a.n_tiles;
    if (pairs > num_sms() / 2) pairs = num_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MVFB_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<BN, CL>, tmX, tmB, a));
    count_launch();
    return MVFB_OK;
  }
  int grid = a.m_tiles * a.n_tiles;
  if (grid > num_sms()) grid = num_sms();
  conv_halo_kernel<BN, CL><<<grid, kThreads, smem_bytes, st>>>(tmX, tmB, a);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

}  // namespace

// The shapes this kernel is the better choice for (measured, profiles/r02_conv_halo.txt: 394 vs 650 us at C = 64 / 56x56,
// 295 vs 320 us at C = 128 / 28x28, but 305 vs 233 us at C = 256 / 14x14); everything else stays on the im2col GEMM.
bool conv_halo_eligible(const mvfb_conv_desc* d) {
  // a tile is R whole image rows: at least 96 of its 128 MMA rows must be real pixels (W = 64 would give one row = 64)
  const int rows = (BM + 2) / (d->W + 2);
  if (rows * d->W < 96) return false;
  return option(OPT_CONV_HALO_OFF) == 0 && d->stride == 1 && d->ksize == 3 && d->W >= 14 && d->W <= 126 && d->Cin % 64 == 0 && d->Cin <= 128 &&
         (d->Cout == 64 || d->Cout % 128 == 0) && (long long)d->F * d->H * (d->W + 2) / BM >= 2 * num_sms();
}

int conv_halo(const mvfb_conv_desc* d, const void* x, const void* w, void* out, float* colsum, float* colsq, cudaStream_t st) {
  if (d->Cout % 128 == 0) {
    // streamed weights: a CTA pair halves the weight bytes per SM -- 243 vs 271 us at C = 128 / 28x28 / 1280 frames;
    // with one channel block (nine k-blocks per tile) the pair's synchronisation costs more: 169 vs 151 us
    if (option(OPT_GEMM_PAIR_OFF) == 0 && d->Cin >= 2 * BK) return launch_halo<128, 2>(d, x, w, out, colsum, colsq, st);
    return launch_halo<128, 1>(d, x, w, out, colsum, colsq, st);
  }
  return launch_halo<64, 1>(d, x, w, out, colsum, colsq, st);
}

}  // namespace mvfb

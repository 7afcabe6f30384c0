// Weight gradient of a 1x1 convolution on the tensor cores:   dW[n, k] = sum_m dY[m, n] * X[m, k]
//
//   dY: (M, N) pixels x output channels, X: (M, K) pixels x input channels, both bf16 row-major (NHWC views);
//   dW: (N, K) fp32 row-major -- torch's (Cout, Cin, 1, 1) gradient, accumulated with fp32 atomics (split-K).
//
// As a GEMM the reduction runs over the pixel axis, which is the OUTER (strided) axis of both operands: both are
// "MN-major" for the MMA (tcgen05 instruction-descriptor bits a_major = b_major = 1).  A TMA box of
// 64 channels x 64 pixels lands in shared memory as 64 pixel-rows of 128 bytes with the 128-byte swizzle, which is
// exactly the canonical MN-major SWIZZLE_128B layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units
// (CUTLASS cute/arch/mma_sm100_desc.hpp): 8 pixel-rows form an atom (SBO = 1024 B), the next 64 channels are the
// next box (LBO = 64 rows * 128 B = 8192 B); one MMA (K = 16 pixels) advances the start address by 2 atoms.
//
// The X operand may be K-split like conv1x1_gemm's A operand: input channels [0, K0) from X0 (the compact MVF
// slab), [K0, K) from X1 -- the weight gradient of MVF's wrapped 1x1 convolution without materialising x'.
//
// Work decomposition: output tiles are few (Cout/128 x Cin/BN) and the reduction is long (M/64 = 10^3..10^4
// k-blocks), so each CTA owns one (tile, k-range) pair; ranges are sized so the grid fills the 148 SMs.
#include <cuda_bf16.h>

#include "common.cuh"
#include "mvf_internal.cuh"
#include "ptx.cuh"

namespace mvfb {

namespace {

constexpr int BM = 128;            // output channels (dY columns) per tile = UMMA M
constexpr int BKP = 64;            // pixels per stage
constexpr int kThreads = 192;

struct WArgs {
  long long M;                     // pixels
  int N, K, K0;                    // Cout, Cin, Cin taken from X0
  int n_tiles, k_tiles, splits;    // tiles along Cout, along Cin (x 9 filter taps for IM2COL); k-range count
  int pblocks;                     // ceil(M / 64)
  float* dw;
  long long lddw;
  int Ho, Wo, stride, ks, pad;     // IM2COL (3x3 / pad 1 or 1x1 / pad 0): output geometry; dw is (Cout, ks, ks, Cin)
};

template <int BN, int CL = 1>
struct WCfg {
  static constexpr int kStages = (BN >= 256 && CL == 1) ? 4 : 6;
  static constexpr int kABytes = BM * BKP * 2;                // 2 boxes of 64 ch x 64 px
  static constexpr int kBBytes = BN * BKP * 2 / CL;           // BN/64 boxes; a CTA pair holds half of them each
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr size_t kSmem = 1024 + (size_t)kStages * kStageBytes + 256;
};

// MN-major, SWIZZLE_128B shared-memory matrix descriptor
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr) {
  return umma_smem_desc_sw128(smem_addr, /*LBO=*/BKP * 128, /*SBO=*/1024);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// CL = 2: CTA pairs (tcgen05 cta_group::2, as in gemm_tn.cu): the pair owns 256 output channels of one (cin tile, pixel range)
// work item; each CTA loads its own 128 dY columns and HALF of the X tile, the leader issues M = 256 MMAs.
template <int BN, bool IM2COL, int CL = 1>
__global__ void __launch_bounds__(kThreads, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX0,
                  const __grid_constant__ CUtensorMap tmX1, const WArgs a) {
  using C = WCfg<BN, CL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)C::kStages * C::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* done = empty + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item -> (cout tile, cin tile, pixel range).  Tiles vary fastest so that co-resident CTAs work on the SAME
  // pixel range: the dY / X blocks they share (all 9 filter taps, all channel tiles) are then served by L2.
  const int crank = CL == 2 ? (int)cluster_ctarank() : 0;
  const int item = (int)blockIdx.x / CL;                          // work item of this CTA (pair)
  const int ntile_all = (a.n_tiles / CL) * a.k_tiles;
  const int split = item / ntile_all, tile = item - split * ntile_all;
  const int nt = CL * (tile / a.k_tiles) + crank, kt = tile % a.k_tiles;
  // IM2COL: kt enumerates (filter tap rs, channel tile ct); the tap's input pixels are gathered by TMA im2col
  const int ctiles = (a.K + BN - 1) / BN;                      // the last tile may run past K: TMA zero-fills, the epilogue skips
  const int rs = IM2COL ? kt / ctiles : 0, ct = IM2COL ? kt - rs * ctiles : kt;
  const int per = (a.pblocks + a.splits - 1) / a.splits;
  const int pb0 = split * per;
  int pb1 = pb0 + per;
  if (pb1 > a.pblocks) pb1 = a.pblocks;
  const int nblocks = pb1 > pb0 ? pb1 - pb0 : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmX1);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
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
  if (CL == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nblocks; ++i) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + (size_t)s * C::kStageBytes;
        const int p = (pb0 + i) * BKP;
        int bw = 0, bh = 0, bn = 0;
        if (IM2COL) {
          const int q = p % a.Wo, pq = p / a.Wo;                // first output pixel of the block -> window origin
          bw = q * a.stride - a.pad; bh = (pq % a.Ho) * a.stride - a.pad; bn = pq / a.Ho;
        }
        if (CL == 2) {
          // both CTAs' boxes complete on the leader's barrier, which expects the bytes of both
          if (crank == 0) mbar_arrive_expect_tx(&full[s], 2 * C::kStageBytes);
          const uint32_t lead = leader_addr(&full[s]);
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d_pair(sa + j * (BKP * 128), &tmG, lead, nt * BM + j * 64, p);
#pragma unroll
          for (int j = 0; j < BN / 128; ++j) {                    // this CTA's half of the X tile
            const int c = ct * BN + crank * (BN / 2) + j * 64;
            if (IM2COL) tma_load_im2col_4d_pair(sa + C::kABytes + j * (BKP * 128), &tmX1, lead, c, bw, bh, bn,
                                                (uint16_t)(rs % a.ks), (uint16_t)(rs / a.ks));
            else tma_load_2d_pair(sa + C::kABytes + j * (BKP * 128), c < a.K0 ? &tmX0 : &tmX1, lead, c, p);
          }
        } else {
          mbar_arrive_expect_tx(&full[s], C::kStageBytes);
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BKP * 128), &tmG, &full[s], nt * BM + j * 64, p);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) {
            const int c = ct * BN + j * 64;
            if (IM2COL) tma_load_im2col_4d(sa + C::kABytes + j * (BKP * 128), &tmX1, &full[s], c, bw, bh, bn,
                                           (uint16_t)(rs % a.ks), (uint16_t)(rs / a.ks));
            else tma_load_2d(sa + C::kABytes + j * (BKP * 128), c < a.K0 ? &tmX0 : &tmX1, &full[s], c, p);
          }
        }
        if (++s == C::kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (CL == 1 || crank == 0) {                               // all lanes run the loop, one elected lane issues (ptx.cuh)
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CL, BN, 1, 1);   // both operands MN-major
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nblocks; ++i) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * C::kStageBytes);
        const uint64_t adesc = desc_mn_sw128(sa), bdesc = desc_mn_sw128(sa + C::kABytes);
        if (CL == 2) {
#pragma unroll
          for (int kk = 0; kk < BKP / 16; ++kk)
            umma_f16_elect_pair(tmem_base, adesc + 128u * kk, bdesc + 128u * kk, idesc, (i | kk) != 0);
          umma_commit_elect_pair(&empty[s]);
        } else {
#pragma unroll
          for (int kk = 0; kk < BKP / 16; ++kk)                 // 16 pixels = 2 atoms = 2048 bytes = 128 descriptor units
            umma_f16_elect(tmem_base, adesc + 128u * kk, bdesc + 128u * kk, idesc, (i | kk) != 0);
          umma_commit_elect(&empty[s]);
        }
        if (++s == C::kStages) { s = 0; ph ^= 1; }
      }
      if (CL == 2) umma_commit_elect_pair(done);
      else umma_commit_elect(done);
    }
  } else if (nblocks > 0) {
    // epilogue: TMEM -> registers -> fp32 atomics on dW (rows = output channels of this tile)
    const int lane_grp = warp & 3;
    const int row = nt * BM + lane_grp * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
    float* dst = a.dw + (size_t)row * a.lddw + (size_t)rs * a.K + (size_t)ct * BN;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + c0, v);
      tmem_ld_wait();
      if (row < a.N) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (ct * BN + c0 + q * 4 < a.K)
            red_add_v4(dst + c0 + q * 4, __uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                     __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    if (CL == 2) tmem_dealloc_pair(tmem_base, C::kTmemCols);
    else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

int map_64x64(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t ld) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {ld * 2};
  const uint32_t box[2] = {64, BKP};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, nullptr,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

template <int BN, bool IM2COL, int CL>
int launch_wgrad_grid(const CUtensorMap& tmG, const CUtensorMap& tmX0, const CUtensorMap& tmX1, const WArgs& a, int ctas,
                      cudaStream_t st) {
  using C = WCfg<BN, CL>;
  static DevOnce once;
  if (once.pending()) {
    MVFB_CUDA(cudaFuncSetAttribute(gemm_wgrad_kernel<BN, IM2COL, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
    once.done();
  }
  if (CL == 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MVFB_CUDA(cudaLaunchKernelEx(&cfg, gemm_wgrad_kernel<BN, IM2COL, CL>, tmG, tmX0, tmX1, a));
    return MVFB_OK;
  }
  gemm_wgrad_kernel<BN, IM2COL, CL><<<ctas, kThreads, C::kSmem, st>>>(tmG, tmX0, tmX1, a);
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
}

template <int BN, bool IM2COL>
int launch_wgrad_kernel(const CUtensorMap& tmG, const CUtensorMap& tmX0, const CUtensorMap& tmX1, WArgs a, size_t dw_elems,
                        cudaStream_t st) {
  a.n_tiles = (a.N + BM - 1) / BM;
  a.k_tiles = ((a.K + BN - 1) / BN) * (IM2COL ? a.ks * a.ks : 1);
  a.pblocks = (int)((a.M + BKP - 1) / BKP);
  const int tiles = a.n_tiles * a.k_tiles;
  int splits = (2 * num_sms() + tiles - 1) / tiles;            // ~2 CTAs worth of work items per SM
  if (splits > a.pblocks) splits = a.pblocks;
  if (splits < 1) splits = 1;
  a.splits = splits;
  MVFB_CUDA(cudaMemsetAsync(a.dw, 0, sizeof(float) * dw_elems, st));
  // CTA pairs: 256 output channels per work item, half of the X tile per CTA (same condition as gemm_tn.cu's pairs: wide
  // tiles, whole 256-channel pairs, a reduction long enough to pay for the pair's synchronisation)
  int rc;
  if (BN == 256 && a.K % 256 == 0 && a.N % 256 == 0 && option(OPT_GEMM_PAIR_OFF) == 0 && a.pblocks / splits >= 8) {
    rc = launch_wgrad_grid<BN, IM2COL, 2>(tmG, tmX0, tmX1, a, tiles * splits, st);
  } else {
    rc = launch_wgrad_grid<BN, IM2COL, 1>(tmG, tmX0, tmX1, a, tiles * splits, st);
  }
  if (rc) return rc;
  count_launch();
  return MVFB_OK;
}

template <int BN>
int launch_wgrad(const mvfb_gemm_desc* d, const void* g, const void* x0, const void* x1, float* dw, cudaStream_t st) {
  CUtensorMap tmG, tmX0, tmX1;
  int rc;
  if ((rc = map_64x64(&tmG, g, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldb))) return rc;
  if ((rc = map_64x64(&tmX1, x1, (uint64_t)d->K, (uint64_t)d->M, (uint64_t)d->lda1))) return rc;
  if (d->K0 > 0) {
    if ((rc = map_64x64(&tmX0, x0, (uint64_t)d->K0, (uint64_t)d->M, (uint64_t)d->lda0))) return rc;
  } else {
    tmX0 = tmX1;
  }
  WArgs a;
  a.M = d->M; a.N = d->N; a.K = d->K; a.K0 = d->K0;
  a.dw = dw; a.lddw = d->ldd;
  a.Ho = a.Wo = a.stride = a.pad = 0;
  a.ks = 1;
  return launch_wgrad_kernel<BN, false>(tmG, tmX0, tmX1, a, (size_t)d->N * d->ldd, st);
}

// 3x3 / pad 1 weight gradient: dw[n, r, s, c] = sum_p g[p, n] * x[window(p) + (r, s), c]
template <int BN>
int launch_wgrad3x3(const mvfb_conv_desc* d, const void* g, const void* x, float* dw, cudaStream_t st) {
  const int Ho = (d->H - 1) / d->stride + 1, Wo = (d->W - 1) / d->stride + 1;
  const long long M = (long long)d->F * Ho * Wo;
  CUtensorMap tmG, tmX;
  int rc;
  if ((rc = map_64x64(&tmG, g, (uint64_t)d->Cout, (uint64_t)M, (uint64_t)d->Cout))) return rc;
  const uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->F};
  const uint64_t strides[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
  const int pad = d->ksize == 3 ? 1 : 0, taps = d->ksize * d->ksize;
  const int lower[2] = {-pad, -pad}, upper[2] = {-pad, -pad};
  const uint32_t estr[4] = {1, (uint32_t)d->stride, (uint32_t)d->stride, 1};
  if ((rc = encode_tmap_im2col(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, lower, upper, 64, BKP, estr,
                               CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  WArgs a;
  a.M = M; a.N = d->Cout; a.K = d->Cin; a.K0 = 0;
  a.dw = dw; a.lddw = (long long)taps * d->Cin;
  a.Ho = Ho; a.Wo = Wo; a.stride = d->stride; a.ks = d->ksize; a.pad = pad;
  return launch_wgrad_kernel<BN, true>(tmG, tmX, tmX, a, (size_t)d->Cout * taps * d->Cin, st);
}

}  // namespace

}  // namespace mvfb

using namespace mvfb;

// d: M = pixels, N = Cout, K = Cin, K0 = Cin taken from x0; lda0/lda1 = leading dims of x0/x1, ldb = leading dim of
// g (dY), ldd = leading dim of dw (fp32, >= K, contiguous rows: the buffer N x ldd is zeroed by this call).
extern "C" int conv1x1_wgrad(const mvfb_gemm_desc* d, const void* g, const void* x0, const void* x1, float* dw,
                             mvfb_stream_t stream) {
  MVFB_CHECK(d && g && x1 && dw, MVFB_ERR_ARG, "null descriptor / operand");
  MVFB_CHECK(d->M > 0 && d->N > 0 && d->K > 0, MVFB_ERR_ARG, "bad shape M=%lld N=%d K=%d", d->M, d->N, d->K);
  MVFB_CHECK(d->N % 64 == 0 && d->K % 64 == 0 && d->K0 % 64 == 0 && d->K0 >= 0 && d->K0 < d->K, MVFB_ERR_UNSUPPORTED,
             "N=%d, K=%d, K0=%d must be multiples of 64 with K0 < K", d->N, d->K, d->K0);
  MVFB_CHECK(d->K0 == 0 || x0, MVFB_ERR_ARG, "K0 > 0 needs the x0 operand");
  MVFB_CHECK(d->lda1 % 8 == 0 && d->ldb % 8 == 0 && d->ldd % 4 == 0 && d->ldd >= d->K && (d->K0 == 0 || d->lda0 % 8 == 0),
             MVFB_ERR_UNSUPPORTED, "bad leading dimensions");
  MVFB_CHECK(!((uintptr_t)g & 15) && !((uintptr_t)x1 & 15) && !((uintptr_t)x0 & 15) && !((uintptr_t)dw & 15),
             MVFB_ERR_UNSUPPORTED, "operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  // K = 192 (the stem's patch rows): ONE 256-wide tile with a zero-filled quarter reads dY once; three 64-wide tiles read
  // it three times and run the N = 64 MMA (2.05 ms for 8.2 GB, profiles/r02_step_by_shape.txt)
  if (d->K % 256 == 0 || (d->K == 192 && d->K0 == 0)) return launch_wgrad<256>(d, g, x0, x1, dw, st);
  if (d->K % 128 == 0) return launch_wgrad<128>(d, g, x0, x1, dw, st);
  return launch_wgrad<64>(d, g, x0, x1, dw, st);
}

// g: (F, Ho, Wo, Cout) bf16; x: (F, H, W, Cin) bf16; dw: (Cout, 3, 3, Cin) fp32, zeroed by the call.
extern "C" int conv3x3_wgrad(const mvfb_conv_desc* d, const void* g, const void* x, float* dw, mvfb_stream_t stream) {
  MVFB_CHECK(d && g && x && dw, MVFB_ERR_ARG, "null descriptor / operand");
  MVFB_CHECK(d->F > 0 && d->H > 0 && d->W > 0 && (d->stride == 1 || d->stride == 2) && (d->ksize == 3 || d->ksize == 1),
             MVFB_ERR_ARG, "bad conv shape F=%d H=%d W=%d stride=%d ksize=%d", d->F, d->H, d->W, d->stride, d->ksize);
  MVFB_CHECK(d->Cin % 64 == 0 && d->Cout % 64 == 0, MVFB_ERR_UNSUPPORTED, "Cin=%d and Cout=%d must be multiples of 64",
             d->Cin, d->Cout);
  MVFB_CHECK(!((uintptr_t)g & 15) && !((uintptr_t)x & 15) && !((uintptr_t)dw & 15), MVFB_ERR_UNSUPPORTED,
             "operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->Cin % 256 == 0) return launch_wgrad3x3<256>(d, g, x, dw, st);
  if (d->Cin % 128 == 0) return launch_wgrad3x3<128>(d, g, x, dw, st);
  return launch_wgrad3x3<64>(d, g, x, dw, st);
}

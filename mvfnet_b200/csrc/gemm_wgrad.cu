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
  int n_tiles, k_tiles, splits;    // tiles along Cout, along Cin; k-range count
  int pblocks;                     // ceil(M / 64)
  float* dw;
  long long lddw;
};

template <int BN>
struct WCfg {
  static constexpr int kStages = BN >= 256 ? 4 : 6;
  static constexpr int kABytes = BM * BKP * 2;                // 2 boxes of 64 ch x 64 px
  static constexpr int kBBytes = BN * BKP * 2;                // BN/64 boxes
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

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX0,
                  const __grid_constant__ CUtensorMap tmX1, const WArgs a) {
  using C = WCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)C::kStages * C::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* done = empty + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item -> (cout tile, cin tile, pixel range)
  const int tile = blockIdx.x / a.splits, split = blockIdx.x - tile * a.splits;
  const int nt = tile / a.k_tiles, kt = tile - nt * a.k_tiles;
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
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nblocks; ++i) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + (size_t)s * C::kStageBytes;
        mbar_arrive_expect_tx(&full[s], C::kStageBytes);
        const int p = (pb0 + i) * BKP;
#pragma unroll
        for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BKP * 128), &tmG, &full[s], nt * BM + j * 64, p);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) {
          const int c = kt * BN + j * 64;
          tma_load_2d(sa + C::kABytes + j * (BKP * 128), c < a.K0 ? &tmX0 : &tmX1, &full[s], c, p);
        }
        if (++s == C::kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 1, 1);      // both operands MN-major
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nblocks; ++i) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * C::kStageBytes);
        const uint32_t sb = sa + C::kABytes;
#pragma unroll
        for (int kk = 0; kk < BKP / 16; ++kk)
          umma_f16(tmem_base, desc_mn_sw128(sa + kk * 2048), desc_mn_sw128(sb + kk * 2048), idesc, (i | kk) != 0);
        umma_commit(&empty[s]);
        if (++s == C::kStages) { s = 0; ph ^= 1; }
      }
      umma_commit(done);
    }
  } else if (nblocks > 0) {
    // epilogue: TMEM -> registers -> fp32 atomics on dW (rows = output channels of this tile)
    const int lane_grp = warp & 3;
    const int row = nt * BM + lane_grp * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
    float* dst = a.dw + (size_t)row * a.lddw + (size_t)kt * BN;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + c0, v);
      tmem_ld_wait();
      if (row < a.N) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          red_add_v4(dst + c0 + q * 4, __uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                     __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

int map_64x64(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t ld) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {ld * 2};
  const uint32_t box[2] = {64, BKP};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, nullptr,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

template <int BN>
int launch_wgrad(const mvfb_gemm_desc* d, const void* g, const void* x0, const void* x1, float* dw, cudaStream_t st) {
  using C = WCfg<BN>;
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
  a.n_tiles = (d->N + BM - 1) / BM;
  a.k_tiles = d->K / BN;
  a.pblocks = (int)((d->M + BKP - 1) / BKP);
  const int tiles = a.n_tiles * a.k_tiles;
  int splits = (2 * num_sms() + tiles - 1) / tiles;            // ~2 CTAs worth of work items per SM
  if (splits > a.pblocks) splits = a.pblocks;
  if (splits < 1) splits = 1;
  a.splits = splits;
  a.dw = dw; a.lddw = d->ldd;
  static bool once = false;
  if (!once) {
    MVFB_CUDA(cudaFuncSetAttribute(gemm_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
    once = true;
  }
  MVFB_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->N * d->ldd, st));
  gemm_wgrad_kernel<BN><<<tiles * splits, kThreads, C::kSmem, st>>>(tmG, tmX0, tmX1, a);
  count_launch();
  MVFB_LAUNCH_CHECK();
  return MVFB_OK;
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
  if (d->K % 256 == 0) return launch_wgrad<256>(d, g, x0, x1, dw, st);
  if (d->K % 128 == 0) return launch_wgrad<128>(d, g, x0, x1, dw, st);
  return launch_wgrad<64>(d, g, x0, x1, dw, st);
}

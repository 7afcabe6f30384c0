// Internal (non-ABI) declarations shared between the MVF translation units.
#pragma once
#include <cuda_runtime.h>

#include "../../include/mvf_b200.h"

namespace mvfb {

void count_launch(int n = 1);

// conv_halo.cu: 3x3 / stride 1 convolution with the A operand read as shifted views of one halo band
bool conv_halo_eligible(const mvfb_conv_desc* d);
int conv_halo(const mvfb_conv_desc* d, const void* x, const void* w, void* out, float* colsum, float* colsq, cudaStream_t st);

// ---- mvf_generic.cu : any layout / dtype / shape
int mvf_generic_fwd(const mvfb_mvf_desc* d, const void* x, void* y, long long y_stride, const float* wt,
                    const float* wh, const float* ww, const float* gamma, const float* beta, float* rm, float* rv,
                    float* save_mean, float* save_rstd, void* ws, cudaStream_t st);
size_t mvf_generic_bwd_ws(const mvfb_mvf_desc* d);
int mvf_generic_bwd(const mvfb_mvf_desc* d, const void* g, long long g_stride, const void* x, void* dx,
                    long long dx_stride, const float* wt, const float* wh, const float* ww, const float* gamma,
                    const float* beta, const float* mean, const float* rstd, float* dwt, float* dwh, float* dww,
                    float* dgamma, float* dbeta, void* ws, cudaStream_t st);

// (mean, rstd) from running statistics (eval-mode backward); mvf_generic.cu
int mvf_eval_stats(const float* rm, const float* rv, float eps, int Cs, float* mean, float* rstd, cudaStream_t st);

// ---- mvf_fast.cu : bf16 NHWC, TMA-staged (T,H,W) tiles in shared memory
bool mvf_fast_supported(const mvfb_mvf_desc* d);
bool mvf_fast_bwd_supported(const mvfb_mvf_desc* d);
size_t mvf_fast_ws(const mvfb_mvf_desc* d);
int mvf_fast_fwd(const mvfb_mvf_desc* d, const void* x, void* y, long long y_stride, const float* wt,
                const float* wh, const float* ww, const float* gamma, const float* beta, float* rm, float* rv,
                float* save_mean, float* save_rstd, void* ws, cudaStream_t st);
int mvf_fast_bwd(const mvfb_mvf_desc* d, const void* g, long long g_stride, const void* x, void* dx,
                long long dx_stride, const float* wt, const float* wh, const float* ww, const float* gamma,
                const float* beta, const float* mean, const float* rstd, float* dwt, float* dwh, float* dww,
                float* dgamma, float* dbeta, void* ws, cudaStream_t st);

// ---- mvf_stream.cu : bf16 NHWC forward, persistent warp-specialised frame stream (preferred forward path)
bool mvf_stream_supported(const mvfb_mvf_desc* d);
size_t mvf_stream_ws(const mvfb_mvf_desc* d);
int mvf_stream_fwd(const mvfb_mvf_desc* d, const void* x, void* y, long long y_stride, const float* wt,
                   const float* wh, const float* ww, const float* gamma, const float* beta, float* rm, float* rv,
                   float* save_mean, float* save_rstd, void* ws, cudaStream_t st);

// ---- mvf_sweep.cu : bf16 NHWC forward, FHFMA arithmetic, train-mode BatchNorm in one cooperative launch (preferred)
bool mvf_sweep_supported(const mvfb_mvf_desc* d);
size_t mvf_sweep_ws(const mvfb_mvf_desc* d);
int mvf_sweep_fwd(const mvfb_mvf_desc* d, const void* x, void* y, long long y_stride, const float* wt,
                  const float* wh, const float* ww, const float* gamma, const float* beta, float* rm, float* rv,
                  float* save_mean, float* save_rstd, void* ws, size_t ws_bytes, cudaStream_t st);

// ---- mvf_stream_bwd.cu : bf16 NHWC backward, persistent frame stream (preferred backward path, whole-frame tiles)
bool mvf_stream_bwd_supported(const mvfb_mvf_desc* d);
size_t mvf_stream_bwd_ws(const mvfb_mvf_desc* d);
int mvf_stream_bwd(const mvfb_mvf_desc* d, const void* g, long long g_stride, const void* x, void* dx,
                   long long dx_stride, const float* wt, const float* wh, const float* ww, const float* gamma,
                   const float* beta, const float* mean, const float* rstd, float* dwt, float* dwh, float* dww,
                   float* dgamma, float* dbeta, void* ws, cudaStream_t st);

// ---- mvf_sweep_bwd.cu : bf16 NHWC backward, FHFMA arithmetic, statistics sweep + dx sweep in one cooperative launch
//      (preferred backward path: whole-frame tiles, 4-channel items, 1 / 2 / 4 items per thread)
bool mvf_sweep_bwd_supported(const mvfb_mvf_desc* d);
size_t mvf_sweep_bwd_ws(const mvfb_mvf_desc* d);
int mvf_sweep_bwd(const mvfb_mvf_desc* d, const void* g, long long g_stride, const void* x, void* dx,
                  long long dx_stride, const float* wt, const float* wh, const float* ww, const float* gamma,
                  const float* beta, const float* mean, const float* rstd, float* dwt, float* dwh, float* dww,
                  float* dgamma, float* dbeta, void* ws, const void* dx_add, cudaStream_t st);

// sums (fp64, [11][Cs]) -> fp32 parameter gradients (mvf_generic.cu)
__global__ void mvf_bwd_finalize(const double* sums, int Cs, int h_shares, int w_shares, int has_h, int has_w,
                                 int use_hs, float* dwt, float* dwh, float* dww, float* dgamma, float* dbeta);

}  // namespace mvfb

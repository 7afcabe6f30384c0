/*
 * mvf_b200.h -- C ABI of libmvf_b200.so: hand-written sm_100a kernels for the MVFNet hot path.
 *
 * The reference (whwu95/MVFNet @ 0ddc7e2) has no native ABI: its hot path is Python calling stock
 * torch.nn modules.  Each entry point below replaces the torch ops executed by the cited reference
 * lines; the Python side of the boundary (mvfnet_b200/mvf.py, resnet.py) keeps the reference's
 * module signatures and calls these through ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers + sizes, no C++/torch types.  All pointers are DEVICE pointers
 *     unless stated otherwise; the caller owns every buffer; the library never frees or retains
 *     them beyond the call (stream-ordered: buffers must stay alive until the stream reaches the
 *     end of the enqueued work).
 *   - Every call only enqueues work on `stream` (a cudaStream_t); no hidden device-wide sync.
 *   - Return value 0 = success; otherwise an MVFB_ERR_* code, message via mvf_b200_last_error().
 *   - dtype: MVFB_F32 or MVFB_BF16 (storage type of activations; arithmetic is always fp32).
 *   - layout: MVFB_NCHW = (F, C, H, W) contiguous; MVFB_NHWC = (F, H, W, C) contiguous
 *     (torch channels_last).  F = N*T frames, clip n owns frames [n*T, (n+1)*T).
 */
#ifndef MVF_B200_H_
#define MVF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVFB_VERSION 200

enum { MVFB_OK = 0, MVFB_ERR_ARG = 1, MVFB_ERR_CUDA = 2, MVFB_ERR_UNSUPPORTED = 3, MVFB_ERR_WORKSPACE = 4 };
enum { MVFB_F32 = 0, MVFB_BF16 = 1 };
enum { MVFB_NCHW = 0, MVFB_NHWC = 1 };
enum { MVFB_MODE_T = 0, MVFB_MODE_TH = 1, MVFB_MODE_THW = 2 };   /* MVF.py:112-129 `mode` */

typedef void* mvfb_stream_t; /* cudaStream_t */

int mvf_b200_version(void);
/* Thread-local, NUL-terminated description of the last failure on this thread ("" if none). */
const char* mvf_b200_last_error(void);
/* Number of kernels launched by this library in this process so far (bench.py `gpu_launches`). */
unsigned long long mvf_b200_launch_count(void);
/* Kernel tier that served the last successful mvf_fwd / mvf_bwd call of this PROCESS (any thread: autograd runs the
 * backward on its own engine thread): "sweep" (persistent
 * frame-stream kernels on the bf16-operand FMA, mvf_sweep.cu / mvf_sweep_bwd.cu), "stream" (mvf_stream*.cu), "ring"
 * (mvf_fast.cu) or "generic" (any layout / dtype, mvf_generic.cu); "" before the first call.  The parity tests assert
 * it, so a silent fall-through to a slower tier fails them. */
const char* mvf_b200_last_kernel(void);
/* Test / tool switches (process-wide; nothing in the library reads the environment).
 *   MVFB_OPT_FORCE_FWD / MVFB_OPT_FORCE_BWD: 0 = automatic tier selection (default), else one of MVFB_KERNEL_*: only
 *       that tier may serve mvf_fwd / mvf_bwd -- a descriptor it cannot serve fails with MVFB_ERR_UNSUPPORTED;
 *   MVFB_OPT_CONV_HALO_OFF: 1 = conv3x3_gemm never takes the halo-band kernel (A/B measurements, tests of the im2col path).
 *   MVFB_OPT_GEMM_PAIR_OFF: 1 = the GEMM kernels never launch as CTA pairs (tcgen05 cta_group::2).
 *   MVFB_OPT_SWEEP_DEBUG: 1 = the sweep forward kernel writes %globaltimer stamps at workspace + 512 KiB (needs a
 *       workspace of >= 1 MiB; tools/stream_timeline.py). */
enum { MVFB_OPT_FORCE_FWD = 0, MVFB_OPT_FORCE_BWD = 1, MVFB_OPT_SWEEP_DEBUG = 2, MVFB_OPT_CONV_HALO_OFF = 3, MVFB_OPT_GEMM_PAIR_OFF = 4 };
enum { MVFB_KERNEL_AUTO = 0, MVFB_KERNEL_SWEEP = 1, MVFB_KERNEL_STREAM = 2, MVFB_KERNEL_RING = 3, MVFB_KERNEL_GENERIC = 4 };
int mvf_b200_set_option(int key, int value);

/* ------------------------------------------------------------------------------------------------
 * MVF module  --  replaces MVF.forward up to (not including) self.net: the view/transpose/split,
 * three depthwise Conv3d, two adds, BatchNorm3d, HardSwish, cat and contiguous of
 * codes/models/modules/MVF.py:109-137 (+ common/se_module.py:5-24).
 *
 *   z = sum_k wt[c,k] x[n,t+k-1,c,h,w] + sum_k wh[c,k] x[n,t,c,h+k-1,w] + sum_k ww[c,k] x[n,t,c,h,w+k-1]
 *   u = gamma (z - mean)/sqrt(var+eps) + beta ;  y = u * clamp(u+3,0,6)/6      (use_hs)
 * for channels c < Cs; zero padding; the T axis never crosses a clip.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int N, T, C, Cs, H, W;  /* x is (N*T, C, H, W); Cs = int(C*alpha) slab channels (MVF.py:59)      */
  int dtype, layout;      /* MVFB_F32|MVFB_BF16, MVFB_NCHW|MVFB_NHWC                                */
  int mode;               /* MVFB_MODE_*; share=True is expressed by passing wh == ww == wt          */
  int use_hs;             /* 1: BatchNorm3d + HardSwish (MVF.py:131-134); 0: y = z                   */
  int training;           /* 1: batch statistics + running-stat update; 0: running statistics        */
  float eps, momentum;    /* BatchNorm3d defaults 1e-5, 0.1                                          */
} mvfb_mvf_desc;

/* Kernel tier that mvf_fwd (backward = 0) / mvf_bwd (backward = 1) selects for this descriptor from its shape alone
 * ("" for an invalid descriptor); host-only, needs no GPU.  A call can still step down a tier at run time (pointers
 * or strides that are not 16-byte aligned, a cooperative launch the device cannot co-schedule): mvf_b200_last_kernel()
 * tells what actually ran. */
const char* mvf_b200_plan(const mvfb_mvf_desc* d, int backward);

/* Bytes of scratch `workspace` the forward / backward need for this descriptor.  Its contents on entry are arbitrary
 * (uninitialised, zeroed, or left behind by any earlier call with any descriptor): the train-mode forward and the
 * backward exchange per-CTA partial sums through it as {value, tag} words whose tag is unique per call (a host call
 * number) and per replay of a captured call (a counter the kernel keeps IN the workspace) -- so the calls are safe to
 * capture in a CUDA graph and to replay, provided each captured call keeps its own workspace, and two calls that may run
 * concurrently must not share one.  Both are cooperative launches (one CTA per SM at most, all resident); if the device
 * cannot co-schedule the grid the call steps down to the two-launch kernels. */
size_t mvf_fwd_workspace_bytes(const mvfb_mvf_desc* d);
size_t mvf_bwd_workspace_bytes(const mvfb_mvf_desc* d);

/*
 * Forward.  Writes ONLY the slab channels [0,Cs) of every pixel to `y`:
 *   NCHW: y[f*y_stride + (c*H + h)*W + w]      (y_stride = C*H*W writes into a full tensor,
 *                                               Cs*H*W into a compact (F,Cs,H,W) slab)
 *   NHWC: y[((f*H + h)*W + w)*y_stride + c]    (y_stride = C or Cs)
 * x is never modified (it is the residual identity, backbones/resnet.py:211).
 * wt/wh/ww: fp32 (Cs,3) taps = shift_conv/h_conv/w_conv weights; wh/ww ignored per `mode`.
 * dtype MVFB_BF16: each of the nine taps is rounded to bf16 before use, in the forward and in the backward
 * alike (what torch.autocast(bfloat16) does to the reference's Conv3d weights); accumulation, BatchNorm and
 * HardSwish are fp32.  dtype MVFB_F32 uses the taps as given.
 * gamma/beta/running_*: fp32 (Cs); save_mean/save_rstd: fp32 (Cs) outputs (training; may be NULL
 * in eval).  training: running_mean/var are updated in place with momentum (unbiased variance).
 */
int mvf_fwd(const mvfb_mvf_desc* d, const void* x, void* y, long long y_stride, const float* wt, const float* wh,
            const float* ww, const float* gamma, const float* beta, float* running_mean, float* running_var,
            float* save_mean, float* save_rstd, void* workspace, size_t workspace_bytes, mvfb_stream_t stream);

/*
 * Backward.  g = dL/dy addressed like y (g_stride); dx receives dL/dx for slab channels only,
 * addressed like y (dx_stride) -- the pass-through channels' gradient is the caller's g itself.
 * Outputs dwt/dwh/dww (Cs,3) fp32, dgamma/dbeta (Cs) fp32 are OVERWRITTEN.  With share (wh==wt and/or
 * ww==wt) the views' tap gradients are summed into dwt and dwh/dww may be NULL.
 * training: save_mean/save_rstd from the forward; eval: running stats.
 * The library owns no device memory: every buffer, the workspace included, is the caller's.
 * dx MAY alias g (same pointer and stride): every kernel reads a frame of g completely before that frame's
 * dx is written -- this is how the caller turns "dL/dx' for all C channels" into dL/dx in place.
 */
int mvf_bwd(const mvfb_mvf_desc* d, const void* g, long long g_stride, const void* x, void* dx, long long dx_stride,
            const float* wt, const float* wh, const float* ww, const float* gamma, const float* beta,
            const float* running_mean, const float* running_var, const float* save_mean, const float* save_rstd,
            float* dwt, float* dwh, float* dww, float* dgamma, float* dbeta, void* workspace,
            size_t workspace_bytes, mvfb_stream_t stream);

/* mvf_bwd with the gradient of the Bottleneck's identity path folded in: dx (slab channels) = mvf_bwd(...) + dx_add, where
 * dx_add is addressed like dx (same stride) -- `out += identity` of backbones/resnet.py:238 sends a second gradient to the
 * block input, and summing it here (and in conv1x1_gemm_add_cols for the other channels) saves autograd's full-tensor add.
 * Served by the sweep tier only (MVFB_ERR_UNSUPPORTED otherwise: ask mvf_b200_plan first).  dx_add must not alias dx. */
int mvf_bwd_add(const mvfb_mvf_desc* d, const void* g, long long g_stride, const void* x, void* dx, long long dx_stride,
                const float* wt, const float* wh, const float* ww, const float* gamma, const float* beta,
                const float* running_mean, const float* running_var, const float* save_mean, const float* save_rstd,
                float* dwt, float* dwh, float* dww, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                const void* dx_add, mvfb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 1x1 convolution on NHWC bf16 activations as a tensor-core GEMM  --  replaces nn.Conv2d(kernel_size=1,
 * bias=False) of Bottleneck.conv1 / conv3 / downsample (backbones/resnet.py:157-162,179-180,299-303) and
 * the wrapped `self.net` of MVF.forward (MVF.py:138); called with (dY, W^T) it is also that layer's
 * input-gradient.
 *
 *   D[m, n] = sum_{k <  K0} A0[m*lda0 + k] * B[n*ldb + k]
 *           + sum_{k >= K0} A1[m*lda1 + k] * B[n*ldb + k]          m < M pixels, n < N, k < K
 *
 * A0 (optional, K0 > 0) is the compact MVF slab (F,H,W,Cs) written by mvf_fwd with y_stride = Cs; A1 is the
 * block input x itself (its first K0 channels are simply never read): MVF.py:135-137's cat + contiguous
 * never touch HBM.  All matrices bf16, K contiguous; accumulation fp32; D is rounded to bf16.
 * colsum/colsq (both or neither; fp32 [N], zeroed by the caller) receive the per-output-channel sum and
 * sum of squares of the ROUNDED D -- the batch statistics of the train-mode BatchNorm that follows.
 * Requirements: K, K0 multiples of 64; N multiple of 64; leading dimensions multiples of 8; 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  long long M;                     /* rows = F*H*W pixels                                                */
  int N, K, K0;                    /* output channels, input channels, channels taken from A0            */
  long long lda0, lda1, ldb, ldd;  /* leading dimensions in elements                                     */
} mvfb_gemm_desc;

int conv1x1_gemm(const mvfb_gemm_desc* d, const void* a0, const void* a1, const void* b, void* out, float* colsum,
                 float* colsq, mvfb_stream_t stream);

/* out = A B^T + res: the same GEMM with an (M, N) bf16 addend (leading dimension ldr) read in the epilogue.  Used as
 * the input-gradient of conv1 of a Bottleneck whose input also feeds the identity path: dL/dx = dY W + dL/d(identity)
 * in one pass instead of a GEMM and a separate full-tensor add (backbones/resnet.py:211-213, 238). */
int conv1x1_gemm_add(const mvfb_gemm_desc* d, const void* a0, const void* a1, const void* b, const void* res,
                     long long ldr, void* out, mvfb_stream_t stream);

/* conv1x1_gemm_add with the addend applied to columns >= first_col only (a multiple of 32): the input-gradient of an
 * MVF-wrapped conv1, whose slab columns [0, Cs) still have to pass through mvf_bwd_add before they meet the identity
 * path's gradient. */
int conv1x1_gemm_add_cols(const mvfb_gemm_desc* d, const void* a0, const void* a1, const void* b, const void* res,
                          long long ldr, int first_col, void* out, mvfb_stream_t stream);

/* Inference form of the same layers: an eval-mode BatchNorm2d (+ residual) (+ ReLU) folded into the convolution's epilogue,
 *   out = [relu]( (A B^T)[m, n] * scale[n] + shift[n] [+ res[m*ldr + n]] ),
 * scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale (fp32 [N], 16-byte aligned): norm1+relu,
 * norm2+relu, norm3 + identity + relu and the down-sample norm of Bottleneck.forward (backbones/resnet.py:213-242) under
 * model.eval() (test_recognizer.py:72-77).  The activation makes ONE trip to HBM per convolution.
 * conv3x3_gemm_bnact: res (optional) is (F, Ho, Wo, Cout) contiguous. */
int conv1x1_gemm_bnact(const mvfb_gemm_desc* d, const void* a0, const void* a1, const void* b, const float* scale,
                       const float* shift, const void* res, long long ldr, int relu, void* out, mvfb_stream_t stream);

/* Weight gradient of the same layer: dw[n, k] = sum_m g[m*ldb + n] * X[m, k], X = [x0[:, :K0] | x1[:, K0:]] as above
 * (M = pixels, N = Cout, K = Cin; lda0 / lda1 / ldb = leading dimensions of x0 / x1 / g; dw is fp32 (N, ldd) and is
 * zeroed by the call, then accumulated with fp32 atomics over split-K partitions of the pixel axis).  Both MMA
 * operands are MN-major views of the NHWC tensors: no transposed copies are made. */
int conv1x1_wgrad(const mvfb_gemm_desc* d, const void* g, const void* x0, const void* x1, float* dw,
                  mvfb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 3x3 / pad 1 / stride 1|2 convolution on NHWC bf16 activations as an implicit GEMM on the tensor cores  --
 * replaces Bottleneck.conv2 (backbones/resnet.py:163-170, stride per :151-153); with the spatially rotated,
 * channel-transposed weights it is also the stride-1 input-gradient.
 *   out[f, ho, wo, n] = sum_{r,s,c} x[f, ho*stride + r - 1, wo*stride + s - 1, c] * w[n, r, s, c]
 * x: (F, H, W, Cin); w: (Cout, 3, 3, Cin) (= a channels_last torch weight); out: (F, Ho, Wo, Cout); all bf16,
 * contiguous.  The A operand is gathered by TMA im2col loads (zero fill = padding); the rest is conv1x1_gemm's
 * kernel, including the optional per-channel (sum, sum of squares) epilogue.  Cin, Cout multiples of 64.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int F, H, W, Cin, Cout, stride;
  int ksize;                       /* 3: 3x3 / pad 1 (Bottleneck.conv2);  1: 1x1 / pad 0 with stride 2 -- the strided
                                      down-sampling convolution of make_res_layer (resnet.py:299-303); w is then (Cout, Cin) */
} mvfb_conv_desc;

int conv3x3_gemm(const mvfb_conv_desc* d, const void* x, const void* w, void* out, float* colsum, float* colsq,
                 mvfb_stream_t stream);

int conv3x3_gemm_bnact(const mvfb_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift,
                       const void* res, int relu, void* out, mvfb_stream_t stream);

/* Input gradient of the STRIDE-2 3x3 convolution (conv2 of the first block of layer2/3/4): `d` is the forward layer
 * (F, H, W, Cin, Cout, stride 2, ksize 3; H, W even), g (F, H/2, W/2, Cout), dx (F, H, W, Cin), all bf16 NHWC.  Four
 * stride-1 implicit GEMMs, one per parity of the output pixel, with 1x1 / 1x2 / 2x1 / 2x2 windows over g (9 taps in total:
 * no wasted MAC), each scattering its rows to that parity's pixels of dx.  wq = the four parity operands back to back,
 * parity p = 2*ph + pw:
 *   wq_p[c][(dh*kw_p + dw)*Cout + n] = w[n, r(ph, dh), s(pw, dw), c]    (Cin rows, kh_p*kw_p*Cout columns)
 * with kh_p = 1 + ph, kw_p = 1 + pw, r(0, 0) = 1, r(1, 0) = 2, r(1, 1) = 0 (the same for s) -- mvfnet_b200/ops.py builds
 * it. */
int conv3x3s2_dgrad(const mvfb_conv_desc* d, const void* g, const void* wq, void* dx, mvfb_stream_t stream);

/* Input gradient of the STRIDE-2 1x1 (down-sampling) convolution: g (F, H/2, W/2, Cout), wT = W^T (Cin, Cout), dx
 * (F, H, W, Cin), bf16 NHWC, H and W even.  The GEMM dY W writes each row to pixel (2i, 2j) of dx and zeroes the three other
 * pixels of that 2 x 2 cell: dx is written completely, no memset, no scatter pass. */
int conv1x1s2_dgrad(const mvfb_conv_desc* d, const void* g, const void* wT, void* dx, mvfb_stream_t stream);

/* Its weight gradient: dw[n, r, s, c] = sum_{f,ho,wo} g[f, ho, wo, n] * x[f, ho*stride + r - 1, wo*stride + s - 1, c]
 * (g bf16 (F, Ho, Wo, Cout); dw fp32 (Cout, 3, 3, Cin), zeroed by the call; MN-major MMA operands, the x operand
 * gathered by TMA im2col, split-K fp32 atomics over the pixel axis). */
int conv3x3_wgrad(const mvfb_conv_desc* d, const void* g, const void* x, float* dw, mvfb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm2d (+ residual add) (+ ReLU) on NHWC bf16 activations viewed as (M = F*H*W, C) rows  --  replaces
 * norm1+relu, norm2+relu, norm3 / downsample norm + `out += identity` + relu of Bottleneck.forward
 * (backbones/resnet.py:213-242; nn.BatchNorm2d semantics as common/norm.py:66 builds it: eps 1e-5,
 * momentum 0.1, biased variance to normalise, unbiased for the running estimate).
 *
 *   bn_stats : sums[0][c] = sum_m x[m,c], sums[1][c] = sum_m x[m,c]^2     (zeroes `sums` itself; the 1x1
 *              GEMM can produce the same two rows in its epilogue, in which case this launch is skipped)
 *   bn_apply : y = [relu]( gamma (x - mean) rstd + beta [+ residual] );  training: mean / var from `sums`,
 *              save_mean/save_rstd written, running statistics updated in place; eval: running statistics
 *   bn_bwd   : g = dL/dy.  g' = g * (y > 0) when relu;  dbeta = sum g', dgamma = sum g' xhat,
 *              dx = gamma rstd (g' - dbeta/M - xhat dgamma/M)  (training)  or  gamma rstd g'  (eval);
 *              dres (optional) = g' -- the gradient of the residual input.  `sums` is [2][C] scratch.
 *   relu_mask (optional, M*C/8 bytes): bn_apply records 1 bit per element (y > 0), byte index m*(C/8) + c/8, bit c%8;
 *              bn_bwd then reads that instead of y (16x less traffic for the mask).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  long long M;       /* rows = F*H*W                          */
  int C;             /* channels, multiple of 8, <= 2048      */
  int relu;          /* 1: ReLU after the affine (+ residual)  */
  int training;      /* 1: batch statistics                    */
  float eps, momentum;
} mvfb_bn_desc;

/* y[r, 0:cols] = x[r, 0:cols] for r < M (bf16, 16-byte vectors): gathers / scatters a channel range of an
 * NHWC tensor, e.g. the MVF slab; also the probe for what HBM sustains on that strided pattern. */
int copy_cols(const void* x, long long ldx, void* y, long long ldy, long long M, int cols, mvfb_stream_t stream);
int bn_stats(const mvfb_bn_desc* d, const void* x, long long ldx, float* sums, mvfb_stream_t stream);
int bn_apply(const mvfb_bn_desc* d, const void* x, long long ldx, const void* residual, long long ldr, void* y,
             long long ldy, const float* sums, const float* gamma, const float* beta, float* running_mean,
             float* running_var, float* save_mean, float* save_rstd, unsigned char* relu_mask, mvfb_stream_t stream);
int bn_bwd(const mvfb_bn_desc* d, const void* g, long long ldg, const void* y, long long ldy, const void* x,
           long long ldx, const float* gamma, const float* mean, const float* rstd, void* dx, long long lddx,
           void* dres, long long lddr, float* dgamma, float* dbeta, float* sums, const unsigned char* relu_mask,
           mvfb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * ResNet stem  --  replaces conv1 = nn.Conv2d(3, 64, 7, stride 2, padding 3, bias=False) and
 * maxpool = nn.MaxPool2d(3, 2, 1) of ResNet.forward (backbones/resnet.py:424-431, 481-484).
 *
 *   stem_im2col      : x (F, H, W, 3) bf16 NHWC -> a (F*Ho*Wo, 192) bf16 row-major, Ho = (H-1)/2+1: the 7x7x3
 *                      patch of every output pixel, column k = kh*24 + kw*3 + c (each kernel row = 21 values + 3
 *                      zeros: three aligned 16-byte chunks; columns 168..191 zero).  conv1x1_gemm(a, W (64,192) in the
 *                      same column order) is then the convolution (BatchNorm sums in its epilogue) and
 *                      conv1x1_wgrad(dY, a) its weight gradient.
 *   maxpool3x3s2_fwd : x (F, H, W, C) bf16 NHWC -> y (F, Ho, Wo, C) and idx (F, Ho, Wo, C) bytes = position of the
 *                      maximum inside the 3x3 window (kh*3 + kw; the first maximum in scan order, NaN wins: the
 *                      rule of ATen's max_pool2d kernels).  C % 8 == 0.
 *   maxpool3x3s2_bwd : dx (F, H, W, C) = g (F, Ho, Wo, C) routed to the recorded positions (gather over the <= 4
 *                      windows containing a pixel; every element of dx is written).
 * ---------------------------------------------------------------------------------------------- */
int stem_im2col(const void* x, void* a, long long F, int H, int W, mvfb_stream_t stream);
int maxpool3x3s2_fwd(const void* x, void* y, void* idx, long long F, int H, int W, int C, mvfb_stream_t stream);
int maxpool3x3s2_bwd(const void* g, const void* idx, void* dx, long long F, int H, int W, int C, mvfb_stream_t stream);

/* norm1 + ReLU + maxpool of ResNet.forward (backbones/resnet.py:481-484) in ONE pass over the stem convolution's output, and
 * the matching backward -- the 112 x 112 x 64 activation and its gradient never exist in memory.
 *
 *   bn_relu_maxpool_fwd : x (F, H, W, C) bf16 = conv1 output, `d` as for bn_apply (M = F*H*W, relu = 1; training: mean / var
 *                         from `sums`, save_mean / save_rstd written, running statistics updated; eval: running statistics)
 *                         -> y (F, Ho, Wo, C) = maxpool3x3s2(relu(bn(x))) bit for bit, idx (F, Ho, Wo, C) bytes = window
 *                         position of the winning x (first in scan order), or 9 where the pooled value is <= 0 (the ReLU
 *                         passes no gradient there).  C / 8 must divide 256.
 *   bn_relu_maxpool_bwd : g (F, Ho, Wo, C) = dL/dy -> dx (F, H, W, C) = dL/d(conv1 output), dgamma, dbeta (bn_bwd's formulas,
 *                         g' = the pooled gradients routed to the recorded positions, summed in fp32).  `sums` is [2][C]
 *                         scratch.  Even H and W.
 */
int bn_relu_maxpool_fwd(const mvfb_bn_desc* d, const void* x, long long F, int H, int W, const float* sums, const float* gamma,
                        const float* beta, float* running_mean, float* running_var, float* save_mean, float* save_rstd, void* y,
                        void* idx, mvfb_stream_t stream);
int bn_relu_maxpool_bwd(const mvfb_bn_desc* d, const void* g, const void* idx, const void* x, long long F, int H, int W,
                        const float* gamma, const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                        float* sums, mvfb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Input pre-processing on the GPU  --  replaces Normalize + FormatShape + ToTensor of the data pipeline
 * (codes/datasets/pipelines/augmentations.py:343-396, formating.py:134-185; configs r50_dense.py:70-75) for frames that
 * arrive as decoded uint8 images: y[p, c] = (float(x[p, to_rgb ? 2 - c : c]) - mean[c]) / std[c], written as bf16 in
 * the NHWC frame layout the stem reads.  x: (pixels, 3) uint8 HWC frames back to back (pixels % 4 == 0);
 * mean / std: HOST pointers to 3 floats, indexed by the output channel as mmcv.imnormalize applies them.
 * ---------------------------------------------------------------------------------------------- */
int preprocess_u8(const void* x, void* y, long long pixels, const float* mean, const float* std_, int to_rgb,
                  mvfb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Classification head and loss  --  replaces TSNClsHead.forward + BaseHead.loss (codes/models/heads/tsn_clshead.py:71-98,
 * heads/base.py:40-45, segmental_consensuses/simple_consensus.py:41-61) around the 2048 -> num_classes projection,
 * which runs on conv1x1_gemm / conv1x1_wgrad with the class axis zero-padded to a multiple of 64 (ldl columns):
 *
 *   head_pool_fwd : feat[f, c] = dropout_p( mean over the HW pixels of x[f, :, c] )   x (F, HW, C) bf16 NHWC -> (F, C) bf16
 *                   (the keep decision of element (f, c) is a counter-based hash of (seed, f*C + c): the backward
 *                   recomputes it, no mask is stored; p = 0 disables dropout.  seed_dev (optional DEVICE pointer) is
 *                   mixed into the seed by the kernel: a step captured in a CUDA graph advances that word on the
 *                   device, so every replay draws a new mask)
 *   head_ce_fwd   : s[b, c] = bias[c] + mean_t logits[b*T + t, c]  (SimpleConsensus 'avg' over the T segments);
 *                   loss = mean_b ( logsumexp(s[b]) - s[b, label[b]] );  ds = d loss / d s  (B, NC) fp32;
 *                   dbias[c] = sum_b ds[b, c];  score (optional) receives s.  loss / dbias are zeroed by the call.
 *   head_ce_bwd   : dlogits[b*T + t, c] = gout * ds[b, c] / T  (bf16, (B*T, ldl), padding columns zero);
 *                   gout = device pointer to d L / d loss (NULL = 1).
 *   head_pool_bwd : dx[f, q, c] = dfeat[f, c] * keep(f, c) / ((1 - p) * HW)  for every pixel q.
 * ---------------------------------------------------------------------------------------------- */
int head_pool_fwd(const void* x, void* feat, long long F, int HW, int C, float p, unsigned long long seed,
                  const unsigned long long* seed_dev, mvfb_stream_t stream);
int head_pool_bwd(const void* dfeat, void* dx, long long F, int HW, int C, float p, unsigned long long seed,
                  const unsigned long long* seed_dev, mvfb_stream_t stream);
int head_ce_fwd(const void* logits, long long ldl, const float* bias, const long long* labels, int B, int T, int NC,
                float* score, float* ds, float* dbias, float* loss, mvfb_stream_t stream);
int head_ce_bwd(const float* ds, const float* gout, void* dlogits, long long ldl, int B, int T, int NC,
                mvfb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer tail  --  replaces what DistOptimizerHook.after_train_iter does after the all-reduce
 * (codes/core/dist_utils.py:29-32, 59-67 with optimizer / grad_clip of r50_dense.py:152-154): divide by the world size,
 * clip the global gradient norm, torch.optim.SGD(momentum, weight_decay, nesterov) -- over FLAT fp32 buffers:
 *
 *   flat_sqnorm       : out (device double, zeroed by the call) = sum_i g[i]^2
 *   sgd_nesterov_step : c = grad_scale * min(1, max_norm / (sqrt(sqnorm) * grad_scale + 1e-6))   (max_norm <= 0: c = grad_scale)
 *                       g' = c g + weight_decay p;  m = momentum m + g';  p -= lr (nesterov ? g' + momentum m : m)
 *                       p_bf16 (optional) receives the updated parameters rounded to bf16 -- the operands the next
 *                       forward's GEMMs read; norm_out (optional, device float) the total norm sqrt(sqnorm) * grad_scale.
 *                       A zero-initialised m reproduces torch's first step (buf = grad).
 * ---------------------------------------------------------------------------------------------- */
int flat_sqnorm(const float* g, long long n, double* out, mvfb_stream_t stream);
/* All transposed weight forms of a step in one launch: `tiles` is a DEVICE array of `ntiles` records
 *   struct { const bf16* src; bf16* dst; int lds, ldd, rows, cols, r0, c0; }   (40 bytes, 8-byte aligned)
 * each naming one 32 x 32 tile at (r0, c0) of a rows x cols bf16 matrix (element (r, c) at src[r*lds + c]) whose transpose
 * goes to dst[c*ldd + r].  The host builds the table once (the flat weight buffers never move): the W^T operands of the
 * 1x1 input-gradient GEMMs and the rotated (Cin, 3, 3, Cout) operands of the 3x3 ones (one record set per filter tap). */
int transpose_tiles(const void* tiles, long long ntiles, mvfb_stream_t stream);
int sgd_nesterov_step(float* p, float* mom, const float* g, void* p_bf16, long long n, const double* sqnorm,
                      float* norm_out, float grad_scale, float max_norm, float lr, float momentum, float weight_decay,
                      int nesterov, mvfb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MVF_B200_H_ */

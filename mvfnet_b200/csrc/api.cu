// C-ABI entry points shared by every kernel family: version / error / launch accounting, TMA tensor-map
// encoding through the driver entry point (no -lcuda link dependency), and the MVF dispatch.
#include <cuda_bf16.h>
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "mvf_internal.cuh"

namespace mvfb {

static thread_local char g_err[512] = {0};
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int num_sms() {
  static std::atomic<int> sms[64];                      // per device: a process may drive several GPUs
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::atomic<int>& slot = sms[dev & 63];
  int n = slot.load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    slot.store(n, std::memory_order_relaxed);
  }
  return n;
}

// process-wide, not thread-local: autograd runs the backward on its own engine thread, the test that asks is on another
static std::atomic<const char*> g_last_kernel{""};
void note_kernel(const char* name) { g_last_kernel.store(name, std::memory_order_relaxed); }

static std::atomic<int> g_options[OPT_COUNT];
int option(int key) { return key >= 0 && key < OPT_COUNT ? g_options[key].load(std::memory_order_relaxed) : 0; }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_tiled() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

static PFN_cuTensorMapEncodeIm2col_v12000 get_encode_im2col() {
  static PFN_cuTensorMapEncodeIm2col_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeIm2col_v12000)p;
  }
  return fn;
}

int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                CUtensorMapSwizzle swizzle, CUtensorMapL2promotion l2) {
  auto fn = get_encode_tiled();
  MVFB_CHECK(fn != nullptr, MVFB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  uint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), (const cuuint64_t*)dims,
                  (const cuuint64_t*)strides_bytes, (const cuuint32_t*)box,
                  (const cuuint32_t*)(elem_strides ? elem_strides : ones), CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, l2,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVFB_CHECK(r == CUDA_SUCCESS, MVFB_ERR_CUDA,
             "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
             (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
             rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
  return MVFB_OK;
}

int encode_tmap_im2col(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                       const uint64_t* strides_bytes, const int* lower, const int* upper, uint32_t channels_per_pixel,
                       uint32_t pixels_per_column, const uint32_t* elem_strides, CUtensorMapSwizzle swizzle) {
  auto fn = get_encode_im2col();
  MVFB_CHECK(fn != nullptr, MVFB_ERR_CUDA, "cuTensorMapEncodeIm2col entry point unavailable");
  uint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), (const cuuint64_t*)dims,
                  (const cuuint64_t*)strides_bytes, lower, upper, channels_per_pixel, pixels_per_column,
                  (const cuuint32_t*)(elem_strides ? elem_strides : ones), CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVFB_CHECK(r == CUDA_SUCCESS, MVFB_ERR_CUDA, "cuTensorMapEncodeIm2col failed (%d)", (int)r);
  return MVFB_OK;
}

static int check_mvf_desc(const mvfb_mvf_desc* d) {
  MVFB_CHECK(d != nullptr, MVFB_ERR_ARG, "null descriptor");
  MVFB_CHECK(d->N > 0 && d->T > 0 && d->C > 0 && d->H > 0 && d->W > 0, MVFB_ERR_ARG,
             "bad shape N=%d T=%d C=%d H=%d W=%d", d->N, d->T, d->C, d->H, d->W);
  MVFB_CHECK(d->Cs > 0 && d->Cs <= d->C, MVFB_ERR_ARG, "Cs=%d must be in (0, C=%d] (Cs==0 is the caller's bypass)", d->Cs,
             d->C);
  MVFB_CHECK(d->dtype == MVFB_F32 || d->dtype == MVFB_BF16, MVFB_ERR_ARG, "bad dtype %d", d->dtype);
  MVFB_CHECK(d->layout == MVFB_NCHW || d->layout == MVFB_NHWC, MVFB_ERR_ARG, "bad layout %d", d->layout);
  MVFB_CHECK(d->mode >= MVFB_MODE_T && d->mode <= MVFB_MODE_THW, MVFB_ERR_ARG, "bad mode %d", d->mode);
  return MVFB_OK;
}

}  // namespace mvfb

using namespace mvfb;

extern "C" {

int mvf_b200_version(void) { return MVFB_VERSION; }
const char* mvf_b200_last_error(void) { return g_err; }
unsigned long long mvf_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char* mvf_b200_last_kernel(void) { return g_last_kernel.load(std::memory_order_relaxed); }
const char* mvf_b200_plan(const mvfb_mvf_desc* d, int backward) {
  if (check_mvf_desc(d)) return "";
  const int force = option(backward ? OPT_FORCE_BWD : OPT_FORCE_FWD);
  const auto on = [&](int k) { return force == 0 || force == k; };
  if (!backward) {
    if (on(MVFB_KERNEL_SWEEP) && mvf_sweep_supported(d)) return "sweep";
    if (on(MVFB_KERNEL_STREAM) && mvf_stream_supported(d)) return "stream";
    if (on(MVFB_KERNEL_RING) && mvf_fast_supported(d)) return "ring";
  } else {
    if (on(MVFB_KERNEL_SWEEP) && mvf_sweep_bwd_supported(d)) return "sweep";
    if (on(MVFB_KERNEL_STREAM) && mvf_stream_bwd_supported(d)) return "stream";
    if (on(MVFB_KERNEL_RING) && mvf_fast_bwd_supported(d)) return "ring";
  }
  return on(MVFB_KERNEL_GENERIC) ? "generic" : "";
}
int mvf_b200_set_option(int key, int value) {
  MVFB_CHECK(key >= 0 && key < OPT_COUNT, MVFB_ERR_ARG, "unknown option %d", key);
  g_options[key].store(value, std::memory_order_relaxed);
  return MVFB_OK;
}

size_t mvf_fwd_workspace_bytes(const mvfb_mvf_desc* d) {
  if (!d) return 0;
  size_t generic = 16 * sizeof(double) * (size_t)d->Cs + 256;
  size_t fast = mvf_fast_supported(d) ? mvf_fast_ws(d) : 0;
  size_t stream = mvf_stream_ws(d);
  if (stream > fast) fast = stream;
  const size_t sweep = mvf_sweep_ws(d);
  if (sweep > fast) fast = sweep;
  return generic > fast ? generic : fast;
}

size_t mvf_bwd_workspace_bytes(const mvfb_mvf_desc* d) {
  if (!d) return 0;
  // both paths are sized: the fast path can still decline at run time (pointer / stride alignment)
  size_t generic = mvf_generic_bwd_ws(d);
  const size_t head = 2 * sizeof(float) * (((size_t)d->Cs + 63) / 64 * 64);
  size_t fast = mvf_fast_bwd_supported(d) ? head + mvf_fast_ws(d) : 0;
  const size_t stream = mvf_stream_bwd_supported(d) ? head + mvf_stream_bwd_ws(d) : 0;
  if (stream > fast) fast = stream;
  const size_t sweep = mvf_sweep_bwd_supported(d) ? head + mvf_sweep_bwd_ws(d) : 0;
  if (sweep > fast) fast = sweep;
  return generic > fast ? generic : fast;
}

int mvf_fwd(const mvfb_mvf_desc* d, const void* x, void* y, long long y_stride, const float* wt, const float* wh,
            const float* ww, const float* gamma, const float* beta, float* running_mean, float* running_var,
            float* save_mean, float* save_rstd, void* workspace, size_t workspace_bytes, mvfb_stream_t stream) {
  int rc = check_mvf_desc(d);
  if (rc) return rc;
  MVFB_CHECK(x && y && wt, MVFB_ERR_ARG, "null x / y / wt");
  MVFB_CHECK(d->mode == MVFB_MODE_T || wh, MVFB_ERR_ARG, "mode needs wh");
  MVFB_CHECK(d->mode != MVFB_MODE_THW || ww, MVFB_ERR_ARG, "mode THW needs ww");
  if (d->use_hs) {
    MVFB_CHECK(gamma && beta, MVFB_ERR_ARG, "use_hs needs gamma/beta");
    MVFB_CHECK(d->training || (running_mean && running_var), MVFB_ERR_ARG, "eval mode needs running statistics");
    MVFB_CHECK(!d->training || (save_mean && save_rstd), MVFB_ERR_ARG, "training needs save_mean/save_rstd");
  }
  MVFB_CHECK(workspace_bytes >= mvf_fwd_workspace_bytes(d) && workspace, MVFB_ERR_WORKSPACE,
             "workspace too small: %zu < %zu", workspace_bytes, mvf_fwd_workspace_bytes(d));
  cudaStream_t st = (cudaStream_t)stream;
  // Kernel tiers, fastest first; each declines (MVFB_ERR_UNSUPPORTED) what it cannot serve.  `force` pins one tier
  // (tests and tools only, mvf_b200_set_option): a forced tier that declines is an error, not a fall-through.
  const int force = option(OPT_FORCE_FWD);
  if ((force == 0 || force == MVFB_KERNEL_SWEEP) && mvf_sweep_supported(d)) {
    rc = mvf_sweep_fwd(d, x, y, y_stride, wt, wh, ww, gamma, beta, running_mean, running_var, save_mean, save_rstd,
                       workspace, workspace_bytes, st);
    if (rc == MVFB_OK) note_kernel("sweep");
    if (rc != MVFB_ERR_UNSUPPORTED) return rc;
  }
  if ((force == 0 || force == MVFB_KERNEL_STREAM) && mvf_stream_supported(d)) {
    rc = mvf_stream_fwd(d, x, y, y_stride, wt, wh, ww, gamma, beta, running_mean, running_var, save_mean, save_rstd,
                        workspace, st);
    if (rc == MVFB_OK) note_kernel("stream");
    if (rc != MVFB_ERR_UNSUPPORTED) return rc;
  }
  if ((force == 0 || force == MVFB_KERNEL_RING) && mvf_fast_supported(d)) {
    rc = mvf_fast_fwd(d, x, y, y_stride, wt, wh, ww, gamma, beta, running_mean, running_var, save_mean, save_rstd,
                      workspace, st);
    if (rc == MVFB_OK) note_kernel("ring");
    if (rc != MVFB_ERR_UNSUPPORTED) return rc;   // unaligned pointers / strides: take the layout-generic kernels
  }
  MVFB_CHECK(force == 0 || force == MVFB_KERNEL_GENERIC, MVFB_ERR_UNSUPPORTED,
             "forced forward kernel tier %d does not serve this descriptor", force);
  rc = mvf_generic_fwd(d, x, y, y_stride, wt, wh, ww, gamma, beta, running_mean, running_var, save_mean, save_rstd,
                       workspace, st);
  if (rc == MVFB_OK) note_kernel("generic");
  return rc;
}

static int mvf_bwd_impl(const mvfb_mvf_desc* d, const void* g, long long g_stride, const void* x, void* dx, long long dx_stride,
            const float* wt, const float* wh, const float* ww, const float* gamma, const float* beta,
            const float* running_mean, const float* running_var, const float* save_mean, const float* save_rstd,
            float* dwt, float* dwh, float* dww, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
            const void* dx_add, mvfb_stream_t stream) {
  int rc = check_mvf_desc(d);
  if (rc) return rc;
  MVFB_CHECK(g && x && dx && wt && dwt, MVFB_ERR_ARG, "null g / x / dx / wt / dwt");
  const bool has_h = d->mode != MVFB_MODE_T, has_w = d->mode == MVFB_MODE_THW;
  MVFB_CHECK(!has_h || wh, MVFB_ERR_ARG, "mode needs wh");
  MVFB_CHECK(!has_w || ww, MVFB_ERR_ARG, "mode THW needs ww");
  MVFB_CHECK(!has_h || wh == wt || dwh, MVFB_ERR_ARG, "dwh required unless wh aliases wt");
  MVFB_CHECK(!has_w || ww == wt || dww, MVFB_ERR_ARG, "dww required unless ww aliases wt");
  const float *mean = nullptr, *rstd = nullptr;
  if (d->use_hs) {
    MVFB_CHECK(gamma && beta && dgamma && dbeta, MVFB_ERR_ARG, "use_hs needs gamma/beta/dgamma/dbeta");
    if (d->training) {
      MVFB_CHECK(save_mean && save_rstd, MVFB_ERR_ARG, "training backward needs save_mean/save_rstd");
    } else {
      MVFB_CHECK(running_mean && running_var, MVFB_ERR_ARG, "eval backward needs running statistics");
    }
  }
  MVFB_CHECK(workspace_bytes >= mvf_bwd_workspace_bytes(d) && workspace, MVFB_ERR_WORKSPACE,
             "workspace too small: %zu < %zu", workspace_bytes, mvf_bwd_workspace_bytes(d));
  cudaStream_t st = (cudaStream_t)stream;
  // eval mode: derive (mean, rstd) from the running statistics into the head of the workspace
  char* ws = (char*)workspace;
  if (d->use_hs) {
    if (d->training) {
      mean = save_mean;
      rstd = save_rstd;
    } else {
      float* tmp = (float*)ws;
      rc = mvf_eval_stats(running_mean, running_var, d->eps, d->Cs, tmp, tmp + d->Cs, st);
      if (rc) return rc;
      mean = tmp;
      rstd = tmp + d->Cs;
    }
  }
  ws += 2 * sizeof(float) * (((size_t)d->Cs + 63) / 64 * 64);
  const int force = option(OPT_FORCE_BWD);
  if ((force == 0 || force == MVFB_KERNEL_SWEEP) && mvf_sweep_bwd_supported(d)) {
    rc = mvf_sweep_bwd(d, g, g_stride, x, dx, dx_stride, wt, wh, ww, gamma, beta, mean, rstd, dwt, dwh, dww, dgamma,
                       dbeta, ws, dx_add, st);
    if (rc == MVFB_OK) note_kernel("sweep");
    if (rc != MVFB_ERR_UNSUPPORTED) return rc;
  }
  MVFB_CHECK(dx_add == nullptr, MVFB_ERR_UNSUPPORTED, "mvf_bwd_add is served by the sweep tier only (see mvf_b200_plan)");
  if ((force == 0 || force == MVFB_KERNEL_STREAM) && mvf_stream_bwd_supported(d)) {
    rc = mvf_stream_bwd(d, g, g_stride, x, dx, dx_stride, wt, wh, ww, gamma, beta, mean, rstd, dwt, dwh, dww, dgamma,
                        dbeta, ws, st);
    if (rc == MVFB_OK) note_kernel("stream");
    if (rc != MVFB_ERR_UNSUPPORTED) return rc;
  }
  if ((force == 0 || force == MVFB_KERNEL_RING) && mvf_fast_bwd_supported(d)) {
    rc = mvf_fast_bwd(d, g, g_stride, x, dx, dx_stride, wt, wh, ww, gamma, beta, mean, rstd, dwt, dwh, dww, dgamma,
                      dbeta, ws, st);
    if (rc == MVFB_OK) note_kernel("ring");
    if (rc != MVFB_ERR_UNSUPPORTED) return rc;
  }
  MVFB_CHECK(force == 0 || force == MVFB_KERNEL_GENERIC, MVFB_ERR_UNSUPPORTED,
             "forced backward kernel tier %d does not serve this descriptor", force);
  rc = mvf_generic_bwd(d, g, g_stride, x, dx, dx_stride, wt, wh, ww, gamma, beta, mean, rstd, dwt, dwh, dww, dgamma,
                       dbeta, ws, st);
  if (rc == MVFB_OK) note_kernel("generic");
  return rc;
}

int mvf_bwd(const mvfb_mvf_desc* d, const void* g, long long g_stride, const void* x, void* dx, long long dx_stride,
            const float* wt, const float* wh, const float* ww, const float* gamma, const float* beta,
            const float* running_mean, const float* running_var, const float* save_mean, const float* save_rstd,
            float* dwt, float* dwh, float* dww, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
            mvfb_stream_t stream) {
  return mvf_bwd_impl(d, g, g_stride, x, dx, dx_stride, wt, wh, ww, gamma, beta, running_mean, running_var, save_mean,
                      save_rstd, dwt, dwh, dww, dgamma, dbeta, workspace, workspace_bytes, nullptr, stream);
}

int mvf_bwd_add(const mvfb_mvf_desc* d, const void* g, long long g_stride, const void* x, void* dx, long long dx_stride,
                const float* wt, const float* wh, const float* ww, const float* gamma, const float* beta,
                const float* running_mean, const float* running_var, const float* save_mean, const float* save_rstd,
                float* dwt, float* dwh, float* dww, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                const void* dx_add, mvfb_stream_t stream) {
  MVFB_CHECK(dx_add != nullptr, MVFB_ERR_ARG, "mvf_bwd_add needs the addend");
  return mvf_bwd_impl(d, g, g_stride, x, dx, dx_stride, wt, wh, ww, gamma, beta, running_mean, running_var, save_mean,
                      save_rstd, dwt, dwh, dww, dgamma, dbeta, workspace, workspace_bytes, dx_add, stream);
}

}  // extern "C"

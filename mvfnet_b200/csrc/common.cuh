// Host-side plumbing shared by the C-ABI entry points: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/mvf_b200.h"

namespace mvfb {

void set_error(const char* fmt, ...);

#define MVFB_CHECK(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::mvfb::set_error(__VA_ARGS__); \
      return (code);                 \
    }                                \
  } while (0)

#define MVFB_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      ::mvfb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return MVFB_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)

#define MVFB_LAUNCH_CHECK() MVFB_CUDA(cudaGetLastError())

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

int num_sms();          // of the CURRENT device (cached per device)

// "Do this once per CUDA device": the shared-memory opt-in of a kernel (cudaFuncSetAttribute) is a per-device
// property, so a process-wide flag would leave a second GPU of the same process without it.
//   static DevOnce once;  if (once.pending()) { ...set attributes...; once.done(); }
// Two threads racing through the same first call both set the attribute, which is harmless.
struct DevOnce {
  std::atomic<unsigned long long> mask{0};
  static int dev() {
    int d = 0;
    (void)cudaGetDevice(&d);
    return d & 63;
  }
  bool pending() const { return !((mask.load(std::memory_order_acquire) >> dev()) & 1ull); }
  void done() { mask.fetch_or(1ull << dev(), std::memory_order_release); }
};

// Kernel family that served the last mvf_fwd / mvf_bwd call of this thread (mvf_b200_last_kernel()).
void note_kernel(const char* name);
// Test / tool knobs set through mvf_b200_set_option() (never read from the environment).
enum { OPT_FORCE_FWD = 0, OPT_FORCE_BWD = 1, OPT_SWEEP_DEBUG = 2, OPT_CONV_HALO_OFF = 3, OPT_GEMM_PAIR_OFF = 4, OPT_COUNT };
int option(int key);

// Encode a tiled TMA descriptor. dims/strides innermost first; strides in BYTES for dims 1..rank-1
// (dim 0 is contiguous).  elem_strides may be NULL (all 1).  Returns 0 or an MVFB_ERR_* code.
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                CUtensorMapSwizzle swizzle, CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B);

// im2col-mode descriptor (NHWC activations, rank 4 or 5): lower/upper = pixel-box corners of the rank-2 spatial dims.
int encode_tmap_im2col(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                       const uint64_t* strides_bytes, const int* lower, const int* upper, uint32_t channels_per_pixel,
                       uint32_t pixels_per_column, const uint32_t* elem_strides, CUtensorMapSwizzle swizzle);

}  // namespace mvfb

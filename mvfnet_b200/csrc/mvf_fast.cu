// placeholder until the TMA-staged kernels land (next commit)
#include "common.cuh"
#include "mvf_internal.cuh"
namespace mvfb {
bool mvf_fast_supported(const mvfb_mvf_desc*) { return false; }
size_t mvf_fast_bwd_ws(const mvfb_mvf_desc*) { return 0; }
int mvf_fast_fwd(const mvfb_mvf_desc*, const void*, void*, long long, const float*, const float*, const float*,
                 const float*, const float*, float*, float*, float*, float*, void*, cudaStream_t) {
  set_error("fast path not built");
  return MVFB_ERR_UNSUPPORTED;
}
int mvf_fast_bwd(const mvfb_mvf_desc*, const void*, long long, const void*, void*, long long, const float*,
                 const float*, const float*, const float*, const float*, const float*, const float*, float*, float*,
                 float*, float*, float*, void*, cudaStream_t) {
  set_error("fast path not built");
  return MVFB_ERR_UNSUPPORTED;
}
}  // namespace mvfb

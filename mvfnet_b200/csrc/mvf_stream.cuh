// Device helpers shared by the persistent frame-stream MVF kernels (mvf_stream.cu forward, mvf_stream_bwd.cu backward).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "ptx.cuh"

namespace mvfb {
namespace stream {

// V fp32 values as V/2 packed pairs (operands of FFMA2)
template <int V>
struct FV {
  float2 p[V / 2];
};
template <int V>
__device__ __forceinline__ FV<V> zerov() {
  FV<V> r;
#pragma unroll
  for (int j = 0; j < V / 2; ++j) r.p[j] = make_float2(0.f, 0.f);
  return r;
}
// V bf16 values from shared memory (32-bit shared-space address), unpacked
template <int V>
__device__ __forceinline__ FV<V> lds_bf16(uint32_t addr);
template <>
__device__ __forceinline__ FV<8> lds_bf16<8>(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  FV<8> r;
  r.p[0] = make_float2(bf16_lo(v.x), bf16_hi(v.x));
  r.p[1] = make_float2(bf16_lo(v.y), bf16_hi(v.y));
  r.p[2] = make_float2(bf16_lo(v.z), bf16_hi(v.z));
  r.p[3] = make_float2(bf16_lo(v.w), bf16_hi(v.w));
  return r;
}
template <>
__device__ __forceinline__ FV<4> lds_bf16<4>(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  FV<4> r;
  r.p[0] = make_float2(bf16_lo(v.x), bf16_hi(v.x));
  r.p[1] = make_float2(bf16_lo(v.y), bf16_hi(v.y));
  return r;
}
template <int V>
__device__ __forceinline__ FV<V> lds_f32(const float* p) {
  FV<V> r;
#pragma unroll
  for (int j = 0; j < V / 4; ++j) {
    const float4 a = *reinterpret_cast<const float4*>(p + 4 * j);
    r.p[2 * j] = make_float2(a.x, a.y);
    r.p[2 * j + 1] = make_float2(a.z, a.w);
  }
  return r;
}
template <int V>
__device__ __forceinline__ void fmav(FV<V>& z, const FV<V>& k, const FV<V>& x) {
#pragma unroll
  for (int j = 0; j < V / 2; ++j) z.p[j] = __ffma2_rn(k.p[j], x.p[j], z.p[j]);
}
template <int V>
__device__ __forceinline__ void store_bf16(__nv_bfloat16* dst, const FV<V>& z);
template <>
__device__ __forceinline__ void store_bf16<8>(__nv_bfloat16* dst, const FV<8>& z) {
  uint4 o;
  o.x = pack_bf16(z.p[0].x, z.p[0].y); o.y = pack_bf16(z.p[1].x, z.p[1].y);
  o.z = pack_bf16(z.p[2].x, z.p[2].y); o.w = pack_bf16(z.p[3].x, z.p[3].y);
  *reinterpret_cast<uint4*>(dst) = o;
}
template <>
__device__ __forceinline__ void store_bf16<4>(__nv_bfloat16* dst, const FV<4>& z) {
  uint2 o;
  o.x = pack_bf16(z.p[0].x, z.p[0].y); o.y = pack_bf16(z.p[1].x, z.p[1].y);
  *reinterpret_cast<uint2*>(dst) = o;
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool try_wait_u32(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
static __device__ __noinline__ void slow_wait_u32(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!try_wait_u32(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void wait_u32(uint32_t bar, uint32_t parity) {
  if (!try_wait_u32(bar, parity)) slow_wait_u32(bar, parity);
}
__device__ __forceinline__ void arrive_u32(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}


__device__ __forceinline__ void consumer_bar_sync(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

}  // namespace stream
}  // namespace mvfb

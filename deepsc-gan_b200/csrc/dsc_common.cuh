// Shared helpers for libdeepsc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/deepsc_b200.h"

namespace dsc {

void set_error(const char* fmt, ...);   // thread-local message, dsc_api.cu

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DSC_ERR_CUDA;
  }
  return DSC_OK;
}

#define DSC_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::dsc::set_error(__VA_ARGS__);      \
      return DSC_ERR_BAD_ARG;             \
    }                                     \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit accesses that do not pollute L1
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

constexpr int kSMs = 148;   // B200

// ---- row helpers of the vocabulary-sized kernels (dsc_rows.cu, dsc_backward.cu)
// A logits row starts at any 4-byte boundary (ld = 22,234): the first (4 - misalignment) % 4 elements and the tail are
// read as scalars, the body as float4.
struct RowSpan {
  int head;        // scalar elements before the aligned body
  int n4;          // float4 of the body
  const float4* body;
};
__device__ __forceinline__ RowSpan row_span(const float* row, int N) {
  RowSpan s;
  s.head = (int)((4u - ((uint32_t)((uintptr_t)row >> 2) & 3u)) & 3u);
  if (s.head > N) s.head = N;
  s.n4 = (N - s.head) >> 2;
  s.body = reinterpret_cast<const float4*>(row + s.head);
  return s;
}

// running (max, sum of exp(v - max)) of a softmax denominator
__device__ __forceinline__ void lse_take(float v, float& m, float& s) {
  if (v > m) { s = s * expf(m - v) + 1.0f; m = v; } else { s += expf(v - m); }
}
__device__ __forceinline__ void lse_merge(float m2, float s2, float& m, float& s) {
  const float mm = fmaxf(m, m2);
  s = s * expf(m - mm) + s2 * expf(m2 - mm);
  m = mm;
}



}  // namespace dsc

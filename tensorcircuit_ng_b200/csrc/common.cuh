// common.cuh — error plumbing and small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace tcb {

void set_error(const char* fmt, ...);

#define TCB_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::tcb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                     \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

#define TCB_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::tcb::set_error(__VA_ARGS__);  \
      return 2;                       \
    }                                 \
  } while (0)

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
  // (no "memory" clobber: callers never read the stored location again before a barrier, and a
  //  clobber would serialise the independent LDS -> STG chains of the streaming loops)
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w));
}

// insert a zero bit at position p of x
__device__ __host__ __forceinline__ uint64_t insert_zero(uint64_t x, int p) {
  return ((x >> p) << (p + 1)) | (x & ((1ull << p) - 1ull));
}

int sm_count();

}  // namespace tcb

// Shared helpers for the fnsm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/fnsm_b200.h"

namespace fnsm {

extern std::atomic<long long> g_launches;   // defined in abi.cu

struct DevInfo { int sms; int max_smem_optin; int cc_major; int cc_minor; };
// cached cudaGetDeviceProperties subset for the current device (abi.cu)
int device_info(DevInfo* out);

inline int post_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? FNSM_OK : (int)cudaGetLastError();
}

template <class T> struct scalar_of;
template <> struct scalar_of<double> { static constexpr int code = FNSM_F64; };
template <> struct scalar_of<float> { static constexpr int code = FNSM_F32; };

// streaming (touch-once) global accesses: keep them out of L1, evict-first in L2
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v2.f64 {%0,%1}, [%2];"
               : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ double ldg_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));   // (L2::evict_first needs a vector type)
  return v;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(double2* p, double2 v) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream(double* p, double v) {
  asm volatile("st.global.cs.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void stg_stream(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

}  // namespace fnsm

// Shared device/host helpers for libb200rl (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "b200rl.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200rl is written for sm_100a (B200) only"
#endif

namespace b200rl {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> B200RL_ECUDA

#define B200RL_REQUIRE(cond, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      ::b200rl::set_error(__VA_ARGS__);    \
      return B200RL_EINVAL;                \
    }                                      \
  } while (0)

#define B200RL_UNSUPPORTED(cond, ...)      \
  do {                                     \
    if (cond) {                            \
      ::b200rl::set_error(__VA_ARGS__);    \
      return B200RL_EUNSUPPORTED;          \
    }                                      \
  } while (0)

struct DeviceInfo {
  int sm_count;
  int max_smem_optin;
};
const DeviceInfo& device_info();

constexpr float kF32Lowest = -3.402823466e+38f;  // torch.finfo(torch.float32).min

// ---- warp / block reductions -----------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of K values per thread; result valid in every thread.  `scratch` holds
// K * 32 elements of T.  Fixed reduction order => deterministic.
template <typename T, int K>
__device__ __forceinline__ void block_sum(T (&v)[K], T* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();  // scratch may still be read from a previous call
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) scratch[k * 32 + warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    T x = lane < nwarps ? scratch[k * 32 + lane] : T(0);
    v[k] = warp_sum(x);
  }
}

// ---- streaming loads / stores (each byte is touched once: keep it out of L1) -----------------
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_u4(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while the kernel ahead of
// it on the stream is still running: its launch latency and prologue overlap that kernel's tail.  pdl_wait()
// blocks until the kernel ahead has completed and its memory is visible (a no-op without the attribute);
// pdl_trigger() in the kernel ahead lets the dependent grid be scheduled once every CTA has called it or exited.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ---- element access by runtime dtype -------------------------------------------------------
template <int DT>
struct IndexLoad;
template <>
struct IndexLoad<B200RL_U8> {
  typedef uint8_t type;
};
template <>
struct IndexLoad<B200RL_I32> {
  typedef int32_t type;
};
template <>
struct IndexLoad<B200RL_I64> {
  typedef int64_t type;
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

}  // namespace b200rl

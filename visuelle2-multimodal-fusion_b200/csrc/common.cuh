// Shared device helpers for libv2f_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "v2f.h"

// host-side launch counter (v2f_launch_count); the only process-global the library keeps
extern long long g_v2f_launches;

#define V2F_CHECK_LAUNCH()                                   \
  do {                                                       \
    ++g_v2f_launches;                                        \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return V2F_ERR_LAUNCH;           \
  } while (0)

#define V2F_TRY(expr)              \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != V2F_OK) return rc__; \
  } while (0)

#define V2F_REQUIRE(cond, code) \
  do {                          \
    if (!(cond)) return (code); \
  } while (0)

namespace v2f {

void prof_begin(int id, cudaStream_t st);
void prof_end(int id, cudaStream_t st);
void prof_bytes(int id, long long bytes);

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// tanh with ~3e-7 absolute error from two MUFU ops (ex2, rcp): 1 - 2/(e^{2x}+1).
// Saturates correctly (e^{2x} -> inf gives 1, -> 0 gives -1).
__device__ __forceinline__ float tanh_acc(float x) {
  float t = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, t + 1.0f);
}
// Lean variants for the streaming attention kernels (issue-bound there): raw ex2/rcp without the
// denormal / huge-argument slow paths (ftz saturates correctly: e^{2x} -> inf gives 1, -> 0 gives -1),
// |err| ~ 4e-7; APPROX = the single-instruction MUFU.TANH (|err| ~ 5e-4, tensor-core mode only).
template <bool APPROX>
__device__ __forceinline__ float tanh_fast(float x) {
  if (APPROX) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  }
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 2.885390081777927f));   // e^{2x}
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}
// sigmoid, same construction.
__device__ __forceinline__ float sigmoid_acc(float x) {
  return __fdividef(1.0f, 1.0f + __expf(-x));
}

// full-precision versions for the (cheap, error-accumulating) recurrent gates
__device__ __forceinline__ float sigmoid_full(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float tanh_full(float x) { return tanhf(x); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// streaming 128-bit load that does not pollute L1 (tiles are read once per step)
__device__ __forceinline__ float4 ld4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace v2f

// mbarrier / bulk-copy (TMA) primitives shared by the tcgen05 GEMM and the streaming attention
// kernels.  Every wait is bounded: a protocol bug traps (kernel error) instead of hanging the GPU.
#pragma once
#include "common.cuh"

namespace v2f {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// Round a pointer into the dynamic shared-memory array up to `align` bytes BY OFFSET: pointer arithmetic on the
// __shared__ symbol keeps the shared state space, so the accesses compile to LDS / STS.  (Rounding through uintptr_t
// yields a generic pointer and every access becomes a generic LD / ST.)
__device__ __forceinline__ uint8_t* smem_align(uint8_t* p, uint32_t align) {
  return p + ((align - (smem_u32(p) & (align - 1))) & (align - 1));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 24); i++)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
// Warp-collective wait: lane 0 polls, the other lanes park at the warp barrier (512 threads spinning on try_wait keep
// the shared-memory synchronisation unit busy while TMA is trying to complete transactions on the same barriers);
// __syncwarp orders the lanes' later shared-memory reads after lane 0's acquire.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (UBLKCP in SASS)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// sub-block barrier over `count` threads (consumer warps only)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

}  // namespace v2f

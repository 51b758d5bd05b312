// tcgen05 / TMA device helpers shared by the tensor-core GEMM (gemm_tc.cu) and the row-team persistent decoder
// (decode_team.cu): tensor-map loads, shared-memory matrix descriptors (K-major, 128-byte swizzle), instruction
// descriptors, MMA issue and commit.  sm_100a.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "async.cuh"

namespace v2f {

constexpr int TC_BM = 128;           // UMMA_M

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address  [0,14)
  d |= (uint64_t)0 << 16;                           // leading byte offset (ignored for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                           // layout type: SWIZZLE_128B
  return d;
}

// instruction descriptor, kind::f16 (bf16 x bf16 -> f32) or kind::tf32, both operands K-major
template <int KIND>
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  uint32_t d = 0;
  d |= 1u << 4;                                     // D format: F32
  const uint32_t fmt = KIND == 0 ? 1u : 2u;         // kind::f16: 1 = BF16; kind::tf32: 2 = TF32
  d |= fmt << 7;                                    // A format
  d |= fmt << 10;                                   // B format
  // bit 15 / 16: A / B major = 0 (K-major)
  d |= (uint32_t)(n >> 3) << 17;                    // N >> 3
  d |= (uint32_t)(TC_BM >> 4) << 24;                // M >> 4
  return d;
}

template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (KIND == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}


// host: tensor map over batch x row-major [rows, cols] (cols contiguous, row stride ld, batch stride bs, in elements),
// box = [1, box_rows, 128 bytes], 128-byte swizzle.  kind 0: bf16, 1: fp32.  (gemm_tc.cu)
int tc_make_map(CUtensorMap* map, int kind, const void* ptr, long long rows, long long cols, long long ld,
                long long batch, long long bs, int box_rows);

}  // namespace v2f

// Micro-benchmark of the access pattern of the row-team decoder's product phases: all CTAs of a team meet at a barrier,
// then every CTA loads the SAME 64 x 512 bf16 block (8 TMA boxes of 8 KB, one per issuing warp) -- optionally after
// every CTA has rewritten its own 16-byte column slice of that block (what the GRU gate phase does to hb).
// Prints the time from the barrier to the first / last box landed.  Cooperative launch, 128 CTAs = 2 teams of 64.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t par) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* map, uint64_t* b, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(s32(dst)), "l"(map), "r"(s32(b)), "r"(c0), "r"(c1), "r"(0) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// mode bit 3: no cross-proxy fence; bit 4: no global barrier at all; bit 5: poll with relaxed loads + nanosleep, one acquire fence at the end
__device__ void team_barrier(unsigned* ctr, unsigned& epoch, int team_size, int mode) {
  __syncthreads();
  ++epoch;
  if (mode & 16) return;
  if (threadIdx.x == 0) {
    if (mode & 128) { __threadfence(); asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory"); }
    else asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned v = 0;
    if (mode & 32) {
      while (true) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= epoch * team_size) break;
        __nanosleep(40);
      }
      if (!(mode & 64)) asm volatile("fence.acquire.gpu;" ::: "memory");
    } else {
      while (v < epoch * team_size) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    }
    if (!(mode & 8)) asm volatile("fence.proxy.async.global;" ::: "memory");
  }
  __syncthreads();
}

// mode bit 0: rewrite the block before each burst (each CTA its 8 columns of all 64 rows); bit 1: each CTA loads its OWN
// block instead of the team's; bit 2: stagger: CTA c waits c * 32 cycles after the barrier before issuing
__global__ void __launch_bounds__(576, 1) burst_kernel(const __grid_constant__ CUtensorMap map, __nv_bfloat16* hb, unsigned* ctr, int iters, int mode,
                                                       unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint8_t* buf = sm + 1024;
  const int team = blockIdx.x / 64, c = blockIdx.x % 64, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; i++) mb_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned epoch = 0;
  unsigned long long first = 0, last = 0, issue = 0;
  const int row0 = (mode & 2) ? blockIdx.x * 64 : team * 64;
  for (int it = 0; it < iters; it++) {
    if ((mode & 1) && threadIdx.x < 512) {          // thread (row nl, column gu) like the gate phase
      const int nl = threadIdx.x >> 3, gu = threadIdx.x & 7;
      hb[(size_t)(team * 64 + nl) * 512 + 8 * c + gu] = __float2bfloat16_rn((float)(it + gu));
    }
    team_barrier(ctr + team * 32, epoch, 64, mode);
    if (mode & 256) { const unsigned long long w0 = gtime(); while (gtime() - w0 < 2000) {} __syncthreads(); }
    unsigned long long tprobe = 0;
    if ((mode & 512) && threadIdx.x == 0) {
      const unsigned long long a0 = gtime();
      unsigned v;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr + 512 + blockIdx.x) : "memory");
      tprobe = gtime() - a0 + (v & 0);
    }
    const unsigned long long t0 = gtime();
    if (mode & 4) { const long long w0 = clock64(); while (clock64() - w0 < c * 32) {} }
    if (warp < 8 && lane == 0) {
      mb_expect(&full[warp], 8192);
      tma3(buf + warp * 8192, &map, &full[warp], warp * 64, row0);
    }
    const unsigned long long t1 = gtime();
    if (threadIdx.x == 8 * 32) {       // a ninth warp watches the boxes land, in order
      unsigned long long tf = 0, tl = 0;
      for (int i = 0; i < 8; i++) {
        mb_wait(&full[i], it & 1);
        const unsigned long long t = gtime();
        if (i == 0) tf = t;
        tl = t;
      }
      first += tf - t0;
      last += tl - t0;
    }
    if (threadIdx.x == 0) issue += (mode & 512) ? tprobe : t1 - t0;
    __syncthreads();
  }
  if (threadIdx.x == 8 * 32) { out[blockIdx.x * 3] = first; out[blockIdx.x * 3 + 1] = last; }
  if (threadIdx.x == 0) out[blockIdx.x * 3 + 2] = issue;
}

int main() {
  const int G = 128, iters = 50;
  __nv_bfloat16* d;
  cudaMalloc(&d, (size_t)G * 64 * 512 * 2);
  cudaMemset(d, 0, (size_t)G * 64 * 512 * 2);
  unsigned* ctr;
  cudaMalloc(&ctr, 4096);
  unsigned long long* out;
  cudaMalloc(&out, G * 3 * 8);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr);
  Fn fn = (Fn)fp;
  CUtensorMap map;
  cuuint64_t dims[3] = {512, (cuuint64_t)G * 64, 1};
  cuuint64_t strides[2] = {1024, (cuuint64_t)G * 64 * 1024};
  cuuint32_t box[3] = {64, 64, 1}, es[3] = {1, 1, 1};
  fn(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const size_t smem = 1024 + 8 * 8192;
  cudaFuncSetAttribute(burst_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[] = {"team block, read-only", "team block, rewritten each round", "own block, read-only", "own block (rewrite n/a)",
                         "team block, read-only, staggered issue", "team block, rewritten, staggered issue"};
  const int modes[] = {0, 16, 256, 512, 528};
  for (int mi = 0; mi < 5; mi++) {
    int mode = modes[mi];
    for (int rep = 0; rep < 2; rep++) {
      cudaMemset(ctr, 0, 4096);
      int it = iters;
      void* args[] = {(void*)&map, (void*)&d, (void*)&ctr, (void*)&it, (void*)&mode, (void*)&out};
      cudaError_t e = cudaLaunchCooperativeKernel((void*)burst_kernel, dim3(G), dim3(576), args, smem, 0);
      cudaDeviceSynchronize();
      if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch error\n"); return 1; }
    }
    unsigned long long h[G * 3];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double f = 0, l = 0, is = 0;
    for (int i = 0; i < G; i++) { f += h[3 * i]; l += h[3 * i + 1]; is += h[3 * i + 2]; }
    printf("mode %d: issue %.0f ns, first box landed after %.0f ns, all 8 boxes (64 KB) after %.0f ns\n", mode, is / G / iters, f / G / iters, l / G / iters);
  }
  return 0;
}

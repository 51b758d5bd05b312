// Micro-benchmark: per-SM delivery rate of cp.async.bulk (global -> shared) from an L2-resident buffer as a function of
// the copy size and the number of copies in flight; and of plain 16-byte LDGs for comparison.  One CTA per SM, each CTA
// streams its own region (no sharing).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_probe bulk_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t par) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(n), "r"(s32(b)) : "memory");
}

// mode 0: bulk copies, `depth` slots of `bytes`; warp 0 lane 0 produces, warp 1 lane 0 consumes (wait + free)
__global__ void __launch_bounds__(64, 1) bulk_kernel(const uint8_t* base, size_t region, int bytes, int depth, int ncopies, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 16;
  uint8_t* buf = sm + 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; i++) { mb_init(&full[i], 1); mb_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* src = base + (size_t)blockIdx.x * region;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int i = 0; i < ncopies; i++) {
      const int s = i % depth, r = i / depth;
      mb_wait(&empty[s], (r & 1) ^ 1);
      mb_expect(&full[s], bytes);
      bulk(buf + (size_t)s * bytes, src + ((size_t)i * bytes) % region, bytes, &full[s]);
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < ncopies; i++) {
      const int s = i % depth, r = i / depth;
      mb_wait(&full[s], r & 1);
      mb_arrive(&empty[s]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

// mode 2: `issuers` producer warps (lane 0 each), copy i belongs to issuer i % issuers; one consumer warp per issuer
__global__ void __launch_bounds__(512, 1) bulk_multi_kernel(const uint8_t* base, size_t region, int bytes, int depth, int ncopies, int issuers, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 64;
  uint8_t* buf = sm + 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth * issuers; i++) { mb_init(&full[i], 1); mb_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* src = base + (size_t)blockIdx.x * region;
  const long long t0 = clock64();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0 && w < issuers) {
    int k = 0;
    for (int i = w; i < ncopies; i += issuers, k++) {
      const int s = w * depth + k % depth, r = k / depth;
      mb_wait(&empty[s], (r & 1) ^ 1);
      mb_expect(&full[s], bytes);
      bulk(buf + (size_t)s * bytes, src + ((size_t)i * bytes) % region, bytes, &full[s]);
    }
  } else if (lane == 0 && w >= 8 && w - 8 < issuers) {
    const int ww = w - 8;
    int k = 0;
    for (int i = ww; i < ncopies; i += issuers, k++) {
      const int s = ww * depth + k % depth, r = k / depth;
      mb_wait(&full[s], r & 1);
      mb_arrive(&empty[s]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

// mode 3: tensor-map TMA loads (box 64 rows x 64 bf16, 128-byte swizzle = 8 KB) from `issuers` warps; rows of a
// [G*rows_per_cta, 512] bf16 matrix; shared != 0: every CTA reads the SAME 64 rows (hot lines)
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* map, uint64_t* b, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(s32(dst)), "l"(map), "r"(s32(b)), "r"(c0), "r"(c1), "r"(0) : "memory");
}
__global__ void __launch_bounds__(512, 1) tma_kernel(const __grid_constant__ CUtensorMap map, int depth, int ncopies, int issuers, int shared, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 64;
  uint8_t* buf = sm + 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth * issuers; i++) { mb_init(&full[i], 1); mb_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int row0 = shared ? 0 : blockIdx.x * 64;
  const long long t0 = clock64();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0 && w < issuers) {
    int k = 0;
    for (int i = w; i < ncopies; i += issuers, k++) {
      const int s = w * depth + k % depth, r = k / depth;
      mb_wait(&empty[s], (r & 1) ^ 1);
      mb_expect(&full[s], 8192);
      tma3(buf + (size_t)s * 8192, &map, &full[s], (i % 8) * 64, row0);
    }
  } else if (lane == 0 && w >= 8 && w - 8 < issuers) {
    const int ww = w - 8;
    int k = 0;
    for (int i = ww; i < ncopies; i += issuers, k++) {
      const int s = ww * depth + k % depth, r = k / depth;
      mb_wait(&full[s], r & 1);
      mb_arrive(&empty[s]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

// mode 1: every thread keeps `depth` independent 16-byte loads in flight (register ring), 512 threads
__global__ void __launch_bounds__(512, 1) ldg_kernel(const uint8_t* base, size_t region, int iters, long long* out, float* sink) {
  const uint4* src = reinterpret_cast<const uint4*>(base + (size_t)blockIdx.x * region);
  const size_t n16 = region / 16;
  const long long t0 = clock64();
  uint4 acc = make_uint4(0, 0, 0, 0);
  size_t idx = threadIdx.x;
  for (int i = 0; i < iters; i++) {
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      v[k] = __ldcg(src + idx);
      idx += 512;
      if (idx >= n16) idx -= n16;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) { acc.x ^= v[k].x; acc.y += v[k].y; acc.z ^= v[k].z; acc.w += v[k].w; }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  if (acc.x == 0x12345678 && acc.y == 1) sink[0] = 1.f;
}

int main() {
  const int G = 128;
  const size_t region = 304 * 1024;                  // per-CTA stream (one row's bf16 tiles)
  uint8_t* d;
  cudaMalloc(&d, G * region);
  cudaMemset(d, 1, G * region);
  long long* out;
  cudaMalloc(&out, G * sizeof(long long));
  float* sink;
  cudaMalloc(&sink, 4);
  long long h[G];
  int dev = 0, khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("# clock %d kHz; region %zu KB per CTA, %d CTAs (total %.1f MB, L2-resident after the first pass)\n", khz, region / 1024, G, G * region / 1e6);
  const int sizes[] = {1024, 2048, 4096, 8192, 16384, 32768};
  const int depths[] = {1, 2, 4, 8, 16};
  for (int bytes : sizes)
    for (int depth : depths) {
      if ((size_t)bytes * depth > 190 * 1024) continue;
      const int ncopies = (int)(4 * region / bytes);      // 4 passes over the region
      for (int rep = 0; rep < 2; rep++) {
        bulk_kernel<<<G, 64, 1024 + (size_t)bytes * depth>>>(d, region, bytes, depth, ncopies, out);
        cudaDeviceSynchronize();
      }
      cudaError_t e = cudaGetLastError();
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double mx = 0, av = 0;
      for (int i = 0; i < G; i++) { av += h[i]; if (h[i] > mx) mx = h[i]; }
      av /= G;
      const double tot = (double)ncopies * bytes;
      printf("bulk  bytes %6d depth %2d : %.1f B/clk/SM avg (%.1f at the slowest SM), %.0f cycles per copy%s\n", bytes, depth,
             tot / av, tot / mx, av / ncopies, e == cudaSuccess ? "" : "  ERROR");
    }
  cudaFuncSetAttribute(bulk_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int bytes : {2048, 8192, 16384})
    for (int issuers : {1, 2, 4, 8})
      for (int depth : {2, 4}) {
        if ((size_t)bytes * depth * issuers > 190 * 1024) continue;
        const int ncopies = (int)(4 * region / bytes);
        for (int rep = 0; rep < 2; rep++) {
          bulk_multi_kernel<<<G, 512, 1024 + (size_t)bytes * depth * issuers>>>(d, region, bytes, depth, ncopies, issuers, out);
          cudaDeviceSynchronize();
        }
        cudaError_t e = cudaGetLastError();
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double av = 0;
        for (int i = 0; i < G; i++) av += h[i];
        av /= G;
        printf("multi bytes %6d issuers %d depth %d : %.1f B/clk/SM, %.0f cycles per copy%s\n", bytes, issuers, depth,
               (double)ncopies * bytes / av, av / ncopies, e == cudaSuccess ? "" : "  ERROR");
      }
  {
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
    cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int shared : {0, 1})
      for (int issuers : {1, 2, 4, 8})
        for (int depth : {1, 2}) {
          const int ncopies = 8 * 40;          // 40 sweeps over the 64 x 512 block
          for (int rep = 0; rep < 2; rep++) {
            tma_kernel<<<G, 512, 1024 + (size_t)8192 * depth * issuers>>>(map, depth, ncopies, issuers, shared, out);
            cudaDeviceSynchronize();
          }
          cudaError_t e = cudaGetLastError();
          cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
          double av = 0;
          for (int i = 0; i < G; i++) av += h[i];
          av /= G;
          printf("tma   8 KB boxes, %s rows, issuers %d depth %d : %.1f B/clk/SM, %.0f cycles per copy%s\n", shared ? "SHARED" : "own   ",
                 issuers, depth, (double)ncopies * 8192 / av, av / ncopies, e == cudaSuccess ? "" : "  ERROR");
        }
  }
  for (int rep = 0; rep < 2; rep++) {
    const int iters = (int)(4 * region / (512 * 16 * 8));
    ldg_kernel<<<G, 512>>>(d, region, iters, out, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double av = 0;
    for (int i = 0; i < G; i++) av += h[i];
    av /= G;
    printf("ldg   512 threads x 8 x 16 B in flight: %.1f B/clk/SM\n", (double)iters * 512 * 16 * 8 / av);
  }
  return 0;
}

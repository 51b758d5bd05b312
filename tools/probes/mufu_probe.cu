// Throughput of the special-function pipe on sm_100a: tanh.approx / ex2.approx / rcp.approx per clock and SM, measured
// with 8 independent dependency chains per thread, 1024 threads per SM resident.  Grounds the "MUFU-bound" statements
// about the energies pass of the decoder and tilegrad_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mufu_probe tools/probes/mufu_probe.cu && tools/probes/mufu_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else y = fmaf(x, 1.0009765625f, 0.25f);        // FMA-pipe reference
  return y;
}

template <int OP>
__global__ void __launch_bounds__(256) probe(float* out, int iters, long long* clocks) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = 0.001f * (threadIdx.x + 1) + 0.1f * i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = op<OP>(v[i]);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) s += v[i];
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const char* name, int sms) {
  const int blocks = sms * 4, iters = 4096;
  float* out;
  long long* clk;
  cudaMalloc(&out, sizeof(float) * blocks * 256);
  cudaMalloc(&clk, sizeof(long long) * blocks);
  probe<OP><<<blocks, 256>>>(out, iters, clk);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  probe<OP><<<blocks, 256>>>(out, iters, clk);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  long long* h = new long long[blocks];
  cudaMemcpy(h, clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; i++) avg += (double)h[i];
  avg /= blocks;
  const double ops_per_sm = 4.0 * 256 * 8.0 * iters;      // 4 resident CTAs per SM
  printf("%-12s %8.3f ms   %6.2f ops/clk/SM (block clocks)   %7.2f G ops/s/SM (wall)\n", name, ms, ops_per_sm / avg,
         ops_per_sm / (ms * 1e-3) / 1e9);
  delete[] h;
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  run<0>("tanh.approx", sms);
  run<1>("ex2.approx", sms);
  run<2>("rcp.approx", sms);
  run<3>("fma", sms);
  return 0;
}

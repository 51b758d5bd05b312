// Device side of the image transform: ToTensor + Normalize from the decoded uint8 pixels.  sm_100a.
//
// Reference: /root/reference/dataset_fusion.py:50-65 -- per item, in a DataLoader worker: PIL decode ->
// Resize((299,299)) -> ToTensor() (uint8 HWC -> float CHW / 255) -> Normalize(mean, std); the batch then crosses
// PCIe as fp32 [B,3,299,299] (137 MB for 128 items).  Decode and resize stay on the host (PIL); the arithmetic moves
// here so that the batch crosses PCIe as uint8 [B,299,299,3] (34 MB) and lands directly in the layout and type the
// bf16 channels_last trunk consumes: NHWC uint8 in, NHWC (= channels_last NCHW) bf16 or fp32 out,
//     out[n,h,w,c] = (float(u8) / 255 - mean[c]) / std[c]        (same operation order as torchvision: bit-identical
// to Normalize(ToTensor(img)) in fp32, then one rounding to bf16).
// HBM-bound elementwise pass: 16 pixels-channels (16 B) in, 32 B (bf16) / 64 B (fp32) out per thread.
#include <cuda_bf16.h>

#include "common.cuh"

namespace v2f {

template <bool BF16>
__global__ void __launch_bounds__(256)
image_prep_kernel(long long n16, long long n, int C, const uint8_t* __restrict__ in, const float* __restrict__ mean,
                  const float* __restrict__ stdv, void* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n16) return;
  const long long e0 = i * 16;
  uint8_t px[16];
  if (e0 + 16 <= n) {
    *reinterpret_cast<uint4*>(px) = *reinterpret_cast<const uint4*>(in + e0);
  } else {
    for (int k = 0; k < 16; k++) px[k] = e0 + k < n ? in[e0 + k] : 0;
  }
  float v[16];
  int c = (int)(e0 % C);
#pragma unroll
  for (int k = 0; k < 16; k++) {
    v[k] = ((float)px[k] / 255.0f - mean[c]) / stdv[c];
    c = c + 1 == C ? 0 : c + 1;
  }
  if (e0 + 16 <= n) {
    if (BF16) {
      __nv_bfloat162 o[8];
#pragma unroll
      for (int k = 0; k < 8; k++) o[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
      uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + e0);
      op[0] = *reinterpret_cast<uint4*>(&o[0]);
      op[1] = *reinterpret_cast<uint4*>(&o[4]);
    } else {
      float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + e0);
#pragma unroll
      for (int k = 0; k < 4; k++) op[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    }
  } else {
    for (int k = 0; k < 16 && e0 + k < n; k++) {
      if (BF16) reinterpret_cast<__nv_bfloat16*>(out)[e0 + k] = __float2bfloat16_rn(v[k]);
      else reinterpret_cast<float*>(out)[e0 + k] = v[k];
    }
  }
}

}  // namespace v2f

// in: uint8 [pixels, C] (NHWC flattened), out: bf16 (out_bf16 = 1) or fp32, same shape; mean/std: device [C].
extern "C" int v2f_image_normalize_u8(long long pixels, int C, const void* in, const float* mean, const float* stdv,
                                      int out_bf16, void* out, void* st) {
  V2F_REQUIRE(pixels > 0 && C > 0 && in && mean && stdv && out, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(v2f::aligned16(in) && v2f::aligned16(out), V2F_ERR_ALIGN);
  const long long n = pixels * C, n16 = (n + 15) / 16;
  const unsigned grid = (unsigned)((n16 + 255) / 256);
  if (out_bf16)
    v2f::image_prep_kernel<true><<<grid, 256, 0, (cudaStream_t)st>>>(n16, n, C, (const uint8_t*)in, mean, stdv, out);
  else
    v2f::image_prep_kernel<false><<<grid, 256, 0, (cudaStream_t)st>>>(n16, n, C, (const uint8_t*)in, mean, stdv, out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

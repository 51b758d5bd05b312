// Scaled-dot-product attention core for short sequences, one CTA per (batch, head).  sm_100a.
//
// Reference arithmetic: torch.nn.MultiheadAttention as used by ts_self_attention
// (/root/reference/models/CrossAttnRNN210.py:126,176-179: 4 heads, L=52, attn-dropout .1) and by
// nn.TransformerEncoderLayer / nn.TransformerDecoderLayer in
// /root/reference/models/GTM_Visuelle2.py:52-53,200-202 (block-diagonal / causal additive masks).
//   S = scale * Q K^T + mask ; P = softmax_row(S) ; O = (P * drop) V
// The whole (Lq x Lk) problem lives in shared memory; P (before dropout) is saved for backward.
#include "common.cuh"

namespace v2f {

constexpr int SD_THREADS = 512;   // 16 warps: the 13 x 32 register tiles of the [52 x 128] products fit one round
constexpr int SD_MAXL = 64;

struct SdpaArgs {
  int B, heads, Lq, Lk, hd;
  const float *q, *k, *v;
  int ldq, ldk, ldv;
  long long bsq, bsk, bsv;
  float* o;
  int ldo;
  long long bso;
  const float *mask, *drop;
  float* P;
  float scale;
};

__device__ __forceinline__ void load_rows(float* dst, const float* src, int L, int hd, int ld, int ldp) {
  if (((hd | ld) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // 128-bit global loads, several in flight per thread (the padded shared rows take scalar stores)
    const int hq = hd >> 2;
    for (int i = threadIdx.x; i < L * hq; i += SD_THREADS) {
      const int l = i / hq, d = (i - l * hq) << 2;
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + (long long)l * ld + d));
      float* o = dst + l * ldp + d;
      o[0] = v.x;
      o[1] = v.y;
      o[2] = v.z;
      o[3] = v.w;
    }
    return;
  }
  for (int i = threadIdx.x; i < L * hd; i += SD_THREADS) {
    const int l = i / hd, d = i - l * hd;
    dst[l * ldp + d] = src[(long long)l * ld + d];
  }
}

// C[m][n] = sum_k A(m,k) B(n,k) over operands in shared memory, A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk]:
// each thread owns a 4 x 4 register tile (8 shared-memory loads per 16 FMAs instead of 2 per FMA), rows tm + r*MT and
// columns tn + c*NT interleaved so that the B loads of a warp fall in consecutive banks and its A loads broadcast.
// Every output keeps the k = 0..K-1 summation order of a plain loop.  epi(m, n, acc) stores the result.
template <class Epi>
__device__ __forceinline__ void smem_mm(int M, int N, int K, const float* __restrict__ A, int sam, int sak,
                                        const float* __restrict__ B, int sbn, int sbk, Epi epi) {
  const int MT = (M + 3) >> 2, NT = (N + 3) >> 2;
  for (int t = threadIdx.x; t < MT * NT; t += SD_THREADS) {
    const int tm = t / NT, tn = t - tm * NT;
    const float* ap[4];
    const float* bp[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      ap[r] = A + min(tm + r * MT, M - 1) * sam;       // clamped: out-of-range rows compute a discarded copy
      bp[r] = B + min(tn + r * NT, N - 1) * sbn;
    }
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[r][c] = 0.f;
#pragma unroll 2
    for (int k = 0; k < K; k++) {
      float a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        a[r] = ap[r][k * sak];
        b[r] = bp[r][k * sbk];
      }
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int c = 0; c < 4; c++)
        if (tm + r * MT < M && tn + c * NT < N) epi(tm + r * MT, tn + c * NT, acc[r][c]);
  }
}

__global__ void __launch_bounds__(SD_THREADS)
sdpa_fwd_kernel(SdpaArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / a.heads, hh = blockIdx.x % a.heads;
  const int Lq = a.Lq, Lk = a.Lk, hd = a.hd, ldp = hd + 1;
  float* Q = sm;
  float* K = Q + Lq * ldp;
  float* V = K + Lk * ldp;
  float* S = V + Lk * ldp;   // [Lq][Lk]
  load_rows(Q, a.q + b * a.bsq + hh * hd, Lq, hd, a.ldq, ldp);
  load_rows(K, a.k + b * a.bsk + hh * hd, Lk, hd, a.ldk, ldp);
  load_rows(V, a.v + b * a.bsv + hh * hd, Lk, hd, a.ldv, ldp);
  __syncthreads();
  smem_mm(Lq, Lk, hd, Q, ldp, 1, K, ldp, 1, [&](int i, int j, float acc) {
    acc *= a.scale;
    if (a.mask) acc += a.mask[i * Lk + j];
    S[i * Lk + j] = acc;
  });
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long pbase = ((long long)blockIdx.x) * Lq * Lk;
  for (int i = warp; i < Lq; i += SD_THREADS / 32) {
    float m = -INFINITY;
    for (int j = lane; j < Lk; j += 32) m = fmaxf(m, S[i * Lk + j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float e = expf(S[i * Lk + j] - m);
      S[i * Lk + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < Lk; j += 32) {
      float p = S[i * Lk + j] * inv;
      a.P[pbase + i * Lk + j] = p;
      if (a.drop) p *= a.drop[pbase + i * Lk + j];
      S[i * Lk + j] = p;
    }
  }
  __syncthreads();
  float* op = a.o + b * a.bso + hh * hd;
  smem_mm(Lq, hd, Lk, S, Lk, 1, V, 1, ldp, [&](int i, int d, float acc) { op[(long long)i * a.ldo + d] = acc; });
}

struct SdpaBwdArgs {
  int B, heads, Lq, Lk, hd;
  const float *q, *k, *v, *dO;
  int ldq, ldk, ldv, ldo;
  long long bsq, bsk, bsv, bso;
  const float *drop, *P;
  float *dq, *dk, *dv;
  int lddq, lddk, lddv;
  long long bsdq, bsdk, bsdv;
  float scale;
};

__global__ void __launch_bounds__(SD_THREADS)
sdpa_bwd_kernel(SdpaBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / a.heads, hh = blockIdx.x % a.heads;
  const int Lq = a.Lq, Lk = a.Lk, hd = a.hd, ldp = hd + 1;
  float* Q = sm;
  float* K = Q + Lq * ldp;
  float* V = K + Lk * ldp;
  float* dO = V + Lk * ldp;    // [Lq][ldp]
  float* Pd = dO + Lq * ldp;   // [Lq][Lk]  P * drop
  float* dS = Pd + Lq * Lk;    // [Lq][Lk]
  load_rows(Q, a.q + b * a.bsq + hh * hd, Lq, hd, a.ldq, ldp);
  load_rows(K, a.k + b * a.bsk + hh * hd, Lk, hd, a.ldk, ldp);
  load_rows(V, a.v + b * a.bsv + hh * hd, Lk, hd, a.ldv, ldp);
  load_rows(dO, a.dO + b * a.bso + hh * hd, Lq, hd, a.ldo, ldp);
  const long long pbase = ((long long)blockIdx.x) * Lq * Lk;
  __syncthreads();
  // dP = (dO V^T) * drop ; keep P*drop for dV
  smem_mm(Lq, Lk, hd, dO, ldp, 1, V, ldp, 1, [&](int i, int j, float acc) {
    const int idx = i * Lk + j;
    const float dr = a.drop ? a.drop[pbase + idx] : 1.f;
    const float p = a.P[pbase + idx];
    Pd[idx] = p * dr;
    dS[idx] = acc * dr;    // dP for now
  });
  __syncthreads();
  // dV[j,d] = sum_i Pd[i,j] dO[i,d]
  float* dvp = a.dv + b * a.bsdv + hh * hd;
  smem_mm(Lk, hd, Lq, Pd, 1, Lk, dO, 1, ldp, [&](int j, int d, float acc) { dvp[(long long)j * a.lddv + d] = acc; });
  // dS = P * (dP - sum_j P dP)   (P without dropout)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < Lq; i += SD_THREADS / 32) {
    float dot = 0.f;
    for (int j = lane; j < Lk; j += 32) dot = fmaf(a.P[pbase + i * Lk + j], dS[i * Lk + j], dot);
    dot = warp_sum(dot);
    for (int j = lane; j < Lk; j += 32)
      dS[i * Lk + j] = a.P[pbase + i * Lk + j] * (dS[i * Lk + j] - dot) * a.scale;
  }
  __syncthreads();
  float* dqp = a.dq + b * a.bsdq + hh * hd;
  smem_mm(Lq, hd, Lk, dS, Lk, 1, K, 1, ldp, [&](int i, int d, float acc) { dqp[(long long)i * a.lddq + d] = acc; });
  float* dkp = a.dk + b * a.bsdk + hh * hd;
  smem_mm(Lk, hd, Lq, dS, 1, Lk, Q, 1, ldp, [&](int j, int d, float acc) { dkp[(long long)j * a.lddk + d] = acc; });
}

}  // namespace v2f

using namespace v2f;

extern "C" int v2f_sdpa_fwd(int B, int heads, int Lq, int Lk, int hd, const float* q, int ldq,
                            long long bsq, const float* k, int ldk, long long bsk, const float* v,
                            int ldv, long long bsv, float* o, int ldo, long long bso,
                            const float* mask, const float* drop, float* P, float scale, void* st) {
  V2F_REQUIRE(B > 0 && heads > 0 && hd > 0, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(Lq > 0 && Lk > 0 && Lq <= SD_MAXL && Lk <= SD_MAXL && hd <= 256, V2F_ERR_UNSUPPORTED);
  V2F_REQUIRE(q && k && v && o && P, V2F_ERR_BAD_ARG);
  const size_t smem = sizeof(float) * ((size_t)(Lq + 2 * Lk) * (hd + 1) + (size_t)Lq * Lk);
  V2F_REQUIRE(smem <= 220 * 1024, V2F_ERR_UNSUPPORTED);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(sdpa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  SdpaArgs a{B, heads, Lq, Lk, hd, q, k, v, ldq, ldk, ldv, bsq, bsk, bsv, o, ldo, bso, mask, drop, P, scale};
  sdpa_fwd_kernel<<<B * heads, SD_THREADS, smem, (cudaStream_t)st>>>(a);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_sdpa_bwd(int B, int heads, int Lq, int Lk, int hd, const float* q, int ldq,
                            long long bsq, const float* k, int ldk, long long bsk, const float* v,
                            int ldv, long long bsv, const float* dO, int ldo, long long bso,
                            const float* drop, const float* P, float* dq, int lddq, long long bsdq,
                            float* dk, int lddk, long long bsdk, float* dv, int lddv,
                            long long bsdv, float scale, void* st) {
  V2F_REQUIRE(B > 0 && heads > 0 && hd > 0, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(Lq > 0 && Lk > 0 && Lq <= SD_MAXL && Lk <= SD_MAXL && hd <= 256, V2F_ERR_UNSUPPORTED);
  V2F_REQUIRE(q && k && v && dO && P && dq && dk && dv, V2F_ERR_BAD_ARG);
  const size_t smem = sizeof(float) * ((size_t)(2 * Lq + 2 * Lk) * (hd + 1) + 2 * (size_t)Lq * Lk);
  V2F_REQUIRE(smem <= 220 * 1024, V2F_ERR_UNSUPPORTED);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(sdpa_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  SdpaBwdArgs a{B, heads, Lq, Lk, hd, q, k, v, dO, ldq, ldk, ldv, ldo, bsq, bsk, bsv, bso, drop, P,
                dq, dk, dv, lddq, lddk, lddv, bsdq, bsdk, bsdv, scale};
  sdpa_bwd_kernel<<<B * heads, SD_THREADS, smem, (cudaStream_t)st>>>(a);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

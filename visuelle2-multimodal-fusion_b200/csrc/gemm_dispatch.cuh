// Host-side GEMM dispatch shared by the recurrent orchestrators: exact fp32 CUDA-core kernel
// (precision 0) or tcgen05 tensor cores on tf32 operands (precision 1), bringing NN / TN products
// into the K-major form with transposed copies held in caller-provided scratch.
#pragma once
#include <cstdlib>

#include "common.cuh"

extern "C" int v2f_gemm_f32(int, int, int, int, int, const float*, int, long long, const float*, int,
                            long long, float*, int, long long, int, const float*, float, int, void*);
extern "C" int v2f_gemm_tc(int, int, int, int, const void*, long long, const void*, long long, float*,
                           long long, const float*, float, int, int, void*);
extern "C" int v2f_transpose(int, int, const void*, long long, int, void*, long long, int, void*);
extern "C" int v2f_colsum_f32(int, int, const float*, int, float*, float, void*);

namespace v2f {

struct GemmCtx {
  int tc;               // 0: fp32 SIMT, 1: tf32 tensor cores
  float* ws;            // scratch for transposed operands (tc mode)
  long long ws_floats;
  void* st;
};

static inline bool tc_ok(const void* p, long long ld) { return aligned16(p) && (ld & 3) == 0; }

// C[M,N] = A[M,K] B[N,K]^T (+bias) (+beta C)
static inline int gemm_nt(const GemmCtx& g, int M, int N, int K, const float* A, int lda, const float* B,
                          int ldb, float* C, int ldc, const float* bias, float beta) {
  if (g.tc && K >= 8 && tc_ok(A, lda) && tc_ok(B, ldb))
    return v2f_gemm_tc(1, M, N, K, A, lda, B, ldb, C, ldc, bias, beta, 0, 1, g.st);
  return v2f_gemm_f32(0, 1, M, N, K, A, lda, 0, B, ldb, 0, C, ldc, 0, 1, bias, beta, 0, g.st);
}

// C[M,N] = A[M,K] B[K,N] (+beta C);  BT = B^T stored [N,K] (ldbt) when available (tc mode).
// The BPTT products have M = 128 rows and a long K (1536, 3072): one wave of 16-wide tiles is only 32 CTAs, each
// walking all of K.  Split-K (partial sums added atomically onto C) spreads them over ~128 CTAs; beta = 1 is the
// natural case (C already holds the value to accumulate onto), beta = 0 zeroes C first.
static inline int gemm_nn(const GemmCtx& g, int M, int N, int K, const float* A, int lda, const float* B,
                          int ldb, const float* BT, int ldbt, float* C, int ldc, float beta) {
  if (g.tc && BT && K >= 8 && tc_ok(A, lda) && tc_ok(BT, ldbt)) {
    const int tiles = ((M + 127) / 128) * ((N + 15) / 16);
    static int enabled = -1;                 // V2F_SPLITK=0: A/B switch
    if (enabled < 0) {
      const char* e = getenv("V2F_SPLITK");
      enabled = (e && e[0] == '0') ? 0 : 1;
    }
    int splits = enabled ? K / 512 : 1;
    if (splits > 148 / tiles) splits = 148 / tiles;
    if (splits > 1 && (beta == 0.f || beta == 1.f) && ldc == N) {
      if (beta == 0.f && cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, (cudaStream_t)g.st) != cudaSuccess)
        return V2F_ERR_LAUNCH;
      return v2f_gemm_tc(1, M, N, K, A, lda, BT, ldbt, C, ldc, nullptr, 0.f, 0, splits, g.st);
    }
    return v2f_gemm_tc(1, M, N, K, A, lda, BT, ldbt, C, ldc, nullptr, beta, 0, 1, g.st);
  }
  return v2f_gemm_f32(0, 0, M, N, K, A, lda, 0, B, ldb, 0, C, ldc, 0, 1, nullptr, beta, 0, g.st);
}

// C[M,N] = A^T B with A stored [K,M], B stored [K,N]  (weight gradients over K = rows)
static inline int gemm_tn(const GemmCtx& g, int M, int N, int K, const float* A, int lda, const float* B,
                          int ldb, float* C, int ldc) {
  const long long Kp = (K + 3) & ~3LL;
  if (g.tc && g.ws && M >= 16 && N >= 16 && K >= 8 && (long long)(M + N) * Kp <= g.ws_floats) {
    float* AT = g.ws;
    float* BT = g.ws + (long long)M * Kp;
    V2F_TRY(v2f_transpose(K, M, A, lda, 1, AT, Kp, 1, g.st));
    V2F_TRY(v2f_transpose(K, N, B, ldb, 1, BT, Kp, 1, g.st));
    return v2f_gemm_tc(1, M, N, K, AT, Kp, BT, Kp, C, ldc, nullptr, 0.f, 0, 1, g.st);
  }
  return v2f_gemm_f32(1, 0, M, N, K, A, lda, 0, B, ldb, 0, C, ldc, 0, 1, nullptr, 0.f, 0, g.st);
}

}  // namespace v2f

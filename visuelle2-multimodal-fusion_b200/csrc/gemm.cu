// fp32 SIMT GEMM (exact-mode dense projections) + column sums.  sm_100a.
//
//   C[b] = op(A[b]) * op(B[b]) (+ bias[n]) (+ beta * C[b]),  optional ReLU
//   op(A) is MxK: ta=0 -> A stored [M,K] (lda), ta=1 -> A stored [K,M] (lda)
//   op(B) is KxN: tb=0 -> B stored [K,N] (ldb), tb=1 -> B stored [N,K] (ldb)   (tb=1 is y = x W^T)
//
// Replaces the cuBLAS calls behind nn.Linear / autograd's mm in the reference hot path
// (e.g. /root/reference/models/CrossAttnRNN210.py:72,84-85,196,208).  The bf16 tcgen05 variant
// for the large hoisted projections lives in gemm_tc.cu.
#include "common.cuh"

namespace v2f {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <int TA, int TB>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(int M, int N, int K, const float* __restrict__ A, int lda, long long sA,
                const float* __restrict__ B, int ldb, long long sB, float* __restrict__ C, int ldc,
                long long sC, const float* __restrict__ bias, float beta, int act) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  A += (long long)blockIdx.z * sA;
  B += (long long)blockIdx.z * sB;
  C += (long long)blockIdx.z * sC;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  const bool a_vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0);
  const bool b_vec = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15u) == 0);

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- A tile -> As[k][m]
    if (TA == 0) {
      const int m = tid >> 2, k4 = (tid & 3) * 4;
      const int gm = m0 + m, gk = k0 + k4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gm < M) {
        const float* p = A + (long long)gm * lda + gk;
        if (a_vec && gk + 3 < K) {
          float4 t = ld4(p);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) if (gk + i < K) v[i] = p[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; i++) As[k4 + i][m] = v[i];
    } else {
      const int k = tid >> 4, m4 = (tid & 15) * 4;
      const int gk = k0 + k, gm = m0 + m4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gk < K) {
        const float* p = A + (long long)gk * lda + gm;
        if (a_vec && gm + 3 < M) {
          float4 t = ld4(p);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) if (gm + i < M) v[i] = p[i];
        }
      }
      st4(&As[k][m4], make_float4(v[0], v[1], v[2], v[3]));
    }
    // ---- B tile -> Bs[k][n]
    if (TB == 1) {
      const int n = tid >> 2, k4 = (tid & 3) * 4;
      const int gn = n0 + n, gk = k0 + k4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gn < N) {
        const float* p = B + (long long)gn * ldb + gk;
        if (b_vec && gk + 3 < K) {
          float4 t = ld4(p);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) if (gk + i < K) v[i] = p[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; i++) Bs[k4 + i][n] = v[i];
    } else {
      const int k = tid >> 4, n4 = (tid & 15) * 4;
      const int gk = k0 + k, gn = n0 + n4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gk < K) {
        const float* p = B + (long long)gk * ldb + gn;
        if (b_vec && gn + 3 < N) {
          float4 t = ld4(p);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) if (gn + i < N) v[i] = p[i];
        }
      }
      st4(&Bs[k][n4], make_float4(v[0], v[1], v[2], v[3]));
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      const float4 a = ld4(&As[kk][ty * 4]);
      const float4 b = ld4(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      float* c = C + (long long)gm * ldc + gn;
      if (beta != 0.f) v += beta * (*c);
      if (act == 1) v = fmaxf(v, 0.f);
      *c = v;
    }
  }
}

// out[n] = sum_m X[m,n] (+ beta*out[n]);  grid = ceil(N/32), block (32,8)
__global__ void colsum_kernel(int M, int N, const float* __restrict__ X, int ldx,
                              float* __restrict__ out, float beta) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (n < N)
    for (int m = threadIdx.y; m < M; m += 8) s += X[(long long)m * ldx + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += red[i][threadIdx.x];
    out[n] = (beta != 0.f ? beta * out[n] : 0.f) + t;
  }
}

}  // namespace v2f

extern "C" int v2f_gemm_f32(int ta, int tb, int M, int N, int K, const float* A, int lda,
                            long long sA, const float* B, int ldb, long long sB, float* C, int ldc,
                            long long sC, int batch, const float* bias, float beta, int act,
                            void* stream) {
  using namespace v2f;
  V2F_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0, V2F_ERR_BAD_ARG);
  if (M == 0 || N == 0 || batch == 0) return V2F_OK;
  V2F_REQUIRE(A && B && C, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(batch <= 65535, V2F_ERR_BAD_ARG);
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch), block(256);
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(TA_, TB_) \
  gemm_f32_kernel<TA_, TB_><<<grid, block, 0, s>>>(M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC, bias, beta, act)
  if (ta == 0 && tb == 0) LAUNCH(0, 0);
  else if (ta == 0 && tb == 1) LAUNCH(0, 1);
  else if (ta == 1 && tb == 0) LAUNCH(1, 0);
  else if (ta == 1 && tb == 1) LAUNCH(1, 1);
  else return V2F_ERR_BAD_ARG;
#undef LAUNCH
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_colsum_f32(int M, int N, const float* X, int ldx, float* out, float beta, void* stream) {
  V2F_REQUIRE(M >= 0 && N > 0 && X && out, V2F_ERR_BAD_ARG);
  v2f::colsum_kernel<<<(N + 31) / 32, dim3(32, 8), 0, (cudaStream_t)stream>>>(M, N, X, ldx, out, beta);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

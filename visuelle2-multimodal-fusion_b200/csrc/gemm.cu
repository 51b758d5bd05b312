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

// out[n] = sum_m X[m,n] (+ beta*out[n]);  grid = (ceil(N/32), chunks), block (32,32).
// chunks > 1: rows are split over gridDim.y and partial sums added atomically to a pre-zeroed out.
__global__ void __launch_bounds__(1024)
colsum_kernel(int M, int N, const float* __restrict__ X, int ldx, float* __restrict__ out, float beta) {
  __shared__ float red[32][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m_lo = blockIdx.y * rows_per, m_hi = min(M, m_lo + rows_per);
  float s = 0.f;
  if (n < N)
    for (int m = m_lo + threadIdx.y; m < m_hi; m += 32) s += X[(long long)m * ldx + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i++) t += red[i][threadIdx.x];
    if (gridDim.y > 1) atomicAdd(out + n, t);
    else out[n] = (beta != 0.f ? beta * out[n] : 0.f) + t;
  }
}

// C[m,n] = sum_k A[k,m] B[k,n] for skinny outputs (N <= 4): weight gradients of scalar inputs
// (decoder w_x, decoder_fc, the 3-feature trend GRU input).  grid = ceil(M/32), block (32,32).
__global__ void __launch_bounds__(1024)
skinny_tn_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B,
                 int ldb, float* __restrict__ C, int ldc) {
  __shared__ float red[4][32][33];
  const int m = blockIdx.x * 32 + threadIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (m < M)
    for (int k = threadIdx.y; k < K; k += 32) {
      const float a = A[(long long)k * lda + m];
#pragma unroll
      for (int n = 0; n < 4; n++)
        if (n < N) acc[n] = fmaf(a, B[(long long)k * ldb + n], acc[n]);
    }
#pragma unroll
  for (int n = 0; n < 4; n++) red[n][threadIdx.y][threadIdx.x] = acc[n];
  __syncthreads();
  if (threadIdx.y < N && m < M) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i++) t += red[threadIdx.y][i][threadIdx.x];
    C[(long long)m * ldc + threadIdx.y] = t;
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
  if (ta == 1 && tb == 0 && N <= 4 && batch == 1 && !bias && beta == 0.f && act == 0 && K >= 256) {
    skinny_tn_kernel<<<(M + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream>>>(M, N, K, A, lda, B, ldb, C, ldc);
    V2F_CHECK_LAUNCH();
    return V2F_OK;
  }
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
  int chunks = 1;
  if (beta == 0.f && M >= 4096) {   // long reductions: split rows over the grid, atomics onto zeros
    chunks = (M + 1023) / 1024;
    if (chunks > 64) chunks = 64;
    cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, (cudaStream_t)stream);
  }
  v2f::colsum_kernel<<<dim3((N + 31) / 32, chunks), dim3(32, 32), 0, (cudaStream_t)stream>>>(M, N, X, ldx, out, beta);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

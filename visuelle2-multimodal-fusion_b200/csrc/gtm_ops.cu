// Row-local operators of the GTM family (GTM_Visuelle2, Proposed_model v1-v4).  sm_100a.
//
// Reference arithmetic (all under /root/reference/models):
//   add + LayerNorm         nn.TransformerEncoderLayer / DecoderLayer post-LN residuals
//                           (GTM_Visuelle2.py:52-53,200-202), GatedResidualBlock.norm
//                           (Proposed_model.py:147-155), fusion_fc LayerNorm (Proposed_model_v4.py:177)
//   BatchNorm1d             GTMFusionNetwork (GTM_Visuelle2.py:158), FusionBlock (Proposed_model_v3.py:163)
//   sigmoid gates           x*s(g), x + x*s(g): Proposed_model.py:154,217; _v2.py:598,635,682;
//                           _v3.py:222-227; _v4.py:186-192
//   attribute stack         AttributeEncoder (GTM_Visuelle2.py:81-96)
//   four 1->E linears       DummyEmbedder (GTM_Visuelle2.py:129-145), TemporalEmbedder (_v3.py:127-145)
//   global average pool     ImageEncoder.pool (GTM_Visuelle2.py:117,123-125), applied BEFORE the
//                           1x1 projection (both linear: SURVEY.md 8a identity 5)
//   + positional encoding   PositionalEncoding.forward (GTM_Visuelle2.py:26-28)
//   repeat_interleave       window replication (GTM_Visuelle2.py:231-235) and its gradient fold
// These are HBM / latency bound elementwise and row-reduction kernels: coalesced rows, warp-shuffle
// reductions, no atomics (every gradient is written once, deterministically).
#include "common.cuh"

extern "C" int v2f_colsum_f32(int, int, const float*, int, float*, float, void*);

namespace v2f {

// ------------------------------------------------------------------ add + LayerNorm
// y = LN(x + a * m) * gamma + beta, one warp per row, row held in registers (PL values per lane).
template <int PL>
__global__ void __launch_bounds__(256)
add_ln_fwd_kernel(int M, int D, const float* __restrict__ x, const float* __restrict__ a,
                  const float* __restrict__ m, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, float* __restrict__ y,
                  float* __restrict__ xhat, float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const long long base = (long long)row * D;
  float v[PL];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PL; i++) {
    const int c = lane + 32 * i;
    float z = 0.f;
    if (c < D) {
      z = x[base + c];
      if (a) {
        float t = a[base + c];
        if (m) t *= m[base + c];
        z += t;
      }
    }
    v[i] = z;
    sum += z;
  }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < PL; i++) {
    const int c = lane + 32 * i;
    const float d = (c < D) ? v[i] - mean : 0.f;
    sq = fmaf(d, d, sq);
  }
  const float var = warp_sum(sq) / (float)D;
  const float r = 1.0f / sqrtf(var + eps);
  if (lane == 0) rstd[row] = r;
#pragma unroll
  for (int i = 0; i < PL; i++) {
    const int c = lane + 32 * i;
    if (c < D) {
      const float xh = (v[i] - mean) * r;
      xhat[base + c] = xh;
      y[base + c] = fmaf(xh, gamma[c], beta[c]);
    }
  }
}

// dz = rstd * (g - mean(g) - xhat * mean(g*xhat)), g = dy*gamma; dx = dz; da = dz*m.
// Persistent grid; per-lane register partials of dgamma/dbeta, reduced over the block's warps in a
// fixed order and written to part[block][2][D] (summed by a column-sum launch).
template <int PL>
__global__ void __launch_bounds__(256)
add_ln_bwd_kernel(int M, int D, const float* __restrict__ dy, const float* __restrict__ xhat,
                  const float* __restrict__ rstd, const float* __restrict__ gamma,
                  const float* __restrict__ m, float* __restrict__ dx, float* __restrict__ da,
                  float* __restrict__ part) {
  extern __shared__ float red[];   // [2][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float gg[PL], gb[PL], gam[PL];
#pragma unroll
  for (int i = 0; i < PL; i++) {
    const int c = lane + 32 * i;
    gg[i] = 0.f;
    gb[i] = 0.f;
    gam[i] = (c < D) ? gamma[c] : 0.f;
  }
  for (int row = blockIdx.x * nwarp + warp; row < M; row += gridDim.x * nwarp) {
    const long long base = (long long)row * D;
    float g[PL], xh[PL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < PL; i++) {
      const int c = lane + 32 * i;
      float d = 0.f, h = 0.f;
      if (c < D) {
        d = dy[base + c];
        h = xhat[base + c];
      }
      gg[i] = fmaf(d, h, gg[i]);
      gb[i] += d;
      g[i] = d * gam[i];
      xh[i] = h;
      s1 += g[i];
      s2 = fmaf(g[i], h, s2);
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
    const float r = rstd[row];
#pragma unroll
    for (int i = 0; i < PL; i++) {
      const int c = lane + 32 * i;
      if (c < D) {
        const float dz = r * (g[i] - s1 - xh[i] * s2);
        dx[base + c] = dz;
        if (da) da[base + c] = m ? dz * m[base + c] : dz;
      }
    }
  }
  for (int w = 0; w < nwarp; w++) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < PL; i++) {
        const int c = lane + 32 * i;
        if (c < D) {
          red[c] = (w == 0 ? 0.f : red[c]) + gg[i];
          red[D + c] = (w == 0 ? 0.f : red[D + c]) + gb[i];
        }
      }
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) part[(long long)blockIdx.x * 2 * D + c] = red[c];
}

// ------------------------------------------------------------------ BatchNorm1d over [B,D]
// grid = ceil(D/32), block (32,32): thread (tx,ty) owns column blk*32+tx, rows ty, ty+32, ...
__device__ __forceinline__ float col_reduce(float v, float (*red)[33]) {
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i++) t += red[i][threadIdx.x];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(1024)
bn1d_fwd_kernel(int B, int D, const float* __restrict__ x, const float* __restrict__ gamma,
                const float* __restrict__ beta, float* __restrict__ run_mean,
                float* __restrict__ run_var, int training, float momentum, float eps,
                float* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < D;
  float mean, var;
  if (training) {
    float s = 0.f;
    if (ok) for (int b = threadIdx.y; b < B; b += 32) s += x[(long long)b * D + c];
    mean = col_reduce(s, red) / (float)B;
    float q = 0.f;
    if (ok) for (int b = threadIdx.y; b < B; b += 32) {
      const float d = x[(long long)b * D + c] - mean;
      q = fmaf(d, d, q);
    }
    var = col_reduce(q, red) / (float)B;
    if (ok && threadIdx.y == 0) {
      const float unb = B > 1 ? var * (float)B / (float)(B - 1) : var;
      run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
      run_var[c] = (1.f - momentum) * run_var[c] + momentum * unb;
    }
  } else {
    mean = ok ? run_mean[c] : 0.f;
    var = ok ? run_var[c] : 1.f;
  }
  const float r = 1.0f / sqrtf(var + eps);
  if (ok) {
    if (threadIdx.y == 0) {
      save_mean[c] = mean;
      save_rstd[c] = r;
    }
    const float g = gamma[c], bt = beta[c];
    for (int b = threadIdx.y; b < B; b += 32)
      y[(long long)b * D + c] = fmaf((x[(long long)b * D + c] - mean) * r, g, bt);
  }
}

__global__ void __launch_bounds__(1024)
bn1d_bwd_kernel(int B, int D, const float* __restrict__ x, const float* __restrict__ dy,
                const float* __restrict__ gamma, const float* __restrict__ save_mean,
                const float* __restrict__ save_rstd, int training, float* __restrict__ dx,
                float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < D;
  const float mean = ok ? save_mean[c] : 0.f, r = ok ? save_rstd[c] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  if (ok) for (int b = threadIdx.y; b < B; b += 32) {
    const float d = dy[(long long)b * D + c];
    s1 += d;
    s2 = fmaf(d, (x[(long long)b * D + c] - mean) * r, s2);
  }
  s1 = col_reduce(s1, red);
  s2 = col_reduce(s2, red);
  if (!ok) return;
  if (threadIdx.y == 0) {
    dgamma[c] = s2;
    dbeta[c] = s1;
  }
  const float g = gamma[c] * r;
  const float m1 = s1 / (float)B, m2 = s2 / (float)B;
  for (int b = threadIdx.y; b < B; b += 32) {
    const float d = dy[(long long)b * D + c];
    if (training) {
      const float xh = (x[(long long)b * D + c] - mean) * r;
      dx[(long long)b * D + c] = g * (d - m1 - xh * m2);
    } else {
      dx[(long long)b * D + c] = g * d;
    }
  }
}

// ------------------------------------------------------------------ BatchNorm1d with batch statistics over a SHARDED batch
// (data parallelism: every rank holds B_local rows of the global batch).  The reference normalises with the statistics
// of the whole batch (models/GTM_Visuelle2.py:158, Proposed_model_v3.py:166), so under batch sharding the per-channel
// sums are exchanged: stats kernel -> all-reduce of [2,D] doubles (host side, NCCL) -> apply kernel; same for the two
// reductions of the backward pass.  Sums are kept in double so that E[x^2] - mean^2 stays at the fp32 rounding level.
__device__ __forceinline__ double col_reduce_d(double v, double (*red)[33]) {
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.y == 0)
    for (int i = 0; i < 32; i++) t += red[i][threadIdx.x];
  __syncthreads();
  return t;
}

// out[0,c] = sum_b x[b,c], out[1,c] = sum_b x[b,c]^2
__global__ void __launch_bounds__(1024)
bn1d_stats_kernel(int B, int D, const float* __restrict__ x, double* __restrict__ out) {
  __shared__ double red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < D;
  double s = 0.0, q = 0.0;
  if (ok) for (int b = threadIdx.y; b < B; b += 32) {
    const double v = x[(long long)b * D + c];
    s += v;
    q += v * v;
  }
  s = col_reduce_d(s, red);
  q = col_reduce_d(q, red);
  if (ok && threadIdx.y == 0) {
    out[c] = s;
    out[D + c] = q;
  }
}

// sums = all-reduced [2,D]; Btot = rows of the global batch
__global__ void __launch_bounds__(1024)
bn1d_apply_kernel(int B, int D, const float* __restrict__ x, const float* __restrict__ gamma,
                  const float* __restrict__ beta, const double* __restrict__ sums, double Btot,
                  float* __restrict__ run_mean, float* __restrict__ run_var, float momentum, float eps,
                  float* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (c >= D) return;
  const double m = sums[c] / Btot;
  double v = sums[D + c] / Btot - m * m;
  if (v < 0.0) v = 0.0;
  const float mean = (float)m, var = (float)v;
  const float r = 1.0f / sqrtf(var + eps);
  if (threadIdx.y == 0) {
    const float unb = Btot > 1.0 ? (float)(v * Btot / (Btot - 1.0)) : var;
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * unb;
    save_mean[c] = mean;
    save_rstd[c] = r;
  }
  const float g = gamma[c], bt = beta[c];
  for (int b = threadIdx.y; b < B; b += 32)
    y[(long long)b * D + c] = fmaf((x[(long long)b * D + c] - mean) * r, g, bt);
}

// out[0,c] = sum_b dy, out[1,c] = sum_b dy * xhat   (local rows); dgamma / dbeta = the local sums
__global__ void __launch_bounds__(1024)
bn1d_bwd_stats_kernel(int B, int D, const float* __restrict__ x, const float* __restrict__ dy,
                      const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                      double* __restrict__ out, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < D;
  const float mean = ok ? save_mean[c] : 0.f, r = ok ? save_rstd[c] : 0.f;
  double s1 = 0.0, s2 = 0.0;
  if (ok) for (int b = threadIdx.y; b < B; b += 32) {
    const float d = dy[(long long)b * D + c];
    s1 += d;
    s2 += (double)d * (double)((x[(long long)b * D + c] - mean) * r);
  }
  s1 = col_reduce_d(s1, red);
  s2 = col_reduce_d(s2, red);
  if (ok && threadIdx.y == 0) {
    out[c] = s1;
    out[D + c] = s2;
    dbeta[c] = (float)s1;
    dgamma[c] = (float)s2;
  }
}

__global__ void __launch_bounds__(1024)
bn1d_bwd_apply_kernel(int B, int D, const float* __restrict__ x, const float* __restrict__ dy,
                      const float* __restrict__ gamma, const float* __restrict__ save_mean,
                      const float* __restrict__ save_rstd, const double* __restrict__ sums, double Btot,
                      float* __restrict__ dx) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (c >= D) return;
  const float mean = save_mean[c], r = save_rstd[c];
  const float g = gamma[c] * r;
  const float m1 = (float)(sums[c] / Btot), m2 = (float)(sums[D + c] / Btot);
  for (int b = threadIdx.y; b < B; b += 32) {
    const float d = dy[(long long)b * D + c];
    const float xh = (x[(long long)b * D + c] - mean) * r;
    dx[(long long)b * D + c] = g * (d - m1 - xh * m2);
  }
}

// ------------------------------------------------------------------ dropout with the keep decisions drawn IN the kernel
// out = x * keep / (1 - p), keep ~ Bernoulli(1 - p) from Philox4x32-10 keyed by two 64-bit words in DEVICE memory (drawn
// by the caller from torch's CUDA generator: graph-safe, fresh per replay).  The backward regenerates the same decisions
// from the same key, so no mask is ever materialised (the reference's nn.Dropout after ImageEncoder.fc alone was a
// 26 MB fp32 mask written by four elementwise launches and read back twice).  Thread = 4 consecutive elements = one
// Philox counter.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

__global__ void dropout_kernel(long long n, const float* __restrict__ x, const unsigned long long* __restrict__ key,
                               float p, float* __restrict__ out) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = i4 * 4;
  if (i >= n) return;
  const unsigned long long k = key[0], sq = key[1];
  uint32_t r[4];
  philox4x32_10((uint32_t)i4, (uint32_t)(i4 >> 32), (uint32_t)sq, (uint32_t)(sq >> 32), (uint32_t)k, (uint32_t)(k >> 32), r);
  const float scale = 1.0f / (1.0f - p);
  const uint32_t thr = (uint32_t)fminf(p * 4294967296.0f, 4294967295.0f);      // keep iff r >= thr
  if (i + 3 < n) {
    float4 v = ld4(x + i);
    v.x = r[0] >= thr ? v.x * scale : 0.f;
    v.y = r[1] >= thr ? v.y * scale : 0.f;
    v.z = r[2] >= thr ? v.z * scale : 0.f;
    v.w = r[3] >= thr ? v.w * scale : 0.f;
    st4(out + i, v);
  } else {
    for (int j = 0; j < 4 && i + j < n; j++) out[i + j] = r[j] >= thr ? x[i + j] * scale : 0.f;
  }
}

// ------------------------------------------------------------------ elementwise
__global__ void gate_fwd_kernel(long long n, const float* __restrict__ x, const float* __restrict__ g,
                                int mode, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = sigmoid_full(g[i]);
  const float xv = x[i];
  out[i] = mode ? fmaf(xv, s, xv) : xv * s;
}
__global__ void gate_bwd_kernel(long long n, const float* __restrict__ x, const float* __restrict__ g,
                                const float* __restrict__ dout, int mode, float* __restrict__ dx,
                                float* __restrict__ dg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = sigmoid_full(g[i]);
  const float d = dout[i];
  dx[i] = mode ? fmaf(d, s, d) : d * s;
  dg[i] = d * x[i] * s * (1.f - s);
}
__global__ void add_kernel(long long n, const float* __restrict__ a, const float* __restrict__ b,
                           float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}
__global__ void relu_fwd_kernel(long long n, const float* __restrict__ x, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fmaxf(x[i], 0.f);
}
__global__ void relu_bwd_kernel(long long n, const float* __restrict__ dy, const float* __restrict__ y,
                                float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = y[i] > 0.f ? dy[i] : 0.f;
}
// out[b, i] = x[b, i] + p[i]
__global__ void add_bcast_kernel(long long total, long long n, const float* __restrict__ x,
                                 const float* __restrict__ p, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = x[i] + p[i % n];
}
__global__ void copy2d_kernel(int rows, int cols, const float* __restrict__ src, long long lds,
                              float* __restrict__ dst, long long ldd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const long long r = i / cols, c = i - r * cols;
  dst[r * ldd + c] = src[r * lds + c];
}
// out[b*W + w, :] = x[b, :]
__global__ void repeat_rows_kernel(long long total, int W, long long D, const float* __restrict__ x,
                                   float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long n = i / D, c = i - n * D;
  out[i] = x[(n / W) * D + c];
}
// dx[b, :] = sum_w dout[b*W + w, :]
__global__ void fold_rows_kernel(long long total, int W, long long D, const float* __restrict__ dout,
                                 float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / D, c = i - b * D;
  float s = 0.f;
  for (int w = 0; w < W; w++) s += dout[(b * W + w) * D + c];
  dx[i] = s;
}

// ------------------------------------------------------------------ attribute stack / four linears
struct Tab4 { const float* t[4]; };
struct DTab4 { float* t[4]; int rows[4]; };

__global__ void gather4_fwd_kernel(int B, int E, Tab4 tb, const long long* __restrict__ idx,
                                   const float* __restrict__ drop, float* __restrict__ out) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < 4 * E; i += blockDim.x) {
    const int k = i / E, e = i - k * E;
    float v = tb.t[k][idx[(long long)k * B + b] * E + e];
    if (drop) v *= drop[(long long)b * 4 * E + i];
    out[(long long)b * 4 * E + i] = v;
  }
}
// one block per table row: deterministic scatter (loop over the batch, no atomics)
__global__ void gather4_bwd_kernel(int B, int E, const long long* __restrict__ idx,
                                   const float* __restrict__ drop, const float* __restrict__ dout,
                                   DTab4 dt) {
  int blk = blockIdx.x, k = 0;
  while (k < 3 && blk >= dt.rows[k]) { blk -= dt.rows[k]; k++; }
  const int r = blk;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float g = 0.f;
    for (int b = 0; b < B; b++)
      if (idx[(long long)k * B + b] == r) {
        float v = dout[((long long)b * 4 + k) * E + e];
        if (drop) v *= drop[((long long)b * 4 + k) * E + e];
        g += v;
      }
    dt.t[k][(long long)r * E + e] = g;
  }
}
// out[b, k, e] = t[b,k] * Wt[k,e] + bt[k,e]
__global__ void feat4_fwd_kernel(int B, int E, const float* __restrict__ t, const float* __restrict__ Wt,
                                 const float* __restrict__ bt, float* __restrict__ out) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < 4 * E; i += blockDim.x)
    out[(long long)b * 4 * E + i] = fmaf(t[b * 4 + i / E], Wt[i], bt[i]);
}
__global__ void feat4_bwd_kernel(int B, int E, const float* __restrict__ t, const float* __restrict__ dout,
                                 float* __restrict__ dWt, float* __restrict__ dbt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * E) return;
  const int k = i / E;
  float gw = 0.f, gb = 0.f;
  for (int b = 0; b < B; b++) {
    const float g = dout[(long long)b * 4 * E + i];
    gw = fmaf(g, t[b * 4 + k], gw);
    gb += g;
  }
  dWt[i] = gw;
  dbt[i] = gb;
}

// ------------------------------------------------------------------ global average pool
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<unsigned short>(unsigned short v) {
  return __uint_as_float(((unsigned)v) << 16);
}
__device__ __forceinline__ unsigned short f_to_bf16(float f) {   // round to nearest even
  unsigned u = __float_as_uint(f);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (unsigned short)(u >> 16);
}
// layout 0: x [B,C,L] -> one warp per (b,c)
template <typename T>
__global__ void __launch_bounds__(256)
pool_cl_kernel(long long BC, int L, const T* __restrict__ x, float* __restrict__ out) {
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= BC) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int l = lane; l < L; l += 32) s += to_f<T>(x[w * L + l]);
  s = warp_sum(s);
  if (lane == 0) out[w] = s / (float)L;
}
// layout 1: x [B,L,C] -> one thread per (b,c), coalesced over c
template <typename T>
__global__ void __launch_bounds__(256)
pool_lc_kernel(int B, int L, int C, const T* __restrict__ x, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C) return;
  const long long b = i / C, c = i - b * C;
  float s = 0.f;
  for (int l = 0; l < L; l++) s += to_f<T>(x[(b * L + l) * C + c]);
  out[i] = s / (float)L;
}
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ unsigned short from_f<unsigned short>(float v) { return f_to_bf16(v); }
template <typename T>
__global__ void __launch_bounds__(256)
pool_bwd_kernel(long long total, int L, int C, int layout, const float* __restrict__ dout,
                T* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long bc;
  if (layout == 0) bc = i / L;                       // [B,C,L]
  else { const long long bl = i / C; bc = (bl / L) * C + (i - bl * C); }   // [B,L,C]
  dx[i] = from_f<T>(dout[bc] / (float)L);
}

}  // namespace v2f

using namespace v2f;

static inline unsigned blocks_for(long long n, int t = 256) { return (unsigned)((n + t - 1) / t); }

extern "C" int v2f_add_ln_fwd(int M, int D, const float* x, const float* a, const float* m,
                              const float* gamma, const float* beta, float eps, float* y,
                              float* xhat, float* rstd, void* st) {
  V2F_REQUIRE(M > 0 && D > 0 && x && gamma && beta && y && xhat && rstd, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(D <= 1024, V2F_ERR_UNSUPPORTED);
  cudaStream_t s = (cudaStream_t)st;
  const unsigned grid = (M + 7) / 8;
#define LN_FWD(PL_) add_ln_fwd_kernel<PL_><<<grid, 256, 0, s>>>(M, D, x, a, m, gamma, beta, eps, y, xhat, rstd)
  if (D <= 32) LN_FWD(1);
  else if (D <= 64) LN_FWD(2);
  else if (D <= 128) LN_FWD(4);
  else if (D <= 256) LN_FWD(8);
  else if (D <= 512) LN_FWD(16);
  else LN_FWD(32);
#undef LN_FWD
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_add_ln_bwd_blocks(int M) {
  int b = (M + 7) / 8;
  return b < 296 ? b : 296;
}

extern "C" int v2f_add_ln_bwd(int M, int D, const float* dy, const float* xhat, const float* rstd,
                              const float* gamma, const float* m, float* dx, float* da, float* part,
                              float* dgb, void* st) {
  V2F_REQUIRE(M > 0 && D > 0 && dy && xhat && rstd && gamma && dx && part && dgb, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(D <= 1024, V2F_ERR_UNSUPPORTED);
  cudaStream_t s = (cudaStream_t)st;
  const int grid = v2f_add_ln_bwd_blocks(M);
  const size_t smem = sizeof(float) * 2 * D;
#define LN_BWD(PL_) add_ln_bwd_kernel<PL_><<<grid, 256, smem, s>>>(M, D, dy, xhat, rstd, gamma, m, dx, da, part)
  if (D <= 32) LN_BWD(1);
  else if (D <= 64) LN_BWD(2);
  else if (D <= 128) LN_BWD(4);
  else if (D <= 256) LN_BWD(8);
  else if (D <= 512) LN_BWD(16);
  else LN_BWD(32);
#undef LN_BWD
  V2F_CHECK_LAUNCH();
  return v2f_colsum_f32(grid, 2 * D, part, 2 * D, dgb, 0.f, st);
}

extern "C" int v2f_bn1d_fwd(int B, int D, const float* x, const float* gamma, const float* beta,
                            float* run_mean, float* run_var, int training, float momentum, float eps,
                            float* y, float* save_mean, float* save_rstd, void* st) {
  V2F_REQUIRE(B > 0 && D > 0 && x && gamma && beta && run_mean && run_var && y && save_mean && save_rstd,
              V2F_ERR_BAD_ARG);
  bn1d_fwd_kernel<<<(D + 31) / 32, dim3(32, 32), 0, (cudaStream_t)st>>>(B, D, x, gamma, beta, run_mean, run_var,
                                                                        training, momentum, eps, y, save_mean,
                                                                        save_rstd);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_bn1d_bwd(int B, int D, const float* x, const float* dy, const float* gamma,
                            const float* save_mean, const float* save_rstd, int training, float* dx,
                            float* dgamma, float* dbeta, void* st) {
  V2F_REQUIRE(B > 0 && D > 0 && x && dy && gamma && save_mean && save_rstd && dx && dgamma && dbeta,
              V2F_ERR_BAD_ARG);
  bn1d_bwd_kernel<<<(D + 31) / 32, dim3(32, 32), 0, (cudaStream_t)st>>>(B, D, x, dy, gamma, save_mean, save_rstd,
                                                                        training, dx, dgamma, dbeta);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_bn1d_stats(int B, int D, const float* x, double* sums, void* st) {
  V2F_REQUIRE(B > 0 && D > 0 && x && sums, V2F_ERR_BAD_ARG);
  bn1d_stats_kernel<<<(D + 31) / 32, dim3(32, 32), 0, (cudaStream_t)st>>>(B, D, x, sums);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_bn1d_apply(int B, int D, const float* x, const float* gamma, const float* beta, const double* sums,
                              double Btot, float* run_mean, float* run_var, float momentum, float eps, float* y,
                              float* save_mean, float* save_rstd, void* st) {
  V2F_REQUIRE(B > 0 && D > 0 && Btot >= B && x && gamma && beta && sums && run_mean && run_var && y && save_mean && save_rstd,
              V2F_ERR_BAD_ARG);
  bn1d_apply_kernel<<<(D + 31) / 32, dim3(32, 32), 0, (cudaStream_t)st>>>(B, D, x, gamma, beta, sums, Btot, run_mean,
                                                                          run_var, momentum, eps, y, save_mean, save_rstd);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_bn1d_bwd_stats(int B, int D, const float* x, const float* dy, const float* save_mean,
                                  const float* save_rstd, double* sums, float* dgamma, float* dbeta, void* st) {
  V2F_REQUIRE(B > 0 && D > 0 && x && dy && save_mean && save_rstd && sums && dgamma && dbeta, V2F_ERR_BAD_ARG);
  bn1d_bwd_stats_kernel<<<(D + 31) / 32, dim3(32, 32), 0, (cudaStream_t)st>>>(B, D, x, dy, save_mean, save_rstd, sums,
                                                                              dgamma, dbeta);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_bn1d_bwd_apply(int B, int D, const float* x, const float* dy, const float* gamma,
                                  const float* save_mean, const float* save_rstd, const double* sums, double Btot,
                                  float* dx, void* st) {
  V2F_REQUIRE(B > 0 && D > 0 && Btot >= B && x && dy && gamma && save_mean && save_rstd && sums && dx, V2F_ERR_BAD_ARG);
  bn1d_bwd_apply_kernel<<<(D + 31) / 32, dim3(32, 32), 0, (cudaStream_t)st>>>(B, D, x, dy, gamma, save_mean, save_rstd,
                                                                              sums, Btot, dx);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_dropout(long long n, const float* x, const unsigned long long* key, float p, float* out, void* st) {
  V2F_REQUIRE(n >= 0 && x && key && out && p >= 0.f && p < 1.f, V2F_ERR_BAD_ARG);
  if (n == 0) return V2F_OK;
  V2F_REQUIRE(aligned16(x) && aligned16(out), V2F_ERR_ALIGN);
  dropout_kernel<<<blocks_for((n + 3) / 4), 256, 0, (cudaStream_t)st>>>(n, x, key, p, out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_gate_fwd(long long n, const float* x, const float* g, int mode, float* out, void* st) {
  V2F_REQUIRE(n >= 0 && x && g && out && (mode == 0 || mode == 1), V2F_ERR_BAD_ARG);
  if (n == 0) return V2F_OK;
  gate_fwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)st>>>(n, x, g, mode, out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_gate_bwd(long long n, const float* x, const float* g, const float* dout, int mode,
                            float* dx, float* dg, void* st) {
  V2F_REQUIRE(n >= 0 && x && g && dout && dx && dg && (mode == 0 || mode == 1), V2F_ERR_BAD_ARG);
  if (n == 0) return V2F_OK;
  gate_bwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)st>>>(n, x, g, dout, mode, dx, dg);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_add_f32(long long n, const float* a, const float* b, float* out, void* st) {
  V2F_REQUIRE(n >= 0 && a && b && out, V2F_ERR_BAD_ARG);
  if (n == 0) return V2F_OK;
  add_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)st>>>(n, a, b, out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_relu_fwd(long long n, const float* x, float* out, void* st) {
  V2F_REQUIRE(n >= 0 && x && out, V2F_ERR_BAD_ARG);
  if (n == 0) return V2F_OK;
  relu_fwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)st>>>(n, x, out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_relu_bwd(long long n, const float* dy, const float* y, float* out, void* st) {
  V2F_REQUIRE(n >= 0 && dy && y && out, V2F_ERR_BAD_ARG);
  if (n == 0) return V2F_OK;
  relu_bwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)st>>>(n, dy, y, out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_add_bcast(long long rows, long long n, const float* x, const float* p, float* out, void* st) {
  V2F_REQUIRE(rows > 0 && n > 0 && x && p && out, V2F_ERR_BAD_ARG);
  add_bcast_kernel<<<blocks_for(rows * n), 256, 0, (cudaStream_t)st>>>(rows * n, n, x, p, out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_copy2d(int rows, int cols, const float* src, long long lds, float* dst, long long ldd,
                          void* st) {
  V2F_REQUIRE(rows >= 0 && cols >= 0 && src && dst && lds >= cols && ldd >= cols, V2F_ERR_BAD_ARG);
  if (rows == 0 || cols == 0) return V2F_OK;
  copy2d_kernel<<<blocks_for((long long)rows * cols), 256, 0, (cudaStream_t)st>>>(rows, cols, src, lds, dst, ldd);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_repeat_rows(int B, int W, long long D, const float* x, float* out, void* st) {
  V2F_REQUIRE(B > 0 && W > 0 && D > 0 && x && out, V2F_ERR_BAD_ARG);
  const long long total = (long long)B * W * D;
  repeat_rows_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)st>>>(total, W, D, x, out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_fold_rows(int B, int W, long long D, const float* dout, float* dx, void* st) {
  V2F_REQUIRE(B > 0 && W > 0 && D > 0 && dout && dx, V2F_ERR_BAD_ARG);
  const long long total = (long long)B * D;
  fold_rows_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)st>>>(total, W, D, dout, dx);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_gather4_fwd(int B, int E, const float* const* tables, const long long* idx,
                               const float* drop, float* out, void* st) {
  V2F_REQUIRE(B > 0 && E > 0 && tables && idx && out, V2F_ERR_BAD_ARG);
  Tab4 tb;
  for (int k = 0; k < 4; k++) {
    V2F_REQUIRE(tables[k], V2F_ERR_BAD_ARG);
    tb.t[k] = tables[k];
  }
  gather4_fwd_kernel<<<B, 128, 0, (cudaStream_t)st>>>(B, E, tb, idx, drop, out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_gather4_bwd(int B, int E, const long long* idx, const float* drop, const float* dout,
                               const int* table_rows, float* const* dtables, void* st) {
  V2F_REQUIRE(B > 0 && E > 0 && idx && dout && table_rows && dtables, V2F_ERR_BAD_ARG);
  DTab4 dt;
  int total = 0;
  for (int k = 0; k < 4; k++) {
    V2F_REQUIRE(dtables[k] && table_rows[k] > 0, V2F_ERR_BAD_ARG);
    dt.t[k] = dtables[k];
    dt.rows[k] = table_rows[k];
    total += table_rows[k];
  }
  gather4_bwd_kernel<<<total, 64, 0, (cudaStream_t)st>>>(B, E, idx, drop, dout, dt);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_feat4_fwd(int B, int E, const float* temporal, const float* Wt, const float* bt,
                             float* out, void* st) {
  V2F_REQUIRE(B > 0 && E > 0 && temporal && Wt && bt && out, V2F_ERR_BAD_ARG);
  feat4_fwd_kernel<<<B, 128, 0, (cudaStream_t)st>>>(B, E, temporal, Wt, bt, out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_feat4_bwd(int B, int E, const float* temporal, const float* dout, float* dWt,
                             float* dbt, void* st) {
  V2F_REQUIRE(B > 0 && E > 0 && temporal && dout && dWt && dbt, V2F_ERR_BAD_ARG);
  feat4_bwd_kernel<<<blocks_for(4LL * E, 128), 128, 0, (cudaStream_t)st>>>(B, E, temporal, dout, dWt, dbt);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_meanpool_fwd(int B, int L, int C, const void* x, int layout, int kind, float* out,
                                void* st) {
  V2F_REQUIRE(B > 0 && L > 0 && C > 0 && x && out, V2F_ERR_BAD_ARG);
  V2F_REQUIRE((layout == 0 || layout == 1) && (kind == 0 || kind == 1), V2F_ERR_BAD_ARG);
  cudaStream_t s = (cudaStream_t)st;
  const long long BC = (long long)B * C;
  if (layout == 0) {
    if (kind == 1) pool_cl_kernel<float><<<blocks_for(BC, 8), 256, 0, s>>>(BC, L, (const float*)x, out);
    else pool_cl_kernel<unsigned short><<<blocks_for(BC, 8), 256, 0, s>>>(BC, L, (const unsigned short*)x, out);
  } else {
    if (kind == 1) pool_lc_kernel<float><<<blocks_for(BC), 256, 0, s>>>(B, L, C, (const float*)x, out);
    else pool_lc_kernel<unsigned short><<<blocks_for(BC), 256, 0, s>>>(B, L, C, (const unsigned short*)x, out);
  }
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}
extern "C" int v2f_meanpool_bwd(int B, int L, int C, const float* dout, int layout, int kind, void* dx,
                                void* st) {
  V2F_REQUIRE(B > 0 && L > 0 && C > 0 && dout && dx, V2F_ERR_BAD_ARG);
  V2F_REQUIRE((layout == 0 || layout == 1) && (kind == 0 || kind == 1), V2F_ERR_BAD_ARG);
  cudaStream_t s = (cudaStream_t)st;
  const long long total = (long long)B * L * C;
  if (kind == 1) pool_bwd_kernel<float><<<blocks_for(total), 256, 0, s>>>(total, L, C, layout, dout, (float*)dx);
  else pool_bwd_kernel<unsigned short><<<blocks_for(total), 256, 0, s>>>(total, L, C, layout, dout,
                                                                          (unsigned short*)dx);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

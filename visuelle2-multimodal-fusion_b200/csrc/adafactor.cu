// Multi-tensor Adafactor step: every trainable tensor of the model in ~10 launches instead of ~15 tiny
// launches per tensor.  sm_100a.
//
// Reference: configure_optimizers of the LightningModules (/root/reference/models/CrossAttnRNN210.py:229-230,
// /root/reference/models/GTM_Visuelle2.py:264-266): fairseq's Adafactor(scale_parameter=True, relative_step=True,
// warmup_init=True, lr=None).  fairseq is not in the image; the same algorithm ships as
// transformers.optimization.Adafactor (the checker of tests/test_gpu_adafactor.py).  Per tensor, per step:
//   RMS = ||p|| / sqrt(numel) ; lr = max(eps2, RMS) * min(1e-6 step, 1/sqrt(step)) ; beta = 1 - step^-0.8
//   upd = g^2 + eps1
//   >= 2-D (factored over the last two dims): row <- beta row + (1-beta) mean_c upd ; col <- beta col + (1-beta) mean_r upd
//            u = g * rsqrt(row / mean_r row) * rsqrt(col)
//   1-D:    sq <- beta sq + (1-beta) upd ; u = g * rsqrt(sq)
//   u /= max(1, RMS(u) / clip) ; p <- p - lr u
// Three shapes of work, each with its own work-unit table built once by the caller:
//   vectors (BatchNorm affine parameters, biases): elementwise chunks;
//   small matrices (convolution weights [O,I,kh,kw]: O*I matrices of kh x kw, kh, kw <= 4): one thread per matrix,
//     everything except the two global norms is matrix-local;
//   big matrices (nn.Linear / GRU / embedding weights): one CTA per row, one thread per column for the column means.
// The two global norms per tensor (||p||^2, ||u||^2) are accumulated in double with atomics (acc[desc][2]).
#include "common.cuh"

namespace v2f {

constexpr int AF_THREADS = 256;
constexpr int AF_DIM = 4;          // a "small" matrix has R <= 4 and C <= 4 (1x1 and 3x3 convolution kernels)
constexpr int AF_COLSTRIP = 32;    // columns per CTA of the column-mean kernel
constexpr int AF_ROWLANES = AF_THREADS / AF_COLSTRIP;

struct AfHyper {
  float beta, omb;        // beta2t, 1 - beta2t
  float eps1, eps2, clip;
  double rel_step;        // min(1e-6 step, 1/sqrt(step)) (relative_step) or the fixed lr
  int scale_parameter;
};

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
  return t;
}
__device__ __forceinline__ float block_sum_f(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
  return t;
}

__device__ __forceinline__ void step_scalars(const v2f_af_desc& d, const double* acc, const AfHyper& h, float& lr,
                                             float& cdiv, float& rms) {
  rms = (float)(sqrt(acc[0]) / sqrt((double)d.numel));
  const float scale = h.scale_parameter ? fmaxf(h.eps2, rms) : 1.0f;
  lr = (float)((double)scale * h.rel_step);
  const float rms_u = (float)(sqrt(acc[1]) / sqrt((double)d.numel));
  cdiv = fmaxf(rms_u / h.clip, 1.0f);
}

// ------------------------------------------------------------------------------------------ vectors
// unit: (desc, first element, count <= AF_VEC_CHUNK)
template <bool APPLY>
__global__ void __launch_bounds__(AF_THREADS)
af_vec_kernel(const v2f_af_desc* __restrict__ descs, const float* const* __restrict__ grads, double* __restrict__ acc,
              const int4* __restrict__ units, AfHyper h) {
  __shared__ double red[AF_THREADS / 32];
  const int4 u = units[blockIdx.x];
  const v2f_af_desc d = descs[u.x];
  const float* g = grads[u.x];
  float lr = 0.f, cdiv = 1.f, rms = 0.f;
  if (APPLY) {
    step_scalars(d, acc + 2 * u.x, h, lr, cdiv, rms);
    if (u.y == 0 && threadIdx.x == 0) *d.rms = rms;
  }
  double sp = 0.0, su = 0.0;
  for (int i = threadIdx.x; i < u.z; i += AF_THREADS) {
    const long long e = (long long)u.y + i;
    const float gv = g[e], pv = d.p[e];
    if (!APPLY) {
      const float sq = d.sq[e] * h.beta + (gv * gv + h.eps1) * h.omb;
      d.sq[e] = sq;
      const float uu = gv * rsqrtf(sq);
      sp += (double)pv * pv;
      su += (double)uu * uu;
    } else {
      const float uu = gv * rsqrtf(d.sq[e]);
      d.p[e] = pv - (uu / cdiv) * lr;
    }
  }
  if (!APPLY) {
    sp = block_sum_d(sp, red);
    su = block_sum_d(su, red);
    if (threadIdx.x == 0) {
      atomicAdd(acc + 2 * u.x, sp);
      atomicAdd(acc + 2 * u.x + 1, su);
    }
  }
}

// ------------------------------------------------------------------------------------------ small matrices
// unit: (desc, first matrix, count <= AF_THREADS); one thread per R x C matrix (R, C <= AF_DIM)
// RR, CC > 0: the matrix shape is known at compile time (1x1 and 3x3 convolution kernels are 99.9 % of the matrices;
// a 1x1 "matrix" is one element: row = col, the row factor is exactly 1); 0: read it from the descriptor.
template <bool APPLY, int RR, int CC>
__global__ void __launch_bounds__(AF_THREADS)
af_small_kernel(const v2f_af_desc* __restrict__ descs, const float* const* __restrict__ grads, double* __restrict__ acc,
                const int4* __restrict__ units, AfHyper h) {
  __shared__ double red[AF_THREADS / 32];
  const int4 u = units[blockIdx.x];
  const v2f_af_desc d = descs[u.x];
  const float* g = grads[u.x];
  constexpr int AF_DIM = RR > 0 ? (RR > CC ? RR : CC) : v2f::AF_DIM;
  const int R = RR > 0 ? RR : d.R, C = CC > 0 ? CC : d.C;
  float lr = 0.f, cdiv = 1.f, rms = 0.f;
  if (APPLY) {
    step_scalars(d, acc + 2 * u.x, h, lr, cdiv, rms);
    if (u.y == 0 && threadIdx.x == 0) *d.rms = rms;
  }
  double sp = 0.0, su = 0.0;
  if ((int)threadIdx.x < u.z) {
    const long long m = (long long)u.y + threadIdx.x;
    const long long base = (m / d.inner) * d.sO + (m % d.inner) * d.sI;
    const float* gm = g + base;
    float* pm = d.p + base;
    const long long sR = d.sR, sC = d.sC;
    float* rowm = d.row + m * R;
    float* colm = d.col + m * C;
    // R, C <= AF_DIM: fully unrolled with predicates so that everything stays in registers
    float gv[AF_DIM][AF_DIM], rowv[AF_DIM], colv[AF_DIM];
#pragma unroll
    for (int r = 0; r < AF_DIM; r++)
#pragma unroll
      for (int c = 0; c < AF_DIM; c++) gv[r][c] = (r < R && c < C) ? gm[r * sR + c * sC] : 0.f;
    if (!APPLY) {
#pragma unroll
      for (int r = 0; r < AF_DIM; r++) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < AF_DIM; c++)
          if (c < C) s += gv[r][c] * gv[r][c] + h.eps1;
        rowv[r] = 1.f;
        if (r < R) {
          rowv[r] = rowm[r] * h.beta + (s / (float)C) * h.omb;
          rowm[r] = rowv[r];
        }
      }
#pragma unroll
      for (int c = 0; c < AF_DIM; c++) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < AF_DIM; r++)
          if (r < R) s += gv[r][c] * gv[r][c] + h.eps1;
        colv[c] = 1.f;
        if (c < C) {
          colv[c] = colm[c] * h.beta + (s / (float)R) * h.omb;
          colm[c] = colv[c];
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < AF_DIM; r++) rowv[r] = r < R ? rowm[r] : 1.f;
#pragma unroll
      for (int c = 0; c < AF_DIM; c++) colv[c] = c < C ? colm[c] : 1.f;
    }
    float rmean = 0.f;
#pragma unroll
    for (int r = 0; r < AF_DIM; r++)
      if (r < R) rmean += rowv[r];
    rmean /= (float)R;
#pragma unroll
    for (int r = 0; r < AF_DIM; r++) rowv[r] = rsqrtf(rowv[r] / rmean);
#pragma unroll
    for (int c = 0; c < AF_DIM; c++) colv[c] = rsqrtf(colv[c]);
#pragma unroll
    for (int r = 0; r < AF_DIM; r++)
#pragma unroll
      for (int c = 0; c < AF_DIM; c++)
        if (r < R && c < C) {
          const long long i = r * sR + c * sC;
          const float uu = gv[r][c] * (rowv[r] * colv[c]);
          if (!APPLY) {
            const float pv = pm[i];
            sp += (double)pv * pv;
            su += (double)uu * uu;
          } else {
            pm[i] = pm[i] - (uu / cdiv) * lr;
          }
        }
  }
  if (!APPLY) {
    sp = block_sum_d(sp, red);
    su = block_sum_d(su, red);
    if (threadIdx.x == 0) {
      atomicAdd(acc + 2 * u.x, sp);
      atomicAdd(acc + 2 * u.x + 1, su);
    }
  }
}

// ------------------------------------------------------------------------------------------ big matrices
// row unit: (desc, matrix, row); column unit: (desc, matrix, first column of a 32-wide strip)
__global__ void __launch_bounds__(AF_THREADS)
af_big_rowstat_kernel(const v2f_af_desc* __restrict__ descs, const float* const* __restrict__ grads,
                      double* __restrict__ acc, const int4* __restrict__ units, AfHyper h) {
  __shared__ double redd[AF_THREADS / 32];
  __shared__ float redf[AF_THREADS / 32];
  const int4 u = units[blockIdx.x];
  const v2f_af_desc d = descs[u.x];
  const long long off = ((long long)u.y * d.R + u.z) * d.C;
  const float* g = grads[u.x] + off;
  const float* p = d.p + off;
  float s = 0.f;
  double sp = 0.0;
  for (int c = threadIdx.x; c < d.C; c += AF_THREADS) {
    const float gv = g[c], pv = p[c];
    s += gv * gv + h.eps1;
    sp += (double)pv * pv;
  }
  s = block_sum_f(s, redf);
  sp = block_sum_d(sp, redd);
  if (threadIdx.x == 0) {
    float* row = d.row + (long long)u.y * d.R + u.z;
    *row = *row * h.beta + (s / (float)d.C) * h.omb;
    atomicAdd(acc + 2 * u.x, sp);
  }
}

// CTA = 32 columns x 8 row lanes: thread (cl, rl) sums the rows r = rl (mod 8) of column u.z + cl, four loads in
// flight; the 8 partials meet in shared memory.  (A thread per column alone left 3 CTAs with 1536 serial rows each
// for the GRU weights.)
__global__ void __launch_bounds__(AF_THREADS)
af_big_colstat_kernel(const v2f_af_desc* __restrict__ descs, const float* const* __restrict__ grads,
                      const int4* __restrict__ units, AfHyper h) {
  __shared__ float part[AF_ROWLANES][AF_COLSTRIP + 1];
  const int4 u = units[blockIdx.x];
  const v2f_af_desc d = descs[u.x];
  const int cl = threadIdx.x % AF_COLSTRIP, rl = threadIdx.x / AF_COLSTRIP;
  const int c = u.z + cl;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < d.C) {
    const float* g = grads[u.x] + (long long)u.y * d.R * d.C + c;
    const long long st = (long long)AF_ROWLANES * d.C;
    int r = rl;
    for (; r + 3 * AF_ROWLANES < d.R; r += 4 * AF_ROWLANES) {
      const float* gp = g + (long long)r * d.C;
      const float a0 = gp[0], a1 = gp[st], a2 = gp[2 * st], a3 = gp[3 * st];
      s0 += a0 * a0 + h.eps1;
      s1 += a1 * a1 + h.eps1;
      s2 += a2 * a2 + h.eps1;
      s3 += a3 * a3 + h.eps1;
    }
    for (; r < d.R; r += AF_ROWLANES) {
      const float a0 = g[(long long)r * d.C];
      s0 += a0 * a0 + h.eps1;
    }
  }
  part[rl][cl] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (rl == 0 && c < d.C) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < AF_ROWLANES; k++) s += part[k][cl];
    float* col = d.col + (long long)u.y * d.C + c;
    *col = *col * h.beta + (s / (float)d.R) * h.omb;
  }
}

// u = g * rsqrt(row[r] / mean(row)) * rsqrt(col[c]); APPLY: p -= lr u / cdiv, else accumulate sum u^2
template <bool APPLY>
__global__ void __launch_bounds__(AF_THREADS)
af_big_update_kernel(const v2f_af_desc* __restrict__ descs, const float* const* __restrict__ grads,
                     double* __restrict__ acc, const int4* __restrict__ units, AfHyper h) {
  __shared__ double redd[AF_THREADS / 32];
  __shared__ float redf[AF_THREADS / 32];
  const int4 u = units[blockIdx.x];
  const v2f_af_desc d = descs[u.x];
  const float* rowm = d.row + (long long)u.y * d.R;
  const float* colm = d.col + (long long)u.y * d.C;
  float rs = 0.f;
  for (int r = threadIdx.x; r < d.R; r += AF_THREADS) rs += rowm[r];
  const float rmean = block_sum_f(rs, redf) / (float)d.R;
  const float rfac = rsqrtf(rowm[u.z] / rmean);
  const long long off = ((long long)u.y * d.R + u.z) * d.C;
  const float* g = grads[u.x] + off;
  float* p = d.p + off;
  float lr = 0.f, cdiv = 1.f, rms = 0.f;
  if (APPLY) {
    step_scalars(d, acc + 2 * u.x, h, lr, cdiv, rms);
    if (u.y == 0 && u.z == 0 && threadIdx.x == 0) *d.rms = rms;
  }
  double su = 0.0;
  for (int c = threadIdx.x; c < d.C; c += AF_THREADS) {
    const float uu = g[c] * (rfac * rsqrtf(colm[c]));
    if (APPLY) p[c] = p[c] - (uu / cdiv) * lr;
    else su += (double)uu * uu;
  }
  if (!APPLY) {
    su = block_sum_d(su, redd);
    if (threadIdx.x == 0) atomicAdd(acc + 2 * u.x + 1, su);
  }
}

}  // namespace v2f

using namespace v2f;

extern "C" int v2f_adafactor_step(const v2f_adafactor_plan* pl, double beta2t, double rel_step, void* st) {
  V2F_REQUIRE(pl && pl->descs && pl->grads && pl->acc && pl->n_desc > 0, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(pl->n_vec == 0 || pl->vec_units, V2F_ERR_BAD_ARG);
  for (int k = 0; k < 3; k++) V2F_REQUIRE(pl->n_small[k] == 0 || pl->small_units[k], V2F_ERR_BAD_ARG);
  V2F_REQUIRE((pl->n_rows == 0 || pl->row_units) && (pl->n_cols == 0 || pl->col_units), V2F_ERR_BAD_ARG);
  cudaStream_t s = (cudaStream_t)st;
  AfHyper h;
  h.beta = (float)beta2t;
  h.omb = (float)(1.0 - beta2t);
  h.eps1 = pl->eps1;
  h.eps2 = pl->eps2;
  h.clip = pl->clip_threshold;
  h.rel_step = rel_step;
  h.scale_parameter = pl->scale_parameter;
  const v2f_af_desc* D = pl->descs;
  const float* const* G = (const float* const*)pl->grads;
  const int4* VU = (const int4*)pl->vec_units;
  const int4* RU = (const int4*)pl->row_units;
  const int4* CU = (const int4*)pl->col_units;
  if (cudaMemsetAsync(pl->acc, 0, sizeof(double) * 2 * (size_t)pl->n_desc, s) != cudaSuccess) return V2F_ERR_LAUNCH;
  // statistics, second-moment states, the two norms
  if (pl->n_vec) {
    af_vec_kernel<false><<<pl->n_vec, AF_THREADS, 0, s>>>(D, G, pl->acc, VU, h);
    V2F_CHECK_LAUNCH();
  }
  if (pl->n_small[0]) {
    af_small_kernel<false, 1, 1><<<pl->n_small[0], AF_THREADS, 0, s>>>(D, G, pl->acc, (const int4*)pl->small_units[0], h);
    V2F_CHECK_LAUNCH();
  }
  if (pl->n_small[1]) {
    af_small_kernel<false, 3, 3><<<pl->n_small[1], AF_THREADS, 0, s>>>(D, G, pl->acc, (const int4*)pl->small_units[1], h);
    V2F_CHECK_LAUNCH();
  }
  if (pl->n_small[2]) {
    af_small_kernel<false, 0, 0><<<pl->n_small[2], AF_THREADS, 0, s>>>(D, G, pl->acc, (const int4*)pl->small_units[2], h);
    V2F_CHECK_LAUNCH();
  }
  if (pl->n_rows) {
    af_big_rowstat_kernel<<<pl->n_rows, AF_THREADS, 0, s>>>(D, G, pl->acc, RU, h);
    V2F_CHECK_LAUNCH();
    af_big_colstat_kernel<<<pl->n_cols, AF_THREADS, 0, s>>>(D, G, CU, h);
    V2F_CHECK_LAUNCH();
    af_big_update_kernel<false><<<pl->n_rows, AF_THREADS, 0, s>>>(D, G, pl->acc, RU, h);
    V2F_CHECK_LAUNCH();
  }
  // parameter update
  if (pl->n_vec) {
    af_vec_kernel<true><<<pl->n_vec, AF_THREADS, 0, s>>>(D, G, pl->acc, VU, h);
    V2F_CHECK_LAUNCH();
  }
  if (pl->n_small[0]) {
    af_small_kernel<true, 1, 1><<<pl->n_small[0], AF_THREADS, 0, s>>>(D, G, pl->acc, (const int4*)pl->small_units[0], h);
    V2F_CHECK_LAUNCH();
  }
  if (pl->n_small[1]) {
    af_small_kernel<true, 3, 3><<<pl->n_small[1], AF_THREADS, 0, s>>>(D, G, pl->acc, (const int4*)pl->small_units[1], h);
    V2F_CHECK_LAUNCH();
  }
  if (pl->n_small[2]) {
    af_small_kernel<true, 0, 0><<<pl->n_small[2], AF_THREADS, 0, s>>>(D, G, pl->acc, (const int4*)pl->small_units[2], h);
    V2F_CHECK_LAUNCH();
  }
  if (pl->n_rows) {
    af_big_update_kernel<true><<<pl->n_rows, AF_THREADS, 0, s>>>(D, G, pl->acc, RU, h);
    V2F_CHECK_LAUNCH();
  }
  return V2F_OK;
}

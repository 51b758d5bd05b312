// Fused BatchNorm2d (+ residual add) (+ ReLU) over bf16 channels_last activations, forward and
// backward.  sm_100a.
//
// Reference arithmetic: the torchvision ResNet-101 trunk of ImageEncoder
// (/root/reference/models/CrossAttnRNN210.py:58-72; layer3/layer4 trainable :63-65, the whole trunk
// in train() mode so every BatchNorm2d normalises with batch statistics and updates its running
// statistics): torchvision Bottleneck.forward = conv1-bn1-relu, conv2-bn2-relu, conv3-bn3,
// (+ downsample conv-bn), add, relu.  The convolutions stay cuDNN; the normalisation, the residual
// add and the ReLU between them are HBM-bound row sweeps over [R = N*H*W, C] and are done here in
// two passes per direction instead of torch's 4-5 separate elementwise / reduction kernels:
//   forward : stats (read x)  ->  finalize (per channel)  ->  apply (read x [,res], write y)
//   backward: reduce (read dy, x [,y]; write dz if a residual consumes it)  ->  finalize  ->
//             elemt (read dz|dy, x; write dx)
// Activations are bf16 (16-byte = 8-channel vectors, C % 8 == 0), statistics / parameters fp32,
// partial sums per block are merged in double by the finalize kernels (no atomics, deterministic).
#include <cuda_bf16.h>

#include <unordered_map>

#include "common.cuh"

namespace v2f {

constexpr int BN_THREADS = 256;

// Programmatic dependent launch inside the stats -> finalize -> apply (and reduce -> finalize -> elemt) chains: a
// kernel of the chain tells the scheduler at its start that its successor may be made resident (launch_dependents),
// and the successor blocks at pdl_wait() until the predecessor grid has completed and its writes are visible.  The
// launch latency and ramp-up of ~370 small launches per step then overlap the predecessor's tail.  Both instructions
// are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static bool g_bn_pdl = true;

template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool dependent,
                                Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (dependent && g_bn_pdl) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

struct __align__(16) bf16x8 { __nv_bfloat162 a, b, c, d; };

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Thread geometry shared by all sweeps: CV = C/8 vectors per row.  A block is viewed as
// [RB rows][CVB vector columns] with CVB = min(CV, 256); when CV > 256 a thread loops over columns
// (not needed for ResNet, C <= 2048).  Rows are dealt to blocks in an interleaved, grid-strided way.
struct Geo {
  int CV, CVB, RB;
};
__host__ __device__ inline Geo make_geo(int C) {
  Geo g;
  g.CV = C / 8;
  g.CVB = g.CV < BN_THREADS ? g.CV : BN_THREADS;
  g.RB = BN_THREADS / g.CVB;
  return g;
}

// ---------------------------------------------------------------- forward: statistics
// part[blk][0][c] = sum x, part[blk][1][c] = sum x^2 over the rows of block blk.
__global__ void __launch_bounds__(BN_THREADS)
bn_stats_kernel(long long R, int C, const uint4* __restrict__ x, float* __restrict__ part) {
  extern __shared__ float sm[];   // [RB][2][CVB*8]
  pdl_launch_dependents();        // the finalize kernel may become resident while this sweep runs
  const Geo g = make_geo(C);
  const int vcol = threadIdx.x % g.CVB, roff = threadIdx.x / g.CVB;
  for (int v0 = 0; v0 < g.CV; v0 += g.CVB) {
    const int v = v0 + vcol;
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = q[i] = 0.f;
    if (roff < g.RB) {
      const long long stride = (long long)gridDim.x * g.RB;
      long long r = (long long)blockIdx.x * g.RB + roff;
      // the last round is predicated instead of falling into a serial tail (small late-layer tensors are 2-3 rounds
      // in total: a tail of dependent round trips was a third of their time)
      // a read-only sweep needs more bytes in flight than the read+write ones to cover the HBM latency: 8 x 16 B per
      // thread, 128 KB per SM (with 4 it stopped at 0.61 of the HBM peak on 369 MB tensors, the apply pass at 0.87)
      constexpr int UN = 8;
      for (; r < R; r += UN * stride) {
        uint4 u[UN];
#pragma unroll
        for (int k = 0; k < UN; k++)
          u[k] = (r + k * stride < R) ? ldg_stream(x + (r + k * stride) * g.CV + v) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int k = 0; k < UN; k++) {
          float f[8];
          unpack8(u[k], f);
#pragma unroll
          for (int i = 0; i < 8; i++) {
            s[i] += f[i];
            q[i] = fmaf(f[i], f[i], q[i]);
          }
        }
      }
    }
    const int W8 = g.CVB * 8;
    if (roff < g.RB) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        sm[(roff * 2 + 0) * W8 + vcol * 8 + i] = s[i];
        sm[(roff * 2 + 1) * W8 + vcol * 8 + i] = q[i];
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * W8; j += BN_THREADS) {
      const int which = j / W8, cc = j - which * W8;
      float t = 0.f;
      for (int rr = 0; rr < g.RB; rr++) t += sm[(rr * 2 + which) * W8 + cc];
      part[((long long)blockIdx.x * 2 + which) * C + v0 * 8 + cc] = t;
    }
    __syncthreads();
  }
}

// Merge of the block partials of one channel group: block (32 channels, 32 slices of the partial
// range), two independent double accumulators per thread so the loads pipeline, fixed order.
// Returns the totals to the threads with threadIdx.y == 0.
constexpr int FIN_X = 8;      // channels per finalize block (one 32-byte sector of a partial row)
constexpr int FIN_Y = 128;    // slices of the partial range per block
constexpr int FIN_R = 5;      // partial rows per thread: FIN_Y * FIN_R >= 148 * 4 blocks of a sweep
// Merge of the block partials of FIN_X channels by one block of FIN_X x FIN_Y threads: every thread issues its (at most
// five) row loads at once, so the merge is one L2 round trip deep; C / 8 blocks instead of C / 32 (the 64-channel layers
// had 2 blocks walking 592 rows).  Fixed summation order, double.  Totals returned to the threads with threadIdx.y == 0.
__device__ __forceinline__ void merge_partials(int C, int c, int nblk, const float* __restrict__ part,
                                               double& S, double& Q) {
  __shared__ double red[2][FIN_Y][FIN_X + 1];
  double s = 0.0, q = 0.0;
  if (c < C) {
    for (int b0 = threadIdx.y; b0 < nblk; b0 += FIN_Y * FIN_R) {
      float vs[FIN_R], vq[FIN_R];
#pragma unroll
      for (int k = 0; k < FIN_R; k++) {
        const int b = b0 + k * FIN_Y;
        vs[k] = b < nblk ? part[((long long)b * 2) * C + c] : 0.f;
        vq[k] = b < nblk ? part[((long long)b * 2 + 1) * C + c] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < FIN_R; k++) {
        s += (double)vs[k];
        q += (double)vq[k];
      }
    }
  }
  red[0][threadIdx.y][threadIdx.x] = s;
  red[1][threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y < 8) {          // 128 slices -> 8 sums of 16
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int i = 0; i < FIN_Y / 8; i++) {
      a += red[0][threadIdx.y + 8 * i][threadIdx.x];
      b += red[1][threadIdx.y + 8 * i][threadIdx.x];
    }
    __syncwarp();
    red[0][threadIdx.y][threadIdx.x] = a;      // rows 0..7 are read above only by their own thread (i = 0)
    red[1][threadIdx.y][threadIdx.x] = b;
  }
  __syncthreads();
  S = Q = 0.0;
  if (threadIdx.y == 0)
#pragma unroll
    for (int i = 0; i < 8; i++) {
      S += red[0][i][threadIdx.x];
      Q += red[1][i][threadIdx.x];
    }
}

// Per channel: merge the block partials, produce scale/shift, the saved mean / rstd and the
// running-statistics update (momentum, unbiased variance).  grid = ceil(C/8), block (8,128).
__global__ void bn_fwd_finalize_kernel(long long R, int C, int nblk, const float* __restrict__ part,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float* __restrict__ run_mean, float* __restrict__ run_var,
                                       int training, float momentum, float eps, float* __restrict__ scale,
                                       float* __restrict__ shift, float* __restrict__ save_mean,
                                       float* __restrict__ save_rstd) {
  const int c = blockIdx.x * FIN_X + threadIdx.x;
  double S = 0.0, Q = 0.0;
  pdl_launch_dependents();
  pdl_wait();                     // the statistics sweep (or the stem convolution) has completed
  if (training) merge_partials(C, c, nblk, part, S, Q);
  if (c >= C || threadIdx.y != 0) return;
  float mean, var;
  if (training) {
    const double m = S / (double)R;
    double vv = Q / (double)R - m * m;
    if (vv < 0.0) vv = 0.0;
    mean = (float)m;
    var = (float)vv;
    if (run_mean) {
      const float unb = R > 1 ? (float)(vv * (double)R / (double)(R - 1)) : var;
      run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
      run_var[c] = (1.f - momentum) * run_var[c] + momentum * unb;
    }
  } else {
    mean = run_mean[c];
    var = run_var[c];
  }
  const float r = 1.0f / sqrtf(var + eps);
  const float sc = gamma[c] * r;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  save_mean[c] = mean;
  save_rstd[c] = r;
}

// ---------------------------------------------------------------- forward: apply
// y = act(x * scale + shift (+ res))
template <bool RES, bool RELU>
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_kernel(long long R, int C, const uint4* __restrict__ x, const uint4* __restrict__ res,
                const float* __restrict__ scale, const float* __restrict__ shift, uint4* __restrict__ y) {
  const Geo g = make_geo(C);
  const int vcol = threadIdx.x % g.CVB, roff = threadIdx.x / g.CVB;
  pdl_wait();                     // scale / shift come from the finalize kernel
  if (roff >= g.RB) return;
  for (int v0 = 0; v0 < g.CV; v0 += g.CVB) {
    const int v = v0 + vcol;
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      sc[i] = scale[v * 8 + i];
      sh[i] = shift[v * 8 + i];
    }
    const long long stride = (long long)gridDim.x * g.RB;
    long long r = (long long)blockIdx.x * g.RB + roff;
    for (; r < R; r += 4 * stride) {
      uint4 u[4], w[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const bool ok = r + k * stride < R;
        u[k] = ok ? ldg_stream(x + (r + k * stride) * g.CV + v) : make_uint4(0u, 0u, 0u, 0u);
        if (RES) w[k] = ok ? ldg_stream(res + (r + k * stride) * g.CV + v) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float f[8], h[8];
        unpack8(u[k], f);
        if (RES) unpack8(w[k], h);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          float t = fmaf(f[i], sc[i], sh[i]);
          if (RES) t += h[i];
          f[i] = RELU ? fmaxf(t, 0.f) : t;
        }
        if (r + k * stride < R) y[(r + k * stride) * g.CV + v] = pack8(f);
      }
    }
  }
}

// ---------------------------------------------------------------- backward: reduce
// dz = dy * [y > 0] (RELU) ; part[blk][0][c] = sum dz, part[blk][1][c] = sum dz * xhat.
// WRITE_DZ: dz is also stored (the residual branch receives it as its gradient).
// ADD2: the output fed two consumers (the next block's conv1 and its residual add); their gradients
// arrive separately (dy, dy2) and are summed here instead of by a separate elementwise kernel.
template <bool RELU, bool WRITE_DZ, bool ADD2>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_reduce_kernel(long long R, int C, const uint4* __restrict__ dy, const uint4* __restrict__ dy2,
                     const uint4* __restrict__ x, const uint4* __restrict__ y, const float* __restrict__ mean,
                     const float* __restrict__ rstd, uint4* __restrict__ dz, float* __restrict__ part) {
  extern __shared__ float sm[];
  const Geo g = make_geo(C);
  const int vcol = threadIdx.x % g.CVB, roff = threadIdx.x / g.CVB;
  pdl_launch_dependents();
  for (int v0 = 0; v0 < g.CV; v0 += g.CVB) {
    const int v = v0 + vcol;
    float s[8], q[8], mu[8], rs[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      s[i] = q[i] = 0.f;
      mu[i] = mean[v * 8 + i];
      rs[i] = rstd[v * 8 + i];
    }
    if (roff < g.RB) {
      const long long stride = (long long)gridDim.x * g.RB;
      long long r = (long long)blockIdx.x * g.RB + roff;
      const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
      for (; r < R; r += 2 * stride) {
        uint4 ud[2], ue[2], ux[2], uy[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
          const bool ok = r + k * stride < R;       // predicated last round: zero gradient adds nothing to the sums
          ud[k] = ok ? ldg_stream(dy + (r + k * stride) * g.CV + v) : z4;
          if (ADD2) ue[k] = ok ? ldg_stream(dy2 + (r + k * stride) * g.CV + v) : z4;
          ux[k] = ok ? ldg_stream(x + (r + k * stride) * g.CV + v) : z4;
          if (RELU) uy[k] = ok ? ldg_stream(y + (r + k * stride) * g.CV + v) : z4;
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
          float d[8], e[8], xv[8], yv[8];
          unpack8(ud[k], d);
          if (ADD2) unpack8(ue[k], e);
          unpack8(ux[k], xv);
          if (RELU) unpack8(uy[k], yv);
#pragma unroll
          for (int i = 0; i < 8; i++) {
            if (ADD2) d[i] += e[i];
            if (RELU && !(yv[i] > 0.f)) d[i] = 0.f;
            s[i] += d[i];
            q[i] = fmaf(d[i], (xv[i] - mu[i]) * rs[i], q[i]);
          }
          if (WRITE_DZ && r + k * stride < R) dz[(r + k * stride) * g.CV + v] = pack8(d);
        }
      }
    }
    const int W8 = g.CVB * 8;
    if (roff < g.RB) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        sm[(roff * 2 + 0) * W8 + vcol * 8 + i] = s[i];
        sm[(roff * 2 + 1) * W8 + vcol * 8 + i] = q[i];
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * W8; j += BN_THREADS) {
      const int which = j / W8, cc = j - which * W8;
      float t = 0.f;
      for (int rr = 0; rr < g.RB; rr++) t += sm[(rr * 2 + which) * W8 + cc];
      part[((long long)blockIdx.x * 2 + which) * C + v0 * 8 + cc] = t;
    }
    __syncthreads();
  }
}

// dgamma = sum dz*xhat, dbeta = sum dz; coefficients of the elementwise pass:
//   dx = a * (dz - m1 - xhat * m2),  a = gamma*rstd, m1 = dbeta/R, m2 = dgamma/R   (training)
//      = a * dz + k1 * x + k0,       k1 = -a*rstd*m2, k0 = -a*m1 - k1*mean        (three coefficients per channel)
//   dx = a * dz                                                                      (eval: k1 = k0 = 0)
__global__ void bn_bwd_finalize_kernel(long long R, int C, int nblk, const float* __restrict__ part,
                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                       const float* __restrict__ rstd,
                                       int training, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ coef) {
  const int c = blockIdx.x * FIN_X + threadIdx.x;
  double S, Q;
  pdl_launch_dependents();
  pdl_wait();
  merge_partials(C, c, nblk, part, S, Q);
  if (c >= C || threadIdx.y != 0) return;
  dbeta[c] = (float)S;
  dgamma[c] = (float)Q;
  const float a = gamma[c] * rstd[c];
  const float m1 = training ? (float)(S / (double)R) : 0.f, m2 = training ? (float)(Q / (double)R) : 0.f;
  const float k1 = -a * rstd[c] * m2;
  coef[c] = a;
  coef[C + c] = k1;
  coef[2 * C + c] = -a * m1 - k1 * mean[c];
}

// dx = a * dz + k1 * x + k0; dz either stored by the reduce pass (HAVE_DZ) or rebuilt from dy & y.
template <bool RELU, bool HAVE_DZ>
__global__ void __launch_bounds__(BN_THREADS, 4)
bn_bwd_elemt_kernel(long long R, int C, const uint4* __restrict__ dy, const uint4* __restrict__ x,
                    const uint4* __restrict__ y, const float* __restrict__ coef, uint4* __restrict__ dx) {
  const Geo g = make_geo(C);
  const int vcol = threadIdx.x % g.CVB, roff = threadIdx.x / g.CVB;
  pdl_wait();                     // coefficients (and dz) come from the finalize / reduce kernels
  if (roff >= g.RB) return;
  constexpr int UN = (RELU && !HAVE_DZ) ? 2 : 3;      // 16-byte loads in flight per thread: 6
  for (int v0 = 0; v0 < g.CV; v0 += g.CVB) {
    const int v = v0 + vcol;
    float a[8], k1[8], k0[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int c = v * 8 + i;
      a[i] = coef[c];
      k1[i] = coef[C + c];
      k0[i] = coef[2 * C + c];
    }
    const long long stride = (long long)gridDim.x * g.RB;
    long long r = (long long)blockIdx.x * g.RB + roff;
    for (; r + (UN - 1) * stride < R; r += UN * stride) {
      uint4 ud[UN], ux[UN], uy[UN];
#pragma unroll
      for (int k = 0; k < UN; k++) {
        ud[k] = ldg_stream(dy + (r + k * stride) * g.CV + v);
        ux[k] = ldg_stream(x + (r + k * stride) * g.CV + v);
        if (RELU && !HAVE_DZ) uy[k] = ldg_stream(y + (r + k * stride) * g.CV + v);
      }
#pragma unroll
      for (int k = 0; k < UN; k++) {
        float d[8], xv[8], yv[8];
        unpack8(ud[k], d);
        unpack8(ux[k], xv);
        if (RELU && !HAVE_DZ) unpack8(uy[k], yv);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (RELU && !HAVE_DZ && !(yv[i] > 0.f)) d[i] = 0.f;
          d[i] = fmaf(a[i], d[i], fmaf(k1[i], xv[i], k0[i]));
        }
        dx[(r + k * stride) * g.CV + v] = pack8(d);
      }
    }
    for (; r < R; r += stride) {       // (a predicated last round costs registers this kernel does not have: 64, 4 CTAs/SM)
      float d[8], xv[8], yv[8];
      unpack8(ldg_stream(dy + r * g.CV + v), d);
      unpack8(ldg_stream(x + r * g.CV + v), xv);
      if (RELU && !HAVE_DZ) unpack8(ldg_stream(y + r * g.CV + v), yv);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (RELU && !HAVE_DZ && !(yv[i] > 0.f)) d[i] = 0.f;
        d[i] = fmaf(a[i], d[i], fmaf(k1[i], xv[i], k0[i]));
      }
      dx[r * g.CV + v] = pack8(d);
    }
  }
}

// ---------------------------------------------------------------- stem: BN + ReLU + 3x3/2 max-pool (forward only)
// y[n,oh,ow,:] = max over the valid 3x3 window (stride 2, pad 1) of relu(x*scale+shift).  One thread per
// (output pixel, 8 channels); the overlapping window reads are served by L1/L2, DRAM sees x once.
__global__ void __launch_bounds__(BN_THREADS)
bn_relu_maxpool_kernel(int N, int H, int W, int C, int OH, int OW, const uint4* __restrict__ x,
                       const float* __restrict__ scale, const float* __restrict__ shift, uint4* __restrict__ y) {
  const int CV = C / 8;
  const long long total = (long long)N * OH * OW * CV;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();
  if (i >= total) return;
  const int v = (int)(i % CV);
  long long p = i / CV;
  const int ow = (int)(p % OW);
  p /= OW;
  const int oh = (int)(p % OH);
  const int n = (int)(p / OH);
  float sc[8], sh[8], m[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    sc[k] = scale[v * 8 + k];
    sh[k] = shift[v * 8 + k];
    m[k] = 0.f;                       // relu output >= 0 and every window holds a valid pixel
  }
#pragma unroll
  for (int dh = 0; dh < 3; dh++) {
    const int ih = 2 * oh - 1 + dh;
    if (ih < 0 || ih >= H) continue;
#pragma unroll
    for (int dw = 0; dw < 3; dw++) {
      const int iw = 2 * ow - 1 + dw;
      if (iw < 0 || iw >= W) continue;
      float f[8];
      unpack8(__ldg(x + (((long long)n * H + ih) * W + iw) * CV + v), f);
#pragma unroll
      for (int k = 0; k < 8; k++) m[k] = fmaxf(m[k], fmaf(f[k], sc[k], sh[k]));
    }
  }
  y[i] = pack8(m);
}

// Resident CTAs per SM of one kernel instantiation (cached): the sweeps are sized to exactly one wave --
// a kernel that needs 80-98 registers fits 2-3 CTAs of 256 threads per SM, and a grid of 4 per SM would then
// run a second, quarter-full wave.
template <typename K>
static int resident_per_sm(K kern, size_t smem) {
  // keyed by the kernel's ADDRESS: the instantiations of one template share a function type, so a function-local
  // static would hand the first variant's occupancy to all of them
  static std::unordered_map<const void*, int> cache;
  const void* key = reinterpret_cast<const void*>(kern);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, BN_THREADS, smem) != cudaSuccess || n < 1) n = 1;
  if (n > 4) n = 4;
  cache[key] = n;
  return n;
}
template <typename K>
static int wave_blocks(K kern, size_t smem, long long R, int C) {
  const Geo g = make_geo(C);
  const long long need = (R + g.RB - 1) / g.RB;
  const long long cap = 148LL * resident_per_sm(kern, smem);
  return (int)(need < cap ? need : cap);
}

static inline int sweep_blocks(long long R, int C) {
  const Geo g = make_geo(C);
  long long need = (R + g.RB - 1) / g.RB;           // one block-iteration per RB rows
  // 4 CTAs of 256 threads per SM (6 and 8 were measured slower for the statistics sweep: tools/bn_probe.py)
  long long cap = 148 * 4;
  return (int)(need < cap ? need : cap);
}

}  // namespace v2f

using namespace v2f;

// A/B switch (default 1): 0 launches the chains without programmatic dependent launch.
extern "C" int v2f_bn2d_pdl_enable(int on) {
  g_bn_pdl = on != 0;
  return V2F_OK;
}

extern "C" int v2f_bn2d_blocks(long long R, int C) {
  if (R <= 0 || C <= 0 || (C & 7)) return 0;
  return sweep_blocks(R, C);
}

extern "C" int v2f_bn2d_act_fwd(long long R, int C, const void* x, const void* res, const float* gamma,
                                const float* beta, float* run_mean, float* run_var, int training,
                                float momentum, float eps, int relu, void* y, float* save_mean,
                                float* save_rstd, float* scale_shift, float* part, void* st) {
  V2F_REQUIRE(R > 0 && C > 0 && (C & 7) == 0, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(x && gamma && beta && y && save_mean && save_rstd && scale_shift && part, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(training || (run_mean && run_var), V2F_ERR_BAD_ARG);
  V2F_REQUIRE(aligned16(x) && aligned16(y) && (!res || aligned16(res)), V2F_ERR_ALIGN);
  cudaStream_t s = (cudaStream_t)st;
  const Geo g = make_geo(C);
  const int nblk = sweep_blocks(R, C);
  const size_t smem = sizeof(float) * (size_t)g.RB * 2 * g.CVB * 8;
  const long long tb = R * (long long)C * 2;     // bytes of one bf16 [R,C] tensor
  if (training) {
    prof_begin(V2F_K_BN_STATS, s);
    prof_bytes(V2F_K_BN_STATS, tb);
    bn_stats_kernel<<<nblk, BN_THREADS, smem, s>>>(R, C, (const uint4*)x, part);
    prof_end(V2F_K_BN_STATS, s);
    V2F_CHECK_LAUNCH();
  }
  launch_chain(bn_fwd_finalize_kernel, dim3((C + FIN_X - 1) / FIN_X), dim3(FIN_X, FIN_Y), 0, s, training != 0, R, C, nblk,
               part, gamma, beta, run_mean, run_var, training, momentum, eps, scale_shift, scale_shift + C, save_mean,
               save_rstd);
  V2F_CHECK_LAUNCH();
  prof_begin(V2F_K_BN_APPLY, s);
  prof_bytes(V2F_K_BN_APPLY, tb * (res ? 3 : 2));
  const uint4 *xp = (const uint4*)x, *rp = (const uint4*)res;
  uint4* yp = (uint4*)y;
  const float *sc = scale_shift, *sh = scale_shift + C;
#define BN_APPLY(RES_, RELU_)                                                                                          \
  launch_chain(bn_apply_kernel<RES_, RELU_>, dim3(wave_blocks(bn_apply_kernel<RES_, RELU_>, 0, R, C)), dim3(BN_THREADS), 0, s, \
               true, R, C, xp, rp, sc, sh, yp)
  if (res && relu) BN_APPLY(true, true);
  else if (res) BN_APPLY(true, false);
  else if (relu) BN_APPLY(false, true);
  else BN_APPLY(false, false);
#undef BN_APPLY
  prof_end(V2F_K_BN_APPLY, s);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_bn2d_act_bwd(long long R, int C, const void* dy, const void* dy2, const void* x, const void* y,
                                const float* gamma, const float* save_mean, const float* save_rstd,
                                int training, int relu, void* dz, void* dx, float* dgamma, float* dbeta,
                                float* coef, float* part, void* st) {
  V2F_REQUIRE(R > 0 && C > 0 && (C & 7) == 0, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(dy && x && gamma && save_mean && save_rstd && dx && dgamma && dbeta && coef && part, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(!relu || y, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(!dy2 || (dz && aligned16(dy2)), V2F_ERR_BAD_ARG);     // the summed gradient has to be stored
  V2F_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(dx) && (!y || aligned16(y)) && (!dz || aligned16(dz)),
              V2F_ERR_ALIGN);
  cudaStream_t s = (cudaStream_t)st;
  const Geo g = make_geo(C);
  const int nblk = sweep_blocks(R, C);
  const size_t smem = sizeof(float) * (size_t)g.RB * 2 * g.CVB * 8;
  const uint4 *dyp = (const uint4*)dy, *dy2p = (const uint4*)dy2, *xp = (const uint4*)x, *yp = (const uint4*)y;
  uint4 *dzp = (uint4*)dz, *dxp = (uint4*)dx;
  const long long tb = R * (long long)C * 2;
  prof_begin(V2F_K_BN_BWD_REDUCE, s);
  const bool have_dz = dz && (relu || dy2);
  prof_bytes(V2F_K_BN_BWD_REDUCE, tb * (2 + (relu ? 1 : 0) + (have_dz ? 1 : 0) + (dy2 ? 1 : 0)));
  int nred = nblk;       // blocks of the reduce pass = partial rows the finalize merges (<= sweep_blocks = rows of ``part``)
#define BN_RED(RELU_, WDZ_, ADD2_)                                                                    \
  do {                                                                                                \
    nred = wave_blocks(bn_bwd_reduce_kernel<RELU_, WDZ_, ADD2_>, smem, R, C);                         \
    bn_bwd_reduce_kernel<RELU_, WDZ_, ADD2_><<<nred, BN_THREADS, smem, s>>>(R, C, dyp, dy2p, xp, yp, save_mean, \
                                                                           save_rstd, dzp, part);     \
  } while (0)
  if (dy2 && relu) BN_RED(true, true, true);
  else if (dy2) BN_RED(false, true, true);
  else if (relu && dz) BN_RED(true, true, false);
  else if (relu) BN_RED(true, false, false);
  else BN_RED(false, false, false);
#undef BN_RED
  prof_end(V2F_K_BN_BWD_REDUCE, s);
  V2F_CHECK_LAUNCH();
  launch_chain(bn_bwd_finalize_kernel, dim3((C + FIN_X - 1) / FIN_X), dim3(FIN_X, FIN_Y), 0, s, true, R, C, nred, part, gamma,
               save_mean, save_rstd, training, dgamma, dbeta, coef);
  V2F_CHECK_LAUNCH();
  prof_begin(V2F_K_BN_BWD_ELEMT, s);
  prof_bytes(V2F_K_BN_BWD_ELEMT, tb * (3 + ((relu && !dz) ? 1 : 0)));
#define BN_ELEMT(RELU_, HAVE_, SRC_)                                                                                  \
  launch_chain(bn_bwd_elemt_kernel<RELU_, HAVE_>, dim3(wave_blocks(bn_bwd_elemt_kernel<RELU_, HAVE_>, 0, R, C)),        \
               dim3(BN_THREADS), 0, s, true, R, C, SRC_, xp, yp, coef, dxp)
  if (have_dz && !relu) BN_ELEMT(false, false, (const uint4*)dz);
  else if (relu && dz) BN_ELEMT(true, true, (const uint4*)dz);
  else if (relu) BN_ELEMT(true, false, dyp);
  else BN_ELEMT(false, false, dyp);
#undef BN_ELEMT
  prof_end(V2F_K_BN_BWD_ELEMT, s);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

// Stem of the trunk: y = maxpool3x3/2(relu(BN(x))), x bf16 [N,H,W,C], y bf16 [N,OH,OW,C] with
// OH = (H-1)/2+1, OW = (W-1)/2+1.  Forward only (the stem is frozen in the reference).
// part_blocks > 0: the statistics arrive as that many partial rows from the producer of x (stem_conv.cu).
static int stem_bn_pool(int N, int H, int W, int C, const void* x, const float* gamma, const float* beta,
                        float* run_mean, float* run_var, int training, float momentum, float eps, void* y,
                        float* save_mean, float* save_rstd, float* scale_shift, float* part, int part_blocks,
                        cudaStream_t s) {
  const long long R = (long long)N * H * W;
  const Geo g = make_geo(C);
  int nblk = part_blocks;
  if (training && part_blocks <= 0) {
    nblk = sweep_blocks(R, C);
    const size_t smem = sizeof(float) * (size_t)g.RB * 2 * g.CVB * 8;
    bn_stats_kernel<<<nblk, BN_THREADS, smem, s>>>(R, C, (const uint4*)x, part);
    V2F_CHECK_LAUNCH();
  }
  // after the stem convolution's epilogue statistics (part_blocks > 0) the finalize kernel follows a kernel that does
  // not signal: a plain launch there
  launch_chain(bn_fwd_finalize_kernel, dim3((C + FIN_X - 1) / FIN_X), dim3(FIN_X, FIN_Y), 0, s,
               training != 0 && part_blocks <= 0, R, C, nblk, part, gamma, beta, run_mean, run_var, training, momentum, eps,
               scale_shift, scale_shift + C, save_mean, save_rstd);
  V2F_CHECK_LAUNCH();
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const long long total = (long long)N * OH * OW * (C / 8);
  launch_chain(bn_relu_maxpool_kernel, dim3((unsigned)((total + BN_THREADS - 1) / BN_THREADS)), dim3(BN_THREADS), 0, s, true,
               N, H, W, C, OH, OW, (const uint4*)x, (const float*)scale_shift, (const float*)(scale_shift + C), (uint4*)y);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_bn2d_relu_maxpool_fwd(int N, int H, int W, int C, const void* x, const float* gamma,
                                         const float* beta, float* run_mean, float* run_var, int training,
                                         float momentum, float eps, void* y, float* save_mean,
                                         float* save_rstd, float* scale_shift, float* part, void* st) {
  V2F_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && (C & 7) == 0, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(x && gamma && beta && y && save_mean && save_rstd && scale_shift && part, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(training || (run_mean && run_var), V2F_ERR_BAD_ARG);
  V2F_REQUIRE(aligned16(x) && aligned16(y), V2F_ERR_ALIGN);
  return stem_bn_pool(N, H, W, C, x, gamma, beta, run_mean, run_var, training, momentum, eps, y, save_mean, save_rstd,
                      scale_shift, part, 0, (cudaStream_t)st);
}

extern "C" int v2f_bn2d_relu_maxpool_fwd_parts(int N, int H, int W, int C, const void* x, const float* gamma,
                                               const float* beta, float* run_mean, float* run_var, float momentum,
                                               float eps, void* y, float* save_mean, float* save_rstd,
                                               float* scale_shift, const float* part, int part_blocks, void* st) {
  V2F_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && (C & 7) == 0 && part_blocks > 0, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(x && gamma && beta && y && save_mean && save_rstd && scale_shift && part, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(aligned16(x) && aligned16(y), V2F_ERR_ALIGN);
  return stem_bn_pool(N, H, W, C, x, gamma, beta, run_mean, run_var, 1, momentum, eps, y, save_mean, save_rstd,
                      scale_shift, const_cast<float*>(part), part_blocks, (cudaStream_t)st);
}

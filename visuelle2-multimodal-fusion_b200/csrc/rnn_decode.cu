// Fused recurrent-attention decoder of the CrossAttnRNN family, forward and BPTT.  sm_100a.
//
// Reference arithmetic: /root/reference/models/CrossAttnRNN210.py:191-225 (loop), :83-89
// (AdditiveAttention), :135-140,210-211 (decoder GRU), :141,212-225 (decoder_fc + teacher
// forcing); CrossAttnRNN21.py:183-206; CrossAttnRNNDemand.py:285-347,134-149.
//
// Per decode step the step-invariant tiles Himg/Vimg/Htr/Ptr of every row are streamed exactly
// once (forward) and once more (backward); their gradients are produced after the time loop
// in one pass that re-derives tanh from the saved per-step scalars, so each gradient tile is
// written once (SURVEY.md section 8d byte model).
#include "attn.cuh"
#include "gemm_dispatch.cuh"

namespace v2f {

constexpr int ATT_THREADS = 512;
constexpr int ATT_WARPS = ATT_THREADS / 32;
constexpr int MAXL = 128;     // max positions per attention
constexpr int MAXV = 8;       // max float4 per lane per row: E <= 1024

// sum NV values over the block; result valid in every thread.  red: [32][NV] floats.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; i++) red[warp * NV + i] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; i++) {
    float t = 0.f;
    for (int w = 0; w < nw; w++) t += red[w * NV + i];
    v[i] = t;
  }
}

// One CTA per (row, modality): energies -> softmax -> context.  Tiles streamed with 128-bit loads.
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int E = a.E;
  float* s_sh = sm;            // [E]
  float* w_sh = sm + E;        // [E]
  float* e_sh = sm + 2 * E;    // [MAXL]
  float* red = e_sh + MAXL;    // [ATT_WARPS][E]
  const int n = blockIdx.x, mod = blockIdx.y + a.mod_first, b = n / a.W;
  const int L = mod ? a.Lt : a.Li;
  const float* Hp = (mod ? a.Htr : a.Himg) + (long long)b * L * E;
  const float* Vp = (mod ? a.Ptr : a.Vimg) + (long long)b * L * E;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < E; i += ATT_THREADS) {
    s_sh[i] = a.S[(long long)n * a.ldS + mod * E + i];
    w_sh[i] = a.w_att[mod * E + i];
  }
  __syncthreads();
  const float beta = a.beta_att[mod];
  for (int j = warp; j < L; j += ATT_WARPS) {
    const float* hp = Hp + (long long)j * E;
    float acc = 0.f;
    for (int x = lane * 4; x < E; x += 128) {
      const float4 h = ld4_stream(hp + x);
      const float4 s = ld4(s_sh + x), w = ld4(w_sh + x);
      acc = fmaf(w.x, tanh_acc(h.x + s.x), acc);
      acc = fmaf(w.y, tanh_acc(h.y + s.y), acc);
      acc = fmaf(w.z, tanh_acc(h.z + s.z), acc);
      acc = fmaf(w.w, tanh_acc(h.w + s.w), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) e_sh[j] = acc + beta;
  }
  __syncthreads();
  if (warp == 0) {
    float m = -INFINITY;
    for (int j = lane; j < L; j += 32) m = fmaxf(m, e_sh[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float ex = expf(e_sh[j] - m);
      e_sh[j] = ex;
      sum += ex;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float* al = (mod ? a.alpha_tr : a.alpha_img) + (long long)n * L;
    for (int j = lane; j < L; j += 32) {
      const float p = e_sh[j] * inv;
      e_sh[j] = p;
      al[j] = p;
    }
  }
  __syncthreads();
  float4 cacc[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; i++) cacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = warp; j < L; j += ATT_WARPS) {
    const float p = e_sh[j];
    const float* vp = Vp + (long long)j * E;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      const int x = lane * 4 + i * 128;
      if (x < E) {
        const float4 v = ld4_stream(vp + x);
        cacc[i].x = fmaf(p, v.x, cacc[i].x);
        cacc[i].y = fmaf(p, v.y, cacc[i].y);
        cacc[i].z = fmaf(p, v.z, cacc[i].z);
        cacc[i].w = fmaf(p, v.w, cacc[i].w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; i++) {
    const int x = lane * 4 + i * 128;
    if (x < E) st4(red + warp * E + x, cacc[i]);
  }
  __syncthreads();
  for (int x = tid; x < E; x += ATT_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; w++) t += red[w * E + x];
    if (mod) t += a.b_tl[x];
    a.C[((long long)n * 2 + mod) * E + x] = t;
  }
}

__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_kernel(AttnBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int E = a.E;
  float* s_sh = sm;              // [E]
  float* w_sh = sm + E;          // [E]
  float* dc_sh = sm + 2 * E;     // [E]
  float* al_sh = sm + 3 * E;     // [MAXL]
  float* de_sh = al_sh + MAXL;   // [MAXL]
  float* red = de_sh + MAXL;     // [ATT_WARPS][E]
  const int n = blockIdx.x, mod = blockIdx.y + a.mod_first, b = n / a.W;
  const int L = mod ? a.Lt : a.Li;
  const float* Hp = (mod ? a.Htr : a.Himg) + (long long)b * L * E;
  const float* Vp = (mod ? a.Ptr : a.Vimg) + (long long)b * L * E;
  const float* al = (mod ? a.alpha_tr : a.alpha_img) + (long long)n * L;
  float* deg = (mod ? a.DE_tr : a.DE_img) + (long long)n * L;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < E; i += ATT_THREADS) {
    s_sh[i] = a.S[(long long)n * a.ldS + mod * E + i];
    w_sh[i] = a.w_att[mod * E + i];
    dc_sh[i] = a.DC[((long long)n * 2 + mod) * E + i];
  }
  for (int j = tid; j < L; j += ATT_THREADS) al_sh[j] = al[j];
  __syncthreads();
  // d alpha_j = dc . V_j
  for (int j = warp; j < L; j += ATT_WARPS) {
    const float* vp = Vp + (long long)j * E;
    float acc = 0.f;
    for (int x = lane * 4; x < E; x += 128) {
      const float4 v = ld4_stream(vp + x);
      const float4 d = ld4(dc_sh + x);
      acc = fmaf(v.x, d.x, acc);
      acc = fmaf(v.y, d.y, acc);
      acc = fmaf(v.z, d.z, acc);
      acc = fmaf(v.w, d.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) de_sh[j] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    float dot = 0.f;
    for (int j = lane; j < L; j += 32) dot = fmaf(al_sh[j], de_sh[j], dot);
    dot = warp_sum(dot);
    for (int j = lane; j < L; j += 32) {
      const float de = al_sh[j] * (de_sh[j] - dot);
      de_sh[j] = de;
      deg[j] = de;
    }
  }
  __syncthreads();
  // ds[a] = sum_j de_j w_a (1 - q^2),  dw[a] = sum_j de_j q,   q = tanh(H[j,a] + s_a)
  float4 sacc[MAXV], wacc[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; i++) {
    sacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    wacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int j = warp; j < L; j += ATT_WARPS) {
    const float de = de_sh[j];
    const float* hp = Hp + (long long)j * E;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      const int x = lane * 4 + i * 128;
      if (x < E) {
        const float4 h = ld4_stream(hp + x);
        const float4 s = ld4(s_sh + x);
        const float qx = tanh_acc(h.x + s.x), qy = tanh_acc(h.y + s.y), qz = tanh_acc(h.z + s.z),
                    qw = tanh_acc(h.w + s.w);
        sacc[i].x = fmaf(de, 1.f - qx * qx, sacc[i].x);
        sacc[i].y = fmaf(de, 1.f - qy * qy, sacc[i].y);
        sacc[i].z = fmaf(de, 1.f - qz * qz, sacc[i].z);
        sacc[i].w = fmaf(de, 1.f - qw * qw, sacc[i].w);
        wacc[i].x = fmaf(de, qx, wacc[i].x);
        wacc[i].y = fmaf(de, qy, wacc[i].y);
        wacc[i].z = fmaf(de, qz, wacc[i].z);
        wacc[i].w = fmaf(de, qw, wacc[i].w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; i++) {
    const int x = lane * 4 + i * 128;
    if (x < E) st4(red + warp * E + x, sacc[i]);
  }
  __syncthreads();
  for (int x = tid; x < E; x += ATT_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; w++) t += red[w * E + x];
    a.DS[(long long)n * a.ldS + mod * E + x] = t * w_sh[x];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MAXV; i++) {
    const int x = lane * 4 + i * 128;
    if (x < E) st4(red + warp * E + x, wacc[i]);
  }
  __syncthreads();
  for (int x = tid; x < E; x += ATT_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; w++) t += red[w * E + x];
    a.dw_acc[((long long)n * 3 + mod) * E + x] += t;
  }
}

// ------------------------------------------------------------------------------------------
// multimodal attention over the (up to) four modality rows of one decoder row.  Row-local.
struct MmArgs {
  int N, W, E, ldS, byproj, mod_mask;
  const float *Mst, *HMst;   // [B,2,E]
  const float *C, *HC;       // [N,2,E] this step
  const float* S;            // this step
  const float* w_att;        // [3,E]
  const float* beta_att;
  float* alpha_mm;           // [N,4] this step
  float* U;                  // [N,E] this step
};

__device__ __forceinline__ const float* mm_row(const float* st, const float* dyn, int b, int n,
                                               int k, int E) {
  // k: 0 date (static 0), 1 image ctx (dynamic 0), 2 attributes (static 1), 3 trend ctx (dynamic 1)
  return (k & 1) ? dyn + ((long long)n * 2 + (k >> 1)) * E : st + ((long long)b * 2 + (k >> 1)) * E;
}

__global__ void __launch_bounds__(256)
mm_fwd_kernel(MmArgs a) {
  __shared__ float red[32 * 4];
  __shared__ float al_sh[4];
  const int n = blockIdx.x, b = n / a.W, E = a.E, tid = threadIdx.x;
  const float* s = a.S + (long long)n * a.ldS + 2 * E;
  const float* w = a.w_att + 2 * E;
  const float* Mk[4];
  const float* HMk[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    Mk[k] = mm_row(a.Mst, a.C, b, n, k, E);
    HMk[k] = mm_row(a.HMst, a.HC, b, n, k, E);
  }
  float e[4] = {0.f, 0.f, 0.f, 0.f};
  for (int x = tid; x < E; x += blockDim.x) {
    const float sx = s[x], wx = w[x];
#pragma unroll
    for (int k = 0; k < 4; k++)
      if ((a.mod_mask >> k) & 1) e[k] = fmaf(wx, tanh_acc(HMk[k][x] + sx), e[k]);
  }
  block_sum<4>(e, red);
  if (tid == 0) {
    const float beta = a.beta_att[2];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; k++)
      if ((a.mod_mask >> k) & 1) { e[k] += beta; m = fmaxf(m, e[k]); }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      e[k] = ((a.mod_mask >> k) & 1) ? expf(e[k] - m) : 0.f;
      sum += e[k];
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      al_sh[k] = e[k] * inv;
      a.alpha_mm[(long long)n * 4 + k] = e[k] * inv;
    }
  }
  __syncthreads();
  for (int x = tid; x < E; x += blockDim.x) {
    float u = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++)
      if ((a.mod_mask >> k) & 1) {
        const float mv = Mk[k][x];
        u += mv + al_sh[k] * (a.byproj ? HMk[k][x] : mv);
      }
    a.U[(long long)n * E + x] = u;
  }
}

struct MmBwdArgs {
  int N, W, E, ldS, byproj, mod_mask;
  const float *Mst, *HMst, *C, *HC, *S, *w_att, *alpha_mm;
  const float* dU;     // [N,E]
  float* DS;           // [N,ldS] this step: writes cols 2E..3E
  float* DHC;          // [N,2,E] this step: d(We_mm c_img), d(We_mm c_tr)
  float* DC;           // [N,2,E] this step: direct part of d c_img, d c_tr
  float *dMst_acc, *dHMst_acc, *dw_acc;   // [N,2,E],[N,2,E],[N,3,E]
};

__global__ void __launch_bounds__(256)
mm_bwd_kernel(MmBwdArgs a) {
  __shared__ float red[32 * 4];
  const int n = blockIdx.x, b = n / a.W, E = a.E, tid = threadIdx.x;
  const float* s = a.S + (long long)n * a.ldS + 2 * E;
  const float* w = a.w_att + 2 * E;
  const float* du = a.dU + (long long)n * E;
  const float* Mk[4];
  const float* HMk[4];
  float al[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    Mk[k] = mm_row(a.Mst, a.C, b, n, k, E);
    HMk[k] = mm_row(a.HMst, a.HC, b, n, k, E);
    al[k] = a.alpha_mm[(long long)n * 4 + k];
  }
  float da[4] = {0.f, 0.f, 0.f, 0.f};
  for (int x = tid; x < E; x += blockDim.x) {
    const float d = du[x];
#pragma unroll
    for (int k = 0; k < 4; k++)
      if ((a.mod_mask >> k) & 1) da[k] = fmaf(d, a.byproj ? HMk[k][x] : Mk[k][x], da[k]);
  }
  block_sum<4>(da, red);
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < 4; k++) dot = fmaf(al[k], da[k], dot);
  float de[4];
#pragma unroll
  for (int k = 0; k < 4; k++) de[k] = al[k] * (da[k] - dot);
  for (int x = tid; x < E; x += blockDim.x) {
    const float sx = s[x], wx = w[x], d = du[x];
    float ds = 0.f, dw = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float dhm = 0.f, dm = 0.f;
      if ((a.mod_mask >> k) & 1) {
        const float q = tanh_acc(HMk[k][x] + sx);
        const float dpre = de[k] * wx * (1.f - q * q);
        ds += dpre;
        dw = fmaf(de[k], q, dw);
        dhm = dpre + (a.byproj ? al[k] * d : 0.f);
        dm = a.byproj ? d : d * (1.f + al[k]);
      }
      const long long dyn = ((long long)n * 2 + (k >> 1)) * E + x;
      if (k & 1) {
        a.DHC[dyn] = dhm;
        a.DC[dyn] = dm;
      } else {
        a.dHMst_acc[dyn] += dhm;
        a.dMst_acc[dyn] += dm;
      }
    }
    a.DS[(long long)n * a.ldS + 2 * E + x] = ds;
    a.dw_acc[((long long)n * 3 + 2) * E + x] += dw;
  }
}

// ------------------------------------------------------------------------------------------
// GRU gates + decoder_fc + teacher-forcing select.  Row-local.
struct GateArgs {
  int N, H, T, t, ldS, goff, forced;
  const unsigned* mask_dev;   // optional device-resident teacher-forcing mask (overrides ``forced``)
  const float* GI;      // [N,3H] = ctx W_ihc^T + b_ih
  const float* S;       // this step; gh at column goff (includes b_hh)
  const float* hprev;   // [N,H]
  const float* xin;     // [N]
  const float *w_x, *w_fc, *b_fc, *y;
  float* RZN;           // [N,3H] this step
  float* hnext;         // [N,H]
  float* yhat;          // [N,T]
  float* xnext;         // [N]
};

__global__ void __launch_bounds__(256)
gates_fwd_kernel(GateArgs a) {
  __shared__ float red[32];
  const int n = blockIdx.x, H = a.H, tid = threadIdx.x;
  const float* gi = a.GI + (long long)n * 3 * H;
  const float* gh = a.S + (long long)n * a.ldS + a.goff;
  const float x = a.xin[n];
  float part[1] = {0.f};
  for (int u = tid; u < H; u += blockDim.x) {
    const float r = sigmoid_full(gi[u] + x * a.w_x[u] + gh[u]);
    const float z = sigmoid_full(gi[H + u] + x * a.w_x[H + u] + gh[H + u]);
    const float c = tanh_full(gi[2 * H + u] + x * a.w_x[2 * H + u] + r * gh[2 * H + u]);
    const float hp = a.hprev[(long long)n * H + u];
    const float hn = (1.f - z) * c + z * hp;
    float* rzn = a.RZN + (long long)n * 3 * H;
    rzn[u] = r;
    rzn[H + u] = z;
    rzn[2 * H + u] = c;
    a.hnext[(long long)n * H + u] = hn;
    part[0] = fmaf(a.w_fc[u], hn, part[0]);
  }
  block_sum<1>(part, red);
  if (tid == 0) {
    const float yh = part[0] + a.b_fc[0];
    a.yhat[(long long)n * a.T + a.t] = yh;
    const int forced = a.mask_dev ? (int)((*a.mask_dev >> a.t) & 1u) : a.forced;
    a.xnext[n] = (forced && a.y) ? a.y[(long long)n * a.T + a.t] : yh;
  }
}

struct GateBwdArgs {
  int N, H, T, t, ldS, goff, forced;   // forced: x_{t+1} was y[:,t] (yhat_t did not feed forward)
  const unsigned* mask_dev;            // optional device-resident mask (overrides ``forced`` for t < T-1)
  const float *RZN, *S, *hprev, *w_x, *w_fc, *dY;
  float* dh;      // [N,H] in: dL/dh_{t+1}; out: direct part of dL/dh_t
  float* dxn;     // [N] in: dL/dx_{t+1}; out: dL/dx_t
  float* DGI;     // [N,3H] this step
  float* DS;      // [N,ldS] this step: writes cols goff..goff+3H
  float* DYH;     // [N] this step
};

__global__ void __launch_bounds__(256)
gates_bwd_kernel(GateBwdArgs a) {
  __shared__ float red[32];
  const int n = blockIdx.x, H = a.H, tid = threadIdx.x;
  const int forced = (a.mask_dev && a.t < a.T - 1) ? (int)((*a.mask_dev >> a.t) & 1u) : a.forced;
  const float dyh = a.dY[(long long)n * a.T + a.t] + (forced ? 0.f : a.dxn[n]);
  const float* rzn = a.RZN + (long long)n * 3 * H;
  const float* gh = a.S + (long long)n * a.ldS + a.goff;
  float part[1] = {0.f};
  for (int u = tid; u < H; u += blockDim.x) {
    const float r = rzn[u], z = rzn[H + u], c = rzn[2 * H + u];
    const float hp = a.hprev[(long long)n * H + u];
    const float dhp = a.dh[(long long)n * H + u] + dyh * a.w_fc[u];
    const float dc = dhp * (1.f - z);
    const float dz = dhp * (hp - c);
    const float dan = dc * (1.f - c * c);
    const float dar = dan * gh[2 * H + u] * r * (1.f - r);
    const float daz = dz * z * (1.f - z);
    float* dgi = a.DGI + (long long)n * 3 * H;
    dgi[u] = dar;
    dgi[H + u] = daz;
    dgi[2 * H + u] = dan;
    float* dgh = a.DS + (long long)n * a.ldS + a.goff;
    dgh[u] = dar;
    dgh[H + u] = daz;
    dgh[2 * H + u] = dan * r;
    a.dh[(long long)n * H + u] = dhp * z;
    part[0] += dar * a.w_x[u] + daz * a.w_x[H + u] + dan * a.w_x[2 * H + u];
  }
  block_sum<1>(part, red);
  if (tid == 0) {
    a.dxn[n] = part[0];
    a.DYH[n] = dyh;
  }
}

// variant 1 (CrossAttnRNN21): yhat = w_fc . ctx + b_fc  and its backward d ctx = dyhat * w_fc
__global__ void __launch_bounds__(256)
fc_fwd_kernel(int E, const float* __restrict__ ctx, const float* __restrict__ w_fc,
              const float* __restrict__ b_fc, float* __restrict__ yhat) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  float part[1] = {0.f};
  for (int x = threadIdx.x; x < E; x += blockDim.x) part[0] = fmaf(w_fc[x], ctx[(long long)n * E + x], part[0]);
  block_sum<1>(part, red);
  if (threadIdx.x == 0) yhat[n] = part[0] + b_fc[0];
}
__global__ void fc_bwd_kernel(int N, int E, const float* __restrict__ dY, const float* __restrict__ w_fc,
                              float* __restrict__ dctx, float* __restrict__ DYH) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)N * E) {
    const int n = (int)(i / E), x = (int)(i % E);
    dctx[i] = dY[n] * w_fc[x];
    if (x == 0) DYH[n] = dY[n];
  }
}

// ------------------------------------------------------------------------------------------
// After the time loop: gradients of the step-invariant tiles, each written exactly once.
//   dV[b,j,:] = sum_{rows n of b, t} alpha[t,n,j] dC[t,n,mod,:]
//   dH[b,j,:] = sum_{n,t} de[t,n,j] w (1 - tanh^2(H[b,j,:] + s[t,n,:]))     (+ dV if byproj)
constexpr int TG_J = 13;   // positions per CTA
struct TileGradArgs {
  int N, W, E, L, T, ldS, mod, byproj;
  const float* H;       // [B,L,E]
  const float* S_all;   // [T,N,ldS]
  const float* w;       // [E]
  const float* alpha;   // [T,N,L]
  const float* DE;      // [T,N,L]
  const float* DC_all;  // [T,N,2,E]
  float *dH, *dV;       // [B,L,E]; dV may be NULL when byproj
};

// Thread = up to CPTG columns (x, x + 256, ...): the 13 tile values of each stay in registers for all T*W (step, row)
// pairs of the item, so H is read once and dH / dV are written once.  Its floor is the MUFU pipe, not HBM: B*L*E*T tanh
// evaluations at 16 per clock and SM (tools/probes/mufu_probe.cu: tanh.approx = ex2.approx = 31.2 G/s/SM) = 21.6 us for
// the two launches of a step (measured: 68 us); in tensor-core mode the single-instruction tanh.approx (the forward's choice there) halves
// the MUFU work; the exact mode keeps the two-instruction form.
constexpr int TG_CHUNK = 32;   // (step, row) pairs whose attention scalars are staged at once
template <bool APPROX, int CPTG>
__global__ void __launch_bounds__(256)
tilegrad_kernel(TileGradArgs a) {
  // alpha / de of the CTA's 13 positions for a whole chunk of (step, row) pairs are staged up front, so the loop over
  // the pairs has no barrier and the compiler can keep the next pairs' s / dC loads in flight under the tanh work
  // (one barrier + two dependent global loads per pair left the MUFU pipe 25 % busy)
  __shared__ float al_sh[TG_CHUNK][TG_J], de_sh[TG_CHUNK][TG_J];
  const int b = blockIdx.x, j0 = blockIdx.y * TG_J, E = a.E, L = a.L;
  const int nj = min(TG_J, L - j0);
  const int iters = a.T * a.W;
  for (int x0 = 0; x0 < E; x0 += 256 * CPTG) {
    float h[CPTG][TG_J], accH[CPTG][TG_J], accV[CPTG][TG_J], wx[CPTG];
    bool act[CPTG];
#pragma unroll
    for (int i = 0; i < CPTG; i++) {
      const int x = x0 + threadIdx.x + 256 * i;
      act[i] = x < E;
      wx[i] = act[i] ? a.w[x] : 0.f;
#pragma unroll
      for (int jj = 0; jj < TG_J; jj++) {
        h[i][jj] = (act[i] && jj < nj) ? a.H[((long long)b * L + j0 + jj) * E + x] : 0.f;
        accH[i][jj] = 0.f;
        accV[i][jj] = 0.f;
      }
    }
    for (int c0 = 0; c0 < iters; c0 += TG_CHUNK) {
      const int cn = min(TG_CHUNK, iters - c0);
      __syncthreads();                     // the previous chunk / x0 pass is done with the staging buffers
      for (int k = threadIdx.x; k < cn * TG_J; k += 256) {
        const int ii = k / TG_J, jj = k - ii * TG_J;
        const int it = c0 + ii, t = it / a.W, wi = it - t * a.W;
        const long long tn = (long long)t * a.N + b * a.W + wi;
        al_sh[ii][jj] = jj < nj ? a.alpha[tn * L + j0 + jj] : 0.f;
        de_sh[ii][jj] = jj < nj ? a.DE[tn * L + j0 + jj] : 0.f;
      }
      __syncthreads();
      // (loading pair ii + 1 into a second register set before computing pair ii was measured SLOWER: 41.9 vs 34.2 us)
#pragma unroll 2
      for (int ii = 0; ii < cn; ii++) {
        const int it = c0 + ii, t = it / a.W, wi = it - t * a.W;
        const long long tn = (long long)t * a.N + b * a.W + wi;
        float s[CPTG], dc[CPTG];
#pragma unroll
        for (int i = 0; i < CPTG; i++) {
          const int x = x0 + threadIdx.x + 256 * i;
          s[i] = act[i] ? __ldg(a.S_all + tn * a.ldS + a.mod * E + x) : 0.f;
          dc[i] = act[i] ? __ldg(a.DC_all + (tn * 2 + a.mod) * E + x) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < CPTG; i++)
#pragma unroll
          for (int jj = 0; jj < TG_J; jj++) {
            const float q = tanh_fast<APPROX>(h[i][jj] + s[i]);
            accH[i][jj] = fmaf(de_sh[ii][jj], fmaf(-q, q, 1.f), accH[i][jj]);
            accV[i][jj] = fmaf(al_sh[ii][jj], dc[i], accV[i][jj]);
          }
      }
    }
#pragma unroll
    for (int i = 0; i < CPTG; i++) {
      const int x = x0 + threadIdx.x + 256 * i;
      if (!act[i]) continue;
#pragma unroll
      for (int jj = 0; jj < TG_J; jj++)
        if (jj < nj) {
          const long long o = ((long long)b * L + j0 + jj) * E + x;
          if (a.byproj) {
            a.dH[o] = accH[i][jj] * wx[i] + accV[i][jj];
          } else {
            a.dH[o] = accH[i][jj] * wx[i];
            a.dV[o] = accV[i][jj];
          }
        }
    }
  }
}

// out[b,:] = sum_{w<W} acc[b*W+w,:]   (rows of width D)
__global__ void window_reduce_kernel(int B, int W, int D, const float* __restrict__ acc,
                                     float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)B * D) {
    const int b = (int)(i / D), x = (int)(i % D);
    float t = 0.f;
    for (int w = 0; w < W; w++) t += acc[((long long)(b * W + w)) * D + x];
    out[i] = t;
  }
}

__global__ void copy_kernel(long long n, const float* __restrict__ src, float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

static int check_params(const v2f_decode_params* p) {
  V2F_REQUIRE(p, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(p->N > 0 && p->B > 0 && p->W > 0 && p->N == p->B * p->W, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(p->E > 0 && (p->E & 3) == 0 && p->E <= 128 * MAXV, V2F_ERR_UNSUPPORTED);
  V2F_REQUIRE(p->H > 0 && p->T > 0 && p->T <= 32, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(p->Li > 0 && p->Li <= MAXL && p->Lt > 0 && p->Lt <= MAXL, V2F_ERR_UNSUPPORTED);
  V2F_REQUIRE(p->variant >= 0 && p->variant <= 2, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(p->variant != 1 || p->T == 1, V2F_ERR_BAD_ARG);
  V2F_REQUIRE((p->mod_mask & 1) == 1, V2F_ERR_BAD_ARG);
  const void* ptrs[] = {p->Himg, p->Vimg, p->Htr, p->Ptr};
  for (const void* q : ptrs) V2F_REQUIRE(q && aligned16(q), V2F_ERR_ALIGN);
  return V2F_OK;
}

static size_t attn_smem(int E, bool bwd) {
  return sizeof(float) * ((bwd ? 3 : 2) * (size_t)E + (bwd ? 2 : 1) * MAXL + (size_t)ATT_WARPS * E);
}

}  // namespace v2f

using namespace v2f;

namespace v2f {
int decode_persist_fwd(const v2f_decode_params* p, cudaStream_t s);   // decode_persist.cu
int decode_team_fwd(const v2f_decode_params* p, cudaStream_t s);      // decode_team.cu
int decode_team_bwd(const v2f_decode_params* p, float* bws, long long bws_floats, cudaStream_t s);
}

#define NT(M, N, K, A, lda, B, ldb, C, ldc, bias, beta) V2F_TRY(gemm_nt(gx, M, N, K, A, lda, B, ldb, C, ldc, bias, beta))
#define NN(M, N, K, A, lda, B, ldb, BT, ldbt, C, ldc, beta) \
  V2F_TRY(gemm_nn(gx, M, N, K, A, lda, B, ldb, BT, ldbt, C, ldc, beta))
#define TN(M, N, K, A, lda, B, ldb, C, ldc) V2F_TRY(gemm_tn(gx, M, N, K, A, lda, B, ldb, C, ldc))

extern "C" int v2f_decode_fwd(const v2f_decode_params* p, void* st) {
  V2F_TRY(check_params(p));
  cudaStream_t s = (cudaStream_t)st;
  const int N = p->N, E = p->E, H = p->H, T = p->T;
  const bool gru = p->variant != 1;
  const int G = gru ? 3 * H : 0, ldS = 3 * E + G;
  const bool use_img = (p->mod_mask >> 1) & 1, use_tr = (p->mod_mask >> 3) & 1;
  const GemmCtx gx{p->precision, nullptr, 0, st};
  const size_t smem = attn_smem(E, false);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // h_all[0] = h0, xin[0] = x0
  copy_kernel<<<(unsigned)(((long long)N * H + 255) / 256), 256, 0, s>>>((long long)N * H, p->h0, p->h_all);
  V2F_CHECK_LAUNCH();
  if (gru) {
    copy_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, p->x0, p->xin);
    V2F_CHECK_LAUNCH();
  }
  {   // the whole loop as one persistent cooperative launch when the configuration allows: the row-team tcgen05
      // kernel (decode_team.cu) at the default dims in tensor-core mode, else the column-split kernel (decode_persist.cu)
    int rc = gru ? decode_team_fwd(p, s) : V2F_ERR_UNSUPPORTED;
    if (rc != V2F_ERR_UNSUPPORTED) return rc;
    rc = decode_persist_fwd(p, s);
    if (rc != V2F_ERR_UNSUPPORTED) return rc;
  }
  for (int t = 0; t < T; t++) {
    const float* h = p->h_all + (long long)t * N * H;
    float* S = p->S_all + (long long)t * N * ldS;
    float* C = p->C + (long long)t * N * 2 * E;
    float* HC = p->HC + (long long)t * N * 2 * E;
    float* U = p->U + (long long)t * N * E;
    float* CTX = p->CTX + (long long)t * N * E;
    // S = h Wcat^T + bcat  (s_img | s_tr | s_mm | gh)
    NT(N, ldS, H, h, H, p->Wcat, H, S, ldS, p->bcat, 0.f);
    if (use_img || use_tr) {
      AttnArgs a{N, p->W, E, p->Li, p->Lt, ldS, p->Himg, p->Vimg, p->Htr, p->Ptr, S, p->w_att,
                 p->beta_att, p->b_tl, C, p->alpha_img + (long long)t * N * p->Li,
                 p->alpha_tr + (long long)t * N * p->Lt, use_img ? 0 : 1, p->precision != 0};
      if (p->attn_ws && attn_stream_supported(E)) {
        V2F_TRY(attn_stream_fwd(a, use_img, use_tr, p->attn_ws, s));
      } else {
        prof_begin(V2F_K_ATTN_FWD, s);
        attn_fwd_kernel<<<dim3(N, (use_img ? 1 : 0) + (use_tr ? 1 : 0)), ATT_THREADS, smem, s>>>(a);
        prof_end(V2F_K_ATTN_FWD, s);
        V2F_CHECK_LAUNCH();
      }
      // HC = C We_mm^T  ([2N,E] view)
      NT(2 * N, E, E, C, E, p->We_mm, E, HC, E, nullptr, 0.f);
    }
    MmArgs m{N, p->W, E, ldS, p->variant == 2, p->mod_mask, p->Mst, p->HMst, C, HC, S, p->w_att,
             p->beta_att, p->alpha_mm + (long long)t * N * 4, U};
    mm_fwd_kernel<<<N, 256, 0, s>>>(m);
    V2F_CHECK_LAUNCH();
    NT(N, E, E, U, E, p->W_me, E, CTX, E, p->b_me, 0.f);
    if (gru) {
      NT(N, 3 * H, E, CTX, E, p->W_ihc, E, p->GI, 3 * H, p->b_ih, 0.f);
      GateArgs g{N, H, T, t, ldS, 3 * E, (int)((p->tf_mask >> t) & 1u), p->y ? p->tf_mask_dev : nullptr, p->GI, S, h,
                 p->xin + (long long)t * N, p->w_x, p->w_fc, p->b_fc, p->y,
                 p->RZN + (long long)t * N * 3 * H, p->h_all + (long long)(t + 1) * N * H, p->yhat,
                 p->xin + (long long)(t + 1) * N};
      gates_fwd_kernel<<<N, 256, 0, s>>>(g);
      V2F_CHECK_LAUNCH();
    } else {
      fc_fwd_kernel<<<N, 256, 0, s>>>(E, CTX, p->w_fc, p->b_fc, p->yhat);
      V2F_CHECK_LAUNCH();
    }
  }
  return V2F_OK;
}

extern "C" int v2f_decode_bwd(const v2f_decode_params* p, void* st) {
  V2F_TRY(check_params(p));
  cudaStream_t s = (cudaStream_t)st;
  const int N = p->N, E = p->E, H = p->H, T = p->T, B = p->B;
  const bool gru = p->variant != 1;
  const int G = gru ? 3 * H : 0, ldS = 3 * E + G;
  const bool use_img = (p->mod_mask >> 1) & 1, use_tr = (p->mod_mask >> 3) & 1;
  const int byproj = p->variant == 2;
  const GemmCtx gx{p->precision, p->ws, p->ws_floats, st};
  const bool tc = p->precision != 0 && p->WcatT && p->W_meT && p->We_mmT && (!gru || p->W_ihcT);
  if (tc) {  // transposed weight copies for the dx-type products (K-major operands for tcgen05)
    V2F_TRY(v2f_transpose(ldS, H, p->Wcat, H, 1, p->WcatT, ldS, 1, st));
    V2F_TRY(v2f_transpose(E, E, p->W_me, E, 1, p->W_meT, E, 1, st));
    V2F_TRY(v2f_transpose(E, E, p->We_mm, E, 1, p->We_mmT, E, 1, st));
    if (gru) V2F_TRY(v2f_transpose(3 * H, E, p->W_ihc, E, 1, p->W_ihcT, 3 * H, 1, st));
  }
  const float* WcatT = tc ? p->WcatT : nullptr;
  const float* W_meT = tc ? p->W_meT : nullptr;
  const float* We_mmT = tc ? p->We_mmT : nullptr;
  const float* W_ihcT = tc ? p->W_ihcT : nullptr;
  const size_t smem = attn_smem(E, true);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // the whole BPTT loop as one persistent cooperative launch when the forward ran on the row-team kernel
  // (decode_team.cu); dCTX = DGI W_ihc for all T*N rows is then one product after the loop
  int team_rc = (gru && tc) ? decode_team_bwd(p, p->ws, p->ws_floats, s) : V2F_ERR_UNSUPPORTED;
  if (team_rc != V2F_OK && team_rc != V2F_ERR_UNSUPPORTED) return team_rc;
  if (team_rc == V2F_OK)
    NN(T * N, E, 3 * H, p->DGI, 3 * H, p->W_ihc, E, W_ihcT, 3 * H, p->DCTX, E, 0.f);
  for (int t = T - 1; t >= 0 && team_rc != V2F_OK; t--) {
    const float* h = p->h_all + (long long)t * N * H;
    const float* S = p->S_all + (long long)t * N * ldS;
    float* DS = p->DScat + (long long)t * N * ldS;
    float* DCTX = p->DCTX + (long long)t * N * E;
    float* DHC = p->DHC + (long long)t * N * 2 * E;
    float* DC = p->DC + (long long)t * N * 2 * E;
    if (gru) {
      // forced flag of THIS step's output: bit t says x_{t+1} = y[:,t]; the last step feeds nothing
      const int forced = (t == T - 1) ? 1 : (int)((p->tf_mask >> t) & 1u);
      GateBwdArgs g{N, H, T, t, ldS, 3 * E, forced, p->y ? p->tf_mask_dev : nullptr,
                    p->RZN + (long long)t * N * 3 * H, S, h, p->w_x,
                    p->w_fc, p->dY, p->dh, p->dxn, p->DGI + (long long)t * N * 3 * H, DS,
                    p->DYH + (long long)t * N};
      gates_bwd_kernel<<<N, 256, 0, s>>>(g);
      V2F_CHECK_LAUNCH();
      // dCTX = DGI W_ihc
      NN(N, E, 3 * H, p->DGI + (long long)t * N * 3 * H, 3 * H, p->W_ihc, E, W_ihcT, 3 * H, DCTX, E, 0.f);
    } else {
      fc_bwd_kernel<<<(unsigned)(((long long)N * E + 255) / 256), 256, 0, s>>>(N, E, p->dY, p->w_fc, DCTX, p->DYH);
      V2F_CHECK_LAUNCH();
    }
    // dU = dCTX W_me
    NN(N, E, E, DCTX, E, p->W_me, E, W_meT, E, p->dU, E, 0.f);
    MmBwdArgs m{N, p->W, E, ldS, byproj, p->mod_mask, p->Mst, p->HMst, p->C + (long long)t * N * 2 * E,
                p->HC + (long long)t * N * 2 * E, S, p->w_att, p->alpha_mm + (long long)t * N * 4,
                p->dU, DS, DHC, DC, p->dMst_acc, p->dHMst_acc, p->dw_acc};
    mm_bwd_kernel<<<N, 256, 0, s>>>(m);
    V2F_CHECK_LAUNCH();
    if (use_img || use_tr) {
      // dC += DHC We_mm
      NN(2 * N, E, E, DHC, E, p->We_mm, E, We_mmT, E, DC, E, 1.f);
      AttnBwdArgs a{N, p->W, E, p->Li, p->Lt, ldS, p->Himg, p->Vimg, p->Htr, p->Ptr, S, p->w_att, DC,
                    p->alpha_img + (long long)t * N * p->Li, p->alpha_tr + (long long)t * N * p->Lt,
                    p->DE_img + (long long)t * N * p->Li, p->DE_tr + (long long)t * N * p->Lt, DS,
                    p->dw_acc, use_img ? 0 : 1};
      if (p->attn_ws && attn_stream_supported(E)) {
        V2F_TRY(attn_stream_bwd(a, p->C + (long long)t * N * 2 * E, p->b_tl, use_img, use_tr, p->precision != 0,
                                p->attn_ws, s));
      } else {
        prof_begin(V2F_K_ATTN_BWD, s);
        attn_bwd_kernel<<<dim3(N, (use_img ? 1 : 0) + (use_tr ? 1 : 0)), ATT_THREADS, smem, s>>>(a);
        prof_end(V2F_K_ATTN_BWD, s);
        V2F_CHECK_LAUNCH();
      }
    }
    // dh (+)= DS Wcat      (gru: dh holds the direct z-path part; else dh is overwritten)
    NN(N, H, ldS, DS, ldS, p->Wcat, H, WcatT, ldS, p->dh, H, gru ? 1.f : 0.f);
  }
  // ---- parameter gradients: one GEMM each over all T*N rows
  const int TNr = T * N;
  TN(ldS, H, TNr, p->DScat, ldS, p->h_all, H, p->dWcat, H);
  V2F_TRY(v2f_colsum_f32(TNr, ldS, p->DScat, ldS, p->dbcat, 0.f, st));
  TN(E, E, TNr, p->DCTX, E, p->U, E, p->dW_me, E);
  V2F_TRY(v2f_colsum_f32(TNr, E, p->DCTX, E, p->db_me, 0.f, st));
  if (use_img || use_tr) {
    TN(E, E, 2 * TNr, p->DHC, E, p->C, E, p->dWe_mm, E);
  } else {
    cudaMemsetAsync(p->dWe_mm, 0, sizeof(float) * E * E, s);
  }
  if (gru) {
    TN(3 * H, E, TNr, p->DGI, 3 * H, p->CTX, E, p->dW_ihc, E);
    V2F_TRY(v2f_colsum_f32(TNr, 3 * H, p->DGI, 3 * H, p->db_ih, 0.f, st));
    TN(3 * H, 1, TNr, p->DGI, 3 * H, p->xin, 1, p->dw_x, 1);
    TN(H, 1, TNr, p->h_all + (long long)N * H, H, p->DYH, 1, p->dw_fc, 1);
  } else {
    TN(E, 1, TNr, p->CTX, E, p->DYH, 1, p->dw_fc, 1);
  }
  V2F_TRY(v2f_colsum_f32(TNr, 1, p->DYH, 1, p->db_fc, 0.f, st));
  V2F_TRY(v2f_colsum_f32(N, 3 * E, p->dw_acc, 3 * E, p->dw_att, 0.f, st));
  if (use_tr) {
    // db_tl = sum_{t,n} dC[t,n,1,:]
    V2F_TRY(v2f_colsum_f32(TNr, E, p->DC + E, 2 * E, p->db_tl, 0.f, st));
  } else {
    cudaMemsetAsync(p->db_tl, 0, sizeof(float) * E, s);
  }
  // ---- tile gradients, written once
  for (int mod = 0; mod < 2; mod++) {
    if (!(mod ? use_tr : use_img)) continue;
    const int L = mod ? p->Lt : p->Li;
    TileGradArgs tg{N, p->W, E, L, T, ldS, mod, byproj, mod ? p->Htr : p->Himg, p->S_all,
                    p->w_att + mod * E, mod ? p->alpha_tr : p->alpha_img, mod ? p->DE_tr : p->DE_img,
                    p->DC, mod ? p->dHtr : p->dHimg, mod ? p->dPtr : (byproj ? nullptr : p->dVimg)};
    // trend context always comes from Ptr (never from Htr), so byproj only affects the image tile
    if (mod) tg.byproj = 0;
    prof_begin(V2F_K_TILEGRAD, s);
    {
      const dim3 grid(B, (L + TG_J - 1) / TG_J);
      if (p->precision != 0) {
        if (E > 256) tilegrad_kernel<true, 2><<<grid, 256, 0, s>>>(tg);
        else tilegrad_kernel<true, 1><<<grid, 256, 0, s>>>(tg);
      } else {
        if (E > 256) tilegrad_kernel<false, 2><<<grid, 256, 0, s>>>(tg);
        else tilegrad_kernel<false, 1><<<grid, 256, 0, s>>>(tg);
      }
    }
    prof_end(V2F_K_TILEGRAD, s);
    V2F_CHECK_LAUNCH();
  }
  const unsigned g2 = (unsigned)(((long long)B * 2 * E + 255) / 256);
  window_reduce_kernel<<<g2, 256, 0, s>>>(B, p->W, 2 * E, p->dMst_acc, p->dMst);
  V2F_CHECK_LAUNCH();
  window_reduce_kernel<<<g2, 256, 0, s>>>(B, p->W, 2 * E, p->dHMst_acc, p->dHMst);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

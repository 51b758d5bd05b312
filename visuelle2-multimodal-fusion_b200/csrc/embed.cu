// Static embedders: TemporalFeatureEncoder + AttributeEncoder, forward and backward.  sm_100a.
//
// Reference arithmetic: /root/reference/models/CrossAttnRNN210.py:26-56
//   date[b,:]  = sum_k drop_k( temporal[b,k] * W_k[:,0] + b_k )            k = day,week,month,year
//   attrs[b,:] = sum_k drop_k( table_k[idx_k[b], :] )                      k = cat,col,fab,store
// (the Demand copy, CrossAttnRNNDemand.py:47-68, sends all four features through day_embedding;
// the caller passes the day weights four times and folds the four gradient rows).
// drop: optional keep-mask [B,8,E], already scaled by 1/(1-p).  Backward is deterministic: one
// thread per output element loops over the batch, no atomics.
#include "common.cuh"

long long g_v2f_launches = 0;
extern "C" int v2f_version(void) { return 4; }   // 4: v2f_decode_params gained team_ws / team_ws_floats
extern "C" long long v2f_launch_count(void) { return g_v2f_launches; }

namespace v2f {

struct Tables4 { const float* t[4]; };
struct DTables4 { float* t[4]; int rows[4]; };

__global__ void embed_fwd_kernel(int B, int E, const float* __restrict__ temporal,
                                 const float* __restrict__ Wt, const float* __restrict__ bt,
                                 Tables4 tb, const long long* __restrict__ idx,
                                 const float* __restrict__ drop, float* __restrict__ out) {
  const int b = blockIdx.x;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float d = 0.f, at = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float v = fmaf(temporal[b * 4 + k], Wt[k * E + e], bt[k * E + e]);
      float w = tb.t[k][idx[(long long)k * B + b] * E + e];
      if (drop) {
        v *= drop[((long long)b * 8 + k) * E + e];
        w *= drop[((long long)b * 8 + 4 + k) * E + e];
      }
      d += v;
      at += w;
    }
    out[((long long)b * 2) * E + e] = d;
    out[((long long)b * 2 + 1) * E + e] = at;
  }
}

// blockIdx.x < 4: temporal linear k; else one embedding-table row
__global__ void embed_bwd_kernel(int B, int E, const float* __restrict__ temporal,
                                 const long long* __restrict__ idx, const float* __restrict__ drop,
                                 const float* __restrict__ dout, float* __restrict__ dWt,
                                 float* __restrict__ dbt, DTables4 dt) {
  int blk = blockIdx.x;
  if (blk < 4) {
    const int k = blk;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      float gw = 0.f, gb = 0.f;
      for (int b = 0; b < B; b++) {
        float g = dout[((long long)b * 2) * E + e];
        if (drop) g *= drop[((long long)b * 8 + k) * E + e];
        gw = fmaf(g, temporal[b * 4 + k], gw);
        gb += g;
      }
      dWt[k * E + e] = gw;
      dbt[k * E + e] = gb;
    }
    return;
  }
  blk -= 4;
  int k = 0;
  while (k < 3 && blk >= dt.rows[k]) { blk -= dt.rows[k]; k++; }
  const int r = blk;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float g = 0.f;
    for (int b = 0; b < B; b++) {
      if (idx[(long long)k * B + b] == r) {
        float v = dout[((long long)b * 2 + 1) * E + e];
        if (drop) v *= drop[((long long)b * 8 + 4 + k) * E + e];
        g += v;
      }
    }
    dt.t[k][(long long)r * E + e] = g;
  }
}

__global__ void mul_kernel(long long n, const float* __restrict__ x, const float* __restrict__ m,
                           float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[i] * m[i];
}

}  // namespace v2f

using namespace v2f;

// out = x * m (dropout keep-mask application; mask already scaled by 1/(1-p))
extern "C" int v2f_mul_f32(long long n, const float* x, const float* m, float* out, void* st) {
  V2F_REQUIRE(n >= 0 && x && m && out, V2F_ERR_BAD_ARG);
  if (n == 0) return V2F_OK;
  mul_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)st>>>(n, x, m, out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_embed_fwd(int B, int E, const float* temporal, const float* Wt, const float* bt,
                             const float* const* tables, const long long* idx, const float* drop,
                             float* out, void* st) {
  V2F_REQUIRE(B > 0 && E > 0 && temporal && Wt && bt && tables && idx && out, V2F_ERR_BAD_ARG);
  Tables4 tb;
  for (int k = 0; k < 4; k++) {
    V2F_REQUIRE(tables[k], V2F_ERR_BAD_ARG);
    tb.t[k] = tables[k];
  }
  embed_fwd_kernel<<<B, 256, 0, (cudaStream_t)st>>>(B, E, temporal, Wt, bt, tb, idx, drop, out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

extern "C" int v2f_embed_bwd(int B, int E, const float* temporal, const long long* idx,
                             const float* drop, const float* dout, const int* table_rows,
                             float* dWt, float* dbt, float* const* dtables, void* st) {
  V2F_REQUIRE(B > 0 && E > 0 && temporal && idx && dout && table_rows && dWt && dbt && dtables, V2F_ERR_BAD_ARG);
  DTables4 dt;
  int total = 4;
  for (int k = 0; k < 4; k++) {
    V2F_REQUIRE(dtables[k] && table_rows[k] > 0, V2F_ERR_BAD_ARG);
    dt.t[k] = dtables[k];
    dt.rows[k] = table_rows[k];
    total += table_rows[k];
  }
  embed_bwd_kernel<<<total, 256, 0, (cudaStream_t)st>>>(B, E, temporal, idx, drop, dout, dWt, dbt, dt);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

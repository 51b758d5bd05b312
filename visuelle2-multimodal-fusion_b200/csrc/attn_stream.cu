// Streaming implementation of the per-step additive cross-attention (the fused
// recurrent-attention kernel of the decode loop).  sm_100a.
//
// Reference arithmetic: AdditiveAttention, /root/reference/models/CrossAttnRNN210.py:83-89, applied
// to the image map (:192-193) and the trend sequence (:195-196, trend_linear re-associated into
// the P tiles), CrossAttnRNNDemand.py:134-149 for the alpha*h_j variant.
//
// Design (B200): the step-invariant tiles are read exactly once per step, straight from HBM (80 MB
// per step at B=128 does not survive in L2 between steps).  One persistent CTA per SM; the 8-position
// chunks of all (row, modality) segments are split evenly over the grid, so all 148 SMs stream the
// same number of bytes.  Warp 8 is the producer: one lane issues 1-D bulk copies (TMA engine,
// cp.async.bulk -> UBLKCP) of the H and V rows of a chunk into a shared-memory ring, completion
// counted on mbarriers; warps 0..7 consume: warp-per-position energies (w . tanh(H_j + s), float4
// LDS, warp-shuffle reduce), online softmax (running max / sum, flash style), thread-per-column
// context accumulation.  A segment cut by a CTA boundary leaves partial (max, sum, context) triples
// that attn_combine_kernel merges; it also turns the raw energies into the softmax weights that the
// backward pass and the Demand model's attention maps need.
#include "async.cuh"
#include "attn.cuh"

namespace v2f {

constexpr int ST_CH = 8;             // positions per chunk
constexpr int ST_CONS = 256;         // consumer threads (8 warps)
constexpr int ST_THREADS = ST_CONS + 32;

struct StreamGeom {
  int cpi, cpt, cpr;     // chunks per row: image, trend, total
  long long total;       // N * cpr
};
__host__ __device__ inline StreamGeom stream_geom(int N, int Li, int Lt, bool use_img, bool use_tr) {
  StreamGeom g;
  g.cpi = use_img ? (Li + ST_CH - 1) / ST_CH : 0;
  g.cpt = use_tr ? (Lt + ST_CH - 1) / ST_CH : 0;
  g.cpr = g.cpi + g.cpt;
  g.total = (long long)N * g.cpr;
  return g;
}

struct StreamArgs {
  AttnArgs a;
  int use_img, use_tr;
  float *PM, *PL, *PC;   // partials: [N,cpr], [N,cpr] (PL zeroed before launch), [N,cpr,E]
};

// ring depth: 2 CTAs per SM (16 consumer warps hide the MUFU / LDS latency) with 3 stages each
// for E <= 512; wider rows fall back to 1 CTA per SM
__host__ __device__ constexpr int stream_stages(int E) { return E <= 512 ? 3 : (E <= 768 ? 4 : 3); }
__host__ __device__ constexpr int stream_ctas(int E) { return E <= 512 ? 2 : 1; }

template <int CPT, bool APPROX>
__global__ void __launch_bounds__(ST_THREADS, stream_ctas(256 * CPT))
attn_stream_fwd_kernel(StreamArgs sa) {
  constexpr int E = 256 * CPT, KV = 2 * CPT;
  constexpr int STG = stream_stages(E);
  constexpr int TILE = ST_CH * E;                       // floats per operand per stage
  extern __shared__ uint8_t raw[];
  float* ring = reinterpret_cast<float*>(smem_align(raw, 128));
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)STG * 2 * TILE);
  uint64_t* empty = full + STG;
  float* e_sh = reinterpret_cast<float*>(empty + STG);  // [2][ST_CH]

  const AttnArgs& a = sa.a;
  const StreamGeom gm = stream_geom(a.N, a.Li, a.Lt, sa.use_img, sa.use_tr);
  const long long g_lo = gm.total * blockIdx.x / gridDim.x, g_hi = gm.total * (blockIdx.x + 1) / gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

  if (tid == 0) {
    for (int s = 0; s < STG; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ST_CONS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == ST_CONS / 32) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      int it = 0;
      for (long long g = g_lo; g < g_hi; g++, it++) {
        const int s = it % STG, r = it / STG;
        const int n = (int)(g / gm.cpr), c = (int)(g - (long long)n * gm.cpr);
        const int mod = c >= gm.cpi, cc = mod ? c - gm.cpi : c;
        const int L = mod ? a.Lt : a.Li, j0 = cc * ST_CH, nj = min(ST_CH, L - j0);
        const long long off = ((long long)(n / a.W) * L + j0) * E;
        const uint32_t bytes = (uint32_t)nj * E * 4u;
        mbar_wait(&empty[s], (r & 1) ^ 1);
        mbar_expect_tx(&full[s], 2u * bytes);
        bulk_g2s(ring + (size_t)s * 2 * TILE, (mod ? a.Htr : a.Himg) + off, bytes, &full[s]);
        bulk_g2s(ring + (size_t)s * 2 * TILE + TILE, (mod ? a.Ptr : a.Vimg) + off, bytes, &full[s]);
      }
    }
    return;
  }
  // -------------------------------------------------------------------- consumers
  float4 sreg[KV], wreg[KV];
  float cacc[CPT];
  float m_run = -INFINITY, l_run = 0.f, beta = 0.f;
  int cur_n = -1, cur_mod = 0, slot0 = 0;
#pragma unroll
  for (int i = 0; i < CPT; i++) cacc[i] = 0.f;

  auto flush = [&]() {
    const long long slot = (long long)cur_n * gm.cpr + (cur_mod ? gm.cpi : 0) + slot0;
    if (tid == 0) {
      sa.PM[slot] = m_run;
      sa.PL[slot] = l_run;
    }
#pragma unroll
    for (int i = 0; i < CPT; i++) sa.PC[slot * E + tid + 256 * i] = cacc[i];
  };

  int it = 0;
  for (long long g = g_lo; g < g_hi; g++, it++) {
    const int s = it % STG, r = it / STG;
    const int n = (int)(g / gm.cpr), c = (int)(g - (long long)n * gm.cpr);
    const int mod = c >= gm.cpi, cc = mod ? c - gm.cpi : c;
    const int L = mod ? a.Lt : a.Li, j0 = cc * ST_CH, nj = min(ST_CH, L - j0);
    if (n != cur_n || mod != cur_mod) {
      if (cur_n >= 0) flush();
      cur_n = n;
      cur_mod = mod;
      slot0 = cc;
      m_run = -INFINITY;
      l_run = 0.f;
#pragma unroll
      for (int i = 0; i < CPT; i++) cacc[i] = 0.f;
      const float* sp = a.S + (long long)n * a.ldS + mod * E;
      const float* wp = a.w_att + mod * E;
#pragma unroll
      for (int k = 0; k < KV; k++) {
        sreg[k] = ld4(sp + lane * 4 + 128 * k);
        wreg[k] = ld4(wp + lane * 4 + 128 * k);
      }
      beta = a.beta_att[mod];
    }
    const float* Hs = ring + (size_t)s * 2 * TILE;
    const float* Vs = Hs + TILE;
    mbar_wait(&full[s], r & 1);
    // energies: warp w <-> position j0 + w
    float* eb = e_sh + (it & 1) * ST_CH;
    if (warp < nj) {
      const float* hp = Hs + warp * E;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < KV; k++) {
        const float4 h = ld4(hp + lane * 4 + 128 * k);
        acc = fmaf(wreg[k].x, tanh_fast<APPROX>(h.x + sreg[k].x), acc);
        acc = fmaf(wreg[k].y, tanh_fast<APPROX>(h.y + sreg[k].y), acc);
        acc = fmaf(wreg[k].z, tanh_fast<APPROX>(h.z + sreg[k].z), acc);
        acc = fmaf(wreg[k].w, tanh_fast<APPROX>(h.w + sreg[k].w), acc);
      }
      acc = warp_sum(acc) + beta;
      if (lane == 0) {
        eb[warp] = acc;
        (mod ? a.alpha_tr : a.alpha_img)[(long long)n * L + j0 + warp] = acc;   // raw energy for now
      }
    } else if (lane == 0) {
      eb[warp] = -INFINITY;
    }
    named_bar_sync(1, ST_CONS);
    // online softmax over the chunk (every warp keeps its own identical copy of m, l)
    const float ej = lane < ST_CH ? eb[lane] : -INFINITY;
    const float m_new = fmaxf(m_run, warp_max(ej));
    const float scale = expf(m_run - m_new);
    const float pj = expf(ej - m_new);              // exp(-inf) = 0 for absent positions
    l_run = l_run * scale + warp_sum(pj);
    m_run = m_new;
#pragma unroll
    for (int i = 0; i < CPT; i++) cacc[i] *= scale;
    if (nj == ST_CH) {     // full chunk: fully unrolled, loads batched ahead of the FMAs
#pragma unroll
      for (int j = 0; j < ST_CH; j++) {
        const float p = __shfl_sync(FULL, pj, j);
#pragma unroll
        for (int i = 0; i < CPT; i++) cacc[i] = fmaf(p, Vs[j * E + tid + 256 * i], cacc[i]);
      }
    } else {
      for (int j = 0; j < nj; j++) {
        const float p = __shfl_sync(FULL, pj, j);
#pragma unroll
        for (int i = 0; i < CPT; i++) cacc[i] = fmaf(p, Vs[j * E + tid + 256 * i], cacc[i]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  if (cur_n >= 0) flush();
}

// Merge the partials of one (row, modality): context, and raw energies -> softmax weights.
__global__ void __launch_bounds__(256)
attn_combine_kernel(StreamArgs sa) {
  const AttnArgs& a = sa.a;
  const StreamGeom gm = stream_geom(a.N, a.Li, a.Lt, sa.use_img, sa.use_tr);
  const int n = blockIdx.x, mod = blockIdx.y + (sa.use_img ? 0 : 1), E = a.E;
  const int L = mod ? a.Lt : a.Li;
  const int cnt = mod ? gm.cpt : gm.cpi;
  const long long base = (long long)n * gm.cpr + (mod ? gm.cpi : 0);
  float M = -INFINITY;
  for (int i = 0; i < cnt; i++)
    if (sa.PL[base + i] > 0.f) M = fmaxf(M, sa.PM[base + i]);
  float Lsum = 0.f;
  for (int i = 0; i < cnt; i++) {
    const float l = sa.PL[base + i];
    if (l > 0.f) Lsum += l * expf(sa.PM[base + i] - M);
  }
  const float inv = 1.0f / Lsum;
  for (int x = threadIdx.x; x < E; x += blockDim.x) {
    float c = 0.f;
    for (int i = 0; i < cnt; i++) {
      const float l = sa.PL[base + i];
      if (l > 0.f) c = fmaf(expf(sa.PM[base + i] - M), sa.PC[(base + i) * E + x], c);
    }
    c *= inv;
    if (mod) c += a.b_tl[x];
    a.C[((long long)n * 2 + mod) * E + x] = c;
  }
  float* al = (mod ? a.alpha_tr : a.alpha_img) + (long long)n * L;
  for (int j = threadIdx.x; j < L; j += blockDim.x) al[j] = expf(al[j] - M) * inv;
}

// ------------------------------------------------------------------------------------------
// Backward of the same step, same streaming skeleton.  With c = sum_j alpha_j V_j saved by the
// forward, the softmax-backward inner product is segment-local:  sum_j alpha_j (dc . V_j) = dc . c,
// so a chunk needs nothing from other CTAs:
//   phase A (warp per position):  dalpha_j = dc . V_j ;  de_j = alpha_j (dalpha_j - dc.c)
//   phase B (thread per column):  q = tanh(H_j + s) ;  ds += de_j (1-q^2) ;  dw += de_j q
// Partial (ds, dw) of a cut segment are summed by attn_combine_bwd_kernel.
struct StreamBwdArgs {
  AttnBwdArgs a;
  const float* C;        // [N,2,E] forward contexts of this step (trend includes b_tl)
  const float* b_tl;
  int use_img, use_tr;
  float *PV, *PS, *PW;   // partials: valid flag [N,cpr] (zeroed), ds [N,cpr,E], dw [N,cpr,E]
};

template <int CPT, bool APPROX>
__global__ void __launch_bounds__(ST_THREADS, stream_ctas(256 * CPT))
attn_stream_bwd_kernel(StreamBwdArgs sa) {
  constexpr int E = 256 * CPT, KV = 2 * CPT;
  constexpr int STG = stream_stages(E);
  constexpr int TILE = ST_CH * E;
  extern __shared__ uint8_t raw[];
  float* ring = reinterpret_cast<float*>(smem_align(raw, 128));
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)STG * 2 * TILE);
  uint64_t* empty = full + STG;
  float* e_sh = reinterpret_cast<float*>(empty + STG);  // [2][ST_CH]
  float* red = e_sh + 2 * ST_CH;                        // [8]

  const AttnBwdArgs& a = sa.a;
  const StreamGeom gm = stream_geom(a.N, a.Li, a.Lt, sa.use_img, sa.use_tr);
  const long long g_lo = gm.total * blockIdx.x / gridDim.x, g_hi = gm.total * (blockIdx.x + 1) / gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

  if (tid == 0) {
    for (int s = 0; s < STG; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ST_CONS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == ST_CONS / 32) {
    if (lane == 0) {
      int it = 0;
      for (long long g = g_lo; g < g_hi; g++, it++) {
        const int s = it % STG, r = it / STG;
        const int n = (int)(g / gm.cpr), c = (int)(g - (long long)n * gm.cpr);
        const int mod = c >= gm.cpi, cc = mod ? c - gm.cpi : c;
        const int L = mod ? a.Lt : a.Li, j0 = cc * ST_CH, nj = min(ST_CH, L - j0);
        const long long off = ((long long)(n / a.W) * L + j0) * E;
        const uint32_t bytes = (uint32_t)nj * E * 4u;
        mbar_wait(&empty[s], (r & 1) ^ 1);
        mbar_expect_tx(&full[s], 2u * bytes);
        bulk_g2s(ring + (size_t)s * 2 * TILE, (mod ? a.Htr : a.Himg) + off, bytes, &full[s]);
        bulk_g2s(ring + (size_t)s * 2 * TILE + TILE, (mod ? a.Ptr : a.Vimg) + off, bytes, &full[s]);
      }
    }
    return;
  }
  float4 dcreg[KV];
  float scol[CPT], sacc[CPT], wacc[CPT];
  float dot = 0.f;
  int cur_n = -1, cur_mod = 0, slot0 = 0;

  auto flush = [&]() {
    const long long slot = (long long)cur_n * gm.cpr + (cur_mod ? gm.cpi : 0) + slot0;
    if (tid == 0) sa.PV[slot] = 1.f;
#pragma unroll
    for (int i = 0; i < CPT; i++) {
      sa.PS[slot * E + tid + 256 * i] = sacc[i];
      sa.PW[slot * E + tid + 256 * i] = wacc[i];
    }
  };

  int it = 0;
  for (long long g = g_lo; g < g_hi; g++, it++) {
    const int s = it % STG, r = it / STG;
    const int n = (int)(g / gm.cpr), c = (int)(g - (long long)n * gm.cpr);
    const int mod = c >= gm.cpi, cc = mod ? c - gm.cpi : c;
    const int L = mod ? a.Lt : a.Li, j0 = cc * ST_CH, nj = min(ST_CH, L - j0);
    if (n != cur_n || mod != cur_mod) {
      if (cur_n >= 0) flush();
      cur_n = n;
      cur_mod = mod;
      slot0 = cc;
      const float* dcp = a.DC + ((long long)n * 2 + mod) * E;
      const float* cp = sa.C + ((long long)n * 2 + mod) * E;
      const float* sp = a.S + (long long)n * a.ldS + mod * E;
      float part = 0.f;
#pragma unroll
      for (int k = 0; k < KV; k++) dcreg[k] = ld4(dcp + lane * 4 + 128 * k);
#pragma unroll
      for (int i = 0; i < CPT; i++) {
        const int x = tid + 256 * i;
        scol[i] = sp[x];
        sacc[i] = 0.f;
        wacc[i] = 0.f;
        part = fmaf(dcp[x], cp[x] - (mod ? sa.b_tl[x] : 0.f), part);
      }
      // dot = dc . c over the 256 consumer threads
      part = warp_sum(part);
      named_bar_sync(2, ST_CONS);          // previous readers of red[] are done
      if (lane == 0) red[warp] = part;
      named_bar_sync(2, ST_CONS);
      dot = 0.f;
#pragma unroll
      for (int w = 0; w < ST_CONS / 32; w++) dot += red[w];
    }
    const float* al = (mod ? a.alpha_tr : a.alpha_img) + (long long)n * L + j0;
    const float alv = lane < nj ? al[lane] : 0.f;     // issued before the wait: latency hidden
    const float* Hs = ring + (size_t)s * 2 * TILE;
    const float* Vs = Hs + TILE;
    mbar_wait(&full[s], r & 1);
    float* eb = e_sh + (it & 1) * ST_CH;
    if (warp < nj) {
      const float* vp = Vs + warp * E;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < KV; k++) {
        const float4 v = ld4(vp + lane * 4 + 128 * k);
        acc = fmaf(v.x, dcreg[k].x, acc);
        acc = fmaf(v.y, dcreg[k].y, acc);
        acc = fmaf(v.z, dcreg[k].z, acc);
        acc = fmaf(v.w, dcreg[k].w, acc);
      }
      acc = warp_sum(acc);
      const float de = __shfl_sync(FULL, alv, warp) * (acc - dot);
      if (lane == 0) {
        eb[warp] = de;
        (mod ? a.DE_tr : a.DE_img)[(long long)n * L + j0 + warp] = de;
      }
    } else if (lane == 0) {
      eb[warp] = 0.f;
    }
    named_bar_sync(1, ST_CONS);
    if (nj == ST_CH) {
#pragma unroll
      for (int j = 0; j < ST_CH; j++) {
        const float de = eb[j];
#pragma unroll
        for (int i = 0; i < CPT; i++) {
          const float q = tanh_fast<APPROX>(Hs[j * E + tid + 256 * i] + scol[i]);
          sacc[i] = fmaf(de, fmaf(-q, q, 1.f), sacc[i]);
          wacc[i] = fmaf(de, q, wacc[i]);
        }
      }
    } else {
      for (int j = 0; j < nj; j++) {
        const float de = eb[j];
#pragma unroll
        for (int i = 0; i < CPT; i++) {
          const float q = tanh_fast<APPROX>(Hs[j * E + tid + 256 * i] + scol[i]);
          sacc[i] = fmaf(de, fmaf(-q, q, 1.f), sacc[i]);
          wacc[i] = fmaf(de, q, wacc[i]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  if (cur_n >= 0) flush();
}

__global__ void __launch_bounds__(256)
attn_combine_bwd_kernel(StreamBwdArgs sa) {
  const AttnBwdArgs& a = sa.a;
  const StreamGeom gm = stream_geom(a.N, a.Li, a.Lt, sa.use_img, sa.use_tr);
  const int n = blockIdx.x, mod = blockIdx.y + (sa.use_img ? 0 : 1), E = a.E;
  const int cnt = mod ? gm.cpt : gm.cpi;
  const long long base = (long long)n * gm.cpr + (mod ? gm.cpi : 0);
  for (int x = threadIdx.x; x < E; x += blockDim.x) {
    float ds = 0.f, dw = 0.f;
    for (int i = 0; i < cnt; i++)
      if (sa.PV[base + i] != 0.f) {
        ds += sa.PS[(base + i) * E + x];
        dw += sa.PW[(base + i) * E + x];
      }
    a.DS[(long long)n * a.ldS + mod * E + x] = ds * a.w_att[mod * E + x];
    a.dw_acc[((long long)n * 3 + mod) * E + x] += dw;
  }
}

bool attn_stream_supported(int E) { return E % 256 == 0 && E >= 256 && E <= 1024; }

long long attn_stream_ws_floats(int N, int Li, int Lt, int E) {
  const StreamGeom g = stream_geom(N, Li, Lt, true, true);
  return g.total * (2 * E + 2);
}

template <int CPT, bool APPROX>
static int launch_stream(const StreamArgs& sa, int sms, cudaStream_t s) {
  constexpr int E = 256 * CPT;
  constexpr int STG = stream_stages(E);
  constexpr size_t smem = 128 + (size_t)STG * 2 * ST_CH * E * 4 + 2 * STG * 8 + 2 * ST_CH * 4 + 16;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_stream_fwd_kernel<CPT, APPROX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return V2F_ERR_LAUNCH;
    attr = true;
  }
  const StreamGeom g = stream_geom(sa.a.N, sa.a.Li, sa.a.Lt, sa.use_img, sa.use_tr);
  const long long slots = (long long)sms * stream_ctas(E);
  const int grid = (int)(g.total < slots ? g.total : slots);
  prof_begin(V2F_K_ATTN_FWD, s);
  attn_stream_fwd_kernel<CPT, APPROX><<<grid, ST_THREADS, smem, s>>>(sa);
  prof_end(V2F_K_ATTN_FWD, s);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

int attn_stream_fwd(const AttnArgs& a, bool use_img, bool use_tr, float* ws, cudaStream_t s) {
  V2F_REQUIRE(ws && attn_stream_supported(a.E), V2F_ERR_UNSUPPORTED);
  const StreamGeom g = stream_geom(a.N, a.Li, a.Lt, use_img, use_tr);
  if (g.total == 0) return V2F_OK;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  StreamArgs sa{a, use_img ? 1 : 0, use_tr ? 1 : 0, ws, ws + g.total, ws + 2 * g.total};
  cudaMemsetAsync(sa.PL, 0, sizeof(float) * (size_t)g.total, s);
#define ST_DISPATCH(AP)                                             \
  switch (a.E / 256) {                                              \
    case 1: V2F_TRY((launch_stream<1, AP>(sa, sms, s))); break;     \
    case 2: V2F_TRY((launch_stream<2, AP>(sa, sms, s))); break;     \
    case 3: V2F_TRY((launch_stream<3, AP>(sa, sms, s))); break;     \
    default: V2F_TRY((launch_stream<4, AP>(sa, sms, s))); break;    \
  }
  if (a.approx) {
    ST_DISPATCH(true)
  } else {
    ST_DISPATCH(false)
  }
#undef ST_DISPATCH
  attn_combine_kernel<<<dim3(a.N, (use_img ? 1 : 0) + (use_tr ? 1 : 0)), 256, 0, s>>>(sa);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

template <int CPT, bool APPROX>
static int launch_stream_bwd(const StreamBwdArgs& sa, int sms, cudaStream_t s) {
  constexpr int E = 256 * CPT;
  constexpr int STG = stream_stages(E);
  constexpr size_t smem = 128 + (size_t)STG * 2 * ST_CH * E * 4 + 2 * STG * 8 + 2 * ST_CH * 4 + 8 * 4 + 16;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_stream_bwd_kernel<CPT, APPROX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return V2F_ERR_LAUNCH;
    attr = true;
  }
  const StreamGeom g = stream_geom(sa.a.N, sa.a.Li, sa.a.Lt, sa.use_img, sa.use_tr);
  const long long slots = (long long)sms * stream_ctas(E);
  const int grid = (int)(g.total < slots ? g.total : slots);
  prof_begin(V2F_K_ATTN_BWD, s);
  attn_stream_bwd_kernel<CPT, APPROX><<<grid, ST_THREADS, smem, s>>>(sa);
  prof_end(V2F_K_ATTN_BWD, s);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

int attn_stream_bwd(const AttnBwdArgs& a, const float* C, const float* b_tl, bool use_img, bool use_tr,
                    int approx, float* ws, cudaStream_t s) {
  V2F_REQUIRE(ws && attn_stream_supported(a.E), V2F_ERR_UNSUPPORTED);
  const StreamGeom g = stream_geom(a.N, a.Li, a.Lt, use_img, use_tr);
  if (g.total == 0) return V2F_OK;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  StreamBwdArgs sa{a, C, b_tl, use_img ? 1 : 0, use_tr ? 1 : 0, ws, ws + g.total, ws + g.total + g.total * a.E};
  cudaMemsetAsync(sa.PV, 0, sizeof(float) * (size_t)g.total, s);
#define ST_DISPATCH(AP)                                               \
  switch (a.E / 256) {                                                \
    case 1: V2F_TRY((launch_stream_bwd<1, AP>(sa, sms, s))); break;   \
    case 2: V2F_TRY((launch_stream_bwd<2, AP>(sa, sms, s))); break;   \
    case 3: V2F_TRY((launch_stream_bwd<3, AP>(sa, sms, s))); break;   \
    default: V2F_TRY((launch_stream_bwd<4, AP>(sa, sms, s))); break;  \
  }
  if (approx) {
    ST_DISPATCH(true)
  } else {
    ST_DISPATCH(false)
  }
#undef ST_DISPATCH
  attn_combine_bwd_kernel<<<dim3(a.N, (use_img ? 1 : 0) + (use_tr ? 1 : 0)), 256, 0, s>>>(sa);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

}  // namespace v2f

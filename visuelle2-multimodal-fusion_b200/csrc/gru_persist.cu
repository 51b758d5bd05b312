// Persistent GRU over a sequence: ONE launch for all L steps, the recurrent weights resident in
// shared memory for the whole horizon.  sm_100a.
//
// Reference arithmetic: torch.nn.GRU (gate order r,z,n) of TSEmbedder
// (/root/reference/models/CrossAttnRNN210.py:13-24, 52 steps), sales_encoder_gru (:123,182) and
// SalesEncoder (/root/reference/models/GTM_Visuelle2.py:99-107); same cell equations as gru.cu.
//
// Why: the per-step product h_{t-1} W_hh^T is [N<=128, H] x [H, 3H] -- 0.2 GFLOP, far too small to
// fill the GPU -- and the step-per-launch path spends ~20 us per step on launch + pipeline ramp of a
// GEMM and a gate kernel (52 steps forward and backward = 2 ms of the step).  Here a cooperative grid
// of (H/16 unit blocks) x (ceil(rows/32) row blocks) CTAs runs the whole recurrence:
//   * CTA (ub, rb) owns hidden units [16 ub, 16 ub+16) of rows [32 rb, 32 rb+32): its 48 rows of
//     W_hh (forward) or its 16 columns of W_hh (backward) -- 96 KB at H=512 -- are loaded into
//     shared memory ONCE and reused by every step;
//   * per step it reads the 32 x H slice of h_{t-1} (forward) / the 32 x 3H slice of dgh_t
//     (backward) that the other unit blocks of its row block produced, through L2 (ld.global.cg),
//     does the 32x48xH (32x16x3H) product in exact fp32 FMAs out of shared memory, applies the gate
//     equations and publishes its slice;
//   * the 32 CTAs of a row block synchronise once per step on a monotonic global counter
//     (release: __syncthreads + __threadfence + atomicAdd; acquire: spin on ld.acquire, bounded --
//     a protocol bug traps instead of hanging the GPU).  Row blocks never wait for each other.
// The kernel is launched cooperatively so that all CTAs are co-resident.  Exact fp32 (no tensor
// cores): it serves both precision modes and removes the tf32 error from the recurrence.
#include <type_traits>

#include "common.cuh"

namespace v2f {

constexpr int GP_THREADS = 256;
constexpr int GP_UN = 16;      // hidden units per CTA
constexpr int GP_RB = 32;      // rows per CTA
constexpr int GP_SLOTS = 16;   // barrier-counter slots (calls in flight on different streams)
constexpr int GP_MAXRB = 8;    // row blocks per launch (<= 4 at H=512 to stay co-resident)

__device__ unsigned g_gp_bar[GP_SLOTS][GP_MAXRB];

__device__ __forceinline__ void row_block_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned v;
    long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > (1LL << 31)) __trap();   // ~1 s: never hang the device
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}

struct GpFwdArgs {
  int N, L, H, rows0;   // rows0: first row of this launch (row blocks are relative to it)
  const float *GI, *h0, *w_hh, *b_hh;
  float *out, *RZN, *GHN;
  unsigned* bar;
};

// shared: Ws [H/4][48] float4 (k-quad major), hs [32][H/4 + 1] float4, red [2][32][48] float
__global__ void __launch_bounds__(GP_THREADS, 1)
gru_persist_fwd_kernel(GpFwdArgs a) {
  extern __shared__ float4 sm4[];
  const int H = a.H, H4 = H >> 2, HP = H4 + 1;
  float4* Ws = sm4;
  float4* hs = Ws + H4 * 48;
  float* red = reinterpret_cast<float*>(hs + GP_RB * HP);
  const int tid = threadIdx.x;
  const int u0 = blockIdx.x * GP_UN;
  const int r0 = a.rows0 + blockIdx.y * GP_RB;
  const int nrows = min(GP_RB, a.N - r0);
  unsigned* ctr = a.bar + blockIdx.y;
  const unsigned nub = gridDim.x;

  // ---- one-time: this CTA's 48 rows of W_hh -> Ws[k4][g*16+u]
  for (int i = tid; i < 48 * H4; i += GP_THREADS) {
    const int col = i / H4, k4 = i - col * H4;
    const int g = col >> 4, u = col & 15;
    Ws[k4 * 48 + col] = ld4(a.w_hh + (long long)(g * H + u0 + u) * H + 4 * k4);
  }
  const int kh = tid >> 7;            // K half
  const int tr = (tid >> 4) & 7;      // 4-row group
  const int tu = tid & 15;            // unit
  for (int t = 0; t < a.L; t++) {
    if (t > 0) row_block_barrier(ctr, (unsigned)t * nub);
    // ---- h_{t-1} slice [32, H] through L2 (written by other SMs: bypass L1)
    const float* hp = t == 0 ? a.h0 : a.out + (long long)(t - 1) * H;
    const long long ldh = t == 0 ? H : (long long)a.L * H;
    for (int i = tid; i < GP_RB * H4; i += GP_THREADS) {
      const int r = i / H4, k4 = i - r * H4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows) v = __ldcg(reinterpret_cast<const float4*>(hp + (long long)(r0 + r) * ldh) + k4);
      hs[r * HP + k4] = v;
    }
    __syncthreads();
    // ---- gh[32 x 48] = hs Ws^T, K split in two halves
    float acc[4][3];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int g = 0; g < 3; g++) acc[r][g] = 0.f;
    const int kb = kh * (H4 >> 1), ke = kb + (H4 >> 1);
#pragma unroll 2
    for (int k4 = kb; k4 < ke; k4++) {
      float4 w[3], h[4];
#pragma unroll
      for (int g = 0; g < 3; g++) w[g] = Ws[k4 * 48 + g * 16 + tu];
#pragma unroll
      for (int r = 0; r < 4; r++) h[r] = hs[(tr * 4 + r) * HP + k4];
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int g = 0; g < 3; g++) acc[r][g] = dot4(h[r], w[g], acc[r][g]);
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int g = 0; g < 3; g++) red[(kh * GP_RB + tr * 4 + r) * 48 + g * 16 + tu] = acc[r][g];
    __syncthreads();
    // ---- gates: 32 rows x 16 units, two per thread
    for (int i = tid; i < GP_RB * GP_UN; i += GP_THREADS) {
      const int r = i >> 4, u = i & 15;
      if (r >= nrows) continue;
      const int n = r0 + r, gu = u0 + u;
      float gh[3];
#pragma unroll
      for (int g = 0; g < 3; g++)
        gh[g] = red[r * 48 + g * 16 + u] + red[(GP_RB + r) * 48 + g * 16 + u] + a.b_hh[g * H + gu];
      const float* gi = a.GI + ((long long)n * a.L + t) * 3 * H;
      const float rg = sigmoid_full(gi[gu] + gh[0]);
      const float zg = sigmoid_full(gi[H + gu] + gh[1]);
      const float cg = tanh_full(gi[2 * H + gu] + rg * gh[2]);
      const float4 hq = hs[r * HP + (gu >> 2)];
      const float hprev = (gu & 3) == 0 ? hq.x : (gu & 3) == 1 ? hq.y : (gu & 3) == 2 ? hq.z : hq.w;
      float* rzn = a.RZN + ((long long)t * a.N + n) * 3 * H;
      rzn[gu] = rg;
      rzn[H + gu] = zg;
      rzn[2 * H + gu] = cg;
      a.GHN[((long long)t * a.N + n) * H + gu] = gh[2];
      a.out[((long long)n * a.L + t) * H + gu] = (1.f - zg) * cg + zg * hprev;
    }
    // the next iteration's barrier (or kernel end) publishes out[:, t, own units]
  }
}

struct GpBwdArgs {
  int N, L, H, rows0;
  const float *h0, *w_hh, *out, *RZN, *GHN, *dOut, *dhL;
  float *DGI, *DGH, *Hprev, *dh_out;   // dh_out [N,H]: gradient w.r.t. h0
  unsigned* bar;
};

// shared: Wc [3H/4][16] float4 (k = gate row, k-quad major), dg [32][H/4 + 1] float4 (one gate chunk),
//         red [4][32][16] float, dhs [32][16] float (the gradient carried through time)
__global__ void __launch_bounds__(GP_THREADS, 1)
gru_persist_bwd_kernel(GpBwdArgs a) {
  extern __shared__ float4 sm4[];
  const int H = a.H, H4 = H >> 2, HP = H4 + 1, K4 = 3 * H4;
  float4* Wc = sm4;
  float4* dg = Wc + K4 * 16;
  float* red = reinterpret_cast<float*>(dg + GP_RB * HP);
  float* dhs = red + 4 * GP_RB * GP_UN;
  const int tid = threadIdx.x;
  const int u0 = blockIdx.x * GP_UN;
  const int r0 = a.rows0 + blockIdx.y * GP_RB;
  const int nrows = min(GP_RB, a.N - r0);
  unsigned* ctr = a.bar + blockIdx.y;
  const unsigned nub = gridDim.x;

  // ---- one-time: this CTA's 16 columns of W_hh -> Wc[j4][u] = (W[4j4..4j4+3][u0+u])
  for (int i = tid; i < K4 * 16; i += GP_THREADS) {
    const int j4 = i >> 4, u = i & 15;
    const float* p = a.w_hh + (long long)(4 * j4) * H + u0 + u;
    Wc[i] = make_float4(p[0], p[H], p[2 * (long long)H], p[3 * (long long)H]);
  }
  for (int i = tid; i < GP_RB * GP_UN; i += GP_THREADS) {
    const int r = i >> 4, u = i & 15;
    dhs[i] = (a.dhL && r < nrows) ? a.dhL[(long long)(r0 + r) * H + u0 + u] : 0.f;
  }
  __syncthreads();
  const int ks = tid >> 6;            // K quarter within a chunk
  const int tr = (tid >> 4) & 3;      // 8-row group
  const int tu = tid & 15;
  for (int t = a.L - 1; t >= 0; t--) {
    // ---- gate backward for the own (rows, units); publishes DGH[t, rows, own gate rows]
    const float* hp = t == 0 ? a.h0 : a.out + (long long)(t - 1) * H;
    const long long ldh = t == 0 ? H : (long long)a.L * H;
    for (int i = tid; i < GP_RB * GP_UN; i += GP_THREADS) {
      const int r = i >> 4, u = i & 15;
      if (r >= nrows) continue;
      const int n = r0 + r, gu = u0 + u;
      const float* rzn = a.RZN + ((long long)t * a.N + n) * 3 * H;
      const float rg = rzn[gu], zg = rzn[H + gu], cg = rzn[2 * H + gu];
      const float ghn = a.GHN[((long long)t * a.N + n) * H + gu];
      const float hprev = hp[(long long)n * ldh + gu];
      float dhp = dhs[i];
      if (a.dOut) dhp += a.dOut[((long long)n * a.L + t) * H + gu];
      const float dan = dhp * (1.f - zg) * (1.f - cg * cg);
      const float daz = dhp * (hprev - cg) * zg * (1.f - zg);
      const float dar = dan * ghn * rg * (1.f - rg);
      float* dgi = a.DGI + ((long long)n * a.L + t) * 3 * H;
      dgi[gu] = dar;
      dgi[H + gu] = daz;
      dgi[2 * H + gu] = dan;
      float* dgh = a.DGH + ((long long)t * a.N + n) * 3 * H;
      dgh[gu] = dar;
      dgh[H + gu] = daz;
      dgh[2 * H + gu] = dan * rg;
      a.Hprev[((long long)t * a.N + n) * H + gu] = hprev;
      dhs[i] = dhp * zg;
    }
    row_block_barrier(ctr, (unsigned)(a.L - t) * nub);
    // ---- dh[rows, own units] += DGH[t, rows, :] W_hh[:, own units], three gate chunks of H
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; r++) acc[r] = 0.f;
    for (int c = 0; c < 3; c++) {
      if (c > 0) __syncthreads();     // previous chunk fully consumed
      for (int i = tid; i < GP_RB * H4; i += GP_THREADS) {
        const int r = i / H4, k4 = i - r * H4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrows)
          v = __ldcg(reinterpret_cast<const float4*>(a.DGH + ((long long)t * a.N + r0 + r) * 3 * H + c * H) + k4);
        dg[r * HP + k4] = v;
      }
      __syncthreads();
      const int q = H4 >> 2, kb = ks * q, ke = kb + q;
#pragma unroll 2
      for (int k4 = kb; k4 < ke; k4++) {
        const float4 w = Wc[(c * H4 + k4) * 16 + tu];
#pragma unroll
        for (int r = 0; r < 8; r++) acc[r] = dot4(dg[(tr * 8 + r) * HP + k4], w, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 8; r++) red[(ks * GP_RB + tr * 8 + r) * GP_UN + tu] = acc[r];
    __syncthreads();
    for (int i = tid; i < GP_RB * GP_UN; i += GP_THREADS)
      dhs[i] += red[i] + red[GP_RB * GP_UN + i] + red[2 * GP_RB * GP_UN + i] + red[3 * GP_RB * GP_UN + i];
    __syncthreads();
  }
  if (a.dh_out)
    for (int i = tid; i < GP_RB * GP_UN; i += GP_THREADS) {
      const int r = i >> 4, u = i & 15;
      if (r < nrows) a.dh_out[(long long)(r0 + r) * H + u0 + u] = dhs[i];
    }
}

static int g_gp_slot = 0;

static size_t gp_smem(int H) {
  // fwd: Ws 48*H/4 + hs 32*(H/4+1) float4 + red 2*32*48 floats ; bwd: Wc 3H/4*16 + dg 32*(H/4+1) float4
  //      + red 4*32*16 + dhs 32*16 floats.  Same leading terms; take the max of the tails.
  const size_t f4 = (size_t)48 * (H / 4) + (size_t)GP_RB * (H / 4 + 1);
  return f4 * 16 + sizeof(float) * (2 * GP_RB * 48 + 4 * GP_RB * GP_UN + GP_RB * GP_UN);
}

static bool g_gp_enabled = true;
void gru_persist_enable(bool on) { g_gp_enabled = on; }

bool gru_persist_supported(int N, int L, int H) {
  if (!g_gp_enabled) return false;
  if (H % 16 != 0 || H < 16 || H > 512 || L < 2 || N < 1) return false;
  int dev = 0, coop = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return coop && (H / GP_UN) <= sms;
}

template <typename Args, typename Kern>
static int gp_launch(Kern kern, Args a, int N, int H, cudaStream_t s) {
  static bool attr_done[2] = {false, false};
  const size_t smem = gp_smem(H);
  const int which = std::is_same<Args, GpFwdArgs>::value ? 0 : 1;
  if (!attr_done[which]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return V2F_ERR_LAUNCH;
    attr_done[which] = true;
  }
  int dev = 0, sms = 148, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GP_THREADS, smem) != cudaSuccess || per_sm < 1)
    return V2F_ERR_UNSUPPORTED;
  const int nub = H / GP_UN;
  int rb_per_launch = (sms * per_sm) / nub;          // co-resident row blocks
  if (rb_per_launch > GP_MAXRB) rb_per_launch = GP_MAXRB;
  if (rb_per_launch < 1) return V2F_ERR_UNSUPPORTED;
  const int total_rb = (N + GP_RB - 1) / GP_RB;
  unsigned* bar_base = nullptr;
  if (cudaGetSymbolAddress((void**)&bar_base, g_gp_bar) != cudaSuccess) return V2F_ERR_LAUNCH;
  for (int rb0 = 0; rb0 < total_rb; rb0 += rb_per_launch) {
    const int nrb = total_rb - rb0 < rb_per_launch ? total_rb - rb0 : rb_per_launch;
    const int slot = g_gp_slot;
    g_gp_slot = (g_gp_slot + 1) % GP_SLOTS;
    a.rows0 = rb0 * GP_RB;
    a.bar = bar_base + slot * GP_MAXRB;
    cudaMemsetAsync(a.bar, 0, sizeof(unsigned) * GP_MAXRB, s);
    void* params[] = {&a};
    if (cudaLaunchCooperativeKernel((void*)kern, dim3(nub, nrb), dim3(GP_THREADS), params, smem, s) != cudaSuccess)
      return V2F_ERR_LAUNCH;
    ++g_v2f_launches;
  }
  return V2F_OK;
}

int gru_persist_fwd(int N, int L, int H, const float* GI, const float* h0, const float* w_hh, const float* b_hh,
                    float* out, float* RZN, float* GHN, cudaStream_t s) {
  GpFwdArgs a{N, L, H, 0, GI, h0, w_hh, b_hh, out, RZN, GHN, nullptr};
  return gp_launch(gru_persist_fwd_kernel, a, N, H, s);
}

int gru_persist_bwd(int N, int L, int H, const float* h0, const float* w_hh, const float* out, const float* RZN,
                    const float* GHN, const float* dOut, const float* dhL, float* DGI, float* DGH, float* Hprev,
                    float* dh_out, cudaStream_t s) {
  GpBwdArgs a{N, L, H, 0, h0, w_hh, out, RZN, GHN, dOut, dhL, DGI, DGH, Hprev, dh_out, nullptr};
  return gp_launch(gru_persist_bwd_kernel, a, N, H, s);
}

}  // namespace v2f

// Debug / A-B switch: 0 routes v2f_gru_seq_* through the step-per-launch path again (default 1).
extern "C" int v2f_gru_persistent_enable(int on) {
  v2f::gru_persist_enable(on != 0);
  return V2F_OK;
}

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
//   * per step it pulls the 32 x H slice of h_{t-1} (forward) / the 32 x 3H slice of dgh_t
//     (backward) that the other unit blocks of its row block produced, through L2 with cp.async.cg,
//     does the 32x48xH (32x16x3H) product out of shared memory -- exact fp32 FMAs (precision 0) or
//     warp-level tf32 tensor-core MMAs with round-to-nearest operands (precision 1) -- applies the gate
//     equations and publishes its slice; the step's saved activations are prefetched into registers
//     before the barrier so their latency hides behind it;
//   * the unit blocks of a row block synchronise once per step on a monotonic global counter
//     (release: __syncthreads + __threadfence + atomicAdd; acquire: spin on ld.acquire, bounded --
//     a protocol bug traps instead of hanging the GPU).  Row blocks never wait for each other.
// The kernel is launched cooperatively so that all CTAs are co-resident.
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
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// D += A(16x8, row) * B(8x8, col), tf32 inputs, fp32 accumulate
__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Asynchronous copy of a [GP_RB, W] fp32 tile (global row pitch ld floats, shared row pitch P floats)
// through L2 only (cp.async.cg: the data was written by other SMs).  Rows >= nrows are zero-filled.
__device__ __forceinline__ void tile_copy_async(float* dst, int P, const float* src, long long ld, int W, int nrows) {
  const int W4 = W >> 2;
  for (int i = threadIdx.x; i < GP_RB * W4; i += GP_THREADS) {
    const int r = i / W4, k4 = i - r * W4;
    float* d = dst + r * P + 4 * k4;
    if (r < nrows) {
      const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(d));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + (long long)r * ld + 4 * k4)
                   : "memory");
    } else {
      *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tile_copy_wait() {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}

struct GpFwdArgs {
  int N, L, H, rows0;   // rows0: first row of this launch (row blocks are relative to it)
  const float *GI, *h0, *w_hh, *b_hh;
  float *out, *RZN, *GHN;
  unsigned* bar;
};

// shared (floats): Wn [48][H+4] (row = gate*16 + unit), hs [32][H+4], red [4][32][48]
template <bool TC>
__global__ void __launch_bounds__(GP_THREADS, 1)
gru_persist_fwd_kernel(GpFwdArgs a) {
  extern __shared__ __align__(16) float smf[];
  const int H = a.H, P = H + 4;
  float* Wn = smf;
  float* hs = Wn + 48 * P;
  float* red = hs + GP_RB * P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u0 = blockIdx.x * GP_UN;
  const int r0 = a.rows0 + blockIdx.y * GP_RB;
  const int nrows = min(GP_RB, a.N - r0);
  unsigned* ctr = a.bar + blockIdx.y;
  const unsigned nub = gridDim.x;

  // ---- one-time: this CTA's 48 rows of W_hh (rounded to tf32 once in tensor-core mode)
  for (int i = tid; i < 48 * (H >> 2); i += GP_THREADS) {
    const int col = i / (H >> 2), k4 = i - col * (H >> 2);
    float4 w = ld4(a.w_hh + (long long)((col >> 4) * H + u0 + (col & 15)) * H + 4 * k4);
    if (TC) {
      w.x = __uint_as_float(tf32_rna(w.x));
      w.y = __uint_as_float(tf32_rna(w.y));
      w.z = __uint_as_float(tf32_rna(w.z));
      w.w = __uint_as_float(tf32_rna(w.w));
    }
    *reinterpret_cast<float4*>(Wn + col * P + 4 * k4) = w;
  }
  // the two (row, unit) pairs this thread owns in the gate phase, and their recurrent biases
  const int pr[2] = {tid >> 4, (tid >> 4) + 16};
  const int pu = tid & 15, gu = u0 + pu;
  float bh[3];
#pragma unroll
  for (int g = 0; g < 3; g++) bh[g] = a.b_hh[g * H + gu];

  for (int t = 0; t < a.L; t++) {
    // ---- prefetch the input projections of this step (independent of h): hides behind the barrier
    float gi[2][3];
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
      for (int g = 0; g < 3; g++)
        gi[p][g] = pr[p] < nrows ? a.GI[((long long)(r0 + pr[p]) * a.L + t) * 3 * H + g * H + gu] : 0.f;
    if (t > 0) row_block_barrier(ctr, (unsigned)t * nub);
    const float* hp = t == 0 ? a.h0 : a.out + (long long)(t - 1) * H;
    const long long ldh = t == 0 ? H : (long long)a.L * H;
    tile_copy_async(hs, P, hp + (long long)r0 * ldh, ldh, H, nrows);
    tile_copy_wait();
    // ---- gh[32 x 48] = hs Wn^T ; partial sums over K quarters / halves go to red[part][row][col]
    if (TC) {
      const int mt = warp & 1, kq = warp >> 1, g = lane >> 2, tt = lane & 3;
      float acc[6][4];
#pragma unroll
      for (int n = 0; n < 6; n++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[n][j] = 0.f;
      const float* ha = hs + (16 * mt + g) * P + tt;
      const float* wb = Wn + g * P + tt;
      const int kb = kq * (H >> 2), ke = kb + (H >> 2);
#pragma unroll 2
      for (int k0 = kb; k0 < ke; k0 += 8) {
        uint32_t af[4];
        af[0] = tf32_rna(ha[k0]);
        af[1] = tf32_rna(ha[8 * P + k0]);
        af[2] = tf32_rna(ha[k0 + 4]);
        af[3] = tf32_rna(ha[8 * P + k0 + 4]);
#pragma unroll
        for (int n = 0; n < 6; n++)
          mma_tf32(acc[n], af, __float_as_uint(wb[n * 8 * P + k0]), __float_as_uint(wb[n * 8 * P + k0 + 4]));
      }
#pragma unroll
      for (int n = 0; n < 6; n++) {
        float* rp = red + (kq * GP_RB + 16 * mt + g) * 48 + n * 8 + 2 * tt;
        rp[0] = acc[n][0];
        rp[1] = acc[n][1];
        rp[8 * 48] = acc[n][2];
        rp[8 * 48 + 1] = acc[n][3];
      }
    } else {
      const int kh = tid >> 7, tr = (tid >> 4) & 7, tu = tid & 15;
      float acc[4][3];
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int g = 0; g < 3; g++) acc[r][g] = 0.f;
      const int kb = kh * (H >> 1), ke = kb + (H >> 1);
#pragma unroll 2
      for (int k = kb; k < ke; k += 4) {
        float4 w[3], h[4];
#pragma unroll
        for (int g = 0; g < 3; g++) w[g] = *reinterpret_cast<const float4*>(Wn + (g * 16 + tu) * P + k);
#pragma unroll
        for (int r = 0; r < 4; r++) h[r] = *reinterpret_cast<const float4*>(hs + (tr * 4 + r) * P + k);
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
          for (int g = 0; g < 3; g++) acc[r][g] = dot4(h[r], w[g], acc[r][g]);
      }
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int g = 0; g < 3; g++) {
          red[(kh * GP_RB + tr * 4 + r) * 48 + g * 16 + tu] = acc[r][g];
          red[((kh + 2) * GP_RB + tr * 4 + r) * 48 + g * 16 + tu] = 0.f;
        }
    }
    __syncthreads();
    // ---- gates: 32 rows x 16 units, two per thread
#pragma unroll
    for (int p = 0; p < 2; p++) {
      const int r = pr[p];
      if (r >= nrows) continue;
      const int n = r0 + r;
      float gh[3];
#pragma unroll
      for (int g = 0; g < 3; g++) {
        const int c = g * 16 + pu;
        gh[g] = red[r * 48 + c] + red[(GP_RB + r) * 48 + c] + red[(2 * GP_RB + r) * 48 + c] +
                red[(3 * GP_RB + r) * 48 + c] + bh[g];
      }
      const float rg = sigmoid_full(gi[p][0] + gh[0]);
      const float zg = sigmoid_full(gi[p][1] + gh[1]);
      const float cg = tanh_full(gi[p][2] + rg * gh[2]);
      const float hprev = hs[r * P + gu];
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

struct SavedStep { float rg, zg, cg, ghn, hprev, dout; };

// shared (floats): Wc [16][3H+4] (row = own unit, column = gate row j), dg [32][H+4] (one gate chunk of
// dgh_t), red [4][32][16], dhs [32][16] (the gradient carried through time)
template <bool TC>
__global__ void __launch_bounds__(GP_THREADS, 1)
gru_persist_bwd_kernel(GpBwdArgs a) {
  extern __shared__ __align__(16) float smf[];
  const int H = a.H, P = H + 4, PW = 3 * H + 4;
  float* Wc = smf;
  float* dg = Wc + 16 * PW;
  float* red = dg + GP_RB * P;
  float* dhs = red + 4 * GP_RB * GP_UN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u0 = blockIdx.x * GP_UN;
  const int r0 = a.rows0 + blockIdx.y * GP_RB;
  const int nrows = min(GP_RB, a.N - r0);
  unsigned* ctr = a.bar + blockIdx.y;
  const unsigned nub = gridDim.x;

  // ---- one-time: this CTA's 16 columns of W_hh, transposed: Wc[u][j] = W_hh[j][u0+u]
  for (int i = tid; i < 3 * H * 16; i += GP_THREADS) {
    const int j = i >> 4, u = i & 15;
    float w = a.w_hh[(long long)j * H + u0 + u];
    if (TC) w = __uint_as_float(tf32_rna(w));
    Wc[u * PW + j] = w;
  }
  const int pr[2] = {tid >> 4, (tid >> 4) + 16};
  const int pu = tid & 15, gu = u0 + pu;
#pragma unroll
  for (int p = 0; p < 2; p++)
    dhs[pr[p] * GP_UN + pu] = (a.dhL && pr[p] < nrows) ? a.dhL[(long long)(r0 + pr[p]) * H + gu] : 0.f;

  SavedStep sv[2];
  auto load_saved = [&](int t) {
    const float* hp = t == 0 ? a.h0 : a.out + (long long)(t - 1) * H;
    const long long ldh = t == 0 ? H : (long long)a.L * H;
#pragma unroll
    for (int p = 0; p < 2; p++) {
      if (pr[p] >= nrows) continue;
      const int n = r0 + pr[p];
      const float* rzn = a.RZN + ((long long)t * a.N + n) * 3 * H;
      sv[p].rg = rzn[gu];
      sv[p].zg = rzn[H + gu];
      sv[p].cg = rzn[2 * H + gu];
      sv[p].ghn = a.GHN[((long long)t * a.N + n) * H + gu];
      sv[p].hprev = hp[(long long)n * ldh + gu];
      sv[p].dout = a.dOut ? a.dOut[((long long)n * a.L + t) * H + gu] : 0.f;
    }
  };
  load_saved(a.L - 1);
  __syncthreads();
  for (int t = a.L - 1; t >= 0; t--) {
    // ---- gate backward for the own (rows, units); publishes DGH[t, rows, own gate rows]
#pragma unroll
    for (int p = 0; p < 2; p++) {
      if (pr[p] >= nrows) continue;
      const int n = r0 + pr[p], i = pr[p] * GP_UN + pu;
      const float rg = sv[p].rg, zg = sv[p].zg, cg = sv[p].cg;
      const float dhp = dhs[i] + sv[p].dout;
      const float dan = dhp * (1.f - zg) * (1.f - cg * cg);
      const float daz = dhp * (sv[p].hprev - cg) * zg * (1.f - zg);
      const float dar = dan * sv[p].ghn * rg * (1.f - rg);
      float* dgi = a.DGI + ((long long)n * a.L + t) * 3 * H;
      dgi[gu] = dar;
      dgi[H + gu] = daz;
      dgi[2 * H + gu] = dan;
      float* dgh = a.DGH + ((long long)t * a.N + n) * 3 * H;
      dgh[gu] = dar;
      dgh[H + gu] = daz;
      dgh[2 * H + gu] = dan * rg;
      a.Hprev[((long long)t * a.N + n) * H + gu] = sv[p].hprev;
      dhs[i] = dhp * zg;
    }
    if (t > 0) load_saved(t - 1);          // consumed after this step's product: latency hidden
    row_block_barrier(ctr, (unsigned)(a.L - t) * nub);
    // ---- dh[rows, own units] += DGH[t, rows, :] W_hh[:, own units], three gate chunks of H
    float accf[8];
    float acct[2][4];
#pragma unroll
    for (int r = 0; r < 8; r++) accf[r] = 0.f;
#pragma unroll
    for (int n = 0; n < 2; n++)
#pragma unroll
      for (int j = 0; j < 4; j++) acct[n][j] = 0.f;
    for (int c = 0; c < 3; c++) {
      if (c > 0) __syncthreads();     // previous chunk fully consumed
      tile_copy_async(dg, P, a.DGH + ((long long)t * a.N + r0) * 3 * H + c * H, 3LL * H, H, nrows);
      tile_copy_wait();
      if (TC) {
        const int mt = warp & 1, kq = warp >> 1, g = lane >> 2, tt = lane & 3;
        const float* da = dg + (16 * mt + g) * P + tt;
        const float* wb = Wc + g * PW + c * H + tt;
        const int kb = kq * (H >> 2), ke = kb + (H >> 2);
#pragma unroll 2
        for (int k0 = kb; k0 < ke; k0 += 8) {
          uint32_t af[4];
          af[0] = tf32_rna(da[k0]);
          af[1] = tf32_rna(da[8 * P + k0]);
          af[2] = tf32_rna(da[k0 + 4]);
          af[3] = tf32_rna(da[8 * P + k0 + 4]);
#pragma unroll
          for (int n = 0; n < 2; n++)
            mma_tf32(acct[n], af, __float_as_uint(wb[n * 8 * PW + k0]), __float_as_uint(wb[n * 8 * PW + k0 + 4]));
        }
      } else {
        const int ks = tid >> 6, tr = (tid >> 4) & 3, tu = tid & 15;
        const int kb = ks * (H >> 2), ke = kb + (H >> 2);
#pragma unroll 2
        for (int k = kb; k < ke; k += 4) {
          const float4 w = *reinterpret_cast<const float4*>(Wc + tu * PW + c * H + k);
#pragma unroll
          for (int r = 0; r < 8; r++)
            accf[r] = dot4(*reinterpret_cast<const float4*>(dg + (tr * 8 + r) * P + k), w, accf[r]);
        }
      }
    }
    if (TC) {
      const int mt = warp & 1, kq = warp >> 1, g = lane >> 2, tt = lane & 3;
#pragma unroll
      for (int n = 0; n < 2; n++) {
        float* rp = red + (kq * GP_RB + 16 * mt + g) * GP_UN + n * 8 + 2 * tt;
        rp[0] = acct[n][0];
        rp[1] = acct[n][1];
        rp[8 * GP_UN] = acct[n][2];
        rp[8 * GP_UN + 1] = acct[n][3];
      }
    } else {
      const int ks = tid >> 6, tr = (tid >> 4) & 3, tu = tid & 15;
#pragma unroll
      for (int r = 0; r < 8; r++) red[(ks * GP_RB + tr * 8 + r) * GP_UN + tu] = accf[r];
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 2; p++) {
      const int i = pr[p] * GP_UN + pu;
      dhs[i] += red[i] + red[GP_RB * GP_UN + i] + red[2 * GP_RB * GP_UN + i] + red[3 * GP_RB * GP_UN + i];
    }
    __syncthreads();
  }
  if (a.dh_out)
#pragma unroll
    for (int p = 0; p < 2; p++)
      if (pr[p] < nrows) a.dh_out[(long long)(r0 + pr[p]) * H + gu] = dhs[pr[p] * GP_UN + pu];
}

static int g_gp_slot = 0;
static bool g_gp_enabled = true;
void gru_persist_enable(bool on) { g_gp_enabled = on; }

static size_t gp_smem(int H, bool fwd) {
  if (fwd) return sizeof(float) * ((size_t)(48 + GP_RB) * (H + 4) + 4 * GP_RB * 48);
  return sizeof(float) * ((size_t)16 * (3 * H + 4) + (size_t)GP_RB * (H + 4) + 4 * GP_RB * GP_UN + GP_RB * GP_UN);
}

bool gru_persist_supported(int N, int L, int H) {
  if (!g_gp_enabled) return false;
  if (H % 32 != 0 || H < 32 || H > 512 || L < 2 || N < 1) return false;
  int dev = 0, coop = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return coop && (H / GP_UN) <= sms;
}

template <typename Args, typename Kern>
static int gp_launch(Kern kern, bool* attr_done, Args a, int N, int H, size_t smem, cudaStream_t s) {
  if (!*attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return V2F_ERR_LAUNCH;
    *attr_done = true;
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
                    float* out, float* RZN, float* GHN, int precision, cudaStream_t s) {
  static bool done[2] = {false, false};
  GpFwdArgs a{N, L, H, 0, GI, h0, w_hh, b_hh, out, RZN, GHN, nullptr};
  if (precision) return gp_launch(gru_persist_fwd_kernel<true>, &done[1], a, N, H, gp_smem(H, true), s);
  return gp_launch(gru_persist_fwd_kernel<false>, &done[0], a, N, H, gp_smem(H, true), s);
}

int gru_persist_bwd(int N, int L, int H, const float* h0, const float* w_hh, const float* out, const float* RZN,
                    const float* GHN, const float* dOut, const float* dhL, float* DGI, float* DGH, float* Hprev,
                    float* dh_out, int precision, cudaStream_t s) {
  static bool done[2] = {false, false};
  GpBwdArgs a{N, L, H, 0, h0, w_hh, out, RZN, GHN, dOut, dhL, DGI, DGH, Hprev, dh_out, nullptr};
  if (precision) return gp_launch(gru_persist_bwd_kernel<true>, &done[1], a, N, H, gp_smem(H, false), s);
  return gp_launch(gru_persist_bwd_kernel<false>, &done[0], a, N, H, gp_smem(H, false), s);
}

}  // namespace v2f

// Debug / A-B switch: 0 routes v2f_gru_seq_* through the step-per-launch path again (default 1).
extern "C" int v2f_gru_persistent_enable(int on) {
  v2f::gru_persist_enable(on != 0);
  return V2F_OK;
}

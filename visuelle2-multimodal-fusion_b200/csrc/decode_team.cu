// Row-team persistent decoder (forward): ONE cooperative launch for the whole T-step horizon of the
// CrossAttnRNN210 / CrossAttnRNNDemand decode loop at the reference's default dims (E = A = H = 512), products on
// tcgen05 tensor cores with the weights resident in shared memory, attention tiles streamed as bf16.  sm_100a.
//
// Reference arithmetic: /root/reference/models/CrossAttnRNN210.py:191-225 (loop), :83-89 (AdditiveAttention),
// :135-140,210-211 (decoder nn.GRU cell), :141,212-225 (decoder_fc + teacher forcing);
// /root/reference/models/CrossAttnRNNDemand.py:285-347,134-149.  Same equations and the same saved activations as
// decode_persist.cu / rnn_decode.cu (the backward consumes them unchanged).
//
// Why a second persistent design.  decode_persist.cu splits the WEIGHT columns over all 148 CTAs, so every product
// needs every row's activations in every CTA (38 MB of L2->SM traffic per phase), runs on legacy mma.sync, needs a
// 148-way grid barrier per phase and a combine pass for attention partials: 64 us per step, of which 19 us stream
// tiles.  Here rows are dealt to TEAMS of 64 CTAs, 64 rows per team:
//   * a CTA owns 1/64 of the weight rows -- 8 query columns of each attention, 8 hidden units x 3 gates of W_hh,
//     8 rows of We_mm, 8 units x 3 gates of W' = W_ihc W_me: 80 rows x 512 bf16 = 80 KB, loaded once by TMA
//     (128-byte swizzle) and kept for all T steps -- and is the owner of ONE row for the row-local phases;
//   * a product is D[weight row, batch row] = W_slice (A operand, M = 128 tile over the resident slice; rows
//     beyond the slice are don't-care) x X^T (B operand: the team's 64 activation rows, bf16, TMA-loaded in 64-wide
//     K chunks through an 8-slot mbarrier ring), tcgen05.mma kind::f16 issued by one thread, fp32 accumulator in
//     TMEM, read back with tcgen05.ld: 64 KB of activations per CTA per product instead of 256 KB;
//   * the attention sweep of a row runs entirely inside its owner CTA, in two passes over bf16 tiles fed by one bulk-copy
//     ring: energies (warp per position, no inter-warp dependency), softmax over the 152 stored energies, then the
//     weighted sums (thread per column pair, plain FMAs): no online softmax, no combine phase, no partial traffic
//     through L2, 304 KB per row and step instead of 608 KB;
//   * barriers are per team (64 arrivals) and there are five per step.
// GI = CTX W_ihc^T is re-associated to U (W_ihc W_me)^T (exact); CTX itself (saved for the backward) is one GEMM
// over all T*N rows after the loop.  The GRU state, softmax, gates and all saved activations stay fp32; bf16 is
// the operand format of the products and the storage format of the streamed tiles (tensor-core mode, 2e-2 contract).
#include "attn.cuh"
#include "gemm_dispatch.cuh"
#include "tc.cuh"

namespace v2f {

constexpr int DT_E = 512;                    // E = A = H = 512 only (train_dl.py:197-199)
constexpr int DT_NG = 64;                    // rows per team = MMA N
constexpr int DT_CG = 64;                    // CTAs per team
constexpr int DT_CONS = 512;                 // 16 consumer warps = 2 groups of 8
constexpr int DT_GRP = 256;
constexpr int DT_THREADS = DT_CONS + 64;     // + warp 16 (weight TMA, bulk-copy producer of the sweep), warp 17 (TMEM owner; backward: MMA issuer)
constexpr int DT_WROWS = 80;                 // resident weight rows per CTA: 24 W' | 8 We_mm | 48 Wcat
constexpr int DT_R5 = 0, DT_R3 = 24, DT_R1 = 32;
constexpr int DT_KCH = DT_E / 64;            // 8 K chunks of 64 bf16 (one 128-byte swizzle span)
constexpr int DT_WCHUNK = DT_WROWS * 128;    // bytes of one K chunk of the weight slice
constexpr int DT_WBYTES = DT_KCH * DT_WCHUNK;
constexpr int DT_BSLOT = DT_NG * 128;        // 8 KB: 64 activation rows x 64 bf16
constexpr int DT_NBS = 8;
constexpr int DT_CH = 8;                     // positions per attention chunk = warps of a consumer group
constexpr int DT_TSLOT = DT_CH * DT_E * 2;   // 8 KB: one operand (H or V) of one chunk, bf16
constexpr int DT_NTS = 8;                    // tile-ring slots
constexpr int DT_RING = DT_NBS * DT_BSLOT;   // 64 KB, shared by the B ring (products) and the tile rings (sweep)
constexpr int DT_P1 = 65, DT_P3 = 129;       // staging pitches (odd: conflict-free transposed writes)
constexpr int DT_NCTR = 8;
constexpr int DT_BARW = DT_NCTR * 32;        // unsigned words of barrier counters per team
constexpr int DT_STAMPS = 32;
constexpr int DT_NI = 1;                      // MMA-issuing warps (consumer warps 8..11): a tcgen05 instruction costs its issuing thread ~85 cycles
constexpr uint32_t DT_TMEM_COLS = 512;       // P1 / P5: 4 partial accumulators x 64 columns at 0; P3: 4 x 64 at 256
constexpr int DT_MAXTEAMS = 2;
constexpr int DT_MAXL = 256;                 // positions per attention (image map 100, trends 52)

static_assert(DT_RING == DT_NTS * DT_TSLOT, "the two rings alias");

struct DtArgs {
  v2f_decode_params p;
  const float* bp;                                  // [3H] = W_ihc b_me + b_ih
  const __nv_bfloat16 *Himg, *Vimg, *Htr, *Ptr;     // bf16 copies of the tiles
  __nv_bfloat16 *hb, *Cb, *Ub;                      // [Np,512], [2,Np,512], [Np,512]: B operands of the products
  float* ypart;                                     // [64, Np] per-CTA partial decoder_fc dot products
  unsigned* bar;                                    // [teams, DT_BARW] (zeroed before launch)
  unsigned long long* stamps;                       // optional [T, DT_STAMPS]
  int Np;                                           // teams * 64
  int dbg;                                          // timing experiments only (results invalid): bit 0 skip the energy arithmetic, bit 1 skip the context arithmetic
};

__device__ __forceinline__ unsigned long long dt_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Team barrier: 64 arrivals spread over DT_NCTR monotonic counters on separate 128-byte lines; lanes 0..7 of warp 0
// poll one each (a sum of monotonic counters read one by one is a lower bound of the arrivals).  Bounded spin.
__device__ __forceinline__ void dt_team_barrier(unsigned* ctr, int c, unsigned& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0)
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr + (c % DT_NCTR) * 32) : "memory");
    const unsigned target = epoch * DT_CG;
    const unsigned* cp = ctr + (threadIdx.x % DT_NCTR) * 32;
    const long long t0 = clock64();
    for (;;) {
      unsigned v = 0;
      if (threadIdx.x < DT_NCTR) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cp) : "memory");
#pragma unroll
      for (int o = DT_NCTR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      v = __shfl_sync(FULL, v, 0);
      if (v >= target) break;
      if (clock64() - t0 > (1LL << 31)) __trap();   // ~1 s: a protocol bug must not hang the device
    }
    // the other CTAs' generic-proxy stores (hb / Cb / Ub / the backward's operand buffers) acquired above are read next
    // by TMA (async proxy): one cross-proxy fence here, ordered before every issuing thread by the barrier below.
    // (A fence in each of the 8 issuing threads serialised their copies: ~600 cycles apiece.)
    if (threadIdx.x == 0) asm volatile("fence.proxy.async.global;" ::: "memory");
  }
  __syncthreads();
}

// 16 accumulator columns of this warp's 32 TMEM lanes
__device__ __forceinline__ void dt_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void dt_tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void dt_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void dt_tmem_ld8_nowait(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
// sum of NP partial accumulators (64 columns apart) of 8 columns of this warp's lanes
template <int NP>
__device__ __forceinline__ void dt_tmem_ld8_sum(uint32_t taddr, float (&out)[8]) {
  uint32_t v[NP][8];
#pragma unroll
  for (int k = 0; k < NP; k++) dt_tmem_ld8_nowait(taddr + (uint32_t)(k * DT_NG), v[k]);
  dt_tmem_wait_ld();
#pragma unroll
  for (int q = 0; q < 8; q++) {
    float t = __uint_as_float(v[0][q]);
#pragma unroll
    for (int k = 1; k < NP; k++) t += __uint_as_float(v[k][q]);
    out[q] = t;
  }
}

__device__ __forceinline__ float2 dt_bf2(uint32_t w) {   // two packed bf16 -> two fp32
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t dt_pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// global column (of S / bcat) of weight row j (0..47) of the P1 slice of CTA c
__device__ __forceinline__ int dt_col1(int j, int c) {
  return j < 24 ? (j >> 3) * DT_E + 8 * c + (j & 7) : 3 * DT_E + ((j - 24) >> 3) * DT_E + 8 * c + ((j - 24) & 7);
}

__global__ void __launch_bounds__(DT_THREADS, 1)
decode_team_fwd_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapH,
                       const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapU,
                       const __grid_constant__ DtArgs a) {
  constexpr int E = DT_E, H = DT_E, ldS = 6 * DT_E;
  extern __shared__ uint8_t raw[];
  const v2f_decode_params& p = a.p;
  const int N = p.N, T = p.T, Li = p.Li, Lt = p.Lt, Wn = p.W, Np = a.Np;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int team = blockIdx.x / DT_CG, c = blockIdx.x % DT_CG;
  const int n0 = team * DT_NG;
  const int n_own = n0 + c;                       // the row this CTA owns in the row-local phases
  const bool own = n_own < N;

  uint8_t* sm = smem_align(raw, 1024);
  uint8_t* Wsm = sm;                                            // [8 chunks][80 rows][128 B], swizzled by TMA
  uint8_t* ring = sm + DT_WBYTES;                               // B ring / tile rings
  float* stage1 = reinterpret_cast<float*>(ring + DT_RING);     // [48][65]: S columns and GH of the team's rows
  float* stage5 = stage1 + 48 * DT_P1;                          // [24][65] GI  (P3: [8][129] HC)
  float* cvec = stage5 + 24 * DT_P1 + 8;                        // [2][512] contexts of the own row
  float* pacc = cvec + 2 * E;                                   // [2 groups][2 modalities][512] partial contexts
  float* e_all = pacc + 4 * E;                                  // [2][DT_MAXL] energies of the own row
  float* al_all = e_all + 2 * DT_MAXL;                          // [2][DT_MAXL] softmax weights
  float* red = al_all + 2 * DT_MAXL;                            // [64]
  uint64_t* bfull = reinterpret_cast<uint64_t*>(red + 64);      // [8] B ring
  uint64_t* bempty = bfull + DT_NBS;                            // [8]
  uint64_t* tfull = bempty + DT_NBS;                            // [8] tile ring (same 64 KB as the B ring)
  uint64_t* tempty = tfull + DT_NBS;                            // [8]
  uint64_t* mma_done = tempty + DT_NBS;
  uint64_t* wfull = mma_done + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wfull + 1);

  // ------------------------------------------------------------------ one-time set-up
  if (tid == 0) {
    for (int s = 0; s < DT_NBS; s++) {
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], 1);
    }
    for (int s = 0; s < DT_NTS; s++) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], DT_GRP / 32);
    }
    mbar_init(mma_done, DT_NI);
    mbar_init(wfull, 1);
    mbar_fence_init();
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(DT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;
  if (warp == 16 && lane == 0) {                    // the weight slice: 8 boxes of [80 rows x 128 B]
    mbar_expect_tx(wfull, DT_WBYTES);
    for (int kc = 0; kc < DT_KCH; kc++) tma_load_3d(Wsm + kc * DT_WCHUNK, &mapW, wfull, kc * 64, c * DT_WROWS, 0);
  }

  const __nv_bfloat16* Vimg_b = a.Vimg;
  const bool byproj = p.variant == 2;
  const int mod_mask = p.mod_mask;
  const unsigned* mask_dev = p.y ? p.tf_mask_dev : nullptr;
  unsigned* bar = a.bar + team * DT_BARW;
  unsigned epoch = 0;
  uint32_t bq = 0;                 // B-ring chunks issued (producer) / consumed (MMA thread) so far
  uint32_t md = 0;                 // completed waits on mma_done
  const int grp = warp < 8 ? 0 : (warp < 16 ? 1 : warp - 16);
  const int gt = tid & (DT_GRP - 1), gw = (tid >> 5) & 7;
  uint32_t tq = 0;                 // tile-ring slots used so far (all steps)
  const int cpi = (Li + DT_CH - 1) / DT_CH, cpt = (Lt + DT_CH - 1) / DT_CH, ctot = cpi + cpt;
  const uint32_t idesc = umma_idesc<0>(DT_NG);
  auto stamp = [&](int t, int k) {
    if (a.stamps && blockIdx.x == 0 && tid == 0) a.stamps[t * DT_STAMPS + k] = dt_globaltimer();
  };

  // gate threads: (row nl of the team, hidden unit u of this CTA); the state h lives in a register for all T steps
  const int nl = tid >> 3, gu = tid & 7;
  const int ng = n0 + nl;
  const bool gact = tid < DT_CONS && ng < N;
  float hreg = 0.f;
  if (gact) {
    hreg = p.h_all[(long long)ng * H + 8 * c + gu];
    a.hb[(long long)ng * H + 8 * c + gu] = __float2bfloat16_rn(hreg);
  }
  // B ring helpers (warp 16 lane 0 produces, warp 17 lane 0 issues the MMAs)
  // A thread that issues TMA / bulk copies sustains one copy per ~600 cycles, whatever its size (tools/probes/
  // bulk_probe.cu: 14 B/clk/SM from one issuer at 8 KB per copy, 76 B/clk/SM from eight) -- so copies are issued by MANY
  // warps: chunk i of a product by lane 0 of consumer warp i % 8 (ring slot w is always filled by warp w, in order: a
  // warp that skipped a use of its slot would see the mbarrier parity of two phases ago and overwrite live data), the
  // tile chunks of the sweep by a warp of the group that will consume them.
  auto load_chunk = [&](const CUtensorMap* map, int row0, int i, uint32_t q) {      // chunk i of a phase, ring sequence q
    const uint32_t s = q % DT_NBS, r = q / DT_NBS;
    mbar_wait(&bempty[s], (r & 1) ^ 1);
    mbar_expect_tx(&bfull[s], DT_BSLOT);
    tma_load_3d(ring + s * DT_BSLOT, map, &bfull[s], (i % DT_KCH) * 64, row0 + (i / DT_KCH) * Np, 0);
  };
  // chunks [i0, i0 + n) of a product phase accumulate into the 64 columns at dcol (one of DT_NI partial accumulators,
  // issued by DT_NI different warps in parallel: the sum is taken by the epilogue)
  auto issue = [&](int r0, uint32_t dcol, int i0, int n, int t, int st0, bool last = true) {
    for (int i = i0; i < i0 + n; i++) {
      const uint32_t q = bq + (uint32_t)i, s = q % DT_NBS, r = q / DT_NBS;
      const int kc = i % DT_KCH;
      mbar_wait(&bfull[s], r & 1);
      if (a.stamps && blockIdx.x == 0 && st0 && (i == i0 || i == i0 + n - 1)) a.stamps[t * DT_STAMPS + st0 + (i != i0)] = dt_globaltimer();
      tc_fence_after();
      const uint64_t ad = umma_desc_sw128(smem_u32(Wsm + kc * DT_WCHUNK + r0 * 128));
      const uint64_t bd = umma_desc_sw128(smem_u32(ring + s * DT_BSLOT));
#pragma unroll
      for (int k = 0; k < 4; k++) umma<0>(tmem_d + dcol, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, ((i - i0) | k) ? 1u : 0u);
      umma_commit(&bempty[s]);
    }
    if (last) umma_commit(mma_done);
    if (a.stamps && blockIdx.x == 0 && st0) a.stamps[t * DT_STAMPS + st0 + 2] = dt_globaltimer();
  };
  if (warp >= 8 && warp < 8 + DT_NI && lane == 0) mbar_wait(wfull, 0);
  dt_team_barrier(bar, c, epoch);                  // hb of step 0 is complete

  for (int t = 0; t < T; t++) {
    float* S = p.S_all + (long long)t * N * ldS;
    float* C = p.C + (long long)t * N * 2 * E;
    float* HC = p.HC + (long long)t * N * 2 * E;
    float* U = p.U + (long long)t * N * E;
    float* al_img = p.alpha_img + (long long)t * N * Li;
    float* al_tr = p.alpha_tr + (long long)t * N * Lt;
    stamp(t, 0);
    // ================================================================ P1: S^T slice = Wcat_slice h^T (+ bcat)
    if (warp < 16) {
      if (warp < DT_KCH) {
        if (lane == 0) load_chunk(&mapH, n0, warp, bq + (uint32_t)warp);
        __syncwarp();
      } else if (warp < 8 + DT_NI) {
        const int k = warp - 8;
        if (lane == 0) issue(DT_R1, (uint32_t)(k * DT_NG), k * (DT_KCH / DT_NI), DT_KCH / DT_NI, t, k == 0 ? 20 : 0);
        __syncwarp();
      }
      stamp(t, 16);
      if (warp < 2) {
        mbar_wait(mma_done, md & 1);
        stamp(t, 17);
        tc_fence_after();
        const int j = warp * 32 + lane;            // TMEM lane = weight row of the slice
        const float bias = j < 48 ? p.bcat[dt_col1(j, c)] : 0.f;
#pragma unroll
        for (int c0 = 0; c0 < DT_NG; c0 += 8) {
          float v[8];
          dt_tmem_ld8_sum<DT_NI>(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
          if (j < 48)
#pragma unroll
            for (int q = 0; q < 8; q++) stage1[j * DT_P1 + c0 + q] = v[q] + bias;
        }
        tc_fence_before();
      }
      md++;
      stamp(t, 18);
      named_bar_sync(3, DT_CONS);
      stamp(t, 19);
      {   // 8 consecutive columns (32 B) per (row, segment): s_img | s_tr | s_mm | gh_r | gh_z | gh_n
        const int seg = tid & 7;
        if (seg < 6 && ng < N) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; q++) v[q] = stage1[(seg * 8 + q) * DT_P1 + nl];
          float* dst = S + (long long)ng * ldS + dt_col1(seg * 8, c);
          st4(dst, make_float4(v[0], v[1], v[2], v[3]));
          st4(dst + 4, make_float4(v[4], v[5], v[6], v[7]));
        }
      }
    }
    bq += DT_KCH;
    stamp(t, 1);
    dt_team_barrier(bar, c, epoch);
    stamp(t, 2);
    // ================================================================ P2: additive attention of the own row
    // Two passes over the row's tiles, no inter-warp dependency inside a pass.  Pass 1 streams the H tiles: warp gw of
    // group g computes the energy of position gw of the chunks g, g+2, ... on its own (no group barrier, the slot is
    // released when its 8 warps have arrived).  Then the 152 energies are normalised (every warp redundantly, 5
    // values per lane).  Pass 2 streams the V tiles: thread = 2 columns, c += alpha_j V_j, plain FMAs (the weights
    // are final: no online-softmax rescaling).  The producer runs ahead through the 8-slot ring across both passes.
    if (own) {
      const int n = n_own, b = n / Wn;
      if (warp == 16) {
        if (lane == 0) {
          uint32_t q = tq;
          for (int pass = 0; pass < 2; pass++)
            for (int cc0 = 0; cc0 < ctot; cc0++, q++) {
              const int mod = cc0 >= cpi, cc = mod ? cc0 - cpi : cc0;
              const int L = mod ? Lt : Li, j0 = cc * DT_CH, nj = min(DT_CH, L - j0);
              const long long off = ((long long)b * L + j0) * E;
              const uint32_t bytes = (uint32_t)nj * E * 2u;
              const uint32_t sl = q % DT_NTS, r = q / DT_NTS;
              const __nv_bfloat16* src = pass == 0 ? (mod ? a.Htr : a.Himg) : (mod ? a.Ptr : Vimg_b);
              mbar_wait(&tempty[sl], (r & 1) ^ 1);
              mbar_expect_tx(&tfull[sl], bytes);
              bulk_g2s(ring + sl * DT_TSLOT, src + off, bytes, &tfull[sl]);
            }
        }
        __syncwarp();
      } else if (warp < 16) {
        // ---- pass 1: energies
        float wmax0 = -INFINITY, wmax1 = -INFINITY;      // running maxima of this warp's energies, per modality
        {
          float4 sreg[4], wreg[4];        // columns 8*lane + 256*k + [0,8) for k = 0,1: two float4 each
          float beta = 0.f;
          int cur_mod = -1;
          for (int cc0 = grp; cc0 < ctot; cc0 += 2) {
            const int mod = cc0 >= cpi, cc = mod ? cc0 - cpi : cc0;
            const int L = mod ? Lt : Li, j0 = cc * DT_CH, nj = min(DT_CH, L - j0);
            if (mod != cur_mod) {
              cur_mod = mod;
              const float* sp = S + (long long)n * ldS + mod * E;
              const float* wp = p.w_att + mod * E;
#pragma unroll
              for (int k = 0; k < 2; k++) {
                sreg[2 * k] = __ldcg(reinterpret_cast<const float4*>(sp + 8 * lane + 256 * k));
                sreg[2 * k + 1] = __ldcg(reinterpret_cast<const float4*>(sp + 8 * lane + 256 * k + 4));
                wreg[2 * k] = ld4(wp + 8 * lane + 256 * k);
                wreg[2 * k + 1] = ld4(wp + 8 * lane + 256 * k + 4);
              }
              beta = p.beta_att[mod];
            }
            const uint32_t q = tq + (uint32_t)cc0, sl = q % DT_NBS, r = q / DT_NBS;
            mbar_wait(&tfull[sl], r & 1);
            if (gw < nj) {
              const uint8_t* hp = ring + sl * DT_TSLOT + gw * (E * 2);
              float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
              for (int k = 0; k < 2; k++) {
                const uint4 hv = *reinterpret_cast<const uint4*>(hp + (8 * lane + 256 * k) * 2);
                const float2 h0 = dt_bf2(hv.x), h1 = dt_bf2(hv.y), h2 = dt_bf2(hv.z), h3 = dt_bf2(hv.w);
                const float4 s0 = sreg[2 * k], s1 = sreg[2 * k + 1], w0 = wreg[2 * k], w1 = wreg[2 * k + 1];
                acc0 = fmaf(w0.x, tanh_fast<true>(h0.x + s0.x), acc0);
                acc1 = fmaf(w0.y, tanh_fast<true>(h0.y + s0.y), acc1);
                acc0 = fmaf(w0.z, tanh_fast<true>(h1.x + s0.z), acc0);
                acc1 = fmaf(w0.w, tanh_fast<true>(h1.y + s0.w), acc1);
                acc0 = fmaf(w1.x, tanh_fast<true>(h2.x + s1.x), acc0);
                acc1 = fmaf(w1.y, tanh_fast<true>(h2.y + s1.y), acc1);
                acc0 = fmaf(w1.z, tanh_fast<true>(h3.x + s1.z), acc0);
                acc1 = fmaf(w1.w, tanh_fast<true>(h3.y + s1.w), acc1);
              }
              const float e = warp_sum(acc0 + acc1) + beta;
              if (lane == 0) e_all[mod * DT_MAXL + j0 + gw] = e;
              if (mod) wmax1 = fmaxf(wmax1, e);
              else wmax0 = fmaxf(wmax0, e);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[sl]);
          }
        }
        if (lane == 0) {
          red[warp] = wmax0;
          red[16 + warp] = wmax1;
        }
        stamp(t, 11);
        named_bar_sync(3, DT_CONS);
        stamp(t, 12);
        // ---- softmax: p_j = exp(e_j - max) in shared memory, 1 / sum per modality; thread j < 256: image position j,
        // thread 256 + j: trend position j (warps 0..7 / 8..15); the maximum comes from the warps' running maxima
        float inv0, inv1;
        {
          float m0 = red[lane & 15], m1 = red[16 + (lane & 15)];
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
            m0 = fmaxf(m0, __shfl_xor_sync(FULL, m0, o));
            m1 = fmaxf(m1, __shfl_xor_sync(FULL, m1, o));
          }
          const int mod = tid >> 8, jj = tid & 255;
          float pj = 0.f;
          if (jj < (mod ? Lt : Li)) {
            pj = __expf(e_all[mod * DT_MAXL + jj] - (mod ? m1 : m0));
            al_all[mod * DT_MAXL + jj] = pj;
          }
          pj = warp_sum(pj);
          if (lane == 0) red[32 + warp] = pj;
          named_bar_sync(3, DT_CONS);
          float l0 = red[32 + (lane & 7)], l1 = red[40 + (lane & 7)];
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {
            l0 += __shfl_xor_sync(FULL, l0, o);
            l1 += __shfl_xor_sync(FULL, l1, o);
          }
          inv0 = 1.0f / l0;
          inv1 = 1.0f / l1;
          // softmax weights saved for the backward pass / returned as attention maps
          if (jj < (mod ? Lt : Li))
            (mod ? al_tr + (long long)n * Lt : al_img + (long long)n * Li)[jj] = al_all[mod * DT_MAXL + jj] * (mod ? inv1 : inv0);
        }
        stamp(t, 13);
        // ---- pass 2: contexts (thread = columns 2 gt, 2 gt + 1 of both modalities; group g takes chunks g, g+2, ...)
        {
          float ci0 = 0.f, ci1 = 0.f, ct0 = 0.f, ct1 = 0.f;
          for (int cc0 = grp; cc0 < ctot; cc0 += 2) {
            const int mod = cc0 >= cpi, cc = mod ? cc0 - cpi : cc0;
            const int L = mod ? Lt : Li, j0 = cc * DT_CH, nj = min(DT_CH, L - j0);
            const uint32_t q = tq + (uint32_t)(ctot + cc0), sl = q % DT_NBS, r = q / DT_NBS;
            const float* alp = al_all + mod * DT_MAXL + j0;
            mbar_wait(&tfull[sl], r & 1);
            const uint8_t* vp = ring + sl * DT_TSLOT + gt * 4;
            float x0 = 0.f, x1 = 0.f;
            if (nj == DT_CH) {
#pragma unroll
              for (int j = 0; j < DT_CH; j++) {
                const float2 v = dt_bf2(*reinterpret_cast<const uint32_t*>(vp + j * (E * 2)));
                x0 = fmaf(alp[j], v.x, x0);
                x1 = fmaf(alp[j], v.y, x1);
              }
            } else {
              for (int j = 0; j < nj; j++) {
                const float2 v = dt_bf2(*reinterpret_cast<const uint32_t*>(vp + j * (E * 2)));
                x0 = fmaf(alp[j], v.x, x0);
                x1 = fmaf(alp[j], v.y, x1);
              }
            }
            if (mod) {
              ct0 += x0;
              ct1 += x1;
            } else {
              ci0 += x0;
              ci1 += x1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[sl]);
          }
          // partial contexts of this group: pacc[grp][mod][512]
          stamp(t, 14);
          *reinterpret_cast<float2*>(pacc + (grp * 2 + 0) * E + 2 * gt) = make_float2(ci0 * inv0, ci1 * inv0);
          *reinterpret_cast<float2*>(pacc + (grp * 2 + 1) * E + 2 * gt) = make_float2(ct0 * inv1, ct1 * inv1);
        }
        named_bar_sync(3, DT_CONS);
        {
          const int x = tid;
#pragma unroll
          for (int mod = 0; mod < 2; mod++) {
            float cv = pacc[mod * E + x] + pacc[(2 + mod) * E + x];
            if (mod) cv += p.b_tl[x];
            cvec[mod * E + x] = cv;
            C[((long long)n * 2 + mod) * E + x] = cv;
            a.Cb[((long long)mod * Np + n) * E + x] = __float2bfloat16_rn(cv);
          }
        }
      }
      tq += 2u * (uint32_t)ctot;
    }
    stamp(t, 3);
    dt_team_barrier(bar, c, epoch);
    stamp(t, 4);
    // ================================================================ P3: HC^T slice = We_mm_slice [c_img ; c_tr]^T
    if (warp < 16) {
      if (warp >= 8 && warp < 8 + DT_NI) {       // issuer k: modality k / PP, K part k % PP -> accumulator 256 + 64 k
        const int k = warp - 8;
        if (lane == 0) {
          if (DT_NI == 1) {
            issue(DT_R3, 256u, 0, DT_KCH, t, 0, false);
            issue(DT_R3, 256u + DT_NG, DT_KCH, DT_KCH, t, 0, true);
          } else {
            issue(DT_R3, (uint32_t)(256 + k * DT_NG), k * (2 * DT_KCH / DT_NI), 2 * DT_KCH / DT_NI, t, 0);
          }
        }
        __syncwarp();
      }
      if (warp < DT_KCH) {             // chunks 0..7: image contexts, 8..15: trend contexts; slot w always from warp w
        if (lane == 0) {
          load_chunk(&mapC, n0, warp, bq + (uint32_t)warp);
          load_chunk(&mapC, n0, warp + DT_KCH, bq + (uint32_t)(warp + DT_KCH));
        }
        __syncwarp();
      }
      if (warp == 0) {
        mbar_wait(mma_done, md & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 2 * DT_NG; c0 += 8) {         // column c0 = modality * 64 + row: PP = DT_NI / 2 partials at 256 + 64 PP mod + 64 j
          float v[8];
          constexpr int PP = DT_NI >= 2 ? DT_NI / 2 : 1;
          dt_tmem_ld8_sum<PP>(tmem_d + (uint32_t)(256 + (c0 / DT_NG) * PP * DT_NG + (c0 % DT_NG)), v);
          if (lane < 8)
#pragma unroll
            for (int q = 0; q < 8; q++) stage5[lane * DT_P3 + c0 + q] = v[q];
        }
        tc_fence_before();
      }
      md++;
      named_bar_sync(3, DT_CONS);
      {   // (row, modality) pair = tid / 4; 2 of the 8 columns per thread
        const int pair = tid >> 2, qd = tid & 3, mod = pair >> 6, r = pair & 63;
        if (n0 + r < N) {
          const float2 v = make_float2(stage5[(2 * qd) * DT_P3 + pair], stage5[(2 * qd + 1) * DT_P3 + pair]);
          *reinterpret_cast<float2*>(HC + ((long long)(n0 + r) * 2 + mod) * E + 8 * c + 2 * qd) = v;
        }
      }
    }
    bq += 2 * DT_KCH;
    stamp(t, 5);
    dt_team_barrier(bar, c, epoch);
    stamp(t, 6);
    // ================================================================ P4: multimodal attention of the own row -> U ; closes step t-1
    if (own) {
      const int n = n_own, b = n / Wn;
      float e[4] = {0.f, 0.f, 0.f, 0.f};
      float mv[2][4], hv[2][4];
      if (tid < DT_GRP) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const int x = tid + 256 * i;
          const float sx = __ldcg(S + (long long)n * ldS + 2 * E + x), wx = p.w_att[2 * E + x];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const bool on = (mod_mask >> k) & 1;
            float m_ = 0.f, h_ = 0.f;
            if (on) {
              if (k & 1) {
                m_ = cvec[(k >> 1) * E + x];
                h_ = __ldcg(HC + ((long long)n * 2 + (k >> 1)) * E + x);
              } else {
                m_ = p.Mst[((long long)b * 2 + (k >> 1)) * E + x];
                h_ = p.HMst[((long long)b * 2 + (k >> 1)) * E + x];
              }
              e[k] = fmaf(wx, tanh_acc(h_ + sx), e[k]);
            }
            mv[i][k] = m_;
            hv[i][k] = h_;
          }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) e[k] = warp_sum(e[k]);
        if (lane == 0)
#pragma unroll
          for (int k = 0; k < 4; k++) red[warp * 4 + k] = e[k];
      } else if (warp == 8 && t > 0) {
        float yp = __ldcg(a.ypart + (long long)lane * Np + n) + __ldcg(a.ypart + (long long)(lane + 32) * Np + n);
        yp = warp_sum(yp);
        if (lane == 0) {
          const float yh = yp + p.b_fc[0];
          p.yhat[(long long)n * T + t - 1] = yh;
          const int forced = mask_dev ? (int)((*mask_dev >> (t - 1)) & 1u) : (int)((p.tf_mask >> (t - 1)) & 1u);
          p.xin[(long long)t * N + n] = (forced && p.y) ? p.y[(long long)n * T + t - 1] : yh;
        }
      }
      __syncthreads();
      if (tid < DT_GRP) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < 8; w++) v += red[w * 4 + k];
          e[k] = v;
        }
        const float beta = p.beta_att[2];
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < 4; k++)
          if ((mod_mask >> k) & 1) {
            e[k] += beta;
            m = fmaxf(m, e[k]);
          }
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          e[k] = ((mod_mask >> k) & 1) ? expf(e[k] - m) : 0.f;
          sum += e[k];
        }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int k = 0; k < 4; k++) e[k] *= inv;
        if (tid == 0)
#pragma unroll
          for (int k = 0; k < 4; k++) p.alpha_mm[((long long)t * N + n) * 4 + k] = e[k];
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const int x = tid + 256 * i;
          float u = 0.f;
#pragma unroll
          for (int k = 0; k < 4; k++)
            if ((mod_mask >> k) & 1) u += mv[i][k] + e[k] * (byproj ? hv[i][k] : mv[i][k]);
          U[(long long)n * E + x] = u;
          a.Ub[(long long)n * E + x] = __float2bfloat16_rn(u);
        }
      }
    }
    stamp(t, 7);
    dt_team_barrier(bar, c, epoch);
    stamp(t, 8);
    // ================================================================ P5: GI^T slice = W'_slice U^T ; GRU gates of the own hidden units
    if (warp < 16) {
      if (warp < DT_KCH) {
        if (lane == 0) load_chunk(&mapU, n0, warp, bq + (uint32_t)warp);
        __syncwarp();
      } else if (warp < 8 + DT_NI) {
        const int k = warp - 8;
        if (lane == 0) issue(DT_R5, (uint32_t)(k * DT_NG), k * (DT_KCH / DT_NI), DT_KCH / DT_NI, t, 0);
        __syncwarp();
      }
      if (warp == 0) {
        mbar_wait(mma_done, md & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < DT_NG; c0 += 8) {
          float v[8];
          dt_tmem_ld8_sum<DT_NI>(tmem_d + (uint32_t)c0, v);
          if (lane < 24)
#pragma unroll
            for (int q = 0; q < 8; q++) stage5[lane * DT_P1 + c0 + q] = v[q];
        }
        tc_fence_before();
      }
      md++;
      named_bar_sync(3, DT_CONS);
      float yp = 0.f;
      if (gact) {
        const int uu = 8 * c + gu;
        const float x = __ldcg(p.xin + (long long)t * N + ng);
        float gi[3], gh[3];
#pragma unroll
        for (int g3 = 0; g3 < 3; g3++) {
          gi[g3] = stage5[(g3 * 8 + gu) * DT_P1 + nl] + a.bp[g3 * H + uu] + x * p.w_x[g3 * H + uu];
          gh[g3] = stage1[(24 + g3 * 8 + gu) * DT_P1 + nl];
        }
        const float rg = sigmoid_full(gi[0] + gh[0]);
        const float zg = sigmoid_full(gi[1] + gh[1]);
        const float cg = tanh_full(gi[2] + rg * gh[2]);
        const float hn = (1.f - zg) * cg + zg * hreg;
        float* rzn = p.RZN + ((long long)t * N + ng) * 3 * H;
        rzn[uu] = rg;
        rzn[H + uu] = zg;
        rzn[2 * H + uu] = cg;
        p.h_all[((long long)(t + 1) * N + ng) * H + uu] = hn;
        a.hb[(long long)ng * H + uu] = __float2bfloat16_rn(hn);
        hreg = hn;
        yp = p.w_fc[uu] * hn;
      }
      yp += __shfl_xor_sync(FULL, yp, 1);
      yp += __shfl_xor_sync(FULL, yp, 2);
      yp += __shfl_xor_sync(FULL, yp, 4);
      if (gu == 0 && gact) a.ypart[(long long)c * Np + ng] = yp;
    }
    bq += DT_KCH;
    stamp(t, 9);
    dt_team_barrier(bar, c, epoch);
    stamp(t, 10);
  }
  // ------------------------------------------------------------------ close the last step: yhat_{T-1}
  if (own && warp == 0) {
    const int n = n_own;
    float yp = __ldcg(a.ypart + (long long)lane * Np + n) + __ldcg(a.ypart + (long long)(lane + 32) * Np + n);
    yp = warp_sum(yp);
    if (lane == 0) {
      const float yh = yp + p.b_fc[0];
      p.yhat[(long long)n * T + T - 1] = yh;
      const int forced = mask_dev ? (int)((*mask_dev >> (T - 1)) & 1u) : (int)((p.tf_mask >> (T - 1)) & 1u);
      p.xin[(long long)T * N + n] = (forced && p.y) ? p.y[(long long)n * T + T - 1] : yh;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(DT_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ preparation kernels
// the per-CTA weight slices, bf16, row-major [64 CTAs][80 rows][512]: rows 0..23 W' (gate g, unit 8c+u), 24..31
// We_mm rows 8c.., 32..55 Wd_img | Wd_tr | Wd_mm rows 8c.., 56..79 W_hh (gate g, unit 8c+u)
__global__ void dt_pack_weights_kernel(const float* __restrict__ Wp, const float* __restrict__ We_mm,
                                       const float* __restrict__ Wcat, __nv_bfloat16* __restrict__ out) {
  constexpr int E = DT_E;
  const int row = blockIdx.x, c = row / DT_WROWS, r = row % DT_WROWS;
  const float* src;
  if (r < 24) src = Wp + (long long)((r >> 3) * E + 8 * c + (r & 7)) * E;
  else if (r < 32) src = We_mm + (long long)(8 * c + r - 24) * E;
  else src = Wcat + (long long)dt_col1(r - 32, c) * E;
  for (int k = threadIdx.x * 4; k < E; k += blockDim.x * 4) {
    const float4 v = ld4(src + k);
    uint2 o;
    o.x = dt_pack(v.x, v.y);
    o.y = dt_pack(v.z, v.w);
    *reinterpret_cast<uint2*>(out + (long long)row * E + k) = o;
  }
}

static bool g_dt_enabled = true;
static int g_dt_stamps = 0;

static size_t dt_smem() {
  return 1024 + (size_t)DT_WBYTES + DT_RING + sizeof(float) * (48 * DT_P1 + 24 * DT_P1 + 8 + 2 * DT_E + 4 * DT_E + 4 * DT_MAXL + 64) +
         8 * (4 * DT_NBS + 2) + 16;
}

struct DtLayout {
  long long bp, wpk, tiles, hb, cb, ub, ypart, bar, stamps, end;   // float offsets into team_ws
};
static DtLayout dt_layout(int N, int B, int T, int Li, int Lt) {
  const long long Np = (long long)((N + DT_NG - 1) / DT_NG) * DT_NG;
  DtLayout l;
  long long o = 0;
  auto take = [&](long long floats) { const long long at = o; o += (floats + 255) / 256 * 256; return at; };   // 1 KB aligned
  l.bp = take(3 * DT_E);
  l.wpk = take((long long)DT_CG * DT_WROWS * DT_E / 2);
  l.tiles = take((long long)B * (2 * Li + 2 * Lt) * DT_E / 2);
  l.hb = take(Np * DT_E / 2);
  l.cb = take(2 * Np * DT_E / 2);
  l.ub = take(Np * DT_E / 2);
  l.ypart = take((long long)DT_CG * Np);
  l.bar = take((long long)DT_MAXTEAMS * DT_BARW);
  l.stamps = take(2LL * T * DT_STAMPS);
  l.end = o;
  return l;
}

long long decode_team_ws_floats(int N, int B, int T, int Li, int Lt) { return dt_layout(N, B, T, Li, Lt).end; }

static bool dt_supported(const v2f_decode_params* p) {
  if (!g_dt_enabled || !p->team_ws || !p->persist_ws) return false;
  if (p->variant == 1 || p->T < 1 || p->precision == 0) return false;
  if ((p->mod_mask & 0b1010) != 0b1010) return false;
  if (p->E != DT_E || p->H != DT_E) return false;
  if (p->N < 1 || p->N > DT_MAXTEAMS * DT_NG) return false;
  if (p->Li < 1 || p->Lt < 1 || p->Li > DT_MAXL || p->Lt > DT_MAXL) return false;
  if (p->team_ws_floats < decode_team_ws_floats(p->N, p->B, p->T, p->Li, p->Lt)) return false;
  int dev = 0, coop = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return coop && sms >= DT_MAXTEAMS * DT_CG;
}

// Returns V2F_ERR_UNSUPPORTED (and launches nothing) outside the envelope: the caller falls back to
// decode_persist.cu / the step-per-launch path.
int decode_team_fwd(const v2f_decode_params* p, cudaStream_t s) {
  if (!dt_supported(p)) return V2F_ERR_UNSUPPORTED;
  constexpr int E = DT_E, H = DT_E;
  const int N = p->N, B = p->B, T = p->T, Li = p->Li, Lt = p->Lt;
  const int teams = (N + DT_NG - 1) / DT_NG, Np = teams * DT_NG;
  const DtLayout l = dt_layout(N, B, T, Li, Lt);
  float* ws = p->team_ws;
  // W' = W_ihc W_me [3H,E] (fp32, in persist_ws like decode_persist.cu), b' = W_ihc b_me + b_ih
  float* Wp = p->persist_ws;
  float* WmeT = Wp + (long long)3 * H * E + 3 * H;
  float* bp = ws + l.bp;
  V2F_TRY(v2f_transpose(E, E, p->W_me, E, 1, WmeT, E, 1, (void*)s));
  V2F_TRY(v2f_gemm_tc(1, 3 * H, E, E, p->W_ihc, E, WmeT, E, Wp, E, nullptr, 0.f, 4, 1, (void*)s));
  V2F_TRY(v2f_gemm_f32(0, 1, 1, 3 * H, E, p->b_me, E, 0, p->W_ihc, E, 0, bp, 3 * H, 0, 1, p->b_ih, 0.f, 0, (void*)s));
  __nv_bfloat16* wpk = reinterpret_cast<__nv_bfloat16*>(ws + l.wpk);
  dt_pack_weights_kernel<<<DT_CG * DT_WROWS, 128, 0, s>>>(Wp, p->We_mm, p->Wcat, wpk);
  V2F_CHECK_LAUNCH();
  // bf16 copies of the streamed tiles
  __nv_bfloat16* tb = reinterpret_cast<__nv_bfloat16*>(ws + l.tiles);
  const long long ni = (long long)B * Li * E, nt = (long long)B * Lt * E;
  __nv_bfloat16 *Himg_b = tb, *Vimg_b = tb + ni, *Htr_b = tb + 2 * ni, *Ptr_b = tb + 2 * ni + nt;
  V2F_TRY(v2f_cast_bf16(ni, p->Himg, Himg_b, (void*)s));
  if (p->Vimg != p->Himg) V2F_TRY(v2f_cast_bf16(ni, p->Vimg, Vimg_b, (void*)s));
  else Vimg_b = Himg_b;
  V2F_TRY(v2f_cast_bf16(nt, p->Htr, Htr_b, (void*)s));
  V2F_TRY(v2f_cast_bf16(nt, p->Ptr, Ptr_b, (void*)s));
  DtArgs a;
  a.p = *p;
  a.bp = bp;
  a.Himg = Himg_b;
  a.Vimg = Vimg_b;
  a.Htr = Htr_b;
  a.Ptr = Ptr_b;
  a.hb = reinterpret_cast<__nv_bfloat16*>(ws + l.hb);
  a.Cb = reinterpret_cast<__nv_bfloat16*>(ws + l.cb);
  a.Ub = reinterpret_cast<__nv_bfloat16*>(ws + l.ub);
  a.ypart = ws + l.ypart;
  a.bar = reinterpret_cast<unsigned*>(ws + l.bar);
  a.stamps = g_dt_stamps ? reinterpret_cast<unsigned long long*>(ws + l.stamps) : nullptr;
  a.Np = Np;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("V2F_TEAM_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    a.dbg = dbg;
  }
  // pad rows of the B operands are zero; valid rows are rewritten every step
  cudaMemsetAsync(ws + l.hb, 0, sizeof(float) * (size_t)(l.ypart - l.hb), s);
  cudaMemsetAsync(a.bar, 0, sizeof(unsigned) * DT_MAXTEAMS * DT_BARW, s);
  CUtensorMap mW, mH, mC, mU;
  V2F_TRY(tc_make_map(&mW, 0, wpk, (long long)DT_CG * DT_WROWS, E, E, 1, 0, DT_WROWS));
  V2F_TRY(tc_make_map(&mH, 0, a.hb, Np, E, E, 1, 0, DT_NG));
  V2F_TRY(tc_make_map(&mC, 0, a.Cb, 2LL * Np, E, E, 1, 0, DT_NG));
  V2F_TRY(tc_make_map(&mU, 0, a.Ub, Np, E, E, 1, 0, DT_NG));
  static bool attr = false;
  const size_t smem = dt_smem();
  if (!attr) {
    if (cudaFuncSetAttribute(decode_team_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return V2F_ERR_LAUNCH;
    attr = true;
  }
  void* params[] = {(void*)&mW, (void*)&mH, (void*)&mC, (void*)&mU, (void*)&a};
  prof_begin(V2F_K_DECODE_PERSIST_FWD, s);
  const cudaError_t e = cudaLaunchCooperativeKernel((void*)decode_team_fwd_kernel, dim3(teams * DT_CG), dim3(DT_THREADS),
                                                    params, smem, s);
  prof_end(V2F_K_DECODE_PERSIST_FWD, s);
  if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
    cudaGetLastError();
    return V2F_ERR_UNSUPPORTED;
  }
  if (e != cudaSuccess) return V2F_ERR_LAUNCH;
  ++g_v2f_launches;
  // CTX = U W_me^T + b_me for all T*N rows (saved for the backward: dW_ihc = DGI^T CTX)
  V2F_TRY(v2f_gemm_tc(1, T * N, E, E, p->U, E, p->W_me, E, p->CTX, E, p->b_me, 0.f, 4, 1, (void*)s));
  return V2F_OK;
}


// ================================================================================================ backward
// Row-team persistent BPTT: ONE cooperative launch walks the T steps backwards (CrossAttnRNN210.py:191-225 reversed).
// Same teams / ownership as the forward.  The three per-step products are the transposed ones,
//   dU  = DGI  W'                 [64 x 1536] x [1536 x 512]
//   dC += DHC  We_mm              [128 x 512] x [512 x 512]
//   dh += DS   Wcat               [64 x 3072] x [3072 x 512]
// computed output-stationary: CTA c owns output columns 8c..8c+7 of each, i.e. 8 ROWS of W'^T | We_mm^T | Wcat^T
// (8 x 5120 bf16 = 80 KB, resident for the whole launch as 80 swizzled [8 x 128 B] chunks).  The activations are the A
// operand now (M = 128 tile: the team's 64 rows + 64 don't-care rows, TMA-loaded in 64-wide K chunks through the 8-slot
// ring), the weight chunk the B operand (N = 16: 8 rows + 8 don't-care), accumulator [row, column] in TMEM: a thread
// of warps 0/1 reads its row's 8 values directly.  Row-local phases (gate backward, multimodal-attention backward, the
// attention backward sweep over the bf16 tiles of the forward) run in the row's owner CTA; dw_att / dMst / dHMst
// accumulate in the owner's registers over all steps and are written once.  Saved per-step tensors (DScat, DGI, DHC,
// DC, DE, DYH) are the ones the weight-gradient GEMMs and tilegrad_kernel after the loop consume (rnn_decode.cu).
struct DtbArgs {
  v2f_decode_params p;
  const __nv_bfloat16 *Himg, *Vimg, *Htr, *Ptr;     // bf16 tiles left in team_ws by the forward
  __nv_bfloat16 *DGIb, *DSb, *DHCb;                 // [2][Np,1536], [2][Np,3072] (double-buffered by step parity), [2 mod][Np,512]
  float* dxpart;                                    // [64, Np] per-CTA partial d x_t
  unsigned* bar;
  unsigned long long* stamps;
  int Np;
};

constexpr int DTB_KG = 3 * DT_E / 64, DTB_KC = DT_E / 64, DTB_KS = 6 * DT_E / 64;   // K chunks: 24, 8, 48
constexpr int DTB_W2 = 0, DTB_W4 = DTB_KG, DTB_W6 = DTB_KG + DTB_KC;               // first weight chunk of each product
constexpr int DTB_WCH = DTB_KG + DTB_KC + DTB_KS;                                   // 80 chunks of 1 KB
constexpr uint32_t DTB_TMEM_COLS = 64;                                              // D2: 0, D4: 16 / 32, D6: 48

__device__ __forceinline__ void dt_tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int NV>
__device__ __forceinline__ void dt_block_sum(float (&v)[NV], float* red, int tid) {   // over the 512 consumer threads
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = warp_sum(v[i]);
  named_bar_sync(3, DT_CONS);
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; i++) red[warp * NV + i] = v[i];
  named_bar_sync(3, DT_CONS);
#pragma unroll
  for (int i = 0; i < NV; i++) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < DT_CONS / 32; w++) t += red[w * NV + i];
    v[i] = t;
  }
}

__global__ void __launch_bounds__(DT_THREADS, 1)
decode_team_bwd_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapG,
                       const __grid_constant__ CUtensorMap mapS, const __grid_constant__ CUtensorMap mapD,
                       const __grid_constant__ DtbArgs a) {
  constexpr int E = DT_E, H = DT_E, ldS = 6 * DT_E;
  extern __shared__ uint8_t raw[];
  const v2f_decode_params& p = a.p;
  const int N = p.N, T = p.T, Li = p.Li, Lt = p.Lt, Wn = p.W, Np = a.Np;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int team = blockIdx.x / DT_CG, c = blockIdx.x % DT_CG;
  const int n0 = team * DT_NG;
  const int n_own = n0 + c;
  const bool own = n_own < N;

  uint8_t* sm = smem_align(raw, 1024);
  uint8_t* Wsm = sm;                                            // [80 chunks][8 rows][128 B]
  uint8_t* ring = sm + DT_WBYTES;                               // A ring (products) / tile ring (sweep)
  float* stage = reinterpret_cast<float*>(ring + DT_RING);      // [64][9] dh columns of the team's rows; 16 KB slack before it is read by MMA overrun
  float* dcv = stage + 64 * 9 + 64;                             // [2][512] d contexts of the own row
  float* pacc = dcv + 2 * E;                                    // [2 groups][2 modalities][2][512]: (sacc, wacc) partials
  float* de_all = pacc + 8 * E;                                 // [2][DT_MAXL]
  float* red = de_all + 2 * DT_MAXL;                            // [16 * 4]
  uint64_t* bfull = reinterpret_cast<uint64_t*>(red + 64);
  uint64_t* bempty = bfull + DT_NBS;
  uint64_t* tfull = bempty + DT_NBS;
  uint64_t* tempty = tfull + DT_NBS;
  uint64_t* mma_done = tempty + DT_NBS;
  uint64_t* wfull = mma_done + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wfull + 1);

  if (tid == 0) {
    for (int s = 0; s < DT_NBS; s++) {
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], 1);
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], DT_GRP / 32);
    }
    mbar_init(mma_done, 1);
    mbar_init(wfull, 1);
    mbar_fence_init();
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(DTB_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;
  if (warp == 16 && lane == 0) {                    // the transposed weight slice: 80 boxes of [8 rows x 128 B]
    mbar_expect_tx(wfull, DT_WBYTES);
    for (int kc = 0; kc < DTB_WCH; kc++) tma_load_3d(Wsm + kc * 1024, &mapW, wfull, kc * 64, c * 8, 0);
  }

  const __nv_bfloat16* Vimg_b = a.Vimg;
  const bool byproj = p.variant == 2;
  const int mod_mask = p.mod_mask;
  const unsigned* mask_dev = p.y ? p.tf_mask_dev : nullptr;
  unsigned* bar = a.bar + team * DT_BARW;
  unsigned epoch = 0;
  uint32_t bq = 0, md = 0, tq = 0;
  const int grp = warp < 8 ? 0 : (warp < 16 ? 1 : warp - 16);
  const int gt = tid & (DT_GRP - 1), gw = (tid >> 5) & 7;
  const int cpi = (Li + DT_CH - 1) / DT_CH, cpt = (Lt + DT_CH - 1) / DT_CH, ctot = cpi + cpt;
  const uint32_t idesc = umma_idesc<0>(16);
  auto stamp = [&](int t, int k) {
    if (a.stamps && blockIdx.x == 0 && tid == 0) a.stamps[t * DT_STAMPS + k] = dt_globaltimer();
  };
  // chunk i of a product phase (ring sequence q = bq + i) is issued by lane 0 of consumer warp i % 8 = its ring slot (one
  // issuing thread sustains only one copy per ~600 cycles; see the forward kernel)
  auto load_chunk = [&](const CUtensorMap* map, int row0, int kc, uint32_t q) {
    const uint32_t s = q % DT_NBS, r = q / DT_NBS;
    mbar_wait(&bempty[s], (r & 1) ^ 1);
    mbar_expect_tx(&bfull[s], DT_BSLOT);
    tma_load_3d(ring + s * DT_BSLOT, map, &bfull[s], kc * 64, row0, 0);
  };
  auto issue = [&](int wch0, uint32_t dcol, int nch, uint32_t q0, bool last) {
    for (int i = 0; i < nch; i++) {
      const uint32_t q = q0 + (uint32_t)i, s = q % DT_NBS, r = q / DT_NBS;
      mbar_wait(&bfull[s], r & 1);
      tc_fence_after();
      const uint64_t ad = umma_desc_sw128(smem_u32(ring + s * DT_BSLOT));
      const uint64_t bd = umma_desc_sw128(smem_u32(Wsm + (wch0 + i) * 1024));
#pragma unroll
      for (int k = 0; k < 4; k++) umma<0>(tmem_d + dcol, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (i | k) ? 1u : 0u);
      umma_commit(&bempty[s]);
    }
    if (last) umma_commit(mma_done);
  };

  // gate threads: (row nl of the team, hidden unit gu of this CTA); d h lives in a register across the steps
  const int nl = tid >> 3, gu = tid & 7, ng = n0 + nl, uu = 8 * c + gu;
  const bool gact = tid < DT_CONS && ng < N;
  float dhreg = gact ? p.dh[(long long)ng * H + uu] : 0.f;
  // owner accumulators over all steps (thread = column x of the own row)
  float dw0 = 0.f, dw1 = 0.f, dw2 = 0.f, dM0 = 0.f, dM1 = 0.f, dHM0 = 0.f, dHM1 = 0.f;
  if (warp == 17 && lane == 0) mbar_wait(wfull, 0);

  for (int t = T - 1; t >= 0; t--) {
    const int ts = T - 1 - t;                       // stamp row
    const int par = t & 1;
    const float* S = p.S_all + (long long)t * N * ldS;
    float* DS = p.DScat + (long long)t * N * ldS;
    float* DHC = p.DHC + (long long)t * N * 2 * E;
    float* DC = p.DC + (long long)t * N * 2 * E;
    __nv_bfloat16* DGIb = a.DGIb + (long long)par * Np * 3 * H;
    __nv_bfloat16* DSb = a.DSb + (long long)par * Np * ldS;
    stamp(ts, 0);
    // ================================================================ A: GRU gate backward (own hidden units, the team's rows)
    float dhdir = 0.f;
    if (tid < DT_CONS) {
      float dx = 0.f;
      if (gact) {
        const int forced = (t == T - 1) ? 1 : (mask_dev ? (int)((*mask_dev >> t) & 1u) : (int)((p.tf_mask >> t) & 1u));
        const float dyh = p.dY[(long long)ng * T + t] + (forced ? 0.f : __ldcg(p.dxn + ng));
        const float* rzn = p.RZN + ((long long)t * N + ng) * 3 * H;
        const float r = rzn[uu], z = rzn[H + uu], cc = rzn[2 * H + uu];
        const float hp = p.h_all[((long long)t * N + ng) * H + uu];
        const float ghn = S[(long long)ng * ldS + 3 * E + 2 * H + uu];
        const float dhp = dhreg + dyh * p.w_fc[uu];
        const float dc_ = dhp * (1.f - z);
        const float dz = dhp * (hp - cc);
        const float dan = dc_ * (1.f - cc * cc);
        const float dar = dan * ghn * r * (1.f - r);
        const float daz = dz * z * (1.f - z);
        float* dgi = p.DGI + ((long long)t * N + ng) * 3 * H;
        dgi[uu] = dar;
        dgi[H + uu] = daz;
        dgi[2 * H + uu] = dan;
        float* dgh = DS + (long long)ng * ldS + 3 * E;
        dgh[uu] = dar;
        dgh[H + uu] = daz;
        dgh[2 * H + uu] = dan * r;
        __nv_bfloat16* gb = DGIb + (long long)ng * 3 * H;
        gb[uu] = __float2bfloat16_rn(dar);
        gb[H + uu] = __float2bfloat16_rn(daz);
        gb[2 * H + uu] = __float2bfloat16_rn(dan);
        __nv_bfloat16* sb = DSb + (long long)ng * ldS + 3 * E;
        sb[uu] = __float2bfloat16_rn(dar);
        sb[H + uu] = __float2bfloat16_rn(daz);
        sb[2 * H + uu] = __float2bfloat16_rn(dan * r);
        dhdir = dhp * z;
        dx = dar * p.w_x[uu] + daz * p.w_x[H + uu] + dan * p.w_x[2 * H + uu];
        if (c == 0 && gu == 0) p.DYH[(long long)t * N + ng] = dyh;
      }
      dx += __shfl_xor_sync(FULL, dx, 1);
      dx += __shfl_xor_sync(FULL, dx, 2);
      dx += __shfl_xor_sync(FULL, dx, 4);
      if (gu == 0 && gact) a.dxpart[(long long)c * Np + ng] = dx;
    }
    stamp(ts, 1);
    dt_team_barrier(bar, c, epoch);
    stamp(ts, 2);
    // ================================================================ B: dU[:, 8c..8c+8) = DGI W'[:, 8c..]
    if (warp == 17) {
      if (lane == 0) issue(DTB_W2, 0, DTB_KG, bq, true);
      __syncwarp();
    } else if (warp < 16) {
      if (warp < DT_NBS) {
        if (lane == 0)
          for (int i = warp; i < DTB_KG; i += DT_NBS) load_chunk(&mapG, par * Np + n0, i, bq + (uint32_t)i);
        __syncwarp();
      }
      if (warp < 2) {
        mbar_wait(mma_done, md & 1);
        tc_fence_after();
        uint32_t v[8];
        dt_tmem_ld8(tmem_d + ((uint32_t)(warp * 32) << 16), v);
        tc_fence_before();
        const int n = n0 + warp * 32 + lane;
        if (n < N) {
          float* dst = p.dU + (long long)n * E + 8 * c;
          st4(dst, make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])));
          st4(dst + 4, make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7])));
        }
      }
      md++;
    }
    bq += DTB_KG;
    stamp(ts, 3);
    dt_team_barrier(bar, c, epoch);
    stamp(ts, 4);
    // ================================================================ C: multimodal attention backward of the own row
    if (own && tid < DT_CONS) {
      const int n = n_own, b = n / Wn, x = tid;
      const float* Cc = p.C + ((long long)t * N + n) * 2 * E;
      const float* HCc = p.HC + ((long long)t * N + n) * 2 * E;
      float mk[4], hk[4], al[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const bool on = (mod_mask >> k) & 1;
        mk[k] = !on ? 0.f : ((k & 1) ? Cc[(k >> 1) * E + x] : p.Mst[((long long)b * 2 + (k >> 1)) * E + x]);
        hk[k] = !on ? 0.f : ((k & 1) ? HCc[(k >> 1) * E + x] : p.HMst[((long long)b * 2 + (k >> 1)) * E + x]);
        al[k] = p.alpha_mm[((long long)t * N + n) * 4 + k];
      }
      const float d = __ldcg(p.dU + (long long)n * E + x);
      float da[4];
#pragma unroll
      for (int k = 0; k < 4; k++) da[k] = ((mod_mask >> k) & 1) ? d * (byproj ? hk[k] : mk[k]) : 0.f;
      dt_block_sum<4>(da, red, tid);
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < 4; k++) dot = fmaf(al[k], da[k], dot);
      const float sx = S[(long long)n * ldS + 2 * E + x], wx = p.w_att[2 * E + x];
      float ds = 0.f, dw = 0.f;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float dhm = 0.f, dm = 0.f;
        if ((mod_mask >> k) & 1) {
          const float de = al[k] * (da[k] - dot);
          const float q = tanh_acc(hk[k] + sx);
          const float dpre = de * wx * (1.f - q * q);
          ds += dpre;
          dw = fmaf(de, q, dw);
          dhm = dpre + (byproj ? al[k] * d : 0.f);
          dm = byproj ? d : d * (1.f + al[k]);
        }
        if (k & 1) {
          const long long dyn = ((long long)n * 2 + (k >> 1)) * E + x;
          DHC[dyn] = dhm;
          DC[dyn] = dm;
          a.DHCb[((long long)(k >> 1) * Np + n) * E + x] = __float2bfloat16_rn(dhm);
        } else if (k == 0) {
          dHM0 += dhm;
          dM0 += dm;
        } else {
          dHM1 += dhm;
          dM1 += dm;
        }
      }
      DS[(long long)n * ldS + 2 * E + x] = ds;
      DSb[(long long)n * ldS + 2 * E + x] = __float2bfloat16_rn(ds);
      dw2 += dw;
      if (warp == 15) {       // d x_t of the own row = sum of the per-CTA partials (consumed by phase A of step t-1)
        float v = __ldcg(a.dxpart + (long long)lane * Np + n) + __ldcg(a.dxpart + (long long)(lane + 32) * Np + n);
        v = warp_sum(v);
        if (lane == 0) p.dxn[n] = v;
      }
    }
    stamp(ts, 5);
    dt_team_barrier(bar, c, epoch);
    stamp(ts, 6);
    // ================================================================ D: dC[:, :, 8c..8c+8) += DHC We_mm[:, 8c..]
    if (warp == 17) {
      if (lane == 0) {
        issue(DTB_W4, 16, DTB_KC, bq, false);
        issue(DTB_W4, 32, DTB_KC, bq + DTB_KC, true);
      }
      __syncwarp();
    } else if (warp < 16) {
      if (warp < DT_NBS) {             // 8 image + 8 trend chunks
        if (lane == 0) {
          load_chunk(&mapD, n0, warp, bq + (uint32_t)warp);
          load_chunk(&mapD, Np + n0, warp, bq + (uint32_t)(warp + DTB_KC));
        }
        __syncwarp();
      }
      if (warp < 2) {
        mbar_wait(mma_done, md & 1);
        tc_fence_after();
        const int n = n0 + warp * 32 + lane;
#pragma unroll
        for (int mod = 0; mod < 2; mod++) {
          uint32_t v[8];
          dt_tmem_ld8(tmem_d + ((uint32_t)(warp * 32) << 16) + 16u * (mod + 1), v);
          if (n < N) {
            float* dst = DC + ((long long)n * 2 + mod) * E + 8 * c;
            float4 o0 = __ldcg(reinterpret_cast<const float4*>(dst)), o1 = __ldcg(reinterpret_cast<const float4*>(dst + 4));
            o0.x += __uint_as_float(v[0]);
            o0.y += __uint_as_float(v[1]);
            o0.z += __uint_as_float(v[2]);
            o0.w += __uint_as_float(v[3]);
            o1.x += __uint_as_float(v[4]);
            o1.y += __uint_as_float(v[5]);
            o1.z += __uint_as_float(v[6]);
            o1.w += __uint_as_float(v[7]);
            st4(dst, o0);
            st4(dst + 4, o1);
          }
        }
        tc_fence_before();
      }
      md++;
    }
    bq += 2 * DTB_KC;
    stamp(ts, 7);
    dt_team_barrier(bar, c, epoch);
    stamp(ts, 8);
    // ================================================================ E: attention backward of the own row (two passes over the bf16 tiles)
    if (own) {
      const int n = n_own, b = n / Wn;
      const float* al_img = p.alpha_img + ((long long)t * N + n) * Li;
      const float* al_tr = p.alpha_tr + ((long long)t * N + n) * Lt;
      if (warp == 16) {
        if (lane == 0) {
          uint32_t q = tq;
          for (int pass = 0; pass < 2; pass++)
            for (int cc0 = 0; cc0 < ctot; cc0++, q++) {
              const int mod = cc0 >= cpi, cc = mod ? cc0 - cpi : cc0;
              const int L = mod ? Lt : Li, j0 = cc * DT_CH, nj = min(DT_CH, L - j0);
              const long long off = ((long long)b * L + j0) * E;
              const uint32_t bytes = (uint32_t)nj * E * 2u;
              const uint32_t sl = q % DT_NTS, r = q / DT_NTS;
              const __nv_bfloat16* src = pass == 0 ? (mod ? a.Ptr : Vimg_b) : (mod ? a.Htr : a.Himg);
              mbar_wait(&tempty[sl], (r & 1) ^ 1);
              mbar_expect_tx(&tfull[sl], bytes);
              bulk_g2s(ring + sl * DT_TSLOT, src + off, bytes, &tfull[sl]);
            }
        }
        __syncwarp();
      } else if (warp < 16) {
        const int x = tid;
        // d contexts of the own row (complete after phase D) and dot_m = dc_m . (c_m - b_tl)
        float dots[2];
        {
          const float* Cc = p.C + ((long long)t * N + n) * 2 * E;
#pragma unroll
          for (int mod = 0; mod < 2; mod++) {
            const float dcx = __ldcg(DC + ((long long)n * 2 + mod) * E + x);
            dcv[mod * E + x] = dcx;
            dots[mod] = dcx * (Cc[mod * E + x] - (mod ? p.b_tl[x] : 0.f));
          }
          dt_block_sum<2>(dots, red, tid);     // (its barriers also publish dcv)
        }
        // ---- pass 1 (V tiles): d alpha_j = V_j . dc ; de_j = alpha_j (d alpha_j - dot)
        {
          float4 dreg[4];
          int cur_mod = -1;
          for (int cc0 = grp; cc0 < ctot; cc0 += 2) {
            const int mod = cc0 >= cpi, cc = mod ? cc0 - cpi : cc0;
            const int L = mod ? Lt : Li, j0 = cc * DT_CH, nj = min(DT_CH, L - j0);
            if (mod != cur_mod) {
              cur_mod = mod;
#pragma unroll
              for (int k = 0; k < 2; k++) {
                dreg[2 * k] = ld4(dcv + mod * E + 8 * lane + 256 * k);
                dreg[2 * k + 1] = ld4(dcv + mod * E + 8 * lane + 256 * k + 4);
              }
            }
            const uint32_t q = tq + (uint32_t)cc0, sl = q % DT_NBS, r = q / DT_NBS;
            const float alv = gw < nj ? (mod ? al_tr : al_img)[j0 + gw] : 0.f;
            mbar_wait(&tfull[sl], r & 1);
            if (gw < nj) {
              const uint8_t* vp = ring + sl * DT_TSLOT + gw * (E * 2);
              float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
              for (int k = 0; k < 2; k++) {
                const uint4 vv = *reinterpret_cast<const uint4*>(vp + (8 * lane + 256 * k) * 2);
                const float2 v0 = dt_bf2(vv.x), v1 = dt_bf2(vv.y), v2 = dt_bf2(vv.z), v3 = dt_bf2(vv.w);
                const float4 d0 = dreg[2 * k], d1 = dreg[2 * k + 1];
                acc0 = fmaf(v0.x, d0.x, acc0);
                acc1 = fmaf(v0.y, d0.y, acc1);
                acc0 = fmaf(v1.x, d0.z, acc0);
                acc1 = fmaf(v1.y, d0.w, acc1);
                acc0 = fmaf(v2.x, d1.x, acc0);
                acc1 = fmaf(v2.y, d1.y, acc1);
                acc0 = fmaf(v3.x, d1.z, acc0);
                acc1 = fmaf(v3.y, d1.w, acc1);
              }
              const float de = alv * (warp_sum(acc0 + acc1) - (mod ? dots[1] : dots[0]));
              if (lane == 0) {
                de_all[mod * DT_MAXL + j0 + gw] = de;
                (mod ? p.DE_tr + ((long long)t * N + n) * Lt : p.DE_img + ((long long)t * N + n) * Li)[j0 + gw] = de;
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[sl]);
          }
        }
        named_bar_sync(3, DT_CONS);
        // ---- pass 2 (H tiles): ds[x] = w[x] sum_j de_j (1 - q^2), dw[x] += sum_j de_j q, q = tanh(H_j[x] + s[x])
        {
          float sa_i0 = 0.f, sa_i1 = 0.f, wa_i0 = 0.f, wa_i1 = 0.f, sa_t0 = 0.f, sa_t1 = 0.f, wa_t0 = 0.f, wa_t1 = 0.f;
          float s0 = 0.f, s1 = 0.f;
          int cur_mod = -1;
          for (int cc0 = grp; cc0 < ctot; cc0 += 2) {
            const int mod = cc0 >= cpi, cc = mod ? cc0 - cpi : cc0;
            const int L = mod ? Lt : Li, j0 = cc * DT_CH, nj = min(DT_CH, L - j0);
            if (mod != cur_mod) {
              cur_mod = mod;
              const float2 sv = *reinterpret_cast<const float2*>(S + (long long)n * ldS + mod * E + 2 * gt);
              s0 = sv.x;
              s1 = sv.y;
            }
            const uint32_t q = tq + (uint32_t)(ctot + cc0), sl = q % DT_NBS, r = q / DT_NBS;
            const float* dep = de_all + mod * DT_MAXL + j0;
            mbar_wait(&tfull[sl], r & 1);
            const uint8_t* hp = ring + sl * DT_TSLOT + gt * 4;
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
            for (int j = 0; j < nj; j++) {
              const float2 hv = dt_bf2(*reinterpret_cast<const uint32_t*>(hp + j * (E * 2)));
              const float de = dep[j];
              const float q0 = tanh_fast<true>(hv.x + s0), q1 = tanh_fast<true>(hv.y + s1);
              a0 = fmaf(de, fmaf(-q0, q0, 1.f), a0);
              a1 = fmaf(de, fmaf(-q1, q1, 1.f), a1);
              b0 = fmaf(de, q0, b0);
              b1 = fmaf(de, q1, b1);
            }
            if (mod) {
              sa_t0 += a0;
              sa_t1 += a1;
              wa_t0 += b0;
              wa_t1 += b1;
            } else {
              sa_i0 += a0;
              sa_i1 += a1;
              wa_i0 += b0;
              wa_i1 += b1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[sl]);
          }
          float* pp = pacc + grp * 4 * E;         // [mod][sacc | wacc][512]
          *reinterpret_cast<float2*>(pp + 0 * E + 2 * gt) = make_float2(sa_i0, sa_i1);
          *reinterpret_cast<float2*>(pp + 1 * E + 2 * gt) = make_float2(wa_i0, wa_i1);
          *reinterpret_cast<float2*>(pp + 2 * E + 2 * gt) = make_float2(sa_t0, sa_t1);
          *reinterpret_cast<float2*>(pp + 3 * E + 2 * gt) = make_float2(wa_t0, wa_t1);
        }
        named_bar_sync(3, DT_CONS);
        {
#pragma unroll
          for (int mod = 0; mod < 2; mod++) {
            const float sacc = pacc[(2 * mod) * E + x] + pacc[(4 + 2 * mod) * E + x];
            const float wacc = pacc[(2 * mod + 1) * E + x] + pacc[(4 + 2 * mod + 1) * E + x];
            const float ds = sacc * p.w_att[mod * E + x];
            DS[(long long)n * ldS + mod * E + x] = ds;
            DSb[(long long)n * ldS + mod * E + x] = __float2bfloat16_rn(ds);
            if (mod) dw1 += wacc;
            else dw0 += wacc;
          }
        }
      }
      tq += 2u * (uint32_t)ctot;
    }
    stamp(ts, 9);
    dt_team_barrier(bar, c, epoch);
    stamp(ts, 10);
    // ================================================================ F: dh[:, 8c..8c+8) = z-path + DS Wcat[:, 8c..]
    if (warp == 17) {
      if (lane == 0) issue(DTB_W6, 48, DTB_KS, bq, true);
      __syncwarp();
    } else if (warp < 16) {
      if (warp < DT_NBS) {
        if (lane == 0)
          for (int i = warp; i < DTB_KS; i += DT_NBS) load_chunk(&mapS, par * Np + n0, i, bq + (uint32_t)i);
        __syncwarp();
      }
      if (warp < 2) {
        mbar_wait(mma_done, md & 1);
        tc_fence_after();
        uint32_t v[8];
        dt_tmem_ld8(tmem_d + ((uint32_t)(warp * 32) << 16) + 48u, v);
        tc_fence_before();
        const int r = warp * 32 + lane;
#pragma unroll
        for (int i = 0; i < 8; i++) stage[r * 9 + i] = __uint_as_float(v[i]);
      }
      md++;
      named_bar_sync(3, DT_CONS);
      dhreg = dhdir + stage[nl * 9 + gu];
      named_bar_sync(3, DT_CONS);          // stage is rewritten by the next step's phase F only, but keep the readers together
    }
    bq += DTB_KS;
    stamp(ts, 11);
  }
  // ------------------------------------------------------------------ outputs that accumulate over the steps
  if (gact) p.dh[(long long)ng * H + uu] = dhreg;
  if (own && tid < DT_CONS) {
    const int n = n_own, x = tid;
    p.dw_acc[((long long)n * 3 + 0) * E + x] = dw0;
    p.dw_acc[((long long)n * 3 + 1) * E + x] = dw1;
    p.dw_acc[((long long)n * 3 + 2) * E + x] = dw2;
    p.dMst_acc[((long long)n * 2 + 0) * E + x] = dM0;
    p.dMst_acc[((long long)n * 2 + 1) * E + x] = dM1;
    p.dHMst_acc[((long long)n * 2 + 0) * E + x] = dHM0;
    p.dHMst_acc[((long long)n * 2 + 1) * E + x] = dHM1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(DTB_TMEM_COLS) : "memory");
  }
}

// the per-CTA slices of the transposed weights, bf16, row-major [64 CTAs x 8 rows][5120]: row 8c+i holds column 8c+i of
// W' (k < 1536), of We_mm (k < 2048) and of Wcat (rest).  32 x 32 tiles through shared memory: coalesced both ways.
__global__ void dt_pack_weights_t_kernel(const float* __restrict__ Wp, const float* __restrict__ We_mm,
                                         const float* __restrict__ Wcat, __nv_bfloat16* __restrict__ out) {
  constexpr int E = DT_E, KT = 10 * DT_E;
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, x0 = blockIdx.y * 32;      // source rows k0.., source columns x0..
  const float* src = k0 < 3 * E ? Wp + (long long)k0 * E : (k0 < 4 * E ? We_mm + (long long)(k0 - 3 * E) * E : Wcat + (long long)(k0 - 4 * E) * E);
  for (int i = threadIdx.y; i < 32; i += 8) tile[i][threadIdx.x] = src[(long long)i * E + x0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) out[(long long)(x0 + i) * KT + k0 + threadIdx.x] = __float2bfloat16_rn(tile[threadIdx.x][i]);
}

static size_t dtb_smem() {
  return 1024 + (size_t)DT_WBYTES + DT_RING + sizeof(float) * (64 * 9 + 64 + 2 * DT_E + 8 * DT_E + 2 * DT_MAXL + 64) +
         8 * (4 * DT_NBS + 2) + 16;
}

struct DtbLayout {
  long long wt, dgib, dsb, dhcb, dxpart, bar, stamps, end;   // float offsets into the backward scratch (p->ws)
};
static DtbLayout dtb_layout(int N, int T) {
  const long long Np = (long long)((N + DT_NG - 1) / DT_NG) * DT_NG;
  DtbLayout l;
  long long o = 0;
  auto take = [&](long long floats) { const long long at = o; o += (floats + 255) / 256 * 256; return at; };
  l.wt = take((long long)DT_E * 10 * DT_E / 2);
  l.dgib = take(2 * Np * 3 * DT_E / 2);
  l.dsb = take(2 * Np * 6 * DT_E / 2);
  l.dhcb = take(2 * Np * DT_E / 2);
  l.dxpart = take((long long)DT_CG * Np);
  l.bar = take((long long)DT_MAXTEAMS * DT_BARW);
  l.stamps = take(2LL * T * DT_STAMPS);
  l.end = o;
  return l;
}
long long decode_team_bwd_ws_floats(int N, int T) { return dtb_layout(N, T).end; }

static bool g_dtb_enabled = true;

// The loop part of v2f_decode_bwd (rnn_decode.cu) as one launch; V2F_ERR_UNSUPPORTED outside the envelope.  Needs the
// forward to have run through decode_team_fwd (the bf16 tiles and W' are taken from team_ws / persist_ws).
int decode_team_bwd(const v2f_decode_params* p, float* bws, long long bws_floats, cudaStream_t s) {
  if (!g_dtb_enabled || !dt_supported(p) || !bws) return V2F_ERR_UNSUPPORTED;
  constexpr int E = DT_E, H = DT_E;
  const int N = p->N, B = p->B, T = p->T, Li = p->Li, Lt = p->Lt;
  const DtbLayout lb = dtb_layout(N, T);
  if (bws_floats < lb.end) return V2F_ERR_UNSUPPORTED;
  const int teams = (N + DT_NG - 1) / DT_NG, Np = teams * DT_NG;
  const DtLayout l = dt_layout(N, B, T, Li, Lt);
  const float* Wp = p->persist_ws;                       // W' = W_ihc W_me, left there by the forward
  __nv_bfloat16* wt = reinterpret_cast<__nv_bfloat16*>(bws + lb.wt);
  dt_pack_weights_t_kernel<<<dim3(10 * E / 32, E / 32), dim3(32, 8), 0, s>>>(Wp, p->We_mm, p->Wcat, wt);
  V2F_CHECK_LAUNCH();
  const __nv_bfloat16* tb = reinterpret_cast<const __nv_bfloat16*>(p->team_ws + l.tiles);
  const long long ni = (long long)B * Li * E, nt = (long long)B * Lt * E;
  DtbArgs a;
  a.p = *p;
  a.Himg = tb;
  a.Vimg = p->Vimg != p->Himg ? tb + ni : tb;
  a.Htr = tb + 2 * ni;
  a.Ptr = tb + 2 * ni + nt;
  a.DGIb = reinterpret_cast<__nv_bfloat16*>(bws + lb.dgib);
  a.DSb = reinterpret_cast<__nv_bfloat16*>(bws + lb.dsb);
  a.DHCb = reinterpret_cast<__nv_bfloat16*>(bws + lb.dhcb);
  a.dxpart = bws + lb.dxpart;
  a.bar = reinterpret_cast<unsigned*>(bws + lb.bar);
  a.stamps = g_dt_stamps ? reinterpret_cast<unsigned long long*>(bws + lb.stamps) : nullptr;
  a.Np = Np;
  cudaMemsetAsync(bws + lb.dgib, 0, sizeof(float) * (size_t)(lb.dxpart - lb.dgib), s);   // pad rows of the A operands
  cudaMemsetAsync(a.bar, 0, sizeof(unsigned) * DT_MAXTEAMS * DT_BARW, s);
  CUtensorMap mW, mG, mS, mD;
  V2F_TRY(tc_make_map(&mW, 0, wt, E, 10 * E, 10 * E, 1, 0, 8));
  V2F_TRY(tc_make_map(&mG, 0, a.DGIb, 2LL * Np, 3 * H, 3 * H, 1, 0, DT_NG));
  V2F_TRY(tc_make_map(&mS, 0, a.DSb, 2LL * Np, 6 * E, 6 * E, 1, 0, DT_NG));
  V2F_TRY(tc_make_map(&mD, 0, a.DHCb, 2LL * Np, E, E, 1, 0, DT_NG));
  static bool attr = false;
  const size_t smem = dtb_smem();
  if (!attr) {
    if (cudaFuncSetAttribute(decode_team_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return V2F_ERR_LAUNCH;
    attr = true;
  }
  void* params[] = {(void*)&mW, (void*)&mG, (void*)&mS, (void*)&mD, (void*)&a};
  prof_begin(V2F_K_DECODE_PERSIST_BWD, s);
  const cudaError_t e = cudaLaunchCooperativeKernel((void*)decode_team_bwd_kernel, dim3(teams * DT_CG), dim3(DT_THREADS),
                                                    params, smem, s);
  prof_end(V2F_K_DECODE_PERSIST_BWD, s);
  if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
    cudaGetLastError();
    return V2F_ERR_UNSUPPORTED;
  }
  if (e != cudaSuccess) return V2F_ERR_LAUNCH;
  ++g_v2f_launches;
  return V2F_OK;
}

}  // namespace v2f

extern "C" long long v2f_decode_team_ws_floats(int N, int B, int T, int Li, int Lt) {
  if (N <= 0 || B <= 0 || T <= 0 || Li <= 0 || Lt <= 0) return 0;
  return v2f::decode_team_ws_floats(N, B, T, Li, Lt);
}
// A/B switch (default 1): 0 routes the decoder through decode_persist.cu / the step-per-launch path again.
extern "C" int v2f_decode_team_enable(int on) {
  v2f::g_dt_enabled = on != 0;
  return V2F_OK;
}
extern "C" int v2f_decode_team_bwd_enable(int on) {
  v2f::g_dtb_enabled = on != 0;
  return V2F_OK;
}
extern "C" long long v2f_decode_team_bwd_ws_floats(int N, int T) {
  if (N <= 0 || T <= 0) return 0;
  return v2f::decode_team_bwd_ws_floats(N, T);
}
extern "C" int v2f_decode_team_stamps_enable(int on) {
  v2f::g_dt_stamps = on != 0;
  return V2F_OK;
}
// Byte offset of the stamp table [T, 16] (unsigned long long, ns) inside team_ws.
extern "C" long long v2f_decode_team_stamps_offset(int N, int B, int T, int Li, int Lt) {
  return (long long)sizeof(float) * v2f::dt_layout(N, B, T, Li, Lt).stamps;
}

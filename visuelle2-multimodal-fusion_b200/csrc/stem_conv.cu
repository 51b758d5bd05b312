// Stem convolution of the ResNet-101 trunk -- Conv2d(3, 64, kernel 7, stride 2, padding 3, no bias) -- as an
// implicit GEMM on tcgen05 tensor cores, with the BatchNorm batch statistics of its output produced by the epilogue.
// sm_100a.
//
// Reference arithmetic: torchvision resnet101().conv1 / bn1 inside ImageEncoder
// (/root/reference/models/CrossAttnRNN210.py:58-72; the stem is frozen there, :62-65, and the whole trunk runs in
// train() mode, so bn1 normalises with the batch statistics of this convolution's output).  Forward only.
//
// Why this one convolution is taken from cuDNN.  With 3 input channels cuDNN's implicit GEMM runs at ~40 TFLOP/s:
// 1.34 ms per 128-image step, 5 % of the whole training step for 0.5 % of its FLOPs, plus a 137 MB -> 68 MB cast of
// the images in front of it and a 368 MB statistics sweep behind it.  Here:
//   * GEMM view: D[pixel, cout] = sum_k A[pixel, k] W[cout, k], k = (kh, kw, cin) with the 21 values of one kernel
//     row padded to 24 (K = 7 x 24 = 168, + 8 zero columns = 11 k-steps of 16); M tile = 128 consecutive output
//     pixels of the flattened (n, oh, ow) order, N = 64, accumulator in TMEM (64 columns);
//   * the weights (64 x 192 bf16, packed by the caller) are written once per CTA into shared memory in the canonical
//     K-major 128-byte-swizzle layout and stay resident; a CTA walks a contiguous range of tiles;
//   * input rows (NHWC or NCHW -- the layout a DataLoader collates --, fp32 or bf16: no transposing / casting copy of
//     the images runs in front of the kernel) are staged once per CTA as bf16 in a 16-slot ring of whole rows with the zero
//     padding baked in (element e of the row at position e + 9, so that the 24-value window of output column ow starts
//     at the even position 6 ow); the rows the NEXT tile needs are fetched into registers while the current tile is
//     multiplied and written out;
//   * the A tile is gathered from the staged rows with conflict-free 32-bit shared loads (lanes = consecutive pixels,
//     stride 3 words) and written as 16-byte chunks into the swizzled layout, published to the async proxy, and
//     multiplied by 11 tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) issued by one thread;
//   * epilogue: tcgen05.ld (thread = pixel, 32 channels), round to bf16, 64-byte stores into the NHWC output, and the
//     per-channel sum / sum of squares of the ROUNDED values (what a statistics sweep over the stored tensor would see)
//     by a transposing butterfly over the warp (31 shuffles per statistic), accumulated per CTA over all its tiles and
//     written as one partial row per CTA in the layout bn_fwd_finalize_kernel merges (bn_act.cu).
// Two CTAs per SM: one CTA's gather overlaps the other's MMA / epilogue.
#include <cuda_bf16.h>

#include "tc.cuh"

namespace v2f {

constexpr int SC_THREADS = 256;
constexpr int SC_COUT = 64;
constexpr int SC_RUN = 24;                     // 7 kernel columns x 3 channels = 21 values, padded to 24
constexpr int SC_KSTEPS = 11;                  // 176 / 16
constexpr int SC_KPAD = 192;                   // 3 swizzle atoms of 64
constexpr int SC_A_ATOM = 128 * 128;           // bytes: 128 pixels x 64 bf16
constexpr int SC_B_ATOM = SC_COUT * 128;
constexpr int SC_A_BYTES = 3 * SC_A_ATOM;      // 48 KB
constexpr int SC_B_BYTES = 3 * SC_B_ATOM;      // 24 KB
constexpr int SC_NSLOT = 16;                   // staged input rows (9 live + 6 incoming at most)
constexpr int SC_PF_ROWS = 4;                  // rows prefetched into registers for the next tile
constexpr int SC_PF_J = 4;                     // elements per thread and row: 3 W <= 1024
constexpr uint32_t SC_TMEM_COLS = 64;

struct ScArgs {
  const void* x;                 // [N, H, W, 3] (NHWC) or [N, 3, H, W] (NCHW), fp32 or bf16
  const __nv_bfloat16* wpk;      // [64, 192] bf16: k = kh * 24 + kw * 3 + c, zero elsewhere
  __nv_bfloat16* y;              // [N * OH * OW, 64]
  float* part;                   // [grid, 2, 64] or null
  int N, H, W, OH, OW;
  long long P;                   // N * OH * OW
  long long ntiles;
  int rowlen;                    // bf16 elements per staged row
};

// byte offset of the 16-byte chunk j (0..7) of row r inside a K-major SWIZZLE_128B atom column (rows x 128 B)
__device__ __forceinline__ uint32_t sc_swz(int r, int j) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
}

template <bool IN_BF16>
__device__ __forceinline__ float sc_load(const void* x, long long idx) {
  if (IN_BF16) return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(x) + idx));
  return __ldg(reinterpret_cast<const float*>(x) + idx);
}

// A thread's share of one input row: SC_PF_J elements.  NHWC: element e = tid + 256 j of the row's 3 W contiguous
// values, staged at position e.  NCHW: the same e addresses (channel e / W, column e % W) of three separate planes
// (coalesced per plane) and is staged interleaved at position 3 (e % W) + e / W.  `off` / `pos` are fixed per thread.
struct ScLane {
  int off[SC_PF_J];      // source offset inside the row (NHWC) or plane offset c * H * W + iw (NCHW); -1: none
  int pos[SC_PF_J];      // staged position (without the + 9 shift)
};
// element offset of input row (n, ih): NHWC: the row's 3 W contiguous values; NCHW: row ih of the image's first plane
template <bool NCHW>
__device__ __forceinline__ long long sc_row_base(const ScArgs& a, int n, int ih) {
  if (!NCHW) return ((long long)n * a.H + ih) * (3LL * a.W);
  return ((long long)n * 3 * a.H + ih) * (long long)a.W;
}
__device__ __forceinline__ void sc_row_adv(const ScArgs& a, int& g, int& n, int& ih) {
  ++g;
  if (++ih == a.H) {
    ih = 0;
    ++n;
  }
}

struct ScPix {
  int n, oh, ow;
};
__device__ __forceinline__ void sc_pix_init(const ScArgs& a, int p, ScPix& px) {
  const unsigned orow = (unsigned)p / (unsigned)a.OW;
  px.ow = p - (int)orow * a.OW;
  px.n = (int)(orow / (unsigned)a.OH);
  px.oh = (int)orow - px.n * a.OH;
}
// + 128 pixels (OW >= 128: at most one output-row wrap)
__device__ __forceinline__ void sc_pix_adv(const ScArgs& a, ScPix& px) {
  px.ow += 128;
  if (px.ow >= a.OW) {
    px.ow -= a.OW;
    if (++px.oh >= a.OH) {
      px.oh = 0;
      ++px.n;
    }
  }
}

// the three 16-byte chunks of kernel row kh of one pixel: staged row (or the zero row) -> swizzled A tile
__device__ __forceinline__ void sc_gather_row(const __nv_bfloat16* colbase, int rowlen, int H, int ih0, int gbase,
                                              uint8_t* dbase, uint32_t r7s, int kh) {
  const int ih = ih0 + kh;
  const int slot = (ih < 0 || ih >= H) ? SC_NSLOT : ((gbase + kh) & (SC_NSLOT - 1));
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(colbase + slot * rowlen);
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const int kc = kh * 3 + c;
    uint4 v;
    v.x = s32[4 * c + 0];
    v.y = s32[4 * c + 1];
    v.z = s32[4 * c + 2];
    v.w = s32[4 * c + 3];
    *reinterpret_cast<uint4*>(dbase + (kc >> 3) * SC_A_ATOM + (((uint32_t)(kc & 7) << 4) ^ r7s)) = v;
  }
}

template <bool IN_BF16, bool NCHW>
__global__ void __launch_bounds__(SC_THREADS, 2) stem_conv_kernel(const ScArgs a) {
  extern __shared__ uint8_t sc_smem_raw[];
  // 1024-byte alignment by OFFSET from the shared array (pointer arithmetic keeps the shared state space: rounding the
  // pointer through an integer made every access below a generic LD / ST instead of LDS / STS)
  uint8_t* smem = smem_align(sc_smem_raw, 1024);
  uint8_t* sA = smem;
  uint8_t* sB = smem + SC_A_BYTES;
  __nv_bfloat16* rows = reinterpret_cast<__nv_bfloat16*>(smem + SC_A_BYTES + SC_B_BYTES);   // [SC_NSLOT + 1][rowlen]
  const int rowlen = a.rowlen;
  uint8_t* tail = reinterpret_cast<uint8_t*>(rows + (SC_NSLOT + 1) * rowlen);      // slot SC_NSLOT stays zero; rowlen % 8 == 0: 16-byte aligned
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tail + 8);
  float* red = reinterpret_cast<float*>(tail + 16);                                       // [2][4][64]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int W3 = 3 * a.W;
  ScLane ln;
#pragma unroll
  for (int j = 0; j < SC_PF_J; j++) {
    const int e = tid + j * SC_THREADS;
    if (e >= W3) {
      ln.off[j] = -1;
      ln.pos[j] = 0;
    } else if (NCHW) {
      const int c = e / a.W, iw = e - c * a.W;
      ln.off[j] = c * a.H * a.W + iw;
      ln.pos[j] = 3 * iw + c;
    } else {
      ln.off[j] = ln.pos[j] = e;
    }
  }

  // ---- one-time set-up: barrier, TMEM, zeroed A tile / staged rows, resident weights
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(SC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    uint4* pa = reinterpret_cast<uint4*>(sA);
    for (int i = tid; i < SC_A_BYTES / 16; i += SC_THREADS) pa[i] = make_uint4(0u, 0u, 0u, 0u);
    uint32_t* pr = reinterpret_cast<uint32_t*>(rows);
    const int nw = (SC_NSLOT + 1) * rowlen / 2;
    for (int i = tid; i < nw; i += SC_THREADS) pr[i] = 0u;
    // weights: 64 rows x 24 chunks of 8 bf16
    const uint4* wsrc = reinterpret_cast<const uint4*>(a.wpk);
    for (int i = tid; i < SC_COUT * (SC_KPAD / 8); i += SC_THREADS) {
      const int r = i / (SC_KPAD / 8), kc = i - r * (SC_KPAD / 8);
      *reinterpret_cast<uint4*>(sB + (kc >> 3) * SC_B_ATOM + sc_swz(r, kc & 7)) = __ldg(wsrc + i);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;
  const uint32_t idesc = umma_idesc<0>(SC_COUT);

  const int t0 = (int)((long long)blockIdx.x * a.ntiles / gridDim.x);
  const int t1 = (int)((long long)(blockIdx.x + 1) * a.ntiles / gridDim.x);
  float sacc = 0.f, qacc = 0.f;        // lane l of warp (q, half): channel half * 32 + l over the rows of quarter q
  uint32_t phase = 0;
  const int P = (int)a.P;              // < 2^31 (checked by the host): all index arithmetic below is 32-bit
  const int ntiles = (int)a.ntiles;

  // pixel cursors (n, oh, ow), advanced by 128 pixels per tile without divisions (OW >= 128: at most one row wrap):
  // the tile's first and last pixel (-> the input rows it needs) and this thread's own gather pixel
  ScPix pxa, pxb, pxr;
  // staging cursor: the next input row (g = n * H + ih) this CTA has not fetched yet
  int cur_g = 0, cur_n = 0, cur_ih = 0;
  float pf[SC_PF_ROWS][SC_PF_J];
  int pf_g0 = 0, pf_n = 0;
  if (t0 < t1) {
    sc_pix_init(a, t0 * 128, pxa);
    sc_pix_init(a, min(t0 * 128 + 127, P - 1), pxb);
    sc_pix_init(a, min(t0 * 128 + (tid & 127), P - 1), pxr);
    cur_n = pxa.n;
    cur_ih = max(2 * pxa.oh - 3, 0);
    cur_g = cur_n * a.H + cur_ih;
  }

  for (int t = t0; t < t1; t++) {
    const int ghi = pxb.n * a.H + min(2 * pxb.oh + 3, a.H - 1);
    // ---- stage the rows this tile adds: the prefetched ones from registers, any others synchronously
    if (pf_n > 0) {
#pragma unroll
      for (int i = 0; i < SC_PF_ROWS; i++) {
        if (i < pf_n) {
          __nv_bfloat16* dst = rows + ((pf_g0 + i) & (SC_NSLOT - 1)) * rowlen + 9;
#pragma unroll
          for (int j = 0; j < SC_PF_J; j++)
            if (ln.off[j] >= 0) dst[ln.pos[j]] = __float2bfloat16_rn(pf[i][j]);
        }
      }
      pf_n = 0;
    }
    while (cur_g <= ghi) {
      __nv_bfloat16* dst = rows + (cur_g & (SC_NSLOT - 1)) * rowlen + 9;
      const long long src = sc_row_base<NCHW>(a, cur_n, cur_ih);
#pragma unroll
      for (int j = 0; j < SC_PF_J; j++)
        if (ln.off[j] >= 0) dst[ln.pos[j]] = __float2bfloat16_rn(sc_load<IN_BF16>(a.x, src + ln.off[j]));
      sc_row_adv(a, cur_g, cur_n, cur_ih);
    }
    __syncthreads();

    // ---- gather the A tile: thread = (pixel r, kernel rows 0..3 | 4..6); chunk indices are compile-time constants in
    // each half, so a chunk costs 4 shared loads at immediate offsets, one XOR for the swizzle and one 16-byte store
    {
      const int r = tid & 127;
      if (t * 128 + r < P) {
        const int gbase = pxr.n * a.H + 2 * pxr.oh - 3;
        const int ih0 = 2 * pxr.oh - 3;
        const __nv_bfloat16* colbase = rows + 6 * pxr.ow;
        uint8_t* dbase = sA + (r >> 3) * 1024 + (r & 7) * 128;
        const uint32_t r7s = (uint32_t)(r & 7) << 4;
        if (tid < 128) {
#pragma unroll
          for (int kh = 0; kh < 4; kh++) sc_gather_row(colbase, rowlen, a.H, ih0, gbase, dbase, r7s, kh);
        } else {
#pragma unroll
          for (int kh = 4; kh < 7; kh++) sc_gather_row(colbase, rowlen, a.H, ih0, gbase, dbase, r7s, kh);
        }
      }
    }
    __syncthreads();

    // ---- multiply: one thread publishes the tile to the async proxy and issues the 11 k-steps; the rows of the next
    // tile are fetched into registers meanwhile
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < SC_KSTEPS; ks++) {
        const uint64_t ad = umma_desc_sw128(smem_u32(sA + (ks >> 2) * SC_A_ATOM)) + (uint64_t)(2 * (ks & 3));
        const uint64_t bd = umma_desc_sw128(smem_u32(sB + (ks >> 2) * SC_B_ATOM)) + (uint64_t)(2 * (ks & 3));
        umma<0>(tmem_d, ad, bd, idesc, ks ? 1u : 0u);
      }
      umma_commit(bar);
    }
    sc_pix_adv(a, pxa);
    sc_pix_adv(a, pxb);
    sc_pix_adv(a, pxr);
    if (t + 1 < t1) {
      if (t + 2 == ntiles && (t + 1) * 128 + 127 >= P) sc_pix_init(a, P - 1, pxb);      // partial last tile
      const int nhi = pxb.n * a.H + min(2 * pxb.oh + 3, a.H - 1);
      int cnt = nhi + 1 - cur_g;
      if (cnt > SC_PF_ROWS) cnt = SC_PF_ROWS;
      if (cnt > 0) {
        pf_g0 = cur_g;
        pf_n = cnt;
#pragma unroll
        for (int i = 0; i < SC_PF_ROWS; i++) {
          if (i < pf_n) {
            const long long src = sc_row_base<NCHW>(a, cur_n, cur_ih);
#pragma unroll
            for (int j = 0; j < SC_PF_J; j++) pf[i][j] = ln.off[j] >= 0 ? sc_load<IN_BF16>(a.x, src + ln.off[j]) : 0.f;
            sc_row_adv(a, cur_g, cur_n, cur_ih);
          }
        }
      }
    }

    // ---- epilogue: warp (q = warp % 4, half = warp / 4): pixel 32 q + lane, channels 32 half .. + 31
    mbar_wait_warp(bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      const int q = warp & 3, half = warp >> 2;
      const int p = t * 128 + q * 32 + lane;
      const bool ok = p < P;
      uint32_t v[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32);
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 16) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]),
              "=r"(v[c0 + 6]), "=r"(v[c0 + 7]), "=r"(v[c0 + 8]), "=r"(v[c0 + 9]), "=r"(v[c0 + 10]), "=r"(v[c0 + 11]),
              "=r"(v[c0 + 12]), "=r"(v[c0 + 13]), "=r"(v[c0 + 14]), "=r"(v[c0 + 15])
            : "r"(taddr + (uint32_t)c0));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float f[32];
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
        pk[i] = *reinterpret_cast<const uint32_t*>(&b);
        const float2 fb = __bfloat1622float2(b);
        f[2 * i] = ok ? fb.x : 0.f;
        f[2 * i + 1] = ok ? fb.y : 0.f;
      }
      if (ok) {
        uint4* dst = reinterpret_cast<uint4*>(a.y + (long long)p * SC_COUT + half * 32);
#pragma unroll
        for (int i = 0; i < 4; i++) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      }
      if (a.part) {
        float g[32];
#pragma unroll
        for (int i = 0; i < 32; i++) g[i] = f[i] * f[i];
        // transposing butterfly: after the step with offset o a lane keeps the half of its values selected by its bit o
        // (measured against a transposition through the free A buffer -- 8 STS.128 + 32 LDS + a CTA barrier per
        // thread and tile: 71 us per launch instead of 50)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
          const bool up = (lane & o) != 0;
#pragma unroll
          for (int i = 0; i < o; i++) {
            const float sf = up ? f[i] : f[i + o], kf = up ? f[i + o] : f[i];
            f[i] = kf + __shfl_xor_sync(FULL, sf, o);
            const float sg = up ? g[i] : g[i + o], kg = up ? g[i + o] : g[i];
            g[i] = kg + __shfl_xor_sync(FULL, sg, o);
          }
        }
        sacc += f[0];
        qacc += g[0];
      }
    }
    tc_fence_before();
    __syncthreads();          // every TMEM read and staged-row read of this tile is done
  }

  // ---- one partial row per CTA: part[blk][0][c] = sum, part[blk][1][c] = sum of squares
  if (a.part) {
    const int q = warp & 3, half = warp >> 2;
    red[(0 * 4 + q) * 64 + half * 32 + lane] = sacc;
    red[(1 * 4 + q) * 64 + half * 32 + lane] = qacc;
    __syncthreads();
    if (tid < 128) {
      const int which = tid >> 6, c = tid & 63;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 4; k++) s += red[(which * 4 + k) * 64 + c];
      a.part[((long long)blockIdx.x * 2 + which) * SC_COUT + c] = s;
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(SC_TMEM_COLS) : "memory");
  }
}

static inline int sc_rowlen(int W, int OW) {
  int need = 6 * (OW - 1) + SC_RUN;
  if (need < 9 + 3 * W) need = 9 + 3 * W;
  return (need + 7) / 8 * 8;
}
static inline size_t sc_smem(int rowlen) {
  return 1024 + SC_A_BYTES + SC_B_BYTES + (size_t)(SC_NSLOT + 1) * rowlen * 2 + 16 + 16 + 2 * 4 * 64 * 4;
}

template <bool IN_BF16, bool NCHW>
static int sc_grid(size_t smem) {
  static int cached_smem = -1, cached = 0;
  if (cached_smem == (int)smem) return cached;
  auto kern = stem_conv_kernel<IN_BF16, NCHW>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  // two CTAs per SM need the full shared-memory carve-out (2 x 108 KB at W = 299)
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  // residency by the real limits (<= 128 registers x 256 threads, shared memory incl. the 1 KB the system reserves per
  // CTA, 64 of 512 TMEM columns): the runtime's occupancy calculator reports 1 CTA/SM for this kernel at any
  // shared-memory size, the hardware co-schedules two
  const int per_sm = 2 * (smem + 1024) <= 228 * 1024 ? 2 : 1;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cached_smem = (int)smem;
  cached = per_sm * sms;
  return cached;
}

}  // namespace v2f

using namespace v2f;

// Number of CTAs v2f_stem_conv_fwd launches for this problem = rows of ``part``; 0 when the shape is not supported
// (the caller then keeps the library convolution).
extern "C" int v2f_stem_conv_blocks(int N, int H, int W, int x_bf16, int x_nchw) {
  if (N <= 0 || H < 7 || W < 7) return 0;
  const int OW = (W - 1) / 2 + 1, OH = (H - 1) / 2 + 1;
  if (OW < 128 || 3 * W > SC_PF_J * SC_THREADS) return 0;      // a tile spans at most two output rows; register prefetch
  if ((long long)N * OH * OW + 128 >= (1LL << 31) || (long long)N * H >= (1LL << 31)) return 0;   // 32-bit index arithmetic
  const size_t smem = sc_smem(sc_rowlen(W, OW));
  if (smem > 227 * 1024) return 0;
  const long long ntiles = ((long long)N * OH * OW + 127) / 128;
  long long g = x_bf16 ? (x_nchw ? sc_grid<true, true>(smem) : sc_grid<true, false>(smem))
                       : (x_nchw ? sc_grid<false, true>(smem) : sc_grid<false, false>(smem));
  if (g > ntiles) g = ntiles;
  return (int)g;
}

// diagnostics (tools/stem_probe.py): occupancy the runtime reports for the fp32-NCHW instantiation
extern "C" int v2f_stem_conv_occupancy(int W, int* regs, int* smem_bytes, int* per_sm) {
  auto kern = stem_conv_kernel<false, true>;
  const int OW = (W - 1) / 2 + 1;
  const size_t smem = sc_smem(sc_rowlen(W, OW));
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return V2F_ERR_LAUNCH;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, SC_THREADS, smem);
  *regs = fa.numRegs;
  *smem_bytes = (int)smem + (int)fa.sharedSizeBytes;
  *per_sm = e == cudaSuccess ? n : -(int)e;
  return V2F_OK;
}

extern "C" int v2f_stem_conv_fwd(int N, int H, int W, const void* x, int x_bf16, int x_nchw, const void* wpk, void* y,
                                 float* part, void* st) {
  V2F_REQUIRE(x && wpk && y, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(aligned16(wpk) && aligned16(y), V2F_ERR_ALIGN);
  const int grid = v2f_stem_conv_blocks(N, H, W, x_bf16, x_nchw);
  V2F_REQUIRE(grid > 0, V2F_ERR_UNSUPPORTED);
  ScArgs a;
  a.x = x;
  a.wpk = reinterpret_cast<const __nv_bfloat16*>(wpk);
  a.y = reinterpret_cast<__nv_bfloat16*>(y);
  a.part = part;
  a.N = N;
  a.H = H;
  a.W = W;
  a.OH = (H - 1) / 2 + 1;
  a.OW = (W - 1) / 2 + 1;
  a.P = (long long)N * a.OH * a.OW;
  a.ntiles = (a.P + 127) / 128;
  a.rowlen = sc_rowlen(W, a.OW);
  const size_t smem = sc_smem(a.rowlen);
  cudaStream_t s = (cudaStream_t)st;
  prof_begin(V2F_K_STEM_CONV, s);
  prof_bytes(V2F_K_STEM_CONV, (long long)N * H * W * 3 * (x_bf16 ? 2 : 4) + a.P * SC_COUT * 2);
  if (x_bf16 && x_nchw)
    stem_conv_kernel<true, true><<<grid, SC_THREADS, smem, s>>>(a);
  else if (x_bf16)
    stem_conv_kernel<true, false><<<grid, SC_THREADS, smem, s>>>(a);
  else if (x_nchw)
    stem_conv_kernel<false, true><<<grid, SC_THREADS, smem, s>>>(a);
  else
    stem_conv_kernel<false, false><<<grid, SC_THREADS, smem, s>>>(a);
  prof_end(V2F_K_STEM_CONV, s);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

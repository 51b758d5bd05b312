// Tensor-core GEMM for sm_100a: tcgen05.mma (bf16 or tf32 inputs, fp32 accumulate in TMEM),
// operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through an mbarrier ring.
//
//   C[M,N] (fp32) = A[M,K] * B[N,K]^T (+bias[n]) (+beta*C) (ReLU)      both operands K-major
//
// This is the "both K-major" form every dense contraction of the hot path is brought into
// (forward: x W^T; backward: dy (W^T)^T and (dy^T)(x^T)^T with explicitly transposed copies), so
// one kernel serves nn.Linear forward/backward in /root/reference/models/CrossAttnRNN210.py:72,
// 84-85,126,128,132 and the per-step recurrent projections of the decoder / GRU (:135-140).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected
// lane issues tcgen05.mma for the CTA), warps 2..5 = epilogue (tcgen05.ld -> registers -> global;
// warp w owns TMEM lanes 32*(w%4)..+31).  One 128 x BN output tile per CTA.
// Every mbarrier wait is bounded: a protocol bug traps instead of hanging the GPU.
#include "tc.cuh"

namespace v2f {

constexpr int TC_STAGE_BYTES_K = 128;  // one 128-byte swizzle span of K per stage row
constexpr int TC_THREADS = 192;

#ifdef V2F_GEMM_TIMELINE
__device__ long long g_tl[16];
#define TL(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_tl[i] = t_; } } while (0)
#else
#define TL(i) do { } while (0)
#endif

struct TcArgs {
  int M, N, K;
  float* C;
  long long ldc;
  const float* bias;
  float beta;
  int act;
  int k_blocks;   // number of 128-byte K blocks per CTA (split-K: blocks per split)
  int atomic;     // split-K: accumulate with red.add into pre-zeroed / pre-scaled C
  int splits;     // blockIdx.z = batch * splits + split
  long long sC;   // batch stride of C (elements)
};

// KIND 0: bf16 (64 elements per 128-B span, UMMA_K 16); KIND 1: tf32 (32 elements, UMMA_K 8)
template <int KIND, int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcArgs a) {
  constexpr int ELEM = KIND == 0 ? 2 : 4;
  constexpr int BK = TC_STAGE_BYTES_K / ELEM;       // elements of K per stage
  constexpr int UK = 32 / ELEM;                     // UMMA_K
  constexpr int A_BYTES = TC_BM * TC_STAGE_BYTES_K;
  constexpr int B_BYTES = BN * TC_STAGE_BYTES_K;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align(smem_raw, 1024);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* conv = tmem_full + 1;    // tf32 only: stage rounded to tf32 in place, ready for the MMA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(conv + STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TL(0);
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int bz = blockIdx.z / a.splits, split = blockIdx.z - bz * a.splits;
  const int kb0 = split * a.k_blocks;
  int nkb = a.k_blocks;
  const int total_kb = (a.K + BK - 1) / BK;
  if (kb0 + nkb > total_kb) nkb = total_kb - kb0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&conv[s], TC_THREADS - 64);
    }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;
  if (threadIdx.x == 0) TL(1);

  if (warp == 0) {
    if (lane == 0 && nkb > 0) {
      for (int i = 0; i < nkb; i++) {
        const int s = i % STAGES, r = i / STAGES;
        mbar_wait(&empty[s], (r & 1) ^ 1);
        mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
        const int kc = (kb0 + i) * BK;
        tma_load_3d(sA + s * A_BYTES, &mapA, &full[s], kc, m0, bz);
        tma_load_3d(sB + s * B_BYTES, &mapB, &full[s], kc, n0, bz);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nkb > 0) {
      const uint32_t idesc = umma_idesc<KIND>(BN);
      for (int i = 0; i < nkb; i++) {
        const int s = i % STAGES, r = i / STAGES;
        mbar_wait((KIND == 1 && (a.act & 4)) ? &conv[s] : &full[s], r & 1);
        if (i == 0) TL(2);
        if (i == 1) TL(3);
        if (i == nkb - 1) TL(4);
        tc_fence_after();
        const uint64_t ad = umma_desc_sw128(smem_u32(sA + s * A_BYTES));
        const uint64_t bd = umma_desc_sw128(smem_u32(sB + s * B_BYTES));
#ifdef V2F_GEMM_TIMELINE
        if (!(a.act & 8))
#endif
#pragma unroll
        for (int k = 0; k < BK / UK; k++) {
          // advance 32 bytes of K inside the swizzle span: +2 in the (address >> 4) field
          umma<KIND>(tmem_d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (i | k) ? 1u : 0u);
        }
        umma_commit(&empty[s]);     // frees the smem slot when the MMAs above have read it
      }
      umma_commit(tmem_full);       // accumulator complete
      TL(5);
    }
  } else {
    // ---- epilogue: warp w reads TMEM lanes 32*(w%4) .. +31  == output rows m0 + 32*(w%4) + lane
    const int q = warp & 3;
    if (KIND == 1 && (a.act & 4)) {
      // kind::tf32 ignores the low 13 mantissa bits of each fp32 operand, i.e. truncates: a biased error
      // (every product shrinks by ~7e-4) that survives the cancellations of a backward pass through
      // BatchNorm / LayerNorm.  On request (act bit 2) the warps that are idle until the accumulator is
      // complete round each landed stage to nearest tf32 in place (cvt.rna), publish it to the async
      // proxy and hand it to the MMA thread.  Costs one shared-memory read+write of the stage, so it is
      // used for the small GEMMs of the GTM family, not for the streaming projections.
      constexpr int NT = TC_THREADS - 64;
      constexpr int A_V = A_BYTES / 16 / NT;          // float4 per thread, A tile (8)
      constexpr int B_V = (B_BYTES / 16 + NT - 1) / NT;
      const int t = threadIdx.x - 64;
      for (int i = 0; i < nkb; i++) {
        const int s = i % STAGES, r = i / STAGES;
        mbar_wait(&full[s], r & 1);
        float4* pa = reinterpret_cast<float4*>(sA + s * A_BYTES);
        float4* pb = reinterpret_cast<float4*>(sB + s * B_BYTES);
        float4 v[A_V + B_V];
#pragma unroll
        for (int j = 0; j < A_V; j++) v[j] = pa[t + j * NT];
#pragma unroll
        for (int j = 0; j < B_V; j++)
          if (t + j * NT < B_BYTES / 16) v[A_V + j] = pb[t + j * NT];
#pragma unroll
        for (int j = 0; j < A_V + B_V; j++) {
          uint32_t x, y, z, w;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(x) : "f"(v[j].x));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(v[j].y));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(z) : "f"(v[j].z));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(w) : "f"(v[j].w));
          v[j] = make_float4(__uint_as_float(x), __uint_as_float(y), __uint_as_float(z), __uint_as_float(w));
        }
#pragma unroll
        for (int j = 0; j < A_V; j++) pa[t + j * NT] = v[j];
#pragma unroll
        for (int j = 0; j < B_V; j++)
          if (t + j * NT < B_BYTES / 16) pb[t + j * NT] = v[A_V + j];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&conv[s]);
      }
    }
    if (nkb > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    if (threadIdx.x == 64) TL(6);
    // Phase 1: TMEM -> registers -> shared staging tile [128][BN+1] (the operand ring is free: every MMA
    // that read it has completed).  Thread = accumulator row; the odd pitch makes the writes conflict-free.
    constexpr int LDS = BN + 1;
    float* stg = reinterpret_cast<float*>(smem);
    const int rloc = q * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      if (nkb > 0) {
        const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 16; j++) stg[rloc * LDS + c0 + j] = __uint_as_float(v[j]);
    }
    named_bar_sync(1, TC_THREADS - 64);
    if (threadIdx.x == 64) TL(9);
    // Phase 2: coalesced write-out.  A warp instruction covers RPI rows x CPR consecutive columns
    // (one 128-byte line per row when BN >= 32), instead of 32 rows x 4 bytes.
    constexpr int CPR = BN < 32 ? BN : 32;      // lanes per row
    constexpr int RPI = 32 / CPR;               // rows per warp instruction
    const int wq = warp - 2;
    const int rsub = lane / CPR, cl = lane % CPR;
    const int ncols = a.N - n0 < BN ? a.N - n0 : BN;
    for (int r0 = wq * RPI; r0 < TC_BM; r0 += 4 * RPI) {
      const int r = r0 + rsub;
      const int grow = m0 + r;
      if (grow >= a.M) continue;
      const float* srow = stg + r * LDS;
      if (a.act & 2) {
        // bf16 output (e.g. the gradient handed back to the bf16 backbone): no beta / atomics
        __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(a.C) + (long long)bz * a.sC + (long long)grow * a.ldc + n0;
        for (int c = cl; c < ncols; c += CPR) {
          float x = srow[c];
          if (a.bias) x += a.bias[n0 + c];
          if (a.act & 1) x = fmaxf(x, 0.f);
          crow[c] = __float2bfloat16_rn(x);
        }
      } else {
        float* crow = a.C + (long long)bz * a.sC + (long long)grow * a.ldc + n0;
        for (int c = cl; c < ncols; c += CPR) {
          float x = srow[c];
          if (a.atomic) {
            if (a.bias && split == 0) x += a.bias[n0 + c];
            atomicAdd(crow + c, x);
          } else {
            if (a.bias) x += a.bias[n0 + c];
            if (a.beta != 0.f) x += a.beta * crow[c];
            if (a.act & 1) x = fmaxf(x, 0.f);
            crow[c] = x;
          }
        }
      }
    }
  }
  if (threadIdx.x == 64) TL(7);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TL(8);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// batch x row-major [rows, cols] (cols contiguous, row stride ld, batch stride bs, in elements),
// box = [1, box_rows, 128 bytes]
int tc_make_map(CUtensorMap* map, int kind, const void* ptr, long long rows, long long cols, long long ld,
                    long long batch, long long bs, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return V2F_ERR_UNSUPPORTED;
  const int elem = kind == 0 ? 2 : 4;
  if (batch <= 1) { batch = 1; bs = rows * ld; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * elem, (cuuint64_t)bs * elem};
  cuuint32_t box[3] = {(cuuint32_t)(128 / elem), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? V2F_OK : V2F_ERR_BAD_ARG;
}

template <int KIND, int BN>
static int launch_tc(const CUtensorMap& mA, const CUtensorMap& mB, const TcArgs& a, int gz, cudaStream_t s) {
  // two CTAs per SM (<= ~110 KB each) so that one tile's epilogue (a burst of global stores) overlaps the
  // other tile's main loop
  constexpr int STAGES = BN >= 128 ? 3 : (BN >= 64 ? 4 : 6);
  constexpr size_t smem = 1024 + (size_t)STAGES * (TC_BM + BN) * 128 + (3 * STAGES + 1) * 8 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<KIND, BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return V2F_ERR_LAUNCH;
    attr_set = true;
  }
  dim3 grid((a.N + BN - 1) / BN, (a.M + TC_BM - 1) / TC_BM, gz);
  gemm_tc_kernel<KIND, BN, STAGES><<<grid, TC_THREADS, smem, s>>>(mA, mB, a);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

}  // namespace v2f

using namespace v2f;

// C[b][M,N] = A[b][M,K] B[b][N,K]^T (+bias) (+beta C) (act).  kind 0: A,B bf16; kind 1: A,B fp32 read as
// tf32.  ld*/s* in elements.  splits > 1: split-K with atomic accumulation (C must hold the value to
// accumulate onto, e.g. zeros; beta/act are then not applied).
extern "C" int v2f_gemm_tc_batched(int kind, int M, int N, int K, const void* A, long long lda, long long sA,
                                   const void* B, long long ldb, long long sB, float* C, long long ldc,
                                   long long sC, int batch, const float* bias, float beta, int act, int splits,
                                   void* stream) {
  V2F_REQUIRE(kind == 0 || kind == 1, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0 && A && B && C, V2F_ERR_BAD_ARG);
  const int elem = kind == 0 ? 2 : 4;
  V2F_REQUIRE(aligned16(A) && aligned16(B), V2F_ERR_ALIGN);
  V2F_REQUIRE((lda * elem) % 16 == 0 && (ldb * elem) % 16 == 0, V2F_ERR_ALIGN);
  V2F_REQUIRE(batch == 1 || ((sA * elem) % 16 == 0 && (sB * elem) % 16 == 0), V2F_ERR_ALIGN);
  if (splits < 1) splits = 1;
  V2F_REQUIRE(splits == 1 || (beta == 0.f && (act & 3) == 0), V2F_ERR_BAD_ARG);
  V2F_REQUIRE(!(act & 2) || beta == 0.f, V2F_ERR_BAD_ARG);
  V2F_REQUIRE((long long)batch * splits <= 65535, V2F_ERR_BAD_ARG);
  // tile width: keep >= ~1 wave of CTAs on 148 SMs for the small-M recurrent GEMMs
  const int mt = (M + TC_BM - 1) / TC_BM;
  // smallest tile width that still fits the whole problem in one wave of 148 CTAs (every CTA must
  // stream the full 128 x K slab of A, so fewer, wider tiles re-read A less); else 128
  int bn = 128;
  for (int cand = 16; cand < 128; cand <<= 1)
    if ((long long)mt * ((N + cand - 1) / cand) * splits * batch <= 148) { bn = cand; break; }
  const int bk = 128 / elem;
  const int total_kb = (K + bk - 1) / bk;
  TcArgs a{M, N, K, C, ldc, bias, beta, act, (total_kb + splits - 1) / splits, splits > 1 ? 1 : 0, splits, sC};
  CUtensorMap mA, mB;
  V2F_TRY(tc_make_map(&mA, kind, A, M, K, lda, batch, sA, TC_BM));
  V2F_TRY(tc_make_map(&mB, kind, B, N, K, ldb, batch, sB, bn));
  cudaStream_t s = (cudaStream_t)stream;
  const int gz = batch * splits;
  int rc = V2F_OK;
#define DISPATCH(KIND_)                                      \
  switch (bn) {                                              \
    case 128: rc = launch_tc<KIND_, 128>(mA, mB, a, gz, s); break; \
    case 64: rc = launch_tc<KIND_, 64>(mA, mB, a, gz, s); break;   \
    case 32: rc = launch_tc<KIND_, 32>(mA, mB, a, gz, s); break;   \
    default: rc = launch_tc<KIND_, 16>(mA, mB, a, gz, s); break;   \
  }
  // roofline leg of bench.py: CUDA events around the launch, 2 M N K flops per problem in the byte counter
  prof_begin(V2F_K_GEMM_TC, s);
  prof_bytes(V2F_K_GEMM_TC, 2LL * M * N * K * batch);
  if (kind == 0) {
    DISPATCH(0)
  } else {
    DISPATCH(1)
  }
#undef DISPATCH
  prof_end(V2F_K_GEMM_TC, s);
  return rc;
}

extern "C" int v2f_gemm_tc(int kind, int M, int N, int K, const void* A, long long lda, const void* B,
                           long long ldb, float* C, long long ldc, const float* bias, float beta, int act,
                           int splits, void* stream) {
  return v2f_gemm_tc_batched(kind, M, N, K, A, lda, 0, B, ldb, 0, C, ldc, 0, 1, bias, beta, act, splits, stream);
}

// ------------------------------------------------------------------------------------------
// fp32 -> bf16 cast and transposing cast (operand preparation for the K-major GEMM form)
namespace v2f {
__global__ void cast_bf16_kernel(long long n, const float* __restrict__ x, __nv_bfloat16* __restrict__ out) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = ld4(x + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<__nv_bfloat162*>(out + i) = a;
    *reinterpret_cast<__nv_bfloat162*>(out + i + 2) = b;
  } else {
    for (long long j = i; j < n; j++) out[j] = __float2bfloat16_rn(x[j]);
  }
}

// out[c, r] = cast(in[r, c]);  in: [rows, cols] with row stride ld.  32x32 smem tile.
template <typename TIN, typename TOUT>
__global__ void transpose_kernel(int rows, int cols, const TIN* __restrict__ in, long long ld,
                                 TOUT* __restrict__ out, long long ldo) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? (float)in[(long long)r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(long long)c * ldo + r] = (TOUT)tile[threadIdx.x][i];
  }
}
}  // namespace v2f

extern "C" int v2f_cast_bf16(long long n, const float* x, void* out, void* stream) {
  V2F_REQUIRE(n >= 0 && x && out, V2F_ERR_BAD_ARG);
  if (n == 0) return V2F_OK;
  V2F_REQUIRE(aligned16(x) && aligned16(out), V2F_ERR_ALIGN);
  const long long thr = (n + 3) / 4;
  cast_bf16_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, x, (__nv_bfloat16*)out);
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

// out[cols, rows] = in[rows, cols]^T.  in_kind / out_kind: 0 = bf16, 1 = fp32.
extern "C" int v2f_transpose(int rows, int cols, const void* in, long long ld, int in_kind, void* out,
                             long long ldo, int out_kind, void* stream) {
  V2F_REQUIRE(rows > 0 && cols > 0 && in && out, V2F_ERR_BAD_ARG);
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  cudaStream_t s = (cudaStream_t)stream;
  if (in_kind == 1 && out_kind == 1)
    transpose_kernel<float, float><<<grid, block, 0, s>>>(rows, cols, (const float*)in, ld, (float*)out, ldo);
  else if (in_kind == 1 && out_kind == 0)
    transpose_kernel<float, __nv_bfloat16><<<grid, block, 0, s>>>(rows, cols, (const float*)in, ld, (__nv_bfloat16*)out, ldo);
  else if (in_kind == 0 && out_kind == 0)
    transpose_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, s>>>(rows, cols, (const __nv_bfloat16*)in, ld, (__nv_bfloat16*)out, ldo);
  else if (in_kind == 0 && out_kind == 1)
    transpose_kernel<__nv_bfloat16, float><<<grid, block, 0, s>>>(rows, cols, (const __nv_bfloat16*)in, ld, (float*)out, ldo);
  else
    return V2F_ERR_BAD_ARG;
  V2F_CHECK_LAUNCH();
  return V2F_OK;
}

#ifdef V2F_GEMM_TIMELINE
extern "C" int v2f_debug_timeline(long long* out) {
  return cudaMemcpyFromSymbol(out, v2f::g_tl, sizeof(long long) * 16) == cudaSuccess ? 0 : -3;
}
#endif

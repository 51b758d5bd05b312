/* libv2f_b200 - C ABI of the B200 (sm_100a) hot path of the Visuelle 2.0 multimodal forecasters.
 *
 * Plain C: raw device pointers, sizes, a cudaStream_t passed as void*.  No torch types.
 * Every function returns 0 (V2F_OK) or a negative error code; nothing allocates, creates
 * streams or keeps global state.  All tensors are caller-allocated, contiguous row-major fp32
 * (int64 for indices) unless stated, and 16-byte aligned.
 *
 * The reference (jeonghoya/visuelle2-multimodal-fusion) is pure Python: it has no FFI layer,
 * its "plugin interface" for this path is the nn.Module surface of models/*.py.  Each entry
 * point below therefore cites the reference lines whose arithmetic it replaces; the Python
 * binding (ctypes) a maintainer adds is shown in INTEGRATION.md.
 */
#ifndef V2F_H_
#define V2F_H_

#ifdef __cplusplus
extern "C" {
#endif

#define V2F_OK 0
#define V2F_ERR_BAD_ARG (-1)
#define V2F_ERR_ALIGN (-2)
#define V2F_ERR_LAUNCH (-3)
#define V2F_ERR_UNSUPPORTED (-4)

/* ABI version; bumped whenever a struct layout or signature changes. */
int v2f_version(void);
/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
long long v2f_launch_count(void);

/* Kernel ids for the optional event timing below. */
enum { V2F_K_ATTN_FWD = 0, V2F_K_ATTN_BWD = 1, V2F_K_TILEGRAD = 2, V2F_K_BN_STATS = 3, V2F_K_BN_APPLY = 4,
       V2F_K_BN_BWD_REDUCE = 5, V2F_K_BN_BWD_ELEMT = 6, V2F_K_DECODE_PERSIST_FWD = 7,
       V2F_K_DECODE_PERSIST_BWD = 8, V2F_K_STEM_CONV = 9, V2F_K_GEMM_TC = 10, V2F_K_COUNT = 11 };
/* Per-kernel CUDA-event timing on the launching stream (bench.py roofline leg).  Off by default.
 * v2f_prof_read sums the spans recorded for one kernel id since the previous read.            */
int v2f_prof_enable(int on);
int v2f_prof_read(int kernel_id, double* total_ms, long long* launches);
/* Same, plus the algorithmic bytes the recorded launches moved (kernels whose size varies per launch,
 * i.e. the BatchNorm sweeps, report them; 0 for the others).                                   */
int v2f_prof_read_bytes(int kernel_id, double* total_ms, long long* launches, long long* bytes);

/* ------------------------------------------------------------------------------------------
 * Dense projections: C[b] = op(A[b]) op(B[b]) (+bias[n]) (+beta*C[b]), optional ReLU (act=1).
 * ta=0: A stored [M,K]; ta=1: A stored [K,M].  tb=0: B stored [K,N]; tb=1: B stored [N,K].
 * Replaces nn.Linear / autograd mm in e.g. models/CrossAttnRNN210.py:72 (ImageEncoder.fc),
 * :84-85 (encoder_linear / decoder_linear), :196 (trend_linear), :208 (multimodal_embedder).
 * fp32 CUDA-core kernel (exact mode, 1e-5 contract).                                        */
int v2f_gemm_f32(int ta, int tb, int M, int N, int K, const float* A, int lda, long long strideA,
                 const float* B, int ldb, long long strideB, float* C, int ldc, long long strideC,
                 int batch, const float* bias, float beta, int act, void* stream);
/* Tensor-core variant (tcgen05.mma, TMEM accumulator, TMA-staged 128B-swizzled operands):
 *   C[M,N] (fp32) = A[M,K] B[N,K]^T (+bias) (+beta C) (ReLU), both operands K-major.
 * kind 0: A,B bf16; kind 1: A,B fp32 consumed as tf32 (the hardware truncates the low 13 mantissa
 * bits; act bit 2 (value 4) rounds each staged tile to nearest tf32 instead, at the cost of one
 * shared-memory pass).  act bit 0: ReLU; bit 1: C is bf16.  lda/ldb/ldc in elements; A,B 16-byte
 * aligned with 16-byte-multiple row pitch.  splits>1: split-K, partial sums added atomically
 * onto C (pre-zeroed by the caller; beta/act not applied).  Same reference call sites as
 * v2f_gemm_f32; this is the performance path (2e-2 contract).                                 */
int v2f_gemm_tc(int kind, int M, int N, int K, const void* A, long long lda, const void* B,
                long long ldb, float* C, long long ldc, const float* bias, float beta, int act,
                int splits, void* stream);
int v2f_gemm_tc_batched(int kind, int M, int N, int K, const void* A, long long lda, long long strideA,
                        const void* B, long long ldb, long long strideB, float* C, long long ldc,
                        long long strideC, int batch, const float* bias, float beta, int act, int splits,
                        void* stream);
/* Operand preparation for the K-major form: fp32->bf16 cast, and out[cols,rows] = in[rows,cols]^T
 * with optional conversion (kind 0 = bf16, 1 = fp32).                                         */
int v2f_cast_bf16(long long n, const float* x, void* out, void* stream);
int v2f_transpose(int rows, int cols, const void* in, long long ld, int in_kind, void* out,
                  long long ldo, int out_kind, void* stream);
/* out[n] = sum_m X[m,n] (+beta*out[n]) : bias gradients. */
int v2f_colsum_f32(int M, int N, const float* X, int ldx, float* out, float beta, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused recurrent-attention decoder (the hot loop):
 *   CrossAttnRNN210.forward loop   models/CrossAttnRNN210.py:191-225   (variant 0)
 *   CrossAttnRNN21 single fusion   models/CrossAttnRNN21.py:183-206    (variant 1)
 *   CrossAttnRNNDemand loop        models/CrossAttnRNNDemand.py:285-347 (variant 2)
 * with AdditiveAttention (CrossAttnRNN210.py:83-89 / CrossAttnRNNDemand.py:134-149), the
 * nn.GRU decoder cell (CrossAttnRNN210.py:135-140,210-211), decoder_fc and the teacher-forcing
 * select (:212-225).  Step-invariant projections are hoisted by the caller:
 *   Himg = We_img V, Htr = We_tr Vtr, Ptr[j] = W_tl[:, jE:(j+1)E] Vtr[j], HMst = We_mm Mst.
 * Row n belongs to item b = n / W.  E == A (attention_dim) is required, as in the reference.   */
typedef struct v2f_decode_params {
  int N, B, W, E, H, Li, Lt, T;
  int variant;      /* 0: 210, 1: 21 (T must be 1, no GRU), 2: Demand */
  int mod_mask;     /* bit0 date (always), bit1 image, bit2 attributes, bit3 trends */
  unsigned tf_mask; /* bit t: decoder input of step t+1 is y[:,t] instead of yhat_t */
  int precision;    /* 0: fp32 CUDA-core GEMMs (exact); 1: tf32 tcgen05 GEMMs for the per-step products */
  /* step-invariant tiles, per item */
  const float *Himg, *Vimg; /* [B,Li,E] energies source / context source (Demand: Vimg == Himg) */
  const float *Htr, *Ptr;   /* [B,Lt,E] */
  const float *Mst, *HMst;  /* [B,2,E]: (date, attributes) and their We_mm projections */
  const float *h0, *x0, *y; /* [N,H], [N], [N,T] (y may be NULL) */
  /* weights, packed by the caller */
  const float *Wcat, *bcat;        /* [3E+G,H],[3E+G]: rows = Wd_img,Wd_tr,Wd_mm,(W_hh); G=3H or 0 */
  const float *w_att, *beta_att;   /* [3,E],[3]: attn_linear of img, trend, multimodal */
  const float *b_tl;               /* [E] trend_linear.bias */
  const float *We_mm;              /* [E,E] multimodal_attention.encoder_linear */
  const float *W_me, *b_me;        /* [E,E],[E] multimodal_embedder */
  const float *W_ihc, *w_x, *b_ih; /* [3H,E],[3H],[3H]: decoder GRU weight_ih split ctx | scalar */
  const float *w_fc, *b_fc;        /* [H] (variant 1: [E]), [1] */
  /* forward outputs and saved activations */
  float *yhat;                             /* [N,T] */
  float *h_all;                            /* [T+1,N,H] */
  float *S_all;                            /* [T,N,3E+G] */
  float *alpha_img, *alpha_tr, *alpha_mm;  /* [T,N,Li],[T,N,Lt],[T,N,4] */
  float *C, *HC;                           /* [T,N,2,E] contexts (img,trend) and We_mm C */
  float *U, *CTX;                          /* [T,N,E] */
  float *GI;                               /* [N,3H] scratch */
  float *RZN;                              /* [T,N,3H] */
  float *xin;                              /* [T+1,N] decoder scalar inputs */
  /* backward: inputs, scratch (zero-initialised where noted), outputs */
  const float *dY;                         /* [N,T] dL/dyhat */
  float *dh;                               /* [N,H] in: dL/dh_T (zeros); out: dL/dh_0 */
  float *DScat;                            /* [T,N,3E+G] */
  float *DGI;                              /* [T,N,3H] */
  float *DCTX;                             /* [T,N,E] */
  float *dU;                               /* [N,E] scratch */
  float *DHC, *DC;                         /* [T,N,2,E] */
  float *DE_img, *DE_tr;                   /* [T,N,Li],[T,N,Lt] */
  float *DYH;                              /* [T,N] */
  float *dxn;                              /* [N] scratch */
  float *dw_acc;                           /* [N,3,E] zero-init */
  float *dMst_acc, *dHMst_acc;             /* [N,2,E] zero-init */
  float *dHimg, *dVimg, *dHtr, *dPtr;      /* tile gradients, written once */
  float *dMst, *dHMst;                     /* [B,2,E] */
  float *dWcat, *dbcat, *dw_att, *db_tl, *dWe_mm, *dW_me, *db_me, *dW_ihc, *dw_x, *db_ih,
      *dw_fc, *db_fc;
  /* precision 1 only (backward): scratch for transposed weights [H,3E+G],[E,3H],[E,E],[E,E] and
   * for transposed activation stacks of the weight-gradient products                          */
  float *WcatT, *W_ihcT, *W_meT, *We_mmT, *ws;
  long long ws_floats;
  /* optional scratch that enables the streaming (TMA-staged, 148-way balanced) attention kernels:
   * N * (ceil(Li/8)+ceil(Lt/8)) * (2E+2) floats.  NULL selects the simple per-(row,modality) kernels. */
  float *attn_ws;
  /* optional: the teacher-forcing bits in DEVICE memory (one unsigned, same layout as tf_mask).  When non-NULL
   * (and y != NULL) the kernels read it at run time instead of the immediate, so a captured CUDA graph of the
   * step can be replayed with a fresh draw (graphs.GraphedTrainStep).                                        */
  const unsigned* tf_mask_dev;
  /* optional scratch of v2f_decode_persist_ws_floats(N,E,H,T) floats: enables the persistent decoder
   * (csrc/decode_persist.cu): the whole T-step loop as ONE cooperative launch with the recurrent and
   * fusion weights resident in shared memory.  Taken when E is 256 or 512, H % 64 == 0, 148 <= H <= 512 (every CTA of the grid
   * owns at least one hidden unit), image and trend attention are both on, variant != 1 and attn_ws is given; otherwise (or when NULL) the
   * step-per-launch path runs.  Both paths fill the same saved activations.                         */
  float *persist_ws;
  /* optional scratch of v2f_decode_team_ws_floats(N,B,T,Li,Lt) floats (team_ws_floats = its size): enables the row-team
   * persistent decoder (csrc/decode_team.cu): one cooperative launch for the T-step loop, rows dealt to teams of 64
   * CTAs, products on tcgen05 with the bf16 weight slices resident in shared memory, bf16 attention tiles.  Taken in
   * tensor-core mode (precision 1) when E = H = 512, N <= 128, image and trend attention on, variant != 1 and
   * persist_ws is given too; otherwise the paths above run.  Fills the same saved activations.                    */
  float *team_ws;
  long long team_ws_floats;
} v2f_decode_params;

int v2f_decode_fwd(const v2f_decode_params* p, void* stream);
int v2f_decode_bwd(const v2f_decode_params* p, void* stream);
long long v2f_decode_persist_ws_floats(int N, int E, int H, int T);
long long v2f_decode_team_ws_floats(int N, int B, int T, int Li, int Lt);
/* Backward of the row-team decoder: v2f_decode_bwd runs the whole BPTT loop as one cooperative launch when the forward
 * took the row-team kernel (same params, team_ws still holding the forward's bf16 tiles) and the backward scratch `ws`
 * has at least this many floats; 0 disables it (the step-per-launch loop runs instead). */
long long v2f_decode_team_bwd_ws_floats(int N, int T);
int v2f_decode_team_bwd_enable(int on);
/* A/B switch (default 1): 0 keeps the decoder off the row-team kernel. */
int v2f_decode_team_enable(int on);
/* Profiling: CTA 0 of the row-team decoder stamps %globaltimer (ns), [T][16] unsigned long long at byte offset
 * v2f_decode_team_stamps_offset(...) of team_ws: stamp 2k = phase k starts, 2k+1 = CTA 0 finished its work of phase k,
 * k = 0 P1 (S product), 1 P2 (attention sweep + combine), 2 P3 (HC), 3 P4 (multimodal attention), 4 P5 (gates);
 * stamp 10 = step end.                                                                                       */
int v2f_decode_team_stamps_enable(int on);
long long v2f_decode_team_stamps_offset(int N, int B, int T, int Li, int Lt);
/* A/B switch (default 1): 0 forces the step-per-launch path even when persist_ws is given. */
int v2f_decode_persistent_enable(int on);
/* Profiling: CTA 0 of the persistent decoder stamps %globaltimer (ns) around its phases,
 * [T][16] unsigned long long at byte offset v2f_decode_persist_stamps_offset(N,E,H) of persist_ws:
 * stamp 2k = phase k starts (after the previous grid barrier), 2k+1 = CTA 0 finished phase k's work (before the
 * barrier), k = 0 P1 (S product), 1 P2 (attention sweep), 2 combine, 3 P3 (HC), 4 P4 (multimodal attention),
 * 5 P5/P6 (embedder + GRU gates); stamp 12 = step end.                                              */
int v2f_decode_persist_stamps_enable(int on);
/* Timing experiments only (results become invalid): bit 0 skips the activation loads of the products,
 * bit 1 their MMAs.  Default 0.                                                                   */
int v2f_decode_persist_debug(int bits);
/* Host logic of the persistent decoder, callable without a GPU: which output columns CTA c of a G-CTA grid owns.
 * out[11] = a_lo, na (columns of S that feed the attention queries), u_lo, nu (hidden units: gh columns of S and the
 * gate columns of GI), n1 = na + 3 nu, m3 (0 image / 1 trend context rows in P3), e_lo3, n3 (columns of HC),
 * x_lo, nx (columns of CTX), n5 = nx + 3 nu.  The kernel supports a configuration when every CTA has
 * n1 <= 24, n3 <= 8, n5 <= 16, 1 <= nu <= 4.                                                       */
int v2f_decode_persist_ownership(int c, int G, int E, int H, int* out);
long long v2f_decode_persist_stamps_offset(int N, int E, int H);

/* ------------------------------------------------------------------------------------------
 * Single-layer batch-first GRU over a sequence (nn.GRU, gate order r,z,n):
 *   TSEmbedder  models/CrossAttnRNN210.py:13-24  (52 steps, input 3)
 *   sales_encoder_gru :123,182 / SalesEncoder models/GTM_Visuelle2.py:99-107 (2 steps, input 1)
 * x [N,L,I]; h0 [N,H] or NULL (zeros); out [N,L,H]; saved RZN [L,N,3H], GHN [L,N,H];
 * GI [N,L,3H] scratch.  precision 0: fp32 CUDA-core products; 1: tf32 tcgen05 products (bwd then
 * needs scratch w_hhT [H,3H] and ws for transposed stacks).                                   */
int v2f_gru_seq_fwd(int N, int L, int I, int H, const float* x, const float* h0,
                    const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                    float* out, float* GI, float* GH, float* RZN, float* GHN, int precision,
                    void* stream);
/* dOut [N,L,H] (may be NULL), dhL [N,H] (may be NULL).  Scratch: dh [N,H], DGI [N,L,3H],
 * DGH [L,N,3H], Hprev [L,N,H].  Outputs (any may be NULL): dx [N,L,I], dh0 [N,H], dw_ih, dw_hh,
 * db_ih, db_hh.                                                                              */
/* Both run the whole recurrence as ONE cooperative launch with W_hh resident in shared memory
 * (csrc/gru_persist.cu; exact fp32, H % 16 == 0, H <= 512, L >= 2), else one GEMM + gate kernel per
 * step.  v2f_gru_persistent_enable(0) forces the per-step path (A/B tests).                     */
int v2f_gru_persistent_enable(int on);
int v2f_gru_seq_bwd(int N, int L, int I, int H, const float* x, const float* h0,
                    const float* w_ih, const float* w_hh, const float* out, const float* RZN,
                    const float* GHN, const float* dOut, const float* dhL, float* dh, float* DGI,
                    float* DGH, float* Hprev, float* dx, float* dh0, float* dw_ih, float* dw_hh,
                    float* db_ih, float* db_hh, float* w_hhT, float* ws, long long ws_floats,
                    int precision, void* stream);

/* ------------------------------------------------------------------------------------------
 * Scaled-dot-product attention core for short sequences (Lq,Lk <= 64), one CTA per
 * (batch, head): nn.MultiheadAttention ts_self_attention models/CrossAttnRNN210.py:126,176-179;
 * nn.TransformerEncoder/DecoderLayer attention in models/GTM_Visuelle2.py:52-53,200-202.
 * q/k/v/o are addressed as base + b*bstride + l*ld + head*hd.  mask: additive [Lq,Lk] or NULL.
 * drop: [B,heads,Lq,Lk] multiplicative keep-mask already scaled by 1/(1-p), or NULL.
 * P (saved softmax, before dropout): [B,heads,Lq,Lk].                                         */
int v2f_sdpa_fwd(int B, int heads, int Lq, int Lk, int hd, const float* q, int ldq, long long bsq,
                 const float* k, int ldk, long long bsk, const float* v, int ldv, long long bsv,
                 float* o, int ldo, long long bso, const float* mask, const float* drop, float* P,
                 float scale, void* stream);
int v2f_sdpa_bwd(int B, int heads, int Lq, int Lk, int hd, const float* q, int ldq, long long bsq,
                 const float* k, int ldk, long long bsk, const float* v, int ldv, long long bsv,
                 const float* dO, int ldo, long long bso, const float* drop, const float* P,
                 float* dq, int lddq, long long bsdq, float* dk, int lddk, long long bsdk, float* dv,
                 int lddv, long long bsdv, float scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Static embedders: TemporalFeatureEncoder + AttributeEncoder, models/CrossAttnRNN210.py:26-56
 * (Demand copy: CrossAttnRNNDemand.py:47-68 routes all four features through day_embedding:
 * pass the day weights four times).  out [B,2,E] = (date, attributes).
 * Wt,bt [4,E]; tables: 4 embedding tables [rows_k,E]; idx [4,B] int64; drop [B,8,E] keep-mask
 * (already scaled) or NULL.                                                                  */
int v2f_embed_fwd(int B, int E, const float* temporal, const float* Wt, const float* bt,
                  const float* const* tables, const long long* idx, const float* drop, float* out,
                  void* stream);
int v2f_embed_bwd(int B, int E, const float* temporal, const long long* idx, const float* drop,
                  const float* dout, const int* table_rows, float* dWt, float* dbt,
                  float* const* dtables, void* stream);

/* out = x * m : application of a dropout keep-mask (already scaled by 1/(1-p)); nn.Dropout in
 * TSEmbedder / ImageEncoder, models/CrossAttnRNN210.py:21-24,67-72.                           */
int v2f_mul_f32(long long n, const float* x, const float* m, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row-local operators of the GTM family (GTM_Visuelle2, Proposed_model v1-v4); csrc/gtm_ops.cu.
 * ------------------------------------------------------------------------------------------ */
/* y = LayerNorm(x + a*m) * gamma + beta over the last dim D (<= 1024) of [M,D]; a, m optional
 * (m: dropout keep-mask of the residual branch, already scaled).  Saves xhat [M,D], rstd [M].
 * Post-LN residuals of nn.TransformerEncoderLayer/DecoderLayer (models/GTM_Visuelle2.py:52-53,
 * 200-202), GatedResidualBlock.norm (models/Proposed_model.py:147-155), fusion_fc LayerNorm
 * (models/Proposed_model_v4.py:177).                                                          */
int v2f_add_ln_fwd(int M, int D, const float* x, const float* a, const float* m, const float* gamma,
                   const float* beta, float eps, float* y, float* xhat, float* rstd, void* stream);
/* Number of partial-sum blocks the backward uses: part must hold blocks*2*D floats. */
int v2f_add_ln_bwd_blocks(int M);
/* dx [M,D] (= gradient of x), da [M,D] or NULL (= dx*m), dgb [2,D] = (dgamma, dbeta). */
int v2f_add_ln_bwd(int M, int D, const float* dy, const float* xhat, const float* rstd,
                   const float* gamma, const float* m, float* dx, float* da, float* part, float* dgb,
                   void* stream);
/* nn.BatchNorm1d over [B,D] (models/GTM_Visuelle2.py:158; FusionBlock models/Proposed_model_v3.py:163):
 * training!=0: batch statistics (biased variance) and running-stat update (unbiased variance,
 * momentum); else running statistics.  Saves the mean / rstd used, [D] each.                   */
int v2f_bn1d_fwd(int B, int D, const float* x, const float* gamma, const float* beta, float* run_mean,
                 float* run_var, int training, float momentum, float eps, float* y, float* save_mean,
                 float* save_rstd, void* stream);
int v2f_bn1d_bwd(int B, int D, const float* x, const float* dy, const float* gamma,
                 const float* save_mean, const float* save_rstd, int training, float* dx,
                 float* dgamma, float* dbeta, void* stream);
/* The same under batch sharding (data parallelism; the reference normalises over the WHOLE batch): per-channel
 * sums of the local rows, sums[2,D] doubles = (sum x, sum x^2) resp. (sum dy, sum dy*xhat); the caller all-reduces
 * them (NCCL, SUM) between the stats and the apply call and passes the global row count Btot. */
int v2f_bn1d_stats(int B, int D, const float* x, double* sums, void* stream);
int v2f_bn1d_apply(int B, int D, const float* x, const float* gamma, const float* beta, const double* sums,
                   double Btot, float* run_mean, float* run_var, float momentum, float eps, float* y,
                   float* save_mean, float* save_rstd, void* stream);
int v2f_bn1d_bwd_stats(int B, int D, const float* x, const float* dy, const float* save_mean,
                       const float* save_rstd, double* sums, float* dgamma, float* dbeta, void* stream);
int v2f_bn1d_bwd_apply(int B, int D, const float* x, const float* dy, const float* gamma, const float* save_mean,
                       const float* save_rstd, const double* sums, double Btot, float* dx, void* stream);
/* nn.Dropout (models/CrossAttnRNN210.py:72 and every encoder / fusion network): out = x * keep / (1 - p) with the keep
 * decisions generated inside the kernel (Philox4x32-10, counter = element index / 4, key = key[0], key[1] read from
 * DEVICE memory).  Calling it again on the incoming gradient with the same key is the backward pass: no mask tensor. */
int v2f_dropout(long long n, const float* x, const unsigned long long* key, float p, float* out, void* stream);
/* Sigmoid gates: mode 0: out = x*sigmoid(g) (models/Proposed_model.py:217, _v2.py:598,682,
 * _v3.py:222-227); mode 1: out = x + x*sigmoid(g) (Proposed_model.py:154, _v2.py:635, _v4.py:186-192). */
int v2f_gate_fwd(long long n, const float* x, const float* g, int mode, float* out, void* stream);
int v2f_gate_bwd(long long n, const float* x, const float* g, const float* dout, int mode, float* dx,
                 float* dg, void* stream);
/* out = a + b (decoder_input = sales_base + static_context, models/GTM_Visuelle2.py:246-247). */
int v2f_add_f32(long long n, const float* a, const float* b, float* out, void* stream);
/* out = max(x, 0) (ReLU after LayerNorm in fusion_fc, models/Proposed_model_v4.py:176-179) and
 * out = dy where y > 0 else 0 (its backward, also the backward of the ReLU epilogue of v2f_gemm_*). */
int v2f_relu_fwd(long long n, const float* x, float* out, void* stream);
int v2f_relu_bwd(long long n, const float* dy, const float* y, float* out, void* stream);
/* out[r, i] = x[r, i] + p[i], i < n: PositionalEncoding.forward (models/GTM_Visuelle2.py:26-28). */
int v2f_add_bcast(long long rows, long long n, const float* x, const float* p, float* out, void* stream);
/* Strided 2-D copy (column concatenation / slicing: torch.cat of the fusion networks). */
int v2f_copy2d(int rows, int cols, const float* src, long long lds, float* dst, long long ldd, void* stream);
/* repeat_interleave over windows, out[b*W+w,:] = x[b,:] (models/GTM_Visuelle2.py:231-235), and
 * its gradient dx[b,:] = sum_w dout[b*W+w,:].                                                  */
int v2f_repeat_rows(int B, int W, long long D, const float* x, float* out, void* stream);
int v2f_fold_rows(int B, int W, long long D, const float* dout, float* dx, void* stream);
/* AttributeEncoder of the GTM family (models/GTM_Visuelle2.py:81-96): out [B,4,E] = stacked rows of
 * the four embedding tables (idx [4,B] int64), times the optional keep-mask drop [B,4,E].        */
int v2f_gather4_fwd(int B, int E, const float* const* tables, const long long* idx, const float* drop,
                    float* out, void* stream);
int v2f_gather4_bwd(int B, int E, const long long* idx, const float* drop, const float* dout,
                    const int* table_rows, float* const* dtables, void* stream);
/* DummyEmbedder / TemporalEmbedder front (models/GTM_Visuelle2.py:129-145):
 * out [B,4,E], out[b,k,:] = temporal[b,k] * Wt[k,:] + bt[k,:].                                  */
int v2f_feat4_fwd(int B, int E, const float* temporal, const float* Wt, const float* bt, float* out,
                  void* stream);
int v2f_feat4_bwd(int B, int E, const float* temporal, const float* dout, float* dWt, float* dbt,
                  void* stream);
/* Global average pool of the trunk's feature map (ImageEncoder.pool, models/GTM_Visuelle2.py:117,
 * 123-125), taken before the 1x1 projection (SURVEY.md 8a identity 5).  layout 0: x [B,C,L];
 * layout 1: x [B,L,C] (channels_last).  kind 0: bf16, 1: fp32.  out/dout [B,C] fp32; dx as x.    */
int v2f_meanpool_fwd(int B, int L, int C, const void* x, int layout, int kind, float* out, void* stream);
int v2f_meanpool_bwd(int B, int L, int C, const float* dout, int layout, int kind, void* dx, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused BatchNorm2d (+ residual add) (+ ReLU) over bf16 channels_last activations viewed as
 * [R = N*H*W, C] (C % 8 == 0): the normalisation / add / ReLU between the cuDNN convolutions of the
 * torchvision ResNet-101 trunk (ImageEncoder, models/CrossAttnRNN210.py:58-72; Bottleneck =
 * conv-bn-relu, conv-bn-relu, conv-bn, +identity, relu).  csrc/bn_act.cu.
 *   y = act(BN(x) (+ res));  training != 0: batch statistics (biased variance), running statistics
 *   updated with `momentum` (unbiased variance) unless run_mean is NULL; else running statistics.
 * x, res, y, dy, dz, dx: bf16 [R,C].  save_mean/save_rstd [C]; scale_shift [2,C] scratch;
 * part: v2f_bn2d_blocks(R,C)*2*C floats of scratch; coef [3,C] scratch.
 * Backward: dy2 (optional) is a second upstream gradient -- the output fed two consumers (the next
 * block's conv1 and its residual add) -- summed on the fly instead of by a separate kernel (needs dz).
 * dz = (dy+dy2)*[y>0] (relu) is what a residual branch receives; pass dz != NULL to have it
 * stored (bf16 [R,C]).  dx = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)) (training) or
 * gamma*rstd*dz (eval); dgamma = sum dz*xhat, dbeta = sum dz.                                   */
int v2f_bn2d_blocks(long long R, int C);
/* A/B switch (default on): the kernels of one call (statistics -> finalize -> apply; reduce -> finalize -> elemt) are
 * chained by programmatic dependent launch, so each successor's launch overlaps its predecessor's tail. */
int v2f_bn2d_pdl_enable(int on);
int v2f_bn2d_act_fwd(long long R, int C, const void* x, const void* res, const float* gamma,
                     const float* beta, float* run_mean, float* run_var, int training, float momentum,
                     float eps, int relu, void* y, float* save_mean, float* save_rstd,
                     float* scale_shift, float* part, void* stream);
int v2f_bn2d_act_bwd(long long R, int C, const void* dy, const void* dy2, const void* x, const void* y,
                     const float* gamma, const float* save_mean, const float* save_rstd, int training,
                     int relu, void* dz, void* dx, float* dgamma, float* dbeta, float* coef, float* part,
                     void* stream);
/* Stem of the trunk (conv1 -> bn1 -> relu -> maxpool 3x3 stride 2 pad 1, torchvision resnet.py):
 * y [N,OH,OW,C] = maxpool(relu(BN(x))), x bf16 [N,H,W,C], OH=(H-1)/2+1, OW=(W-1)/2+1.  Forward only:
 * the stem is frozen in the reference (models/CrossAttnRNN210.py:62-65).  Scratch as v2f_bn2d_act_fwd. */
int v2f_bn2d_relu_maxpool_fwd(int N, int H, int W, int C, const void* x, const float* gamma,
                              const float* beta, float* run_mean, float* run_var, int training,
                              float momentum, float eps, void* y, float* save_mean, float* save_rstd,
                              float* scale_shift, float* part, void* stream);
/* Same, with the batch statistics already reduced to `part_blocks` partial rows [part_blocks,2,C] (sum, sum of
 * squares over disjoint row sets) by the producer of x -- the epilogue of v2f_stem_conv_fwd -- so no statistics
 * sweep over x runs here (training != 0 only). */
int v2f_bn2d_relu_maxpool_fwd_parts(int N, int H, int W, int C, const void* x, const float* gamma,
                                    const float* beta, float* run_mean, float* run_var, float momentum, float eps,
                                    void* y, float* save_mean, float* save_rstd, float* scale_shift,
                                    const float* part, int part_blocks, void* stream);
/* Stem convolution conv1 = Conv2d(3, 64, 7, stride 2, padding 3, bias=False) of the torchvision ResNet trunk
 * (models/CrossAttnRNN210.py:58-65: frozen, forward only) as an implicit GEMM on tcgen05 (csrc/stem_conv.cu).
 * x: [N,H,W,3] NHWC (a channels_last [N,3,H,W] tensor) or, with x_nchw, plain [N,3,H,W]; fp32 or bf16 (x_bf16); wpk: bf16 [64,192], the weight packed
 * as wpk[o, kh*24 + kw*3 + c] = w[o,c,kh,kw], zero elsewhere; y: bf16 [N,OH,OW,64] NHWC with OH=(H-1)/2+1,
 * OW=(W-1)/2+1 (values: fp32 accumulation of bf16 products, rounded to bf16 -- what a bf16 library convolution
 * returns).  part (optional): [v2f_stem_conv_blocks(...), 2, 64] floats, per-CTA sum / sum of squares of the stored
 * (rounded) outputs for v2f_bn2d_relu_maxpool_fwd_parts.  v2f_stem_conv_blocks returns 0 for unsupported shapes
 * (OW < 128 or 3 W > 1024): keep the library convolution there. */
int v2f_stem_conv_blocks(int N, int H, int W, int x_bf16, int x_nchw);
int v2f_stem_conv_fwd(int N, int H, int W, const void* x, int x_bf16, int x_nchw, const void* wpk, void* y,
                      float* part, void* stream);
/* Diagnostics: registers, dynamic shared memory and the CTAs/SM the runtime's occupancy calculator reports for the
 * stem kernel at image width W (it reports 1; the launch uses 2 per SM, which the hardware co-schedules). */
int v2f_stem_conv_occupancy(int W, int* regs, int* smem_bytes, int* per_sm);

/* ------------------------------------------------------------------------------------------
 * Device side of the image transform of dataset_fusion.py:50-65 (ToTensor + Normalize; decode and Resize stay on
 * the host): uint8 NHWC pixels in, (float(u8)/255 - mean[c]) / std[c] out as bf16 or fp32 in the same NHWC order,
 * i.e. a channels_last [B,3,H,W] tensor -- what the bf16 trunk consumes.  The batch then crosses PCIe as uint8
 * (34 MB per 128 items instead of 137 MB).  in/out 16-byte aligned; mean, std: device float[C].          */
int v2f_image_normalize_u8(long long pixels, int C, const void* in, const float* mean, const float* stdv,
                           int out_bf16, void* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-tensor Adafactor step (csrc/adafactor.cu): configure_optimizers of the LightningModules,
 * models/CrossAttnRNN210.py:229-230, models/GTM_Visuelle2.py:264-266 (fairseq Adafactor with scale_parameter,
 * relative_step, warmup_init; beta1 = None, weight_decay = 0).  One descriptor per trainable tensor that has a
 * gradient; ``kind``: 0 vector (state sq[numel]), 1 small matrices (nmat matrices of R x C, R,C <= 4: convolution
 * weights), 2 big matrices (nmat x R x C, factored states row[nmat,R], col[nmat,C]).  Work-unit tables (int4, device):
 *   vec_units   (desc, first element, count <= 1024, 0)     small_units (desc, first matrix, count <= 256, 0)
 *   row_units   (desc, matrix, row, 0)                      col_units   (desc, matrix, first column of a 32 strip, 0)
 * grads: device array of n_desc gradient pointers, refreshed by the caller every step; acc: [n_desc][2] doubles.
 * beta2t = 1 - step^decay_rate and rel_step = min(1e-6 step, 1/sqrt(step)) (or the fixed lr) are computed by the
 * caller, as the reference does on the host.                                                              */
typedef struct v2f_af_desc {
  float *p, *row, *col, *sq, *rms;
  long long numel;
  int nmat, R, C, kind;
  /* kind 1 only: element (m, r, c) of the parameter AND of its gradient lives at
   * (m / inner) * sO + (m % inner) * sI + r * sR + c * sC  -- row-major [O,I,R,C]: inner = I, (I*R*C, R*C, C, 1);
   * channels_last convolution weights (the bf16 trunk keeps them so): (I*R*C, 1, C*I, I).  States stay row-major. */
  int inner, pad_;
  long long sO, sI, sR, sC;
} v2f_af_desc;
typedef struct v2f_adafactor_plan {
  const v2f_af_desc* descs;
  const void* grads;
  double* acc;
  const void *vec_units, *small_units[3], *row_units, *col_units; /* small: [0] 1x1, [1] 3x3, [2] other shapes */
  int n_desc, n_vec, n_small[3], n_rows, n_cols;
  float eps1, eps2, clip_threshold;
  int scale_parameter;
} v2f_adafactor_plan;
int v2f_adafactor_step(const v2f_adafactor_plan* plan, double beta2t, double rel_step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* V2F_H_ */

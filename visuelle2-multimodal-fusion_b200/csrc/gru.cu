// Single-layer batch-first GRU over a sequence, forward and BPTT.  sm_100a.
//
// Reference arithmetic: torch.nn.GRU (gate order r,z,n) as used by TSEmbedder
// (/root/reference/models/CrossAttnRNN210.py:13-24), sales_encoder_gru (:123,182) and
// SalesEncoder (/root/reference/models/GTM_Visuelle2.py:99-107).
//   r = s(W_ir x + b_ir + W_hr h + b_hr), z likewise, n = tanh(W_in x + b_in + r*(W_hn h + b_hn)),
//   h' = (1-z) n + z h
// The input projection of all steps is one GEMM; the recurrent part is a GEMM + gate kernel per
// step (v0).  Saved for BPTT: r,z,n and gh_n.
#include "gemm_dispatch.cuh"

namespace v2f {

// gru_persist.cu: one cooperative launch for the whole recurrence, W_hh resident in shared memory
bool gru_persist_supported(int N, int L, int H);
int gru_persist_fwd(int N, int L, int H, const float* GI, const float* h0, const float* w_hh, const float* b_hh,
                    float* out, float* RZN, float* GHN, int precision, cudaStream_t s);
int gru_persist_bwd(int N, int L, int H, const float* h0, const float* w_hh, const float* out, const float* RZN,
                    const float* GHN, const float* dOut, const float* dhL, float* DGI, float* DGH, float* Hprev,
                    float* dh_out, int precision, cudaStream_t s);

__global__ void __launch_bounds__(256)
gru_gates_fwd_kernel(int N, int L, int H, int t, const float* __restrict__ GI,
                     const float* __restrict__ GH, const float* __restrict__ hprev, int ldh,
                     float* __restrict__ out, float* __restrict__ RZN, float* __restrict__ GHN) {
  const int n = blockIdx.x;
  const float* gi = GI + ((long long)n * L + t) * 3 * H;
  const float* gh = GH + (long long)n * 3 * H;
  for (int u = threadIdx.x; u < H; u += blockDim.x) {
    const float r = sigmoid_full(gi[u] + gh[u]);
    const float z = sigmoid_full(gi[H + u] + gh[H + u]);
    const float ghn = gh[2 * H + u];
    const float c = tanh_full(gi[2 * H + u] + r * ghn);
    const float hp = hprev[(long long)n * ldh + u];
    float* rzn = RZN + ((long long)t * N + n) * 3 * H;
    rzn[u] = r;
    rzn[H + u] = z;
    rzn[2 * H + u] = c;
    GHN[((long long)t * N + n) * H + u] = ghn;
    out[((long long)n * L + t) * H + u] = (1.f - z) * c + z * hp;
  }
}

__global__ void __launch_bounds__(256)
gru_gates_bwd_kernel(int N, int L, int H, int t, const float* __restrict__ RZN,
                     const float* __restrict__ GHN, const float* __restrict__ hprev, int ldh,
                     const float* __restrict__ dOut, float* __restrict__ dh,
                     float* __restrict__ DGI, float* __restrict__ DGH, float* __restrict__ Hprev) {
  const int n = blockIdx.x;
  const float* rzn = RZN + ((long long)t * N + n) * 3 * H;
  for (int u = threadIdx.x; u < H; u += blockDim.x) {
    const float r = rzn[u], z = rzn[H + u], c = rzn[2 * H + u];
    const float ghn = GHN[((long long)t * N + n) * H + u];
    const float hp = hprev[(long long)n * ldh + u];
    float dhp = dh[(long long)n * H + u];
    if (dOut) dhp += dOut[((long long)n * L + t) * H + u];
    const float dan = dhp * (1.f - z) * (1.f - c * c);
    const float daz = dhp * (hp - c) * z * (1.f - z);
    const float dar = dan * ghn * r * (1.f - r);
    float* dgi = DGI + ((long long)n * L + t) * 3 * H;
    dgi[u] = dar;
    dgi[H + u] = daz;
    dgi[2 * H + u] = dan;
    float* dgh = DGH + ((long long)t * N + n) * 3 * H;
    dgh[u] = dar;
    dgh[H + u] = daz;
    dgh[2 * H + u] = dan * r;
    Hprev[((long long)t * N + n) * H + u] = hp;
    dh[(long long)n * H + u] = dhp * z;
  }
}

}  // namespace v2f

using namespace v2f;

extern "C" int v2f_gru_seq_fwd(int N, int L, int I, int H, const float* x, const float* h0,
                               const float* w_ih, const float* w_hh, const float* b_ih,
                               const float* b_hh, float* out, float* GI, float* GH, float* RZN,
                               float* GHN, int precision, void* st) {
  V2F_REQUIRE(N > 0 && L > 0 && I > 0 && H > 0, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(x && h0 && w_ih && w_hh && b_ih && b_hh && out && GI && GH && RZN && GHN, V2F_ERR_BAD_ARG);
  cudaStream_t s = (cudaStream_t)st;
  const GemmCtx gx{precision, nullptr, 0, st};
  V2F_TRY(gemm_nt(gx, N * L, 3 * H, I, x, I, w_ih, I, GI, 3 * H, b_ih, 0.f));
  if (gru_persist_supported(N, L, H)) return gru_persist_fwd(N, L, H, GI, h0, w_hh, b_hh, out, RZN, GHN, precision, s);
  for (int t = 0; t < L; t++) {
    const float* hp = t == 0 ? h0 : out + (long long)(t - 1) * H;
    const int ldh = t == 0 ? H : L * H;
    V2F_TRY(gemm_nt(gx, N, 3 * H, H, hp, ldh, w_hh, H, GH, 3 * H, b_hh, 0.f));
    gru_gates_fwd_kernel<<<N, 256, 0, s>>>(N, L, H, t, GI, GH, hp, ldh, out, RZN, GHN);
    V2F_CHECK_LAUNCH();
  }
  return V2F_OK;
}

extern "C" int v2f_gru_seq_bwd(int N, int L, int I, int H, const float* x, const float* h0,
                               const float* w_ih, const float* w_hh, const float* out,
                               const float* RZN, const float* GHN, const float* dOut,
                               const float* dhL, float* dh, float* DGI, float* DGH, float* Hprev,
                               float* dx, float* dh0, float* dw_ih, float* dw_hh, float* db_ih,
                               float* db_hh, float* w_hhT, float* ws, long long ws_floats,
                               int precision, void* st) {
  V2F_REQUIRE(N > 0 && L > 0 && I > 0 && H > 0, V2F_ERR_BAD_ARG);
  V2F_REQUIRE(x && h0 && w_ih && w_hh && out && RZN && GHN && dh && DGI && DGH && Hprev, V2F_ERR_BAD_ARG);
  cudaStream_t s = (cudaStream_t)st;
  const GemmCtx gx{precision, ws, ws_floats, st};
  const bool tc = precision != 0 && w_hhT;
  const bool persist = gru_persist_supported(N, L, H);
  if (persist) {
    V2F_TRY(gru_persist_bwd(N, L, H, h0, w_hh, out, RZN, GHN, dOut, dhL, DGI, DGH, Hprev, dh, precision, s));
  } else {
  if (tc) V2F_TRY(v2f_transpose(3 * H, H, w_hh, H, 1, w_hhT, 3 * H, 1, st));
  if (dhL)
    cudaMemcpyAsync(dh, dhL, sizeof(float) * (size_t)N * H, cudaMemcpyDeviceToDevice, s);
  else
    cudaMemsetAsync(dh, 0, sizeof(float) * (size_t)N * H, s);
  for (int t = L - 1; t >= 0; t--) {
    const float* hp = t == 0 ? h0 : out + (long long)(t - 1) * H;
    const int ldh = t == 0 ? H : L * H;
    gru_gates_bwd_kernel<<<N, 256, 0, s>>>(N, L, H, t, RZN, GHN, hp, ldh, dOut, dh, DGI, DGH, Hprev);
    V2F_CHECK_LAUNCH();
    // dh += DGH[t] W_hh
    V2F_TRY(gemm_nn(gx, N, H, 3 * H, DGH + (long long)t * N * 3 * H, 3 * H, w_hh, H, tc ? w_hhT : nullptr,
                    3 * H, dh, H, 1.f));
  }
  }
  if (dh0) cudaMemcpyAsync(dh0, dh, sizeof(float) * (size_t)N * H, cudaMemcpyDeviceToDevice, s);
  const int NL = N * L;
  if (dw_ih) V2F_TRY(gemm_tn(gx, 3 * H, I, NL, DGI, 3 * H, x, I, dw_ih, I));
  if (db_ih) V2F_TRY(v2f_colsum_f32(NL, 3 * H, DGI, 3 * H, db_ih, 0.f, st));
  if (dw_hh) V2F_TRY(gemm_tn(gx, 3 * H, H, NL, DGH, 3 * H, Hprev, H, dw_hh, H));
  if (db_hh) V2F_TRY(v2f_colsum_f32(NL, 3 * H, DGH, 3 * H, db_hh, 0.f, st));
  if (dx) V2F_TRY(gemm_nn(gx, NL, I, 3 * H, DGI, 3 * H, w_ih, I, nullptr, 0, dx, I, 0.f));
  return V2F_OK;
}

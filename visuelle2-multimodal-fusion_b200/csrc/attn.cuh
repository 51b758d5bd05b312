// Argument blocks of the attention step kernels + the streaming (TMA-staged) implementation's API.
#pragma once
#include "common.cuh"

namespace v2f {

struct AttnArgs {
  int N, W, E, Li, Lt, ldS;
  const float *Himg, *Vimg, *Htr, *Ptr;
  const float* S;       // [N, ldS] of this step: s_img | s_tr | s_mm | gh
  const float* w_att;   // [3,E]
  const float* beta_att;
  const float* b_tl;
  float* C;             // [N,2,E] of this step
  float *alpha_img, *alpha_tr;  // [N,Li], [N,Lt] of this step
  int mod_first;        // grid.y index 0 maps to modality mod_first (0 img, 1 trend)
  int approx;           // 1: MUFU.TANH in the streaming kernels (tensor-core precision mode)
};

struct AttnBwdArgs {
  int N, W, E, Li, Lt, ldS;
  const float *Himg, *Vimg, *Htr, *Ptr;
  const float* S;
  const float* w_att;
  const float* DC;                 // [N,2,E] of this step
  const float *alpha_img, *alpha_tr;
  float *DE_img, *DE_tr;           // [N,L] of this step
  float* DS;                       // [N,ldS] of this step (writes cols mod*E ..)
  float* dw_acc;                   // [N,3,E]
  int mod_first;
};


// ---- streaming implementation (attn_stream.cu): one persistent CTA per SM, bulk-copy ring,
// positions of all (row, modality) segments split evenly over the grid, partial softmax combined
// by a second tiny kernel.  Supported when E % 256 == 0 and E <= 1024.
bool attn_stream_supported(int E);
long long attn_stream_ws_floats(int N, int Li, int Lt, int E);
int attn_stream_fwd(const AttnArgs& a, bool use_img, bool use_tr, float* ws, cudaStream_t s);
int attn_stream_bwd(const AttnBwdArgs& a, const float* C, const float* b_tl, bool use_img, bool use_tr,
                    int approx, float* ws, cudaStream_t s);

}  // namespace v2f

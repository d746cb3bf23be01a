// Token-side tcgen05 kernels of the SIM cross-attention (sim_tc.cu).
#pragma once
#include "common.cuh"

namespace sig {

struct SimTcBufs {
  const float* maskf;        // [3][B][L]  (0/1)
  const float* qtatt;        // [B][24][d] folded queries, e = q*8 + h (scaled by 1/sqrt(hd))
  const float* catt;         // [B][24]    q_h . b_k^h / sqrt(hd)
  __nv_bfloat16* DXQT;       // [B][64][d] rows 0..23 dxbar (bwd), 32..55 qt (fwd); other rows zero
  float* S32;                // [B][384][32] logits
  __nv_bfloat16* Ptok;       // [B][384][32] P~ = softmax * mask
  __nv_bfloat16* PT;         // [B][32][384] P~ transposed
  float* xbar;               // [B][24][d]
  __nv_bfloat16* xbarb;      // bf16 shadow of xbar (written with it)
  const float* dxbar;        // [B][24][d]
  float* delta;              // [B][32]
  __nv_bfloat16* PdS;        // [B][384][64] [P~ | dS~]
  __nv_bfloat16* dST;        // [B][32][384] dS~ transposed
  float* dqt;                // [B][24][d]
  __nv_bfloat16* dqtb;       // bf16 shadow of dqt (written with it)
};

int sim_tc_tokens_fwd(const sig_tokens* tok, const SimTcBufs& k, cudaStream_t s);
int sim_tc_tokens_bwd(const sig_tokens* tok, const SimTcBufs& k, const sig_token_grads* dtok, cudaStream_t s);

}  // namespace sig

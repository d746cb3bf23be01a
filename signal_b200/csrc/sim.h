// internal C++ interface of the SIM kernels (sim.cu)
#pragma once
#include "common.cuh"

namespace sig {
size_t sim_ctx_bytes(int B, int L, int d);
size_t sim_ctx_bytes_for(int B, int L, int d, int dtype, unsigned flags);
int sim_fold_selection(const sig_sim_params* p, int d, void* m_hl, float* v, float* u, float* s0, float* ws, cudaStream_t s);
int convert_tokens(const sig_tokens* t, float* Xf, float* clsf, cudaStream_t s);
int write_token_grads(const sig_token_grads* g, int dtype, const float* dXf, const float* dclsf, int B, int L, int d,
                      cudaStream_t s);
int sim_forward(const sig_tokens* tok, const sig_sim_params* p, bool do_select, const float* ext_masks, int k1, int k2,
                int max_keep, void* out, float* masks_out, void* ctx, size_t ctx_bytes, unsigned flags, cudaStream_t s);
int sim_backward(const sig_tokens* tok, const sig_sim_params* p, bool has_masks, const void* dout, const sig_token_grads* dtok,
                 const sig_sim_param_grads* dp, void* ctx, size_t ctx_bytes, unsigned flags, cudaStream_t s);
int sim_select(const sig_tokens* tok, const sig_sim_params* p, int which, int k1, int k2, int max_keep, float* masks,
               void* selected, void* ctx, size_t ctx_bytes, cudaStream_t s);
int select_from_scores(const float* intra, const float* inter, const float* raw, int B, int L, int which, int k1, int k2,
                       int max_keep, float* masks, cudaStream_t s);
int sim_dx_operands(void* ctx, int B, int L, int d, int dtype, unsigned flags, void** pds, void** dxqt);
int mask_mul_bwd(const void* dselected, const float* masks, int dtype, int B, int L, int d, const sig_token_grads* g,
                 cudaStream_t s);
}  // namespace sig

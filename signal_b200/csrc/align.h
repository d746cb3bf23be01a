// internal C++ interface of the AlignmentM kernels (align.cu)
#pragma once
#include "common.cuh"

namespace sig {
size_t align_ctx_bytes(int B, int L, int d);
size_t das_ctx_bytes(int B, int L, int d);
size_t align_ctx_bytes_for(int B, int L, int d, int dtype, unsigned flags);
size_t volume_ws_floats(int B1, int B2);
// mean [3][B][d] fp32 = mean over the L patch rows of each modality (bf16 / fp32 tokens; the pooling kernels of the GAM path)
int align_pool_tokens(const sig_tokens* tok, float* mean, cudaStream_t s);
int align_patch_mean_slot(void* ctx, int B, int L, int d, int dtype, unsigned flags, float** slot);
int align_forward(const sig_tokens* tok, const sig_align_params* p, int h, int w, int do_lam, float* losses, void* ctx,
                  size_t ctx_bytes, unsigned flags, cudaStream_t s);
int align_backward(const sig_tokens* tok, const sig_align_params* p, int h, int w, int do_lam, const float* dlosses,
                   const sig_token_grads* dtok, const sig_align_param_grads* dp, void* ctx, size_t ctx_bytes, unsigned flags,
                   cudaStream_t s);
int das_forward(const void* x, int64_t sb, int64_t sl, int dtype, int B, int h, int w, int d, const sig_align_params* p, int m,
                float* sampled, void* ctx, size_t ctx_bytes, unsigned flags, cudaStream_t s);
int das_backward(const void* x, int64_t sb, int64_t sl, int dtype, int B, int h, int w, int d, const sig_align_params* p, int m,
                 const float* dsampled, void* dx, const sig_align_param_grads* dp, void* ctx, size_t ctx_bytes, unsigned flags,
                 cudaStream_t s);
int volume3_forward(const float* l, const float* v, const float* a, int B1, int B2, int d, float* vol, float* ws, cudaStream_t s);
int volume3_backward(const float* l, const float* v, const float* a, int B1, int B2, int d, const float* dvol, float* dl, float* dv,
                     float* da, float* ws, cudaStream_t s);
}  // namespace sig

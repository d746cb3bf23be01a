// extern "C" surface of libsignal_b200.so (see include/signal_b200.h).
#include "align.h"
#include "common.cuh"
#include "sim.h"
#include "tc_gemm.h"

extern "C" {

int sig_version(void) { return SIG_ABI_VERSION; }

const char* sig_error_string(int code) {
  switch (code) {
    case SIG_OK: return "ok";
    case SIG_ERR_NULL: return "signal_b200: a required pointer is NULL";
    case SIG_ERR_SHAPE: return "signal_b200: unsupported shape (B, L<=128, d%64==0, grid h*w==L with h,w multiples of 4 and >= 8, k >= 1)";
    case SIG_ERR_DTYPE: return "signal_b200: unknown dtype (SIG_F32 or SIG_BF16)";
    case SIG_ERR_ALIGN: return "signal_b200: pointer or stride not 16-byte aligned";
    case SIG_ERR_WORKSPACE: return "signal_b200: ctx/workspace buffer too small (see sig_ctx_bytes)";
    case SIG_ERR_ARCH: return "signal_b200: device is not sm_100";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "signal_b200: unknown error";
}

size_t sig_ctx_bytes(int kind, int B, int L, int d, int dtype, unsigned flags) {
  if (B < 1 || L < 1 || L > sig::kMaxL || d < 64 || d % 64) return 0;
  switch (kind) {
    case SIG_CTX_SIM: return sig::sim_ctx_bytes_for(B, L, d, dtype, flags);
    case SIG_CTX_SELECT: return sig::sim_ctx_bytes(B, L, d);
    case SIG_CTX_ALIGN: return sig::align_ctx_bytes_for(B, L, d, dtype, flags);
    case SIG_CTX_DAS: return sig::das_ctx_bytes(B, L, d);
    default: return 0;
  }
}

int sig_sim_fwd(const sig_tokens* tok, const sig_sim_params* p, int k1, int k2, int max_keep, void* out, float* masks, void* ctx,
                size_t ctx_bytes, unsigned flags, int device, void* stream) {
  SIG_ENTER(device);
  return sig::sim_forward(tok, p, true, nullptr, k1, k2, max_keep, out, masks, ctx, ctx_bytes, flags, (cudaStream_t)stream);
}

int sig_sim_bwd(const sig_tokens* tok, const sig_sim_params* p, const void* dout, const sig_token_grads* dtok,
                const sig_sim_param_grads* dp, void* ctx, size_t ctx_bytes, unsigned flags, int device, void* stream) {
  SIG_ENTER(device);
  return sig::sim_backward(tok, p, true, dout, dtok, dp, ctx, ctx_bytes, flags, (cudaStream_t)stream);
}

int sig_sim_fold_selection(const sig_sim_params* p, int d, void* m_hl, float* v, float* u, float* s0, void* ws, size_t ws_bytes,
                           int device, void* stream) {
  SIG_ENTER(device);
  if (!p || !p->sel_wq || !p->sel_bq || !p->sel_wk || !p->sel_bk || !m_hl || !v || !u || !s0 || !ws) return SIG_ERR_NULL;
  if (d < 64 || d % 64) return SIG_ERR_SHAPE;
  if (ws_bytes < (size_t)d * d * sizeof(float)) return SIG_ERR_WORKSPACE;
  return sig::sim_fold_selection(p, d, m_hl, v, u, s0, static_cast<float*>(ws), (cudaStream_t)stream);
}

int sig_sim_select_fwd(const sig_tokens* tok, const sig_sim_params* p, int which, int k1, int k2, int max_keep, float* masks,
                       void* selected, void* ctx, size_t ctx_bytes, int device, void* stream) {
  SIG_ENTER(device);
  return sig::sim_select(tok, p, which, k1, k2, max_keep, masks, selected, ctx, ctx_bytes, (cudaStream_t)stream);
}

int sig_sim_select_from_scores(const float* intra, const float* inter, const float* raw, int B, int L, int which, int k1, int k2,
                               int max_keep, float* masks, int device, void* stream) {
  SIG_ENTER(device);
  if (!masks) return SIG_ERR_NULL;
  if (which < 1 || which > 3 || B < 1 || L < 1 || L > sig::kMaxL || k1 < 1 || k2 < 1 || max_keep > L) return SIG_ERR_SHAPE;
  if ((which & 1) && !intra) return SIG_ERR_NULL;
  if ((which & 2) && !inter) return SIG_ERR_NULL;
  if (which == 3 && max_keep >= 0 && !raw) return SIG_ERR_NULL;
  return sig::select_from_scores(intra, inter, raw, B, L, which, k1, k2, max_keep, masks, (cudaStream_t)stream);
}

int sig_mask_mul_bwd(const void* dselected, const float* masks, int dtype, int B, int L, int d, const sig_token_grads* dtok,
                     int device, void* stream) {
  SIG_ENTER(device);
  return sig::mask_mul_bwd(dselected, masks, dtype, B, L, d, dtok, (cudaStream_t)stream);
}

int sig_sim_attn_fwd(const sig_tokens* tok, const sig_sim_params* p, const float* masks, void* out, void* ctx, size_t ctx_bytes,
                     unsigned flags, int device, void* stream) {
  SIG_ENTER(device);
  return sig::sim_forward(tok, p, false, masks, 0, 0, -1, out, nullptr, ctx, ctx_bytes, flags, (cudaStream_t)stream);
}

int sig_sim_attn_bwd(const sig_tokens* tok, const sig_sim_params* p, const float* masks, const void* dout,
                     const sig_token_grads* dtok, const sig_sim_param_grads* dp, void* ctx, size_t ctx_bytes, unsigned flags,
                     int device, void* stream) {
  SIG_ENTER(device);
  return sig::sim_backward(tok, p, masks != nullptr, dout, dtok, dp, ctx, ctx_bytes, flags, (cudaStream_t)stream);
}

int sig_align_fwd(const sig_tokens* tok, const sig_align_params* p, int h, int w, int do_lam, float* losses, void* ctx,
                  size_t ctx_bytes, unsigned flags, int device, void* stream) {
  SIG_ENTER(device);
  return sig::align_forward(tok, p, h, w, do_lam, losses, ctx, ctx_bytes, flags, (cudaStream_t)stream);
}

int sig_align_bwd(const sig_tokens* tok, const sig_align_params* p, int h, int w, int do_lam, const float* dlosses,
                  const sig_token_grads* dtok, const sig_align_param_grads* dp, void* ctx, size_t ctx_bytes, unsigned flags,
                  int device, void* stream) {
  SIG_ENTER(device);
  return sig::align_backward(tok, p, h, w, do_lam, dlosses, dtok, dp, ctx, ctx_bytes, flags, (cudaStream_t)stream);
}

int sig_das_fwd(const void* x, int64_t stride_b, int64_t stride_l, int dtype, int B, int h, int w, int d,
                const sig_align_params* p, int m, float* sampled, void* ctx, size_t ctx_bytes, unsigned flags, int device,
                void* stream) {
  SIG_ENTER(device);
  return sig::das_forward(x, stride_b, stride_l, dtype, B, h, w, d, p, m, sampled, ctx, ctx_bytes, flags, (cudaStream_t)stream);
}

int sig_das_bwd(const void* x, int64_t stride_b, int64_t stride_l, int dtype, int B, int h, int w, int d,
                const sig_align_params* p, int m, const float* dsampled, void* dx, const sig_align_param_grads* dp, void* ctx,
                size_t ctx_bytes, unsigned flags, int device, void* stream) {
  SIG_ENTER(device);
  return sig::das_backward(x, stride_b, stride_l, dtype, B, h, w, d, p, m, dsampled, dx, dp, ctx, ctx_bytes, flags,
                           (cudaStream_t)stream);
}

size_t sig_volume3_ws_bytes(int B1, int B2) {
  if (B1 < 1 || B2 < 1) return 0;
  return sig::volume_ws_floats(B1, B2) * sizeof(float);
}

int sig_volume3_fwd(const float* l, const float* v, const float* a, int B1, int B2, int d, float* vol, void* ws, size_t ws_bytes,
                    int device, void* stream) {
  SIG_ENTER(device);
  if (!l || !v || !a || !vol || !ws) return SIG_ERR_NULL;
  if (B1 < 1 || B2 < 1 || d < 1) return SIG_ERR_SHAPE;
  if (ws_bytes < sig_volume3_ws_bytes(B1, B2)) return SIG_ERR_WORKSPACE;
  return sig::volume3_forward(l, v, a, B1, B2, d, vol, static_cast<float*>(ws), (cudaStream_t)stream);
}

int sig_volume3_bwd(const float* l, const float* v, const float* a, int B1, int B2, int d, const float* dvol, float* dl, float* dv,
                    float* da, void* ws, size_t ws_bytes, int device, void* stream) {
  SIG_ENTER(device);
  if (!l || !v || !a || !dvol || !dl || !dv || !da || !ws) return SIG_ERR_NULL;
  if (B1 < 1 || B2 < 1 || d < 1) return SIG_ERR_SHAPE;
  if (ws_bytes < sig_volume3_ws_bytes(B1, B2)) return SIG_ERR_WORKSPACE;
  return sig::volume3_backward(l, v, a, B1, B2, d, dvol, dl, dv, da, static_cast<float*>(ws), (cudaStream_t)stream);
}

int sig_debug_gemm_bf16(const void* A, int a_mode, const int64_t* a_geom, const void* B, int b_mode, const int64_t* b_geom,
                        void* C, int64_t ldc, int out_bf16, const float* bias, int M, int N, int K, float alpha, int act,
                        int ksplit, int bn, int64_t c_stride_b, int64_t c_stride_l, const float* rowvec,
                        const float* rowvec_scale, int accumulate, int device, void* stream) {
  SIG_ENTER(device);
  if (!A || !B || !C || !a_geom || !b_geom) return SIG_ERR_NULL;
  sig::TcGemmDesc g = sig::tc_desc();
  auto fill = [](sig::TcOperand& o, const void* p, int mode, const int64_t* ge) {
    for (int i = 0; i < 8; ++i) o.ptr[i] = p; o.mode = mode; o.ld = ge[0]; o.stride_b = ge[1]; o.stride_l = ge[2];
    o.rows = ge[3]; o.cols = ge[4];
  };
  fill(g.A, A, a_mode, a_geom);
  fill(g.B, B, b_mode, b_geom);
  g.M = M; g.N = N; g.K = K; g.C[0] = C; g.ldc = ldc; g.out_bf16 = out_bf16; g.bias[0] = bias; g.alpha = alpha;
  g.act = act; g.ksplit = ksplit; g.bn = (bn == 512 || bn == 1024) ? 256 : bn; g.mt = bn == 512 ? 2 : 1;   // bn = 512: 256 x 256 units
  g.pair = bn == 1024;                                                                           // bn = 1024: 256 x 256 units on CTA pairs
  if (c_stride_b) { g.c_tok = 1; g.c_stride_b = c_stride_b; g.c_stride_l = c_stride_l; }
  g.rowvec[0] = rowvec; g.rowvec_scale = rowvec_scale; g.accumulate = accumulate;
  return sig::tc_gemm(g, (cudaStream_t)stream);
}

int sig_sim_dx_operands(void* ctx, int B, int L, int d, int dtype, unsigned flags, void** pds, void** dxqt) {
  return sig::sim_dx_operands(ctx, B, L, d, dtype, flags, pds, dxqt);
}

int sig_align_patch_mean_slot(void* ctx, int B, int L, int d, int dtype, unsigned flags, float** slot) {
  return sig::align_patch_mean_slot(ctx, B, L, d, dtype, flags, slot);
}

int sig_debug_tc_stamps(long long* out16) { return sig::tc_read_stamps(out16); }

int sig_infer_features(const void* const cls[3], const int64_t cls_stride_b[3], const void* sim_out, int64_t ld_sim, int dtype, int B,
                       int d, int normalize, float* out, int device, void* stream) {
  SIG_ENTER(device);
  if (!cls || !cls[0] || !cls[1] || !cls[2] || !cls_stride_b || !sim_out || !out) return SIG_ERR_NULL;
  if (dtype != SIG_F32 && dtype != SIG_BF16) return SIG_ERR_DTYPE;
  if (B < 1 || d < 1) return SIG_ERR_SHAPE;
  return sig::infer_features(cls, cls_stride_b, sim_out, ld_sim, dtype, B, d, normalize, out, (cudaStream_t)stream);
}

int sig_euclidean_distmat(const float* qf, const float* gf, int nq, int ng, int D, float* dist, void* ws, size_t ws_bytes, int device,
                          void* stream) {
  SIG_ENTER(device);
  if (!qf || !gf || !dist || !ws) return SIG_ERR_NULL;
  if (nq < 1 || ng < 1 || D < 1) return SIG_ERR_SHAPE;
  if (ws_bytes < (size_t)(nq + ng) * sizeof(float)) return SIG_ERR_WORKSPACE;
  return sig::euclidean_distmat(qf, gf, nq, ng, D, dist, static_cast<float*>(ws), (cudaStream_t)stream);
}

int sig_rank_eval(const float* dist, int64_t ld, const int64_t* q_pids, const int64_t* g_pids, const int64_t* q_camids,
                  const int64_t* g_camids, int nq, int ng, int max_rank, float* cmc, double* map_out, double* stats, int* overflow,
                  int device, void* stream) {
  SIG_ENTER(device);
  if (!dist || !q_pids || !g_pids || !q_camids || !g_camids || !cmc || !map_out || !stats || !overflow) return SIG_ERR_NULL;
  if (nq < 1 || ng < 1 || max_rank < 1 || max_rank > ng || ld < ng) return SIG_ERR_SHAPE;
  return sig::rank_eval(dist, ld, q_pids, g_pids, q_camids, g_camids, nq, ng, max_rank, cmc, map_out, stats, overflow,
                        (cudaStream_t)stream);
}

size_t sig_xchg_flag_bytes(void) { return sig::xchg_flag_bytes(); }

int sig_xchg_allreduce_f32(const sig_xchg_peers* peers, size_t off, size_t count, float scale, int ctas, int device, void* stream) {
  SIG_ENTER(device);
  return sig::xchg_allreduce_f32(peers, off, count, scale, ctas, (cudaStream_t)stream);
}

int sig_convert_half(const void* src, int64_t src_stride_b, int64_t src_stride_l, int src_dtype, void* dst, int64_t dst_stride_b,
                     int64_t dst_stride_l, int dst_dtype, int nb, int nl, int d, int device, void* stream) {
  SIG_ENTER(device);
  if (!src || !dst) return SIG_ERR_NULL;
  if (!((src_dtype == SIG_F16 && dst_dtype == SIG_BF16) || (src_dtype == SIG_BF16 && dst_dtype == SIG_F16))) return SIG_ERR_DTYPE;
  if (nb < 1 || nl < 1 || d < 8 || d % 8) return SIG_ERR_SHAPE;
  if (((uintptr_t)src | (uintptr_t)dst) & 15 || (src_stride_b | src_stride_l | dst_stride_b | dst_stride_l) % 8) return SIG_ERR_ALIGN;
  return sig::convert_half(src, src_stride_b, src_stride_l, dst_dtype == SIG_BF16, dst, dst_stride_b, dst_stride_l, nb, nl, d,
                           (cudaStream_t)stream);
}

}  // extern "C"

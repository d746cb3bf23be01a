/*
 * signal_b200.h -- C ABI of libsignal_b200.so, the sm_100a implementation of the
 * Signal fusion head (SIM + GAM + LAM), forward and backward.
 *
 * The reference (maxingan2412/Signal) is pure PyTorch: it has no FFI for this
 * path.  The functions below are what a binding for the path would call; each
 * one names the reference Python function it replaces.  INTEGRATION.md shows
 * the ctypes stub and the nn.Module shims that sit on top.
 *
 * Conventions (all entry points)
 *   - extern "C", returns int: 0 = work enqueued; <0 = argument rejected
 *     (nothing launched), see sig_error_string; >0 = cudaError_t of a launch.
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch); the
 *     library allocates nothing, frees nothing and keeps no pointer after the
 *     call returns.  No host synchronisation, no default-stream use: all work
 *     is enqueued on `stream` (a cudaStream_t) of device ordinal `device`.
 *     Calls are re-entrant and CUDA-graph capturable.
 *   - token maps are given as strided views (element strides, channel stride
 *     must be 1): patches x[:,1:] and CLS x[:,0] of a [B,1+L,d] map
 *     (modeling/meta_arch.py:108-109) are consumed without a copy.
 *   - parameters are fp32 masters with the reference's state_dict layouts.
 *   - `ctx` is an opaque caller-owned buffer of sig_ctx_bytes(...) bytes that
 *     carries saved activations from a *_fwd call to the matching *_bwd call.
 */
#ifndef SIGNAL_B200_H
#define SIGNAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIG_ABI_VERSION 1

/* SIG_F16 is accepted by sig_convert_half only: fp16 callers (the reference's amp.autocast, engine/processor.py:165) are
 * bridged to the bf16 kernels by one conversion pass per token map at the module boundary. */
enum sig_dtype { SIG_F32 = 0, SIG_BF16 = 1, SIG_F16 = 2 };

enum sig_error {
  SIG_OK = 0,
  SIG_ERR_NULL = -1,       /* a required pointer is NULL */
  SIG_ERR_SHAPE = -2,      /* unsupported B / L / d / grid / k */
  SIG_ERR_DTYPE = -3,      /* unknown dtype enum */
  SIG_ERR_ALIGN = -4,      /* pointer or stride not 16-byte aligned */
  SIG_ERR_WORKSPACE = -5,  /* ctx buffer too small */
  SIG_ERR_ARCH = -6        /* device is not sm_100 (tcgen05 path requested) */
};

enum sig_ctx_kind { SIG_CTX_SIM = 0, SIG_CTX_ALIGN = 1, SIG_CTX_SELECT = 2, SIG_CTX_DAS = 3 };

/* sig_flags bits (argument `flags` of the fwd/bwd entry points) */
#define SIG_FLAG_FORCE_SIMT 1u /* bf16 inputs: use the fp32 SIMT kernels instead of tcgen05 (cross-check) */
/* sig_align_fwd / sig_align_bwd, tensor-core path with LAM: the part of the backward that is linear in the loss weight
 * and needs nothing from the caller's backward call (d(MSE) -> d(samples) -> d(offset logits) -> depthwise/GELU tail)
 * already runs inside the FORWARD call with a unit weight; the backward call, given the same flag and the same ctx,
 * skips it and applies d(loss)/d(lam) where the results are consumed.  Ignored on the fp32 SIMT path.  Only set it
 * when a backward call will follow (training). */
#define SIG_FLAG_EAGER_BWD 2u
/* sig_align_fwd / sig_align_bwd, tensor-core path: the caller runs SIM's calls concurrently on another stream
 * (signal_b200.FusionHead).  AlignM's persistent GEMM / ring kernels then size their grids for SIG_ALIGN_SMS SMs
 * (environment; default 108 of 148) instead of all of them, so the short kernels of SIM's dependency chain -- the critical
 * path of the fused step -- find free SMs; the d(patches) GEMM at the tail of the backward always takes all SMs. */
#define SIG_FLAG_SHARE_SMS 4u
/* sig_align_fwd, tensor-core path (bf16 tokens): the caller has already written the mean over the L patch rows of each
 * modality, fp32 [3][B][d], into the slot sig_align_patch_mean_slot() returns inside `ctx` (e.g. the by-product of
 * sig_tokens_fwd, which produced the token maps): the GAM pooling pass over the tokens (useB.py:84-86) is skipped. */
#define SIG_FLAG_PATCH_MEAN 8u

/* Three modality token maps, order RGB, NI, TI.
 * patch[m] -> element (b=0,l=0,c=0) of the [B,L,d] patch view, cls[m] -> (b=0,c=0)
 * of the [B,d] CLS view.  Strides in ELEMENTS. */
typedef struct sig_tokens {
  const void* patch[3];
  const void* cls[3];            /* may be NULL for the AlignM entry points */
  int64_t patch_stride_b[3];
  int64_t patch_stride_l[3];
  int64_t cls_stride_b[3];
  int32_t dtype;                 /* enum sig_dtype */
  int32_t B, L, d;
} sig_tokens;

/* Gradient destinations with the same geometry (dtype = tokens' dtype).
 * accumulate = 0: every addressed element is overwritten (rows of unselected
 * tokens get zeros); 1: the gradient is added to what is there. */
typedef struct sig_token_grads {
  void* dpatch[3];
  void* dcls[3];                 /* may be NULL (AlignM has no CLS input) */
  int64_t patch_stride_b[3];
  int64_t patch_stride_l[3];
  int64_t cls_stride_b[3];
  int32_t accumulate;
  int32_t zero_cls;              /* AlignM: also write zeros to dcls rows when dcls != NULL */
  /* Optional cross-stream ordering when SIM and AlignM backward run concurrently on two streams and
   * share one gradient map (cudaEvent_t, may be NULL): the call makes its stream wait on `wait_event`
   * before its first write to dpatch and records `done_event` after its last write to dpatch (the CLS
   * rows are only ever written by SIM; AlignM zero-fills them only when accumulate == 0).  The call
   * that records must be issued before the call that waits. */
  void* wait_event;
  void* done_event;
  /* Optional fusion of SIM's token gradient into AlignM's dX GEMM (bf16 tensor-core path, both modules on the
   * same three [B,1+L,d] maps).  SIM's d(patches) is a K = 64 product [P~ | dS~] . [dxbar ; qt] per sample;
   * instead of writing it (and AlignM re-reading it to add its own part) the two operands are handed to
   * AlignM, whose dX GEMM appends them as one more k-block.
   *   in SIM's struct:   fuse_skip_dx = 1 -> sig_sim_bwd prepares the operands in its ctx, does not touch dpatch,
   *                      and records done_event once they are complete;
   *   in AlignM's struct: fuse_pds / fuse_dxqt = the operand pointers (sig_sim_dx_operands) -> added to dpatch. */
  int32_t fuse_skip_dx;
  int32_t reserved_;
  const void* fuse_pds;
  const void* fuse_dxqt;
} sig_token_grads;

/* Optional cache of the frozen token_selection parameters folded together (they never receive a
 * gradient -- selection is not differentiable, useA.py:79-93 -- so the product only changes when a
 * checkpoint is loaded):  M = W_k^T W_q,  v = W_k^T b_q,  u = W_q^T b_k,  s0 = b_q . b_k,  so that
 * W_k^T (W_q cls + b_q) = M cls + v  and  (W_q cls + b_q) . b_k = u . cls + s0  (useA.py:123-128).
 * Filled by sig_sim_fold_selection; the caller re-runs it whenever those four tensors change. */
typedef struct sig_sel_fold {
  const void* m_hl;    /* bf16 [d, 2d]: row n = [hi(M[n,:]) | lo(M[n,:])], M = hi + lo (split bf16) */
  const float* v;      /* [d] */
  const float* u;      /* [d] */
  const float* s0;     /* [1] */
} sig_sel_fold;

/* Select_Interactive_Module parameters (modeling/AddModule/useA.py:33-48,340-361).
 * token_selection.W_v is never used by the reference forward (useA.py:48) and is absent. */
typedef struct sig_sim_params {
  const float* sel_wq;   const float* sel_bq;     /* token_selection.W_q  [d,d],[d] */
  const float* sel_wk;   const float* sel_bk;     /* token_selection.W_k  [d,d],[d] */
  const float* in_proj_w;  const float* in_proj_b;  /* cross_attn.in_proj_*  [3d,d],[3d] */
  const float* out_proj_w; const float* out_proj_b; /* cross_attn.out_proj.* [d,d],[d] */
  const float* ffn0_w; const float* ffn0_b;       /* ffn.0 [2d,d],[2d] */
  const float* ffn2_w; const float* ffn2_b;       /* ffn.2 [d,2d],[d] */
  const float* ln1_w; const float* ln1_b;         /* norm1 [d] */
  const float* ln2_w; const float* ln2_b;         /* norm2 [d] */
  const sig_sel_fold* sel_fold;                   /* optional (NULL: fold on the fly in fp32) */
  /* Optional by-product of the selection-score pass (bf16 tensor-core path): fp32 [3][B][d] mean over the
   * L patch rows of each modality = the GAM mean pool (useB.py:84-86), written while the tokens stream through the score
   * kernel.  Meant for sig_align_fwd's SIG_FLAG_PATCH_MEAN slot when both modules read the same token maps
   * (make_model.py:191,205); pool_event (cudaEvent_t, may be NULL) is recorded on the stream position where it is complete. */
  float* pool_out;
  void* pool_event;
} sig_sim_params;

/* Gradients of the trainable SIM parameters (token_selection.* never receive
 * gradients: selection is not differentiable, useA.py:79-93).  Overwritten. */
typedef struct sig_sim_param_grads {
  float* in_proj_w;  float* in_proj_b;
  float* out_proj_w; float* out_proj_b;
  float* ffn0_w; float* ffn0_b;
  float* ffn2_w; float* ffn2_b;
  float* ln1_w; float* ln1_b;
  float* ln2_w; float* ln2_b;
  /* Optional (cudaEvent_t, may be NULL): recorded by sig_sim_bwd once every gradient above is final EXCEPT
   * in_proj_w rows [0, 2d) (W_q, W_k) and in_proj_b, which need the token-side backward first.  A data-parallel
   * caller starts the exchange of the early part on another stream while the rest of the backward runs
   * (FusionHead.grad_sync).  On the fp32 SIMT path it is recorded at the end of the call. */
  void* early_event;
  /* Optional (cudaEvent_t, may be NULL): recorded once the REST -- in_proj_w rows [0, 2d) and in_proj_b -- is final too, i.e.
   * when the last weight-gradient GEMM of the call has been enqueued (on the library's side stream); the call's own stream
   * still has the CLS-gradient GEMM and the token-gradient writes in front of it.  Lets the exchange of the late piece start
   * before the call's last kernels have run. */
  void* late_event;
} sig_sim_param_grads;

/* AlignmentM parameters (modeling/AddModule/useB.py:44-74, DAS.py:30-72), index = modality r,n,t */
typedef struct sig_align_params {
  const float* contra_temp;                      /* [] */
  const float* proj_q_w[3]; const float* proj_q_b[3];   /* DAS_*.proj_q         [d,d,1,1],[d] */
  const float* off0_w[3];   const float* off0_b[3];     /* DAS_*.conv_offset.0  [d,d,1,1],[d] */
  const float* off2_w[3];   const float* off2_b[3];     /* DAS_*.conv_offset.2  [d,1,4,4],[d] */
  const float* off4_w[3];                                /* DAS_*.conv_offset.4  [1,d,1,1]     */
  /* With SIG_FLAG_PATCH_MEAN: optional cudaEvent_t (may be NULL) the GAM chain waits for before it reads the patch means
   * from the ctx slot (they are being written on another stream, e.g. by sig_sim_fwd's pool_out). */
  void* patch_mean_event;
} sig_align_params;

typedef struct sig_align_param_grads {           /* overwritten */
  float* contra_temp;
  float* proj_q_w[3]; float* proj_q_b[3];
  float* off0_w[3];   float* off0_b[3];
  float* off2_w[3];   float* off2_b[3];
  float* off4_w[3];
  /* Optional (cudaEvent_t, may be NULL): recorded by sig_align_bwd once ALL gradients above are final; on the
   * tensor-core path that is before the d(patches) GEMM runs, so the exchange of this arena overlaps it. */
  void* done_event;
} sig_align_param_grads;

/* ---- library info ------------------------------------------------------ */
int sig_version(void);                       /* SIG_ABI_VERSION */
const char* sig_error_string(int code);      /* static string; cudaGetErrorString for code > 0 */
/* bytes of the ctx buffer for (kind, B, L, d) with tokens of `dtype` and the call's `flags`;
 * 0 if the shape is unsupported */
size_t sig_ctx_bytes(int kind, int B, int L, int d, int dtype, unsigned flags);

/* ---- SIM: Select_Interactive_Module.forward (useA.py:454-476) ---------- */
/* out  [B,3d] in tokens' dtype; masks fp32 [3,B,L] (last_masks, useA.py:323).
 * k1 = TOPK, k2 = 2*TOPK (useA.py:42-43), both clipped to the row length;
 * max_keep = int(L*keep_ratio) or -1 when keep_ratio is None (useA.py:254-256). */
int sig_sim_fwd(const sig_tokens* tok, const sig_sim_params* p, int k1, int k2, int max_keep,
                void* out, float* masks, void* ctx, size_t ctx_bytes,
                unsigned flags, int device, void* stream);
/* dout [B,3d] tokens' dtype.  Writes dpatch/dcls and all sig_sim_param_grads. */
int sig_sim_bwd(const sig_tokens* tok, const sig_sim_params* p, const void* dout,
                const sig_token_grads* dtok, const sig_sim_param_grads* dp,
                void* ctx, size_t ctx_bytes, unsigned flags, int device, void* stream);

/* Fill a sig_sel_fold from p->sel_wq/sel_bq/sel_wk/sel_bk.  m_hl bf16 [d,2d], v/u fp32 [d], s0 fp32 [1];
 * ws: scratch of d*d floats. */
int sig_sim_fold_selection(const sig_sim_params* p, int d, void* m_hl, float* v, float* u, float* s0,
                           void* ws, size_t ws_bytes, int device, void* stream);

/* ---- TokenSelection (useA.py:50-325) ----------------------------------- */
/* which: 1 = intra_modal_token_selection (:50), 2 = inter_modal_token_selection (:98),
 * 3 = forward's union (+ keep_ratio) (:223-314).  masks fp32 [3,B,L].
 * selected (optional, NULL to skip): three [B,L,d] contiguous maps in tokens' dtype,
 * laid out [3,B,L,d] = patches * mask (:318-320). */
int sig_sim_select_fwd(const sig_tokens* tok, const sig_sim_params* p, int which,
                       int k1, int k2, int max_keep, float* masks, void* selected,
                       void* ctx, size_t ctx_bytes, int device, void* stream);
/* Test seam: selection from caller-supplied fp32 scores, ties -> lowest index.
 * intra [3,B,L] (softmax scores, useA.py:72-74), inter [3,B,2L] (D_m, useA.py:136-151),
 * raw [3,B,L] (cls.patch, only read when max_keep >= 0).  masks fp32 [3,B,L] as `which`. */
int sig_sim_select_from_scores(const float* intra, const float* inter, const float* raw,
                               int B, int L, int which, int k1, int k2, int max_keep,
                               float* masks, int device, void* stream);
/* dpatch = dselected * mask (backward of useA.py:318-320); dselected [3,B,L,d] contiguous */
int sig_mask_mul_bwd(const void* dselected, const float* masks, int dtype, int B, int L, int d,
                     const sig_token_grads* dtok, int device, void* stream);

/* ---- ModalInteractive.forward (useA.py:364-411) on arbitrary K/V tokens - */
/* tok->patch = the three "selected" maps, tok->cls = CLS tokens; masks may be NULL
 * (all tokens participate) or fp32 [3,B,L] (token rows are multiplied by it). */
int sig_sim_attn_fwd(const sig_tokens* tok, const sig_sim_params* p, const float* masks,
                     void* out, void* ctx, size_t ctx_bytes, unsigned flags, int device, void* stream);
int sig_sim_attn_bwd(const sig_tokens* tok, const sig_sim_params* p, const float* masks,
                     const void* dout, const sig_token_grads* dtok, const sig_sim_param_grads* dp,
                     void* ctx, size_t ctx_bytes, unsigned flags, int device, void* stream);

/* ---- AlignmentM.forward (useB.py:169-190) ------------------------------ */
/* losses[0] = Cls_Align (GAM, useB.py:76-126), losses[1] = patch_Align (LAM, useB.py:128-167;
 * skipped when do_lam == 0, i.e. stage == "CLS").  grid h x w with h*w == L. */
int sig_align_fwd(const sig_tokens* tok, const sig_align_params* p, int h, int w, int do_lam,
                  float* losses, void* ctx, size_t ctx_bytes, unsigned flags, int device, void* stream);
/* dlosses: device fp32 [2] (upstream gradients of the two scalars). */
/* Where SIG_FLAG_PATCH_MEAN expects the patch means: *slot = fp32 [3][B][d] inside ctx.  SIG_ERR_SHAPE when this
 * configuration does not run on the tensor-core path (the exact fp32 path pools in fp64 itself). */
int sig_align_patch_mean_slot(void* ctx, int B, int L, int d, int dtype, unsigned flags, float** slot);
int sig_align_bwd(const sig_tokens* tok, const sig_align_params* p, int h, int w, int do_lam,
                  const float* dlosses, const sig_token_grads* dtok, const sig_align_param_grads* dp,
                  void* ctx, size_t ctx_bytes, unsigned flags, int device, void* stream);

/* ---- DA_sample.forward (DAS.py:107-165), one modality ------------------- */
/* x: channels-last [B,h*w,d] view (stride_b, stride_l elements); params index m of `p`.
 * sampled: fp32 [B,Hk*Wk,d] contiguous (Hk=h/4, Wk=w/4). */
int sig_das_fwd(const void* x, int64_t stride_b, int64_t stride_l, int dtype, int B, int h, int w, int d,
                const sig_align_params* p, int m, float* sampled,
                void* ctx, size_t ctx_bytes, unsigned flags, int device, void* stream);
/* dsampled fp32 [B,Hk*Wk,d]; dx: same geometry/dtype as x (overwritten); parameter grads index m of dp */
int sig_das_bwd(const void* x, int64_t stride_b, int64_t stride_l, int dtype, int B, int h, int w, int d,
                const sig_align_params* p, int m, const float* dsampled, void* dx,
                const sig_align_param_grads* dp, void* ctx, size_t ctx_bytes,
                unsigned flags, int device, void* stream);

/* ---- utils/volume.py:14 volume_computation3 ---------------------------- */
/* l [B1,d], v,a [B2,d] fp32 contiguous -> vol [B1,B2] fp32 = sqrt|det Gram(l_i, v_j, a_j)| */
size_t sig_volume3_ws_bytes(int B1, int B2);   /* scratch for either call */
int sig_volume3_fwd(const float* l, const float* v, const float* a, int B1, int B2, int d,
                    float* vol, void* ws, size_t ws_bytes, int device, void* stream);
int sig_volume3_bwd(const float* l, const float* v, const float* a, int B1, int B2, int d,
                    const float* dvol, float* dl, float* dv, float* da,
                    void* ws, size_t ws_bytes, int device, void* stream);

/* ---- inference tail -------------------------------------------------------------------
 * sig_infer_features: make_model.py:284-290 -- feat[b] = [RGB_global | NI_global | TI_global | vars_total[b]] as fp32
 *   [B, 6d]; normalize != 0: rows are L2-normalised (F.normalize, eps 1e-12; utils/metrics.py:266-268).  cls[m]: [B,d] views
 *   (stride in elements), sim_out [B,3d] with leading dimension ld_sim, both of `dtype` (SIG_F32 | SIG_BF16).
 * sig_euclidean_distmat: utils/metrics.py:494-501 -- dist[i,j] = |q_i|^2 + |g_j|^2 - 2 q_i.g_j (squared, fp32; the operations in
 *   the reference's order: addmm_(beta=1, alpha=-2)); qf [nq,D], gf [ng,D], dist [nq,ng] contiguous fp32; ws: (nq+ng) floats.
 * sig_rank_eval: utils/metrics.py:111-170 eval_func (market1501 protocol: gallery entries with the query's pid AND camid are
 *   discarded; queries whose identity is absent from the valid gallery are skipped) -> cmc float32 [max_rank] (caller clips
 *   max_rank to ng as the reference does), map_out double [2] = {mAP, number of valid queries}; ranking = stable ascending
 *   order (ties -> lowest gallery index).  pids / camids int64 on the device.  stats: 3*nq doubles scratch; overflow: one
 *   int, set to 1 if a query had more than 2048 matches (results then cover the first 2048). */
int sig_infer_features(const void* const cls[3], const int64_t cls_stride_b[3], const void* sim_out, int64_t ld_sim, int dtype,
                       int B, int d, int normalize, float* out, int device, void* stream);
int sig_euclidean_distmat(const float* qf, const float* gf, int nq, int ng, int D, float* dist, void* ws, size_t ws_bytes,
                          int device, void* stream);
int sig_rank_eval(const float* dist, int64_t ld, const int64_t* q_pids, const int64_t* g_pids, const int64_t* q_camids,
                  const int64_t* g_camids, int nq, int ng, int max_rank, float* cmc, double* map_out, double* stats, int* overflow,
                  int device, void* stream);

/* ---- data-parallel gradient exchange over NVLink peer memory ------------------------------
 * Replaces the DDP all-reduce of the head's gradients (engine/processor.py:100-105) by ONE kernel that needs no shared
 * memory, so it runs next to the persistent compute kernels of the backward instead of waiting for their SMs.
 * Every rank (one process per GPU, same node) owns a SYMMETRIC buffer: `buf[r]` / `flags[r]` are the addresses, mapped
 * into THIS process, of rank r's gradient arena and of its flag region (sig_xchg_flag_bytes() bytes, zeroed once before
 * the first call, never touched by the caller afterwards); `multicast` is the NVLS multicast address of the arena or
 * NULL (plain peer loads/stores then).  The caller obtains them from its IPC mechanism of choice -- the Python side uses
 * torch.distributed._symmetric_memory (signal_b200.parallel.GradExchange).
 * sig_xchg_allreduce_f32: arena[off, off+count) <- scale * sum over ranks, in place on every rank; all ranks must
 * issue the same sequence of calls with the same (off, count, ctas).  off, count: fp32 elements, multiples of 4.
 * Stream-ordered, CUDA-graph capturable, no host synchronisation. */
typedef struct sig_xchg_peers {
  void* buf[8];
  void* flags[8];
  void* multicast;
  int32_t rank, world;
} sig_xchg_peers;
size_t sig_xchg_flag_bytes(void);
int sig_xchg_allreduce_f32(const sig_xchg_peers* peers, size_t off, size_t count, float scale, int ctas, int device, void* stream);

/* ---- fp16 boundary ------------------------------------------------------------ */
/* dst = convert(src) over a strided [nb, nl, d] map (element strides, unit channel stride, d % 8 == 0, 16-byte aligned
 * rows): SIG_F16 -> SIG_BF16 (round to nearest even; every fp16 value is in bf16's range) or SIG_BF16 -> SIG_F16
 * (round to nearest even, saturating at +-65504).  Used by the nn.Module shims for fp16 (autocast) token maps and for
 * the token gradients on the way back. */
int sig_convert_half(const void* src, int64_t src_stride_b, int64_t src_stride_l, int src_dtype,
                     void* dst, int64_t dst_stride_b, int64_t dst_stride_l, int dst_dtype,
                     int nb, int nl, int d, int device, void* stream);

/* ---- measurement aids (no effect on results) ------------------------------ */
/* number of kernel launches the library has enqueued since it was loaded */
unsigned long long sig_debug_launch_count(void);
/* per-phase CUDA-event timing: enable, run, then collect (synchronises the recorded events).
 * names_buf receives '\n'-separated phase names; ms[i]/counts[i] the summed time and scope count. */
/* Device pointers, inside a SIM ctx buffer of the bf16 tensor-core path, of the two operands of SIM's token
 * gradient ([B][384][64] and [B*64][d], bf16) -- see sig_token_grads.fuse_*.  Returns SIG_ERR_DTYPE when
 * this (dtype, L, flags) combination does not run on that path. */
int sig_sim_dx_operands(void* ctx, int B, int L, int d, int dtype, unsigned flags, void** pds, void** dxqt);

/* ---- ID / metric losses behind the head ---------------------------------------------
 * logits/feat: row-major [B, C] / [B, D] with leading dimension ld (elements), dtype SIG_F32 or SIG_BF16; targets and
 * labels int64 [B] on the device (no host copy, unlike layers/softmax_loss.py:30); fp32 arithmetic; ws: sig_loss_ws_bytes(B).
 * CrossEntropyLabelSmooth.forward (layers/softmax_loss.py:23-34): loss = mean_b sum_k -((1-eps) 1[k=y_b] + eps/C) log p_bk;
 * lse [B] is saved for the backward; dlogits has the dtype of logits. */
size_t sig_loss_ws_bytes(int B);
int sig_xent_ls_fwd(const void* logits, int dtype, int64_t ld, const int64_t* targets, int B, int C, float eps, float* loss,
                    float* lse, void* ws, size_t ws_bytes, int device, void* stream);
int sig_xent_ls_bwd(const void* logits, int dtype, int64_t ld, const int64_t* targets, int B, int C, float eps, const float* lse,
                    const float* dloss, void* dlogits, int64_t ldd, int device, void* stream);
/* TripletLoss.__call__ (layers/triplet_loss.py:121-135): euclidean_dist (:16-31, clamp 1e-12 before the sqrt),
 * hard_example_mining (:51-104; the anchor itself counts as a positive, ties -> lowest index), dist_ap *= 1 + hard_factor,
 * dist_an *= 1 - hard_factor, then SoftMarginLoss(an - ap, 1) (soft_margin != 0; margin=None in the reference) or
 * MarginRankingLoss(margin)(an, ap, 1).  Outputs: loss [1], dist_ap/dist_an [B], p_idx/n_idx int32 [B].
 * Backward: dloss [1] and optional cotangents of dist_ap / dist_an (may be NULL) -> dfeat (dtype of feat).  B <= 1024. */
int sig_triplet_fwd(const void* feat, int dtype, int64_t ld, const int64_t* labels, int B, int D, float margin, int soft_margin,
                    float hard_factor, float* loss, float* dist_ap, float* dist_an, int* p_idx, int* n_idx, void* ws, size_t ws_bytes,
                    int device, void* stream);
int sig_triplet_bwd(const void* feat, int dtype, int64_t ld, int B, int D, float margin, int soft_margin, float hard_factor,
                    const float* dist_ap, const float* dist_an, const int* p_idx, const int* n_idx, const float* dloss,
                    const float* d_dist_ap, const float* d_dist_an, void* dfeat, int64_t ldd, int device, void* stream);

/* BNNeck + classifier (modeling/make_model.py:128-131,194-195,212-214: nn.BatchNorm1d(D) -> nn.Linear(D, C, bias=False)).
 * feat [B, D] (ld); bn_out [B, D] and logits [B, C] in the dtype of feat; training != 0: batch statistics, running_mean /
 * running_var (may be NULL) updated in place with `momentum` (unbiased variance), else the running statistics are used.
 * save_mean / save_rstd [D] and y32 [B, D] fp32 are kept by the caller for the backward; ws: sig_bnneck_ws_bytes(B, D, C).
 * Backward: dlogits and/or d_bn_out (either may be NULL) -> dfeat (dtype of feat), d_bn_weight, d_bn_bias [D],
 * d_cls_weight [C, D] fp32 (overwritten). */
size_t sig_bnneck_ws_bytes(int B, int D, int C);
int sig_bnneck_cls_fwd(const void* feat, int dtype, int64_t ld, int B, int D, int C, const float* bn_weight, const float* bn_bias,
                       float* running_mean, float* running_var, float momentum, float eps, int training, const float* cls_weight,
                       void* bn_out, int64_t ldo, void* logits, int64_t ldl, float* save_mean, float* save_rstd, float* y32, void* ws,
                       size_t ws_bytes, int device, void* stream);
int sig_bnneck_cls_bwd(const void* feat, int dtype, int64_t ld, int B, int D, int C, const float* bn_weight, const float* cls_weight,
                       const float* save_mean, const float* save_rstd, int training, const float* y32, const void* dlogits, int64_t ldl,
                       const void* d_bn_out, int64_t lddo, void* dfeat, int64_t ldx, float* d_bn_weight, float* d_bn_bias,
                       float* d_cls_weight, void* ws, size_t ws_bytes, int device, void* stream);

/* ---- Token producer in front of the head ----------------------------------------------------
 * The tail of the CLIP vision tower: x = ln_post(x); tokens = x @ proj (modeling/clip/model.py:485-487; LayerNorm with
 * fp32 statistics :154-160), whose rows 0 / 1.. are the CLS and patch views the head consumes (modeling/meta_arch.py:108-110).
 * x: [B, L1 = 1+L, W] map given by element strides (x_stride_b, x_stride_l; unit channel stride), so the tower's [L1, B, W]
 * layout (clip/model.py:484) is read in place; dtype SIG_F32 (exact path), SIG_BF16 or SIG_F16 (LayerNorm in fp32, result
 * rounded to bf16, tcgen05 GEMM with fp32 accumulation: the reference's autocast data flow).  ln_w / ln_b [W], proj [W, D]
 * fp32 masters.  tokens: contiguous [B, L1, D] of tok_dtype = dtype, or SIG_BF16 / SIG_F16 for fp32 x (an autocast caller
 * whose residual stream stayed fp32: torch's autocast runs layer_norm in fp32 and the matmul in half).  patch_mean (optional, may be NULL): fp32 [B, D] mean of
 * the L patch rows of the rounded tokens -- the GAM mean pool (useB.py:84-86) as a by-product, see sig_align_fwd's
 * SIG_FLAG_PATCH_MEAN.  saved: sig_tokens_ws_bytes(0, ...) bytes kept by the caller for the backward; scratch:
 * sig_tokens_ws_bytes(1, ...) (forward) / (2, ...) (backward) bytes, may be NULL when 0.
 * Backward: dtokens [B, L1, D] of tok_dtype (element strides; rows must have one pitch, dt_stride_b == L1 * dt_stride_l, for
 * SIG_F32) -> dx at (dx_stride_b, dx_stride_l) in the dtype of x, d_ln_w / d_ln_b [W], d_proj [W, D] fp32 (all overwritten;
 * d_proj is a split-K sum of fp32 atomics: run-to-run differences at the 1e-7 level). */
size_t sig_tokens_ws_bytes(int which, int B, int L1, int W, int D, int tok_dtype);
int sig_tokens_fwd(const void* x, int dtype, int64_t x_stride_b, int64_t x_stride_l, int B, int L1, int W, int D, const float* ln_w,
                   const float* ln_b, float eps, const float* proj, void* tokens, int tok_dtype, float* patch_mean, void* saved,
                   size_t saved_bytes, void* scratch, size_t scratch_bytes, int device, void* stream);
int sig_tokens_bwd(const void* x, int dtype, int64_t x_stride_b, int64_t x_stride_l, int B, int L1, int W, int D, const float* ln_w,
                   const float* proj, const void* dtokens, int tok_dtype, int64_t dt_stride_b, int64_t dt_stride_l, const void* saved,
                   size_t saved_bytes, void* dx, int64_t dx_stride_b, int64_t dx_stride_l, float* d_ln_w, float* d_ln_b, float* d_proj,
                   void* scratch, size_t scratch_bytes, int device, void* stream);

int sig_profile_enable(int on);
int sig_profile_collect(char* names_buf, size_t names_bytes, float* ms, int* counts, int max);
/* Time line of the recorded scopes ("name start_us end_us" lines, relative to the earliest start); with
 * SIG_PROF_CAPTURE=1 scopes are also recorded during stream capture, so a graph replay can be laid out. */
int sig_profile_timeline(char* buf, size_t bytes);
/* A caller-side scope on the same time line (e.g. around a collective issued between two library calls):
 * begin returns a handle for sig_profile_scope_end, or NULL when the profiler is off. */
void* sig_profile_scope_begin(const char* name, void* stream);
void sig_profile_scope_end(void* scope);

/* ---- volume_computation4 / volume_computation5 (utils/volume.py:65-116, :119-182) ----------------------------------
 * V[i,j] = sqrt|det G(i,j)|, G = Gram matrix of (language_i, video_j, audio_j, subtitles_j [, depth_j]); n = 4 or 5 (3 is
 * accepted and equals sig_volume3_*).  feats[0] = language [B1,d], feats[1..n-1] = [B2,d], contiguous fp32; vol [B1,B2] fp32.
 * The reference never calls these two functions; they complete the utils/volume.py surface.  The per-pair determinant and
 * its cofactors are evaluated in fp64 (the reference: torch.det on G.float()).  Backward: dvol [B1,B2] -> dfeats[k] with the
 * shapes of feats (overwritten); at det == 0 the gradient is 0 (the reference yields NaN).  ws: sig_volume_n_ws_bytes. */
size_t sig_volume_n_ws_bytes(int n, int B1, int B2);
int sig_volume_n_fwd(int n, const float* const* feats, int B1, int B2, int d, float* vol, void* ws, size_t ws_bytes, int device,
                     void* stream);
int sig_volume_n_bwd(int n, const float* const* feats, int B1, int B2, int d, const float* dvol, float* const* dfeats, void* ws,
                     size_t ws_bytes, int device, void* stream);

/* Unit-test seam for the tcgen05 GEMM core: C = alpha * A . B^T (+bias) (GELU if act), bf16 operands.
 * mode: 0 row-major [rows,K]; 1 token view [B,128,d] with rows=(b,l), K=d; 2 row-major [K,cols];
 * 3 token view with K=(b,l), cols=d.  geom = {ld, stride_b, stride_l, rows, cols} (elements).
 * c_stride_b != 0: C rows are (b,l) token rows at C + b*c_stride_b + l*c_stride_l (+ per-sample rowvec). */
int sig_debug_gemm_bf16(const void* A, int a_mode, const int64_t* a_geom, const void* B, int b_mode,
                        const int64_t* b_geom, void* C, int64_t ldc, int out_bf16, const float* bias,
                        int M, int N, int K, float alpha, int act, int ksplit, int bn,
                        int64_t c_stride_b, int64_t c_stride_l, const float* rowvec, const float* rowvec_scale,
                        int accumulate, int device, void* stream);

/* Debug aid: with SIG_TC_STAMPS=1 in the environment, CTA 0 of every tcgen05 pipeline launch records
 * clock64() at nine points of its life (tc_pipeline.cuh); this copies the 16-slot stamp array out. */
int sig_debug_tc_stamps(long long* out16);

#ifdef __cplusplus
}
#endif
#endif /* SIGNAL_B200_H */

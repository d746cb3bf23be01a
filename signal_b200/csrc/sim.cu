// SIM -- Select_Interactive_Module (modeling/AddModule/useA.py), fp32 SIMT path.
//
// Math (validated against the live reference):
//  selection  q = W_q cls + b_q ; qt = W_k^T q ; c = q.b_k ;
//             S[m,j] = softmax_j((qt_m . x_j + c_m)/sqrt(d)) over the 3L tokens   (useA.py:116-129)
//             intra  = softmax_l(cls_m . x_l / sqrt(d))                            (useA.py:72-74)
//             rank-select (ties -> lowest index), scatter to the other modalities, union (useA.py:155-251)
//  attention  per head h: q_h = (W_q^h cls + b_q^h), qt_h = W_k^hT q_h / sqrt(hd), c_h = q_h.b_k^h / sqrt(hd)
//             s_j = mask_j ? qt_h . x_j + c_h : c_h ; p = softmax_j(s)
//             xbar_h = sum_j p_j mask_j x_j ; o_h = W_v^h xbar_h + b_v^h            (useA.py:383-388)
//             out = LN2(y1 + FFN(y1)), y1 = LN1(cls + W_o o + b_o)                  (useA.py:393-408)
// i.e. the K/V projections are applied to 24 effective queries per sample instead of 384 tokens.
#include "common.cuh"
#include "sim.h"
#include "align.h"
#include "sim_tc.h"
#include "simt_ops.cuh"
#include "tc_gemm.h"
#include "tok_ring.cuh"

namespace sig {

// ---------------------------------------------------------------------------------------------
// token conversion: strided T views -> contiguous fp32  Xf[3][B][L][d], clsf[B][3][d]
// ---------------------------------------------------------------------------------------------
struct TokPtrs {
  const void* patch[3];
  const void* cls[3];
  int64_t psb[3], psl[3], csb[3];
};

template <typename T>
static __global__ void convert_tokens_kernel(TokPtrs tp, int B, int L, int d, float* __restrict__ Xf, float* __restrict__ clsf) {
  pdl_enter();
  const int m = blockIdx.y;
  const int64_t row = blockIdx.x;  // [0, B*L) patches, [B*L, B*L+B) cls
  const T* src;
  float* dst;
  if (row < (int64_t)B * L) {
    const int b = (int)(row / L), l = (int)(row % L);
    src = static_cast<const T*>(tp.patch[m]) + b * tp.psb[m] + l * tp.psl[m];
    dst = Xf + (((int64_t)m * B + b) * L + l) * d;
  } else {
    if (!clsf || !tp.cls[m]) return;
    const int b = (int)(row - (int64_t)B * L);
    src = static_cast<const T*>(tp.cls[m]) + b * tp.csb[m];
    dst = clsf + ((int64_t)b * 3 + m) * d;
  }
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    float v[8];
    load8(src + c, v);
    store8(dst + c, v);
  }
}

int convert_tokens(const sig_tokens* t, float* Xf, float* clsf, cudaStream_t s) {
  SIG_PHASE("convert_tokens");
  TokPtrs tp;
  for (int m = 0; m < 3; ++m) {
    tp.patch[m] = t->patch[m]; tp.cls[m] = t->cls[m];
    tp.psb[m] = t->patch_stride_b[m]; tp.psl[m] = t->patch_stride_l[m]; tp.csb[m] = t->cls_stride_b[m];
  }
  dim3 grid((unsigned)((int64_t)t->B * t->L + t->B), 3);
  const int threads = t->d / 8 >= 128 ? 128 : 64;
  if (t->dtype == SIG_BF16)
    SIG_LAUNCH((convert_tokens_kernel<__nv_bfloat16>), grid, threads, 0, s, tp, t->B, t->L, t->d, Xf, clsf);
  else
    SIG_LAUNCH((convert_tokens_kernel<float>), grid, threads, 0, s, tp, t->B, t->L, t->d, Xf, clsf);
  SIG_CHECK_LAUNCH();
  return 0;
}

// CLS tokens only: strided T views -> clsf[B][3][d] fp32.  grid (B, 3)
template <typename T>
static __global__ void gather_cls_kernel(TokPtrs tp, int d, float* __restrict__ clsf, __nv_bfloat16* __restrict__ clsb,
                                         __nv_bfloat16* __restrict__ clsb2) {
  pdl_enter();
  const int b = blockIdx.x, m = blockIdx.y;
  const T* src = static_cast<const T*>(tp.cls[m]) + b * tp.csb[m];
  float* dst = clsf + ((int64_t)b * 3 + m) * d;
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    float v[8];
    load8(src + c, v);
    store8(dst + c, v);
    if (clsb) store8(clsb + ((int64_t)b * 3 + m) * d + c, v);
    if (clsb2) {  // [cls | cls]: A operand of the split-bf16 product with [M_hi | M_lo]
      store8(clsb2 + ((int64_t)b * 3 + m) * 2 * d + c, v);
      store8(clsb2 + ((int64_t)b * 3 + m) * 2 * d + d + c, v);
    }
  }
}

// csel[r] = cls[r] . u + s0   grid R, 128 threads
static __global__ void __launch_bounds__(128) csel_fold_kernel(const float* __restrict__ clsf, const float* __restrict__ u,
                                                               const float* __restrict__ s0, int d, float* __restrict__ csel) {
  pdl_enter();
  __shared__ float scratch[33];
  const int64_t r = blockIdx.x;
  float a = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) a = fmaf(clsf[r * d + c], u[c], a);
  a = block_sum(a, scratch);
  if (threadIdx.x == 0) csel[r] = a + *s0;
}

// out[n][0:d] = hi(M[n][:]), out[n][d:2d] = lo(M[n][:]);  grid d, 128 threads
static __global__ void split_hl_kernel(const float* __restrict__ M, int d, __nv_bfloat16* __restrict__ out) {
  pdl_enter();
  const int64_t n = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float v = M[n * d + c];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[n * 2 * d + c] = hi;
    out[n * 2 * d + d + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

static __global__ void dot_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ out) {
  pdl_enter();
  __shared__ float scratch[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(a[i], b[i], s);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) *out = s;
}

static TokPtrs tok_ptrs(const sig_tokens* t) {
  TokPtrs tp;
  for (int m = 0; m < 3; ++m) {
    tp.patch[m] = t->patch[m]; tp.cls[m] = t->cls[m];
    tp.psb[m] = t->patch_stride_b[m]; tp.psl[m] = t->patch_stride_l[m]; tp.csb[m] = t->cls_stride_b[m];
  }
  return tp;
}

// ---------------------------------------------------------------------------------------------
// selection scores
// ---------------------------------------------------------------------------------------------
// Same four dot products per token as sim_scores_kernel, reading the strided token views in place.
// grid (3, B), 256 threads: one CTA per (modality, sample) stages its four query vectors once, then
// each warp streams 4 token rows at a time (8 channels = 16 B per lane per row and column step, all
// row loads of a step issued before the first FMA) -- one pass over the tokens, T*s bytes per sample.
constexpr int kScoreRows = 4;
template <typename T>
static __global__ void __launch_bounds__(256) sim_scores_tok_kernel(TokPtrs tp, const float* __restrict__ clsf,
                                                                    const float* __restrict__ qtsel, const float* __restrict__ csel,
                                                                    int B, int L, int d, float* __restrict__ sel_logits,
                                                                    float* __restrict__ intra_raw) {
  pdl_enter();
  extern __shared__ __align__(16) float qv[];  // [4][d]
  const int m = blockIdx.x, b = blockIdx.y;
  for (int i = threadIdx.x * 4; i < 4 * d; i += blockDim.x * 4) {
    const int r = i / d, c = i % d;
    const float* src = r < 3 ? qtsel + ((int64_t)b * 3 + r) * d + c : clsf + ((int64_t)b * 3 + m) * d + c;
    *reinterpret_cast<float4*>(qv + i) = *reinterpret_cast<const float4*>(src);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float inv = rsqrtf((float)d);
  const float cs0 = csel[b * 3 + 0], cs1 = csel[b * 3 + 1], cs2 = csel[b * 3 + 2];
  const T* xb = static_cast<const T*>(tp.patch[m]) + b * tp.psb[m];
  for (int l0 = w * kScoreRows; l0 < L; l0 += nw * kScoreRows) {
    float a[kScoreRows][4];
#pragma unroll
    for (int i = 0; i < kScoreRows; ++i) a[i][0] = a[i][1] = a[i][2] = a[i][3] = 0.f;
    for (int c = lane * 8; c < d; c += 256) {
      float xv[kScoreRows][8];
#pragma unroll
      for (int i = 0; i < kScoreRows; ++i) {
        const int l = min(l0 + i, L - 1);
        load8(xb + l * tp.psl[m] + c, xv[i]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float qq[8];
        load8(qv + r * d + c, qq);
#pragma unroll
        for (int i = 0; i < kScoreRows; ++i)
#pragma unroll
          for (int t = 0; t < 8; ++t) a[i][r] = fmaf(xv[i][t], qq[t], a[i][r]);
      }
    }
#pragma unroll
    for (int i = 0; i < kScoreRows; ++i)
#pragma unroll
      for (int r = 0; r < 4; ++r) a[i][r] = warp_sum(a[i][r]);
    if (lane < kScoreRows && l0 + lane < L) {
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
      for (int i = 0; i < kScoreRows; ++i)
        if (lane == i) { v0 = a[i][0]; v1 = a[i][1]; v2 = a[i][2]; v3 = a[i][3]; }
      const int l = l0 + lane;
      const int64_t base = (int64_t)b * 3 * 3 * L + (int64_t)m * L + l;
      sel_logits[base] = (v0 + cs0) * inv;
      sel_logits[base + 3 * L] = (v1 + cs1) * inv;
      sel_logits[base + 6 * L] = (v2 + cs2) * inv;
      intra_raw[((int64_t)b * 3 + m) * L + l] = v3;
    }
  }
}
// grid (ceil(L/32), 3, B), 256 threads: 8 warps x 4 tokens.  Four dot products per token:
// the three folded inter-modal queries and the token's own CLS.
static __global__ void __launch_bounds__(256) sim_scores_kernel(const float* __restrict__ Xf, const float* __restrict__ clsf,
                                                                const float* __restrict__ qtsel, const float* __restrict__ csel,
                                                                int B, int L, int d, float* __restrict__ sel_logits,
                                                                float* __restrict__ intra_raw) {
  pdl_enter();
  extern __shared__ float qv[];  // [4][d]
  const int m = blockIdx.y, b = blockIdx.z;
  for (int i = threadIdx.x; i < 4 * d; i += blockDim.x) {
    const int r = i / d, c = i % d;
    qv[i] = r < 3 ? qtsel[((int64_t)b * 3 + r) * d + c] : clsf[((int64_t)b * 3 + m) * d + c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float inv = rsqrtf((float)d);
  for (int i = 0; i < 4; ++i) {
    const int l = blockIdx.x * 32 + w * 4 + i;
    if (l >= L) break;
    const float* x = Xf + (((int64_t)m * B + b) * L + l) * d;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int c = lane * 4; c < d; c += 128) {
      const float4 xv = *reinterpret_cast<const float4*>(x + c);
      const float4 q0 = *reinterpret_cast<const float4*>(qv + c);
      const float4 q1 = *reinterpret_cast<const float4*>(qv + d + c);
      const float4 q2 = *reinterpret_cast<const float4*>(qv + 2 * d + c);
      const float4 q3 = *reinterpret_cast<const float4*>(qv + 3 * d + c);
      a0 += xv.x * q0.x + xv.y * q0.y + xv.z * q0.z + xv.w * q0.w;
      a1 += xv.x * q1.x + xv.y * q1.y + xv.z * q1.z + xv.w * q1.w;
      a2 += xv.x * q2.x + xv.y * q2.y + xv.z * q2.z + xv.w * q2.w;
      a3 += xv.x * q3.x + xv.y * q3.y + xv.z * q3.z + xv.w * q3.w;
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
    if (lane == 0) {
      const int64_t base = (int64_t)b * 3 * 3 * L + (int64_t)m * L + l;
      sel_logits[base] = (a0 + csel[b * 3 + 0]) * inv;
      sel_logits[base + 3 * L] = (a1 + csel[b * 3 + 1]) * inv;
      sel_logits[base + 6 * L] = (a2 + csel[b * 3 + 2]) * inv;
      intra_raw[((int64_t)b * 3 + m) * L + l] = a3;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// rank-select (shared by the fused path and the from-scores test seam)
// ---------------------------------------------------------------------------------------------
// intra [3][L], inter [3][2L] (D_m of useA.py:136-151), raw [3][L]; all in shared memory.
// mask_out [3][L] (shared).  which: 1 intra, 2 inter, 3 union (+ keep_ratio when max_keep >= 0).
// rank_i = #{j : s_j > s_i or (s_j == s_i and j < i)}; selected iff rank_i < k.
static __device__ void select_masks(const float* intra, const float* inter, const float* raw, int L, int which, int k1,
                                    int k2, int max_keep, unsigned char* sel_inter /*[3][2L]*/, unsigned char* sel_intra /*[3][L]*/,
                                    float* mask_out /*[3][L]*/) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int kk1 = min(k1, L), kk2 = min(k2, 2 * L);
  for (int it = tid; it < 3 * L; it += nt) {
    const int m = it / L, i = it % L;
    const float* row = intra + m * L;
    const float si = row[i];
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const float sj = row[j];
      rank += (sj > si) || (sj == si && j < i);
    }
    sel_intra[it] = rank < kk1;
  }
  for (int it = tid; it < 6 * L; it += nt) {
    const int q = it / (2 * L), i = it % (2 * L);
    const float* row = inter + q * 2 * L;
    const float si = row[i];
    int rank = 0;
    for (int j = 0; j < 2 * L; ++j) {
      const float sj = row[j];
      rank += (sj > si) || (sj == si && j < i);
    }
    sel_inter[it] = rank < kk2;
  }
  __syncthreads();
  // scatter back (useA.py:166-218): query RGB ranks [NIR|TIR], NIR ranks [RGB|TIR], TIR ranks [RGB|NIR]
  for (int it = tid; it < 3 * L; it += nt) {
    const int m = it / L, l = it % L;
    bool c;
    if (m == 0) c = sel_inter[1 * 2 * L + l] | sel_inter[2 * 2 * L + l];
    else if (m == 1) c = sel_inter[0 * 2 * L + l] | sel_inter[2 * 2 * L + L + l];
    else c = sel_inter[0 * 2 * L + L + l] | sel_inter[1 * 2 * L + L + l];
    const bool s = sel_intra[it];
    const bool r = which == 1 ? s : (which == 2 ? c : (s | c));
    mask_out[it] = r ? 1.f : 0.f;
  }
  __syncthreads();
  if (which == 3 && max_keep >= 0) {
    // useA.py:254-314: order by (selected desc, raw desc, index asc), keep the first max_keep
    unsigned char* keep = sel_intra;  // reuse
    for (int it = tid; it < 3 * L; it += nt) {
      const int m = it / L, i = it % L;
      const float* mk = mask_out + m * L;
      const float* rw = raw + m * L;
      const float mi = mk[i], ri = rw[i];
      int rank = 0;
      for (int j = 0; j < L; ++j) {
        const float mj = mk[j], rj = rw[j];
        rank += (mj > mi) || (mj == mi && (rj > ri || (rj == ri && j < i)));
      }
      keep[it] = rank < max_keep;
    }
    __syncthreads();
    for (int it = tid; it < 3 * L; it += nt) mask_out[it] = keep[it] ? 1.f : 0.f;
    __syncthreads();
  }
}

// grid B, 1024 threads (one CTA per sample: latency-bound, so as many threads as a CTA can have -- one
// rank count per thread).  Softmax the logits into the reference-defined scores, then rank-select.
static __global__ void __launch_bounds__(1024) sim_select_kernel(const float* __restrict__ sel_logits, const float* __restrict__ intra_raw,
                                                                int B, int L, int d, int which, int k1, int k2, int max_keep,
                                                                float* __restrict__ masks, float* __restrict__ masks2) {
  pdl_enter();
  __shared__ float s_inter[3 * 2 * kMaxL], s_intra[3 * kMaxL], s_raw[3 * kMaxL], s_mask[3 * kMaxL];
  __shared__ float s_full[3 * kMaxL];
  __shared__ unsigned char f_inter[3 * 2 * kMaxL], f_intra[3 * kMaxL];
  __shared__ float scratch[33];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float inv = rsqrtf((float)d);
  for (int q = 0; q < 3; ++q) {
    // softmax over all 3L keys of query q (useA.py:129)
    const float* lg = sel_logits + ((int64_t)b * 3 + q) * 3 * L;
    float mx = -INFINITY;
    for (int j = tid; j < 3 * L; j += blockDim.x) mx = fmaxf(mx, lg[j]);
    mx = block_max(mx, scratch);
    float sm = 0.f;
    for (int j = tid; j < 3 * L; j += blockDim.x) {
      const float e = expf(lg[j] - mx);
      s_full[j] = e;
      sm += e;
    }
    sm = block_sum(sm, scratch);
    // D_q: drop the query's own modality, keep order (useA.py:136-151)
    for (int j = tid; j < 3 * L; j += blockDim.x) {
      const int mj = j / L, l = j % L;
      if (mj == q) continue;
      const int slot = (mj > q ? mj - 1 : mj) * L + l;
      s_inter[q * 2 * L + slot] = s_full[j] / sm;
    }
    __syncthreads();
  }
  for (int m = 0; m < 3; ++m) {
    const float* rw = intra_raw + ((int64_t)b * 3 + m) * L;
    float mx = -INFINITY;
    for (int l = tid; l < L; l += blockDim.x) mx = fmaxf(mx, rw[l] * inv);
    mx = block_max(mx, scratch);
    float sm = 0.f;
    for (int l = tid; l < L; l += blockDim.x) {
      const float e = expf(rw[l] * inv - mx);
      s_intra[m * L + l] = e;
      s_raw[m * L + l] = rw[l];
      sm += e;
    }
    sm = block_sum(sm, scratch);
    for (int l = tid; l < L; l += blockDim.x) s_intra[m * L + l] /= sm;
    __syncthreads();
  }
  select_masks(s_intra, s_inter, s_raw, L, which, k1, k2, max_keep, f_inter, f_intra, s_mask);
  for (int it = tid; it < 3 * L; it += blockDim.x) {
    const int m = it / L, l = it % L;
    const float v = s_mask[it];
    masks[((int64_t)m * B + b) * L + l] = v;
    if (masks2) masks2[((int64_t)m * B + b) * L + l] = v;
  }
}

static __global__ void __launch_bounds__(256) select_from_scores_kernel(const float* __restrict__ intra, const float* __restrict__ inter,
                                                                        const float* __restrict__ raw, int B, int L, int which, int k1,
                                                                        int k2, int max_keep, float* __restrict__ masks) {
  pdl_enter();
  __shared__ float s_inter[3 * 2 * kMaxL], s_intra[3 * kMaxL], s_raw[3 * kMaxL], s_mask[3 * kMaxL];
  __shared__ unsigned char f_inter[3 * 2 * kMaxL], f_intra[3 * kMaxL];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int it = tid; it < 3 * L; it += blockDim.x) {
    const int m = it / L, l = it % L;
    s_intra[it] = intra ? intra[((int64_t)m * B + b) * L + l] : 0.f;
    s_raw[it] = raw ? raw[((int64_t)m * B + b) * L + l] : 0.f;
  }
  for (int it = tid; it < 6 * L; it += blockDim.x) {
    const int q = it / (2 * L), i = it % (2 * L);
    s_inter[it] = inter ? inter[((int64_t)q * B + b) * 2 * L + i] : 0.f;
  }
  __syncthreads();
  select_masks(s_intra, s_inter, s_raw, L, which, k1, k2, max_keep, f_inter, f_intra, s_mask);
  for (int it = tid; it < 3 * L; it += blockDim.x) {
    const int m = it / L, l = it % L;
    masks[((int64_t)m * B + b) * L + l] = s_mask[it];
  }
}

int select_from_scores(const float* intra, const float* inter, const float* raw, int B, int L, int which, int k1, int k2,
                       int max_keep, float* masks, cudaStream_t s) {
  SIG_LAUNCH((select_from_scores_kernel), B, 256, 0, s, intra, inter, raw, B, L, which, k1, k2, max_keep, masks);
  SIG_CHECK_LAUNCH();
  return 0;
}

// selected[m][b][l][:] = x * mask  (useA.py:318-320), written contiguous in the token dtype
template <typename T>
static __global__ void mask_mul_kernel(const float* __restrict__ Xf, const float* __restrict__ masks, int64_t rows, int d,
                                       T* __restrict__ selected) {
  pdl_enter();
  const int64_t row = blockIdx.x;
  const float mk = masks[row];
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    float v[8];
    load8(Xf + row * d + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= mk;
    store8(selected + row * d + c, v);
  }
}

// ---------------------------------------------------------------------------------------------
// attention over the (masked) tokens, 24 effective queries per sample
// ---------------------------------------------------------------------------------------------
// effective query index e = q*8 + h (q = CLS modality, h = head).  A CTA handles heads
// {2hp, 2hp+1} x 3 queries: local i = hl*3 + q.
__device__ __forceinline__ int eff_index(int hp, int i) { return (i % 3) * 8 + 2 * hp + i / 3; }

// grid (B, 4), 256 threads, dyn smem: qs[6][d] + lg[6][3L] + msk[3L]
static __global__ void __launch_bounds__(256) sim_attn_fwd_kernel(const float* __restrict__ Xf, const float* __restrict__ maskf,
                                                                  const float* __restrict__ qt, const float* __restrict__ cq, int B,
                                                                  int L, int d, float* __restrict__ xbar, float* __restrict__ amax,
                                                                  float* __restrict__ asum) {
  pdl_enter();
  extern __shared__ float smem[];
  float* qs = smem;               // [6][d]
  float* lg = qs + 6 * d;         // [6][3L]
  float* msk = lg + 6 * 3 * L;    // [3L]
  const int b = blockIdx.x, hp = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int T3 = 3 * L;
  for (int i = tid; i < 6 * d; i += blockDim.x) {
    const int r = i / d, c = i % d;
    qs[i] = qt[((int64_t)b * 24 + eff_index(hp, r)) * d + c];
  }
  for (int j = tid; j < T3; j += blockDim.x) {
    const int m = j / L, l = j % L;
    msk[j] = maskf ? maskf[((int64_t)m * B + b) * L + l] : 1.f;
  }
  __syncthreads();
  // phase A: logits
  for (int j = w; j < T3; j += 8) {
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (msk[j] != 0.f) {
      const int m = j / L, l = j % L;
      const float* x = Xf + (((int64_t)m * B + b) * L + l) * d;
      for (int c = lane * 4; c < d; c += 128) {
        const float4 xv = *reinterpret_cast<const float4*>(x + c);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float4 q = *reinterpret_cast<const float4*>(qs + i * d + c);
          acc[i] += xv.x * q.x + xv.y * q.y + xv.z * q.z + xv.w * q.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) acc[i] = warp_sum(acc[i]) * msk[j];
    }
    if (lane < 6) {
      float v = acc[0];
#pragma unroll
      for (int i = 1; i < 6; ++i) v = lane == i ? acc[i] : v;
      lg[lane * T3 + j] = v + cq[(int64_t)b * 24 + eff_index(hp, lane)];
    }
  }
  __syncthreads();
  // phase B: softmax per effective query (one warp per row)
  if (w < 6) {
    float* row = lg + w * T3;
    float mx = -INFINITY;
    for (int j = lane; j < T3; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sm = 0.f;
    for (int j = lane; j < T3; j += 32) {
      const float e = expf(row[j] - mx);
      row[j] = e;
      sm += e;
    }
    sm = warp_sum(sm);
    const float r = 1.f / sm;
    for (int j = lane; j < T3; j += 32) row[j] *= r;
    if (lane == 0) {
      amax[(int64_t)b * 24 + eff_index(hp, w)] = mx;
      asum[(int64_t)b * 24 + eff_index(hp, w)] = sm;
    }
  }
  __syncthreads();
  // phase C: xbar = sum_j p_j mask_j x_j  (thread per channel)
  for (int c = tid; c < d; c += blockDim.x) {
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < T3; ++j) {
      const float mk = msk[j];
      if (mk == 0.f) continue;
      const int m = j / L, l = j % L;
      const float x = Xf[(((int64_t)m * B + b) * L + l) * d + c] * mk;
#pragma unroll
      for (int i = 0; i < 6; ++i) acc[i] = fmaf(lg[i * T3 + j], x, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) xbar[((int64_t)b * 24 + eff_index(hp, i)) * d + c] = acc[i];
  }
}

// grid B, 256 threads; loops over the four head pairs.
// dyn smem: qs[6][d] + dxs[6][d] + P[6][3L] + dS[6][3L] + msk[3L] + delta[8]
static __global__ void __launch_bounds__(256) sim_attn_bwd_kernel(const float* __restrict__ Xf, const float* __restrict__ maskf,
                                                                  const float* __restrict__ qt, const float* __restrict__ cq,
                                                                  const float* __restrict__ xbar, const float* __restrict__ amax,
                                                                  const float* __restrict__ asum, const float* __restrict__ dxbar,
                                                                  int B, int L, int d, float* __restrict__ dqt, float* __restrict__ dXf) {
  pdl_enter();
  extern __shared__ float smem[];
  const int T3 = 3 * L;
  float* qs = smem;
  float* dxs = qs + 6 * d;
  float* P = dxs + 6 * d;
  float* dS = P + 6 * T3;
  float* msk = dS + 6 * T3;
  float* delta = msk + T3;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int j = tid; j < T3; j += blockDim.x) {
    const int m = j / L, l = j % L;
    msk[j] = maskf ? maskf[((int64_t)m * B + b) * L + l] : 1.f;
  }
  for (int hp = 0; hp < 4; ++hp) {
    __syncthreads();
    for (int i = tid; i < 6 * d; i += blockDim.x) {
      const int r = i / d, c = i % d;
      const int64_t e = (int64_t)b * 24 + eff_index(hp, r);
      qs[i] = qt[e * d + c];
      dxs[i] = dxbar[e * d + c];
    }
    __syncthreads();
    if (w < 6) {  // delta_i = dxbar_i . xbar_i
      const int64_t e = (int64_t)b * 24 + eff_index(hp, w);
      float a = 0.f;
      for (int c = lane; c < d; c += 32) a += dxs[w * d + c] * xbar[e * d + c];
      a = warp_sum(a);
      if (lane == 0) delta[w] = a;
    }
    __syncthreads();
    // phase A: p and dS per token
    for (int j = w; j < T3; j += 8) {
      float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dp[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float mk = msk[j];
      if (mk != 0.f) {
        const int m = j / L, l = j % L;
        const float* x = Xf + (((int64_t)m * B + b) * L + l) * d;
        for (int c = lane * 4; c < d; c += 128) {
          const float4 xv = *reinterpret_cast<const float4*>(x + c);
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const float4 q = *reinterpret_cast<const float4*>(qs + i * d + c);
            const float4 g = *reinterpret_cast<const float4*>(dxs + i * d + c);
            s[i] += xv.x * q.x + xv.y * q.y + xv.z * q.z + xv.w * q.w;
            dp[i] += xv.x * g.x + xv.y * g.y + xv.z * g.z + xv.w * g.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          s[i] = warp_sum(s[i]) * mk;
          dp[i] = warp_sum(dp[i]) * mk;
        }
      }
      if (lane < 6) {
        float sv = s[0], dv = dp[0];
#pragma unroll
        for (int i = 1; i < 6; ++i) {
          sv = lane == i ? s[i] : sv;
          dv = lane == i ? dp[i] : dv;
        }
        const int64_t e = (int64_t)b * 24 + eff_index(hp, lane);
        const float p = expf(sv + cq[e] - amax[e]) / asum[e];
        P[lane * T3 + j] = p;
        dS[lane * T3 + j] = p * (dv - delta[lane]);
      }
    }
    __syncthreads();
    // phase C1: dx_j = mask_j * sum_i (p_ij dxbar_i + dS_ij qt_i), accumulated over head pairs
    for (int j = w; j < T3; j += 8) {
      const int m = j / L, l = j % L;
      float* dx = dXf + (((int64_t)m * B + b) * L + l) * d;
      const float mk = msk[j];
      for (int c = lane * 4; c < d; c += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mk != 0.f) {
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const float p = P[i * T3 + j], ds = dS[i * T3 + j];
            const float4 q = *reinterpret_cast<const float4*>(qs + i * d + c);
            const float4 g = *reinterpret_cast<const float4*>(dxs + i * d + c);
            acc.x += p * g.x + ds * q.x; acc.y += p * g.y + ds * q.y;
            acc.z += p * g.z + ds * q.z; acc.w += p * g.w + ds * q.w;
          }
          acc.x *= mk; acc.y *= mk; acc.z *= mk; acc.w *= mk;
          if (hp > 0) {
            const float4 old = *reinterpret_cast<const float4*>(dx + c);
            acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
          }
        }
        if (mk != 0.f || hp == 0) *reinterpret_cast<float4*>(dx + c) = acc;
      }
    }
    // phase C2: dqt_i = sum_j dS_ij mask_j x_j
    for (int c = tid; c < d; c += blockDim.x) {
      float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < T3; ++j) {
        const float mk = msk[j];
        if (mk == 0.f) continue;
        const int m = j / L, l = j % L;
        const float x = Xf[(((int64_t)m * B + b) * L + l) * d + c] * mk;
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[i] = fmaf(dS[i * T3 + j], x, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) dqt[((int64_t)b * 24 + eff_index(hp, i)) * d + c] = acc[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// gradient writers (fp32 scratch -> strided token-dtype destinations)
// ---------------------------------------------------------------------------------------------
struct GradPtrs {
  void* dpatch[3];
  void* dcls[3];
  int64_t psb[3], psl[3], csb[3];
  int accumulate;
};

// grid (B*L + B, 3).  dXf [3][B][L][d] (may be NULL -> zeros), dclsf [B][3][d] (may be NULL -> zeros)
template <typename T>
static __global__ void write_token_grads_kernel(GradPtrs gp, const float* __restrict__ dXf, const float* __restrict__ dclsf, int B,
                                                int L, int d) {
  pdl_enter();
  const int m = blockIdx.y;
  const int64_t row = blockIdx.x;
  const float* src;
  T* dst;
  if (row < (int64_t)B * L) {
    const int b = (int)(row / L), l = (int)(row % L);
    src = dXf ? dXf + (((int64_t)m * B + b) * L + l) * d : nullptr;
    dst = static_cast<T*>(gp.dpatch[m]) + b * gp.psb[m] + l * gp.psl[m];
  } else {
    if (!gp.dcls[m]) return;
    const int b = (int)(row - (int64_t)B * L);
    src = dclsf ? dclsf + ((int64_t)b * 3 + m) * d : nullptr;
    dst = static_cast<T*>(gp.dcls[m]) + b * gp.csb[m];
  }
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (src) load8(src + c, v);
    if (gp.accumulate) {
      float o[8];
      load8(dst + c, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += o[i];
    }
    store8(dst + c, v);
  }
}

int write_token_grads(const sig_token_grads* g, int dtype, const float* dXf, const float* dclsf, int B, int L, int d,
                      cudaStream_t s) {
  SIG_PHASE("write_token_grads");
  GradPtrs gp;
  for (int m = 0; m < 3; ++m) {
    gp.dpatch[m] = g->dpatch[m]; gp.dcls[m] = g->dcls[m];
    gp.psb[m] = g->patch_stride_b[m]; gp.psl[m] = g->patch_stride_l[m]; gp.csb[m] = g->cls_stride_b[m];
  }
  gp.accumulate = g->accumulate;
  dim3 grid((unsigned)((int64_t)B * L + B), 3);
  const int threads = d / 8 >= 128 ? 128 : 64;
  if (dtype == SIG_BF16)
    SIG_LAUNCH((write_token_grads_kernel<__nv_bfloat16>), grid, threads, 0, s, gp, dXf, dclsf, B, L, d);
  else
    SIG_LAUNCH((write_token_grads_kernel<float>), grid, threads, 0, s, gp, dXf, dclsf, B, L, d);
  SIG_CHECK_LAUNCH();
  return 0;
}

// dcls[m][b] (+)= dclsf[b][m]   grid (B, 3)
template <typename T>
static __global__ void write_cls_grads_kernel(GradPtrs gp, const float* __restrict__ dclsf, int d) {
  pdl_enter();
  const int b = blockIdx.x, m = blockIdx.y;
  if (!gp.dcls[m]) return;
  T* dst = static_cast<T*>(gp.dcls[m]) + b * gp.csb[m];
  const float* src = dclsf + ((int64_t)b * 3 + m) * d;
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    float v[8];
    load8(src + c, v);
    if (gp.accumulate) {
      float o[8];
      load8(dst + c, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += o[i];
    }
    store8(dst + c, v);
  }
}

static __global__ void fill_kernel(float* __restrict__ p, float v, int64_t n) {
  pdl_enter();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// ctx layout
// ---------------------------------------------------------------------------------------------
struct SimCtx {
  float *Xf, *clsf, *maskf;
  float *qsel, *qtsel, *csel, *sel_logits, *intra_raw;
  float *qatt, *qtatt, *catt, *amax, *asum, *xbar, *o, *attn;
  float *r1, *mu1, *rstd1, *y1, *a1, *h1, *f, *r2, *mu2, *rstd2;
  // backward scratch
  float *dr2, *dyx, *dyf, *dh1, *dr1, *dob, *dxbar, *dqt, *dqatt, *dXf;
  // tensor-core (bf16) token path
  bool tc;
  __nv_bfloat16 *DXQT, *Ptok, *PT, *PdS, *dST;
  float *S32, *delta;
  __nv_bfloat16 *clsb2;
  __nv_bfloat16 *Wb, *clsb, *qattb, *xbarb, *ob, *y1b, *h1b, *dr2b, *da1b, *dr1b, *dobb, *dqtb, *dqattb;
  size_t bytes;
};

static SimCtx sim_ctx(void* base, int B, int L, int d, bool tc = false) {
  Arena a(base);
  SimCtx c;
  c.tc = tc;
  const size_t R = (size_t)B * 3;
  c.Xf = a.take<float>(tc ? 0 : (size_t)3 * B * L * d);
  c.clsf = a.take<float>(R * d);
  c.maskf = a.take<float>((size_t)3 * B * L);
  c.qsel = a.take<float>(R * d);
  c.qtsel = a.take<float>(R * d);
  c.csel = a.take<float>(R);
  c.sel_logits = a.take<float>(R * 3 * L);
  c.intra_raw = a.take<float>(R * L);
  c.qatt = a.take<float>(R * d);
  c.qtatt = a.take<float>(R * 8 * d);
  c.catt = a.take<float>(R * 8);
  c.amax = a.take<float>(R * 8);
  c.asum = a.take<float>(R * 8);
  c.xbar = a.take<float>(R * 8 * d);
  c.o = a.take<float>(R * d);
  c.r1 = a.take<float>(R * d);
  c.mu1 = a.take<float>(R);
  c.rstd1 = a.take<float>(R);
  c.y1 = a.take<float>(R * d);
  // [attn | a1 | f] contiguous: the split-K targets of the tensor-core MLP chain are zeroed by ONE memset (sim_mlp_tc.inl)
  c.attn = a.take<float>(R * d);
  c.a1 = a.take<float>(R * 2 * d);
  c.f = a.take<float>(R * d);
  c.h1 = a.take<float>(R * 2 * d);
  c.r2 = a.take<float>(R * d);
  c.mu2 = a.take<float>(R);
  c.rstd2 = a.take<float>(R);
  c.dr2 = a.take<float>(R * d);
  c.dyx = a.take<float>(R * d);
  c.dyf = a.take<float>(R * d);
  c.dh1 = a.take<float>(R * 2 * d);
  c.dr1 = a.take<float>(R * d);
  c.dob = a.take<float>(R * d);
  c.dxbar = a.take<float>(R * 8 * d);
  c.dqt = a.take<float>(R * 8 * d);
  c.dqatt = a.take<float>(R * d);
  c.dXf = a.take<float>(tc ? 0 : (size_t)3 * B * L * d);
  c.DXQT = a.take<__nv_bfloat16>(tc ? (size_t)B * 64 * d : 0);
  c.Ptok = a.take<__nv_bfloat16>(tc ? (size_t)B * 384 * 32 : 0);
  c.PT = a.take<__nv_bfloat16>(tc ? (size_t)B * 32 * 384 : 0);
  c.PdS = a.take<__nv_bfloat16>(tc ? (size_t)B * 384 * 64 : 0);
  c.dST = a.take<__nv_bfloat16>(tc ? (size_t)B * 32 * 384 : 0);
  c.S32 = a.take<float>(tc ? (size_t)B * 384 * 32 : 0);
  c.delta = a.take<float>(tc ? (size_t)B * 32 : 0);
  c.Wb = a.take<__nv_bfloat16>(tc ? (size_t)8 * d * d : 0);
  c.clsb = a.take<__nv_bfloat16>(tc ? R * d : 0);
  c.clsb2 = a.take<__nv_bfloat16>(tc ? R * 2 * d : 0);
  c.qattb = a.take<__nv_bfloat16>(tc ? R * d : 0);
  c.xbarb = a.take<__nv_bfloat16>(tc ? R * 8 * d : 0);
  c.ob = a.take<__nv_bfloat16>(tc ? R * d : 0);
  c.y1b = a.take<__nv_bfloat16>(tc ? R * d : 0);
  c.h1b = a.take<__nv_bfloat16>(tc ? R * 2 * d : 0);
  c.dr2b = a.take<__nv_bfloat16>(tc ? R * d : 0);
  c.da1b = a.take<__nv_bfloat16>(tc ? R * 2 * d : 0);
  c.dr1b = a.take<__nv_bfloat16>(tc ? R * d : 0);
  c.dobb = a.take<__nv_bfloat16>(tc ? R * d : 0);
  c.dqtb = a.take<__nv_bfloat16>(tc ? R * 8 * d : 0);
  c.dqattb = a.take<__nv_bfloat16>(tc ? R * d : 0);
  c.bytes = a.off;
  return c;
}

size_t sim_ctx_bytes(int B, int L, int d) { return sim_ctx(nullptr, B, L, d).bytes; }

static bool sim_tc_ok(int dtype, int L, unsigned flags) {
  return dtype == SIG_BF16 && L == 128 && !(flags & SIG_FLAG_FORCE_SIMT);
}
size_t sim_ctx_bytes_for(int B, int L, int d, int dtype, unsigned flags) {
  return sim_ctx(nullptr, B, L, d, sim_tc_ok(dtype, L, flags)).bytes;
}

static SimTcBufs tc_bufs(const SimCtx& c, const float* maskf) {
  SimTcBufs k{};
  k.maskf = maskf; k.qtatt = c.qtatt; k.catt = c.catt; k.DXQT = c.DXQT; k.S32 = c.S32; k.Ptok = c.Ptok; k.PT = c.PT;
  k.xbar = c.xbar; k.dxbar = c.dxbar; k.delta = c.delta; k.PdS = c.PdS; k.dST = c.dST; k.dqt = c.dqt;
  k.xbarb = c.xbarb; k.dqtb = c.dqtb;
  return k;
}

#include "sim_mlp_tc.inl"

// ---------------------------------------------------------------------------------------------
// orchestration
// ---------------------------------------------------------------------------------------------
static int run_selection(const SimCtx& c, const sig_tokens* tok, const sig_sim_params* p, int B, int L, int d, int which, int k1,
                         int k2, int max_keep, float* masks_out, cudaStream_t s) {
  const int R = 3 * B;
  SIG_PHASE("sim_select");
  if (c.tc && p->sel_fold) {
    // qt = M cls + v on the tensor cores: cls is exact in bf16, M = hi + lo  ->  [cls | cls] . [M_hi | M_lo]^T
    const sig_sel_fold* f = p->sel_fold;
    TcGemmDesc t = tc_desc();
    t.A = tc_k2d(c.clsb2, R, 2 * d, 2 * d);
    t.B = tc_k2d(f->m_hl, d, 2 * d, 2 * d);
    t.M = R; t.N = d; t.K = 2 * d;
    t.C[0] = c.qtsel; t.ldc = d; t.bias[0] = f->v;
    SIG_TRY(tc_gemm(t, s));
    SIG_LAUNCH((csel_fold_kernel), R, 128, 0, s, c.clsf, f->u, f->s0, d, c.csel);
    SIG_CHECK_LAUNCH();
  } else {
    // q = W_q cls + b_q (useA.py:123); qt = W_k^T q; c = q . b_k
    SIG_TRY(launch_gemm(gemm_nt(c.clsf, d, p->sel_wq, d, c.qsel, d, p->sel_bq, R, d, d), s));
    SIG_TRY(launch_gemm(gemm_nn(c.qsel, d, p->sel_wk, d, c.qtsel, d, R, d, d), s));
    SIG_TRY(launch_gemm(gemm_nt(c.qsel, d, p->sel_bk, d, c.csel, 1, nullptr, R, 1, d), s));
  }
  dim3 grid((unsigned)ceil_div(L, 32), 3, (unsigned)B);
  const bool rows_contig = tok->patch_stride_l[0] == d && tok->patch_stride_l[1] == d && tok->patch_stride_l[2] == d;
  bool pool_done = false;
  if (c.tc && tok_ring_enabled() && L == kMaxL && (d == 512 || d == 768) && rows_contig) {
    SIG_PHASE("sim_scores");   // streaming ring: 4 x 48 KB in flight per SM (tok_ring.cuh)
    TokSrc3 src;
    for (int m = 0; m < 3; ++m) { src.patch[m] = tok->patch[m]; src.psb[m] = tok->patch_stride_b[m]; }
    const int n_items = 3 * B * (kMaxL / 32);
    const int ctas = n_items < tc_num_sms() ? n_items : tc_num_sms();
    if (scores_split_enabled()) {
      if (d == 768) {
        ensure_dyn_smem(sim_scores_split_kernel<768>, (int)ScoreSplit<768>::kSmemBytes);
        SIG_LAUNCH((sim_scores_split_kernel<768>), ctas, ScoreSplit<768>::kThreads, ScoreSplit<768>::kSmemBytes, s, src, c.clsf, c.qtsel, c.csel, B,
                   n_items, c.sel_logits, c.intra_raw);
      } else {
        ensure_dyn_smem(sim_scores_split_kernel<512>, (int)ScoreSplit<512>::kSmemBytes);
        SIG_LAUNCH((sim_scores_split_kernel<512>), ctas, ScoreSplit<512>::kThreads, ScoreSplit<512>::kSmemBytes, s, src, c.clsf, c.qtsel, c.csel, B,
                   n_items, c.sel_logits, c.intra_raw);
      }
    } else {
      // p->pool_out (FusionHead): the pass also delivers GAM's mean pool (tok_ring.cuh, kPool)
      float* pool = p->pool_out;
      if (pool) cudaMemsetAsync(pool, 0, (size_t)3 * B * d * sizeof(float), s);   // split groups are added atomically onto zero
#define SIG_SCORES(DD, PP)                                                                                                      \
  do {                                                                                                                          \
    ensure_dyn_smem(sim_scores_ring_kernel<DD, PP>, (int)sim_scores_ring_smem<DD>(PP));                                          \
    SIG_LAUNCH((sim_scores_ring_kernel<DD, PP>), ctas, TokRing<DD>::kThreads, sim_scores_ring_smem<DD>(PP), s, src, c.clsf, c.qtsel, \
               c.csel, B, n_items, c.sel_logits, c.intra_raw, pool);                                                            \
  } while (0)
      if (d == 768 && pool) SIG_SCORES(768, true);
      else if (d == 768) SIG_SCORES(768, false);
      else if (pool) SIG_SCORES(512, true);
      else SIG_SCORES(512, false);
#undef SIG_SCORES
      pool_done = pool != nullptr;
    }
  } else if (c.tc) {
    SIG_PHASE("sim_scores");
    SIG_LAUNCH((sim_scores_tok_kernel<__nv_bfloat16>), dim3(3, (unsigned)B), 256, 4 * d * sizeof(float), s, tok_ptrs(tok), c.clsf, c.qtsel, c.csel, B, L, d,
                                                                                 c.sel_logits, c.intra_raw);
  } else
    SIG_LAUNCH((sim_scores_kernel), grid, 256, 4 * d * sizeof(float), s, c.Xf, c.clsf, c.qtsel, c.csel, B, L, d, c.sel_logits, c.intra_raw);
  SIG_CHECK_LAUNCH();
  if (p->pool_out) {
    // the caller (FusionHead) relies on the pool being delivered: layouts the ring kernel does not take get AlignM's own
    // pooling kernel, on this stream
    if (!pool_done) SIG_TRY(align_pool_tokens(tok, p->pool_out, s));
    if (p->pool_event) cudaEventRecord((cudaEvent_t)p->pool_event, s);
  }
  SIG_LAUNCH((sim_select_kernel), B, 1024, 0, s, c.sel_logits, c.intra_raw, B, L, d, which, k1, k2, max_keep, c.maskf, masks_out);
  SIG_CHECK_LAUNCH();
  return 0;
}

static size_t attn_fwd_smem(int L, int d) { return (size_t)(6 * d + 6 * 3 * L + 3 * L) * sizeof(float); }
static size_t attn_bwd_smem(int L, int d) { return (size_t)(12 * d + 12 * 3 * L + 3 * L + 8) * sizeof(float); }

template <typename OutT>
static int run_attention_fwd(const SimCtx& c, const sig_tokens* tok, const sig_sim_params* p, const float* maskf, int B, int L,
                             int d, OutT* out, const Fork& sel_fork, cudaStream_t s) {
  const int R = 3 * B, hd = d / kHeads;
  const float scale = 1.0f / sqrtf((float)hd);
  const float* wq = p->in_proj_w;
  const float* wk = p->in_proj_w + (size_t)d * d;
  const float* wv = p->in_proj_w + (size_t)2 * d * d;
  const float* bq = p->in_proj_b;
  const float* bk = p->in_proj_b + d;
  const float* bv = p->in_proj_b + 2 * d;
  if (c.tc) {
    {
      SIG_PHASE("sim_attn_prep");
      SIG_TRY(attn_prep_tc(c, p, B, d, s));
    }
    if (sel_fork.ok()) sel_fork.join(s);   // the token kernels need the selection masks
    {
      SIG_PHASE("sim_attn_tokens_fwd");
      SIG_TRY(sim_tc_tokens_fwd(tok, tc_bufs(c, maskf), s));
    }
    SIG_PHASE("sim_post");
    return attn_post_tc<OutT>(c, p, B, d, out, s);
  }
  // q = W_q cls + b_q  (unscaled; the 1/sqrt(hd) of MHA is folded into qt and c)
  {
  SIG_PHASE("sim_attn_prep");
  SIG_TRY(launch_gemm(gemm_nt(c.clsf, d, wq, d, c.qatt, d, bq, R, d, d), s));
  {  // qt[(b,q),h,:] = scale * q_h W_k^h   (batched over heads)
    Gemm g = gemm_nn(c.qatt, d, wk, d, c.qtatt, 8 * (int64_t)d, R, d, hd);
    g.batch = kHeads; g.az = hd; g.bz = (int64_t)hd * d; g.cz = d; g.alpha = scale;
    SIG_TRY(launch_gemm(g, s));
  }
  {  // c[(b,q),h] = scale * q_h . b_k^h
    Gemm g = gemm_nt(c.qatt, d, bk, hd, c.catt, 8, nullptr, R, 1, hd);
    g.batch = kHeads; g.az = hd; g.bz = hd; g.cz = 1; g.alpha = scale;
    SIG_TRY(launch_gemm(g, s));
  }
  }
  {
  SIG_PHASE("sim_attn_tokens_fwd");
  if (c.tc) {
    SIG_TRY(sim_tc_tokens_fwd(tok, tc_bufs(c, maskf), s));
  } else {
    const size_t sm = attn_fwd_smem(L, d);
    cudaFuncSetAttribute(sim_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);   // (d-dependent size: set per call)
    SIG_LAUNCH((sim_attn_fwd_kernel), dim3(B, 4), 256, sm, s, c.Xf, maskf, c.qtatt, c.catt, B, L, d, c.xbar, c.amax, c.asum);
    SIG_CHECK_LAUNCH();
  }
  }
  SIG_PHASE("sim_post");
  {  // o_h = W_v^h xbar_h + b_v^h
    Gemm g = gemm_nt(c.xbar, 8 * (int64_t)d, wv, d, c.o, d, bv, R, hd, d);
    g.batch = kHeads; g.az = d; g.bz = (int64_t)hd * d; g.cz = hd; g.biasz = hd;
    SIG_TRY(launch_gemm(g, s));
  }
  SIG_TRY(launch_gemm(gemm_nt(c.o, d, p->out_proj_w, d, c.attn, d, p->out_proj_b, R, d, d), s));
  SIG_LAUNCH((layernorm_fwd_kernel<float>), R, 256, 0, s, c.attn, c.clsf, p->ln1_w, p->ln1_b, d, c.r1, c.mu1, c.rstd1, c.y1, (__nv_bfloat16*)nullptr);
  SIG_CHECK_LAUNCH();
  {
    Gemm g = gemm_nt(c.y1, d, p->ffn0_w, d, c.h1, 2 * (int64_t)d, p->ffn0_b, R, 2 * d, d);
    g.act = 1; g.pre = c.a1;
    SIG_TRY(launch_gemm(g, s));
  }
  SIG_TRY(launch_gemm(gemm_nt(c.h1, 2 * (int64_t)d, p->ffn2_w, 2 * (int64_t)d, c.f, d, p->ffn2_b, R, d, 2 * d), s));
  SIG_LAUNCH((layernorm_fwd_kernel<OutT>), R, 256, 0, s, c.f, c.y1, p->ln2_w, p->ln2_b, d, c.r2, c.mu2, c.rstd2, out, (__nv_bfloat16*)nullptr);
  SIG_CHECK_LAUNCH();
  return 0;
}

template <typename InT>
static int run_attention_bwd(const SimCtx& c, const sig_tokens* tok, const sig_token_grads* dtok, const sig_sim_params* p,
                             const float* maskf, int B, int L, int d, const InT* dout, const sig_sim_param_grads* g,
                             cudaStream_t s) {
  const int R = 3 * B, hd = d / kHeads;
  const float scale = 1.0f / sqrtf((float)hd);
  const float* wq = p->in_proj_w;
  const float* wk = p->in_proj_w + (size_t)d * d;
  const float* wv = p->in_proj_w + (size_t)2 * d * d;
  float* dwq = g->in_proj_w;
  float* dwk = g->in_proj_w + (size_t)d * d;
  float* dwv = g->in_proj_w + (size_t)2 * d * d;
  if (c.tc) {
    const Fork fk = get_fork(FORK_SIM_BWD);   // weight-gradient GEMMs; joined by sim_backward
    {
      SIG_PHASE("sim_post_bwd");
      SIG_TRY((attn_post_bwd_tc<InT>(c, p, B, d, dout, g, fk, s)));
    }
    {
      SIG_PHASE("sim_attn_tokens_bwd");
      SIG_TRY(sim_tc_tokens_bwd(tok, tc_bufs(c, maskf), dtok, s));
    }
    SIG_PHASE("sim_attn_prep_bwd");
    return attn_prep_bwd_tc(c, p, B, d, g, fk, s);
  }
  {
  SIG_PHASE("sim_post_bwd");
  // LN2
  SIG_LAUNCH((layernorm_bwd_kernel<InT>), R, 256, 0, s, dout, c.r2, p->ln2_w, c.mu2, c.rstd2, nullptr, d, c.dr2, c.dyx, c.dyf, (__nv_bfloat16*)nullptr);
  SIG_CHECK_LAUNCH();
  SIG_TRY(launch_colsum(c.dyx, d, R, d, g->ln2_w, 1.f, s));
  SIG_TRY(launch_colsum(c.dyf, d, R, d, g->ln2_b, 1.f, s));
  // FFN
  SIG_TRY(launch_colsum(c.dr2, d, R, d, g->ffn2_b, 1.f, s));
  SIG_TRY(launch_gemm(gemm_tn(c.dr2, d, c.h1, 2 * (int64_t)d, g->ffn2_w, 2 * (int64_t)d, d, 2 * d, R), s));
  SIG_TRY(launch_gemm(gemm_nn(c.dr2, d, p->ffn2_w, 2 * (int64_t)d, c.dh1, 2 * (int64_t)d, R, 2 * d, d), s));
  {
    const int64_t n = (int64_t)R * 2 * d;
    SIG_LAUNCH((gelu_bwd_kernel), (unsigned)ceil_div(n, 256), 256, 0, s, c.dh1, c.a1, c.dh1, n);  // dh1 := da1
    SIG_CHECK_LAUNCH();
  }
  SIG_TRY(launch_colsum(c.dh1, 2 * (int64_t)d, R, 2 * d, g->ffn0_b, 1.f, s));
  SIG_TRY(launch_gemm(gemm_tn(c.dh1, 2 * (int64_t)d, c.y1, d, g->ffn0_w, d, 2 * d, d, R), s));
  {  // dy1 = dr2 + da1 W1
    Gemm gg = gemm_nn(c.dh1, 2 * (int64_t)d, p->ffn0_w, d, c.dr2, d, R, d, 2 * d);
    gg.accumulate = 1;
    SIG_TRY(launch_gemm(gg, s));
  }
  // LN1
  SIG_LAUNCH((layernorm_bwd_kernel<float>), R, 256, 0, s, c.dr2, c.r1, p->ln1_w, c.mu1, c.rstd1, nullptr, d, c.dr1, c.dyx, c.dyf, (__nv_bfloat16*)nullptr);
  SIG_CHECK_LAUNCH();
  SIG_TRY(launch_colsum(c.dyx, d, R, d, g->ln1_w, 1.f, s));
  SIG_TRY(launch_colsum(c.dyf, d, R, d, g->ln1_b, 1.f, s));
  // out_proj
  SIG_TRY(launch_colsum(c.dr1, d, R, d, g->out_proj_b, 1.f, s));
  SIG_TRY(launch_gemm(gemm_tn(c.dr1, d, c.o, d, g->out_proj_w, d, d, d, R), s));
  SIG_TRY(launch_gemm(gemm_nn(c.dr1, d, p->out_proj_w, d, c.dob, d, R, d, d), s));
  // value projection
  SIG_TRY(launch_colsum(c.dob, d, R, d, g->in_proj_b + 2 * d, 1.f, s));
  {  // dW_v^h = do_h^T xbar_h
    Gemm gg = gemm_tn(c.dob, d, c.xbar, 8 * (int64_t)d, dwv, d, hd, d, R);
    gg.batch = kHeads; gg.az = hd; gg.bz = d; gg.cz = (int64_t)hd * d;
    SIG_TRY(launch_gemm(gg, s));
  }
  {  // dxbar_h = do_h W_v^h
    Gemm gg = gemm_nn(c.dob, d, wv, d, c.dxbar, 8 * (int64_t)d, R, d, hd);
    gg.batch = kHeads; gg.az = hd; gg.bz = (int64_t)hd * d; gg.cz = d;
    SIG_TRY(launch_gemm(gg, s));
  }
  }
  {
  SIG_PHASE("sim_attn_tokens_bwd");
  if (c.tc) {
    SIG_TRY(sim_tc_tokens_bwd(tok, tc_bufs(c, maskf), dtok, s));
  } else {
    const size_t sm = attn_bwd_smem(L, d);
    cudaFuncSetAttribute(sim_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);   // (d-dependent size: set per call)
    SIG_LAUNCH((sim_attn_bwd_kernel), B, 256, sm, s, c.Xf, maskf, c.qtatt, c.catt, c.xbar, c.amax, c.asum, c.dxbar, B, L, d, c.dqt, c.dXf);
    SIG_CHECK_LAUNCH();
  }
  }
  SIG_PHASE("sim_attn_prep_bwd");
  {  // dq_h = scale * dqt_h W_k^hT
    Gemm gg = gemm_nt(c.dqt, 8 * (int64_t)d, wk, d, c.dqatt, d, nullptr, R, hd, d);
    gg.batch = kHeads; gg.az = d; gg.bz = (int64_t)hd * d; gg.cz = hd; gg.alpha = scale;
    SIG_TRY(launch_gemm(gg, s));
  }
  {  // dW_k^h = scale * q_h^T dqt_h
    Gemm gg = gemm_tn(c.qatt, d, c.dqt, 8 * (int64_t)d, dwk, d, hd, d, R);
    gg.batch = kHeads; gg.az = hd; gg.bz = d; gg.cz = (int64_t)hd * d; gg.alpha = scale;
    SIG_TRY(launch_gemm(gg, s));
  }
  cudaMemsetAsync(g->in_proj_b + d, 0, d * sizeof(float), s);  // key bias: softmax shift invariance => exactly 0
  SIG_TRY(launch_colsum(c.dqatt, d, R, d, g->in_proj_b, 1.f, s));
  SIG_TRY(launch_gemm(gemm_tn(c.dqatt, d, c.clsf, d, dwq, d, d, d, R), s));
  {  // dcls = dr1 (residual) + dq W_q
    Gemm gg = gemm_nn(c.dqatt, d, wq, d, c.dr1, d, R, d, d);
    gg.accumulate = 1;
    SIG_TRY(launch_gemm(gg, s));
  }
  return 0;
}

static int check_sim_params(const sig_sim_params* p, bool need_sel, bool need_attn) {
  if (!p) return SIG_ERR_NULL;
  if (need_sel && (!p->sel_wq || !p->sel_bq || !p->sel_wk || !p->sel_bk)) return SIG_ERR_NULL;
  if (need_attn && (!p->in_proj_w || !p->in_proj_b || !p->out_proj_w || !p->out_proj_b || !p->ffn0_w || !p->ffn0_b ||
                    !p->ffn2_w || !p->ffn2_b || !p->ln1_w || !p->ln1_b || !p->ln2_w || !p->ln2_b))
    return SIG_ERR_NULL;
  return 0;
}

int sim_fold_selection(const sig_sim_params* p, int d, void* m_hl, float* v, float* u, float* s0, float* ws, cudaStream_t s) {
  // M[n][k] = sum_j W_k[j][n] W_q[j][k]
  SIG_TRY(launch_gemm(gemm_tn(p->sel_wk, d, p->sel_wq, d, ws, d, d, d, d), s));
  SIG_LAUNCH((split_hl_kernel), d, 128, 0, s, ws, d, static_cast<__nv_bfloat16*>(m_hl));
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((gemv_t_kernel), (unsigned)ceil_div(d, 32), 1024, 0, s, p->sel_wk, d, p->sel_bq, d, d, v);   // v = W_k^T b_q
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((gemv_t_kernel), (unsigned)ceil_div(d, 32), 1024, 0, s, p->sel_wq, d, p->sel_bk, d, d, u);   // u = W_q^T b_k
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((dot_kernel), 1, 256, 0, s, p->sel_bq, p->sel_bk, d, s0);
  SIG_CHECK_LAUNCH();
  return 0;
}

int sim_forward(const sig_tokens* tok, const sig_sim_params* p, bool do_select, const float* ext_masks, int k1, int k2,
                int max_keep, void* out, float* masks_out, void* ctx, size_t ctx_bytes, unsigned flags, cudaStream_t s) {
  SIG_TRY(check_tokens(tok, true));
  SIG_TRY(check_sim_params(p, do_select, true));
  if (!out || !ctx) return SIG_ERR_NULL;
  if (tok->d % (8 * kHeads) != 0) return SIG_ERR_SHAPE;
  const int B = tok->B, L = tok->L, d = tok->d;
  const bool tcp = sim_tc_ok(tok->dtype, L, flags);
  if (ctx_bytes < sim_ctx(nullptr, B, L, d, tcp).bytes) return SIG_ERR_WORKSPACE;
  if (do_select && (k1 < 1 || k2 < 1 || max_keep > L)) return SIG_ERR_SHAPE;
  SimCtx c = sim_ctx(ctx, B, L, d, tcp);
  if (tcp) {
    SIG_PHASE("convert_tokens");
    SIG_LAUNCH((gather_cls_kernel<__nv_bfloat16>), dim3(B, 3), 96, 0, s, tok_ptrs(tok), d, c.clsf, c.clsb, c.clsb2);
    SIG_CHECK_LAUNCH();
  } else {
    SIG_TRY(convert_tokens(tok, c.Xf, c.clsf, s));
  }
  const float* maskf = nullptr;
  Fork sel_fork;
  if (do_select) {
    // the selection chain (scores -> softmax -> rank-select) is independent of the attention prep
    if (tcp) sel_fork = get_fork(FORK_SIM_FWD);
    if (sel_fork.ok()) sel_fork.fork(s);
    SIG_TRY(run_selection(c, tok, p, B, L, d, 3, k1, k2, max_keep, masks_out, sel_fork.ok() ? sel_fork.side : s));
    maskf = c.maskf;
  } else if (ext_masks) {
    cudaMemcpyAsync(c.maskf, ext_masks, (size_t)3 * B * L * sizeof(float), cudaMemcpyDeviceToDevice, s);
    maskf = c.maskf;
  } else if (tcp) {
    SIG_LAUNCH((fill_kernel), (unsigned)ceil_div((int64_t)3 * B * L, 256), 256, 0, s, c.maskf, 1.f, (int64_t)3 * B * L);
    SIG_CHECK_LAUNCH();
    maskf = c.maskf;
  }
  if (tok->dtype == SIG_BF16)
    return run_attention_fwd<__nv_bfloat16>(c, tok, p, maskf, B, L, d, static_cast<__nv_bfloat16*>(out), sel_fork, s);
  return run_attention_fwd<float>(c, tok, p, maskf, B, L, d, static_cast<float*>(out), sel_fork, s);
}

int sim_backward(const sig_tokens* tok, const sig_sim_params* p, bool has_masks, const void* dout, const sig_token_grads* dtok,
                 const sig_sim_param_grads* dp, void* ctx, size_t ctx_bytes, unsigned flags, cudaStream_t s) {
  SIG_TRY(check_tokens(tok, true));
  SIG_TRY(check_sim_params(p, false, true));
  SIG_TRY(check_token_grads(dtok, tok->dtype, true));
  if (!dout || !ctx || !dp) return SIG_ERR_NULL;
  if (!dp->in_proj_w || !dp->in_proj_b || !dp->out_proj_w || !dp->out_proj_b || !dp->ffn0_w || !dp->ffn0_b || !dp->ffn2_w ||
      !dp->ffn2_b || !dp->ln1_w || !dp->ln1_b || !dp->ln2_w || !dp->ln2_b)
    return SIG_ERR_NULL;
  const int B = tok->B, L = tok->L, d = tok->d;
  const bool tcp = sim_tc_ok(tok->dtype, L, flags);
  if (ctx_bytes < sim_ctx(nullptr, B, L, d, tcp).bytes) return SIG_ERR_WORKSPACE;
  SimCtx c = sim_ctx(ctx, B, L, d, tcp);
  const float* maskf = (has_masks || tcp) ? c.maskf : nullptr;
  if (tok->dtype == SIG_BF16)
    SIG_TRY(run_attention_bwd<__nv_bfloat16>(c, tok, dtok, p, maskf, B, L, d, static_cast<const __nv_bfloat16*>(dout), dp, s));
  else
    SIG_TRY(run_attention_bwd<float>(c, tok, dtok, p, maskf, B, L, d, static_cast<const float*>(dout), dp, s));
  if (tcp) {
    // the token-gradient GEMM already wrote d(patches) at the token strides; only the CLS rows remain
    SIG_PHASE("write_token_grads");
    GradPtrs gp;
    for (int m = 0; m < 3; ++m) {
      gp.dpatch[m] = dtok->dpatch[m]; gp.dcls[m] = dtok->dcls[m];
      gp.psb[m] = dtok->patch_stride_b[m]; gp.psl[m] = dtok->patch_stride_l[m]; gp.csb[m] = dtok->cls_stride_b[m];
    }
    gp.accumulate = dtok->accumulate;
    SIG_LAUNCH((write_cls_grads_kernel<__nv_bfloat16>), dim3(B, 3), 96, 0, s, gp, c.dr1, d);
    SIG_CHECK_LAUNCH();
    // (done_event was recorded right after the token-gradient kernel, sim_tc_tokens_bwd)
    {
      const Fork fk = get_fork(FORK_SIM_BWD);
      if (fk.ok()) fk.join(s);   // weight-gradient GEMMs enqueued on the side stream
    }
    return 0;
  }
  if (dtok->wait_event) cudaStreamWaitEvent(s, (cudaEvent_t)dtok->wait_event, 0);
  SIG_TRY(write_token_grads(dtok, tok->dtype, c.dXf, c.dr1, B, L, d, s));
  if (dtok->done_event) cudaEventRecord((cudaEvent_t)dtok->done_event, s);
  if (dp->early_event) cudaEventRecord((cudaEvent_t)dp->early_event, s);   // (SIMT path: no early part)
  if (dp->late_event) cudaEventRecord((cudaEvent_t)dp->late_event, s);
  return 0;
}

int sim_dx_operands(void* ctx, int B, int L, int d, int dtype, unsigned flags, void** pds, void** dxqt) {
  if (!ctx || !pds || !dxqt) return SIG_ERR_NULL;
  if (!sim_tc_ok(dtype, L, flags)) return SIG_ERR_DTYPE;
  const SimCtx c = sim_ctx(ctx, B, L, d, true);
  *pds = c.PdS;
  *dxqt = c.DXQT;
  return 0;
}

int sim_select(const sig_tokens* tok, const sig_sim_params* p, int which, int k1, int k2, int max_keep, float* masks,
               void* selected, void* ctx, size_t ctx_bytes, cudaStream_t s) {
  SIG_TRY(check_tokens(tok, true));
  SIG_TRY(check_sim_params(p, true, false));
  if (!masks || !ctx) return SIG_ERR_NULL;
  if (which < 1 || which > 3 || k1 < 1 || k2 < 1 || max_keep > tok->L) return SIG_ERR_SHAPE;
  const int B = tok->B, L = tok->L, d = tok->d;
  if (ctx_bytes < sim_ctx_bytes(B, L, d)) return SIG_ERR_WORKSPACE;
  SimCtx c = sim_ctx(ctx, B, L, d);
  SIG_TRY(convert_tokens(tok, c.Xf, c.clsf, s));
  SIG_TRY(run_selection(c, tok, p, B, L, d, which, k1, k2, max_keep, masks, s));
  if (selected) {
    const int64_t rows = (int64_t)3 * B * L;
    const int threads = d / 8 >= 128 ? 128 : 64;
    if (tok->dtype == SIG_BF16)
      SIG_LAUNCH((mask_mul_kernel<__nv_bfloat16>), (unsigned)rows, threads, 0, s, c.Xf, c.maskf, rows, d, static_cast<__nv_bfloat16*>(selected));
    else
      SIG_LAUNCH((mask_mul_kernel<float>), (unsigned)rows, threads, 0, s, c.Xf, c.maskf, rows, d, static_cast<float*>(selected));
    SIG_CHECK_LAUNCH();
  }
  return 0;
}

// dpatch = dselected * mask
template <typename T>
static __global__ void mask_mul_bwd_kernel(const T* __restrict__ dsel, const float* __restrict__ masks, GradPtrs gp, int B, int L,
                                           int d) {
  pdl_enter();
  const int m = blockIdx.y;
  const int64_t row = blockIdx.x;
  const int b = (int)(row / L), l = (int)(row % L);
  const int64_t r = ((int64_t)m * B + b) * L + l;
  const float mk = masks[r];
  T* dst = static_cast<T*>(gp.dpatch[m]) + b * gp.psb[m] + l * gp.psl[m];
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    float v[8];
    load8(dsel + r * d + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= mk;
    if (gp.accumulate) {
      float o[8];
      load8(dst + c, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += o[i];
    }
    store8(dst + c, v);
  }
}

int mask_mul_bwd(const void* dselected, const float* masks, int dtype, int B, int L, int d, const sig_token_grads* g,
                 cudaStream_t s) {
  if (!dselected || !masks) return SIG_ERR_NULL;
  if (dtype != SIG_F32 && dtype != SIG_BF16) return SIG_ERR_DTYPE;
  if (B < 1 || L < 1 || d < 8 || d % 8) return SIG_ERR_SHAPE;
  SIG_TRY(check_token_grads(g, dtype, false));
  GradPtrs gp;
  for (int m = 0; m < 3; ++m) {
    gp.dpatch[m] = g->dpatch[m]; gp.dcls[m] = nullptr;
    gp.psb[m] = g->patch_stride_b[m]; gp.psl[m] = g->patch_stride_l[m]; gp.csb[m] = 0;
  }
  gp.accumulate = g->accumulate;
  dim3 grid((unsigned)((int64_t)B * L), 3);
  const int threads = d / 8 >= 128 ? 128 : 64;
  if (dtype == SIG_BF16)
    SIG_LAUNCH((mask_mul_bwd_kernel<__nv_bfloat16>), grid, threads, 0, s, static_cast<const __nv_bfloat16*>(dselected), masks, gp, B, L, d);
  else
    SIG_LAUNCH((mask_mul_bwd_kernel<float>), grid, threads, 0, s, static_cast<const float*>(dselected), masks, gp, B, L, d);
  SIG_CHECK_LAUNCH();
  return 0;
}

}  // namespace sig

// ID / metric losses that follow the fusion head in the training step:
//   CrossEntropyLabelSmooth.forward   layers/softmax_loss.py:23-34
//   TripletLoss.__call__              layers/triplet_loss.py:106-135 (euclidean_dist :16-31, hard_example_mining :51-104)
// The reference builds the one-hot targets on the HOST every call (softmax_loss.py:30: `.data.cpu()` + scatter_, a
// forced synchronisation per loss term) and mines hard examples with boolean-mask reshapes; here both are one forward
// and one backward kernel, no host synchronisation, deterministic (no atomics), fp32 arithmetic on fp32 or bf16 inputs.
#include <cuda_bf16.h>

#include "common.cuh"
#include "signal_b200.h"
#include "simt_ops.cuh"

namespace sig {
namespace {

template <typename T>
__device__ __forceinline__ float ldf(const T* p, int64_t i) { return to_f32<T>(p[i]); }

// ---- label-smoothed cross entropy ---------------------------------------------------------------------------------
// row loss_b = sum_k -t_bk log p_bk,  t = (1 - eps) onehot(y_b) + eps / C
//            = lse_b - (1 - eps) z_b[y_b] - (eps / C) sum_k z_bk        (sum_k t_bk = 1)
// grid B, 256 threads.  rowloss [B], lse [B] (saved for the backward).
template <typename T>
__global__ void __launch_bounds__(256) xent_ls_fwd_kernel(const T* __restrict__ z, int64_t ld, const int64_t* __restrict__ y, int C,
                                                          float eps, float* __restrict__ rowloss, float* __restrict__ lse) {
  pdl_enter();
  __shared__ float scratch[33];
  const int b = blockIdx.x;
  const T* zb = z + (int64_t)b * ld;
  float mx = -INFINITY;
  for (int k = threadIdx.x; k < C; k += blockDim.x) mx = fmaxf(mx, ldf(zb, k));
  mx = block_max(mx, scratch);
  float se = 0.f, sz = 0.f;
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const float v = ldf(zb, k);
    se += expf(v - mx);
    sz += v;
  }
  se = block_sum(se, scratch);
  sz = block_sum(sz, scratch);
  if (threadIdx.x == 0) {
    const float l = mx + logf(se);
    const int64_t t = y[b];
    const float zy = (t >= 0 && t < C) ? ldf(zb, t) : 0.f;
    lse[b] = l;
    rowloss[b] = l - (1.f - eps) * zy - (eps / C) * sz;
  }
}

// out = scale * sum_i x[i]   (fixed order: deterministic)
__global__ void __launch_bounds__(256) sum_scale_kernel(const float* __restrict__ x, int n, float scale, float* __restrict__ out) {
  pdl_enter();
  __shared__ float scratch[33];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += x[i];
  a = block_sum(a, scratch);
  if (threadIdx.x == 0) *out = a * scale;
}

// dz_bk = g / B * (softmax_bk - t_bk)
template <typename T>
__global__ void __launch_bounds__(256) xent_ls_bwd_kernel(const T* __restrict__ z, int64_t ld, const int64_t* __restrict__ y, int B, int C,
                                                          float eps, const float* __restrict__ lse, const float* __restrict__ g,
                                                          T* __restrict__ dz, int64_t ldd) {
  pdl_enter();
  const int b = blockIdx.x;
  const float k0 = *g / B, l = lse[b], off = eps / C;
  const int64_t t = y[b];
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const float p = expf(ldf(z + (int64_t)b * ld, k) - l);
    const float tg = off + (k == t ? 1.f - eps : 0.f);
    dz[(int64_t)b * ldd + k] = from_f32<T>(k0 * (p - tg));
  }
}

// ---- triplet loss with hard example mining ------------------------------------------------------------------------
// grid B (anchor i), 256 threads.  dist_ij = sqrt(max(|x_i|^2 + |x_j|^2 - 2 x_i.x_j, 1e-12)); hardest positive = max over
// {j: y_j == y_i} (the anchor itself included, as in the reference), hardest negative = min over {j: y_j != y_i}; ties go
// to the lowest index.  Per-anchor outputs: dist_ap, dist_an (after the hard_factor scaling), their indices, the raw
// distances (for the backward) and the loss term.
template <typename T>
__global__ void __launch_bounds__(256) triplet_fwd_kernel(const T* __restrict__ x, int64_t ld, const int64_t* __restrict__ y, int B, int D,
                                                          float margin, int soft, float hard_factor, float* __restrict__ dist_ap,
                                                          float* __restrict__ dist_an, int* __restrict__ p_idx, int* __restrict__ n_idx,
                                                          float* __restrict__ rowloss) {
  pdl_enter();
  extern __shared__ float xi[];   // [D]
  __shared__ float s_val[2][8];
  __shared__ int s_idx[2][8];
  __shared__ float scratch[33];
  const int i = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float xx = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float v = ldf(x + (int64_t)i * ld, c);
    xi[c] = v;
    xx = fmaf(v, v, xx);
  }
  xx = block_sum(xx, scratch);   // (ends with a barrier: xi is complete)
  const int64_t yi = y[i];
  float best_p = -INFINITY, best_n = INFINITY;
  int ip = -1, in_ = -1;
  (void)xx;
  for (int j = w; j < B; j += nw) {
    // |x_i - x_j|^2 accumulated from the differences: the reference's expansion |x_i|^2 + |x_j|^2 - 2 x_i.x_j
    // (triplet_loss.py:26-28) cancels when features are close (LayerNorm'd vars_total: |x|^2 = 1536, d^2 ~ 10), which puts
    // ~1e-5 of noise on the distances -- enough to flip the hard-example choice between near-equidistant samples
    float d2 = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float df = ldf(x + (int64_t)j * ld, c) - xi[c];
      d2 = fmaf(df, df, d2);
    }
    d2 = warp_sum(d2);
    const float dist = sqrtf(fmaxf(d2, 1e-12f));
    if (y[j] == yi) {
      if (dist > best_p) { best_p = dist; ip = j; }     // j ascends within a warp: the first maximum is kept
    } else {
      if (dist < best_n) { best_n = dist; in_ = j; }
    }
  }
  if (lane == 0) {
    s_val[0][w] = best_p; s_idx[0][w] = ip;
    s_val[1][w] = best_n; s_idx[1][w] = in_;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 0; q < nw; ++q) {
      const float vp = s_val[0][q], vn = s_val[1][q];
      const int jp = s_idx[0][q], jn = s_idx[1][q];
      if (jp >= 0 && (ip < 0 || vp > best_p || (vp == best_p && jp < ip))) { best_p = vp; ip = jp; }
      if (jn >= 0 && (in_ < 0 || vn < best_n || (vn == best_n && jn < in_))) { best_n = vn; in_ = jn; }
    }
    const float ap = best_p * (1.f + hard_factor), an = (in_ >= 0 ? best_n : 0.f) * (1.f - hard_factor);
    dist_ap[i] = ap;
    dist_an[i] = an;
    p_idx[i] = ip;
    n_idx[i] = in_;
    const float s = an - ap;
    // SoftMarginLoss(s, 1) = log(1 + exp(-s)); MarginRankingLoss(an, ap, 1) = max(0, -(an - ap) + margin)
    rowloss[i] = soft ? (s > 0.f ? log1pf(expf(-s)) : -s + log1pf(expf(s))) : fmaxf(0.f, margin - s);
  }
}

// grid B (row j of dx), 256 threads.  coef_ap[i] = dL/d(dist_ap_i) etc. are recomputed from the saved distances:
//   dL/ds_i = g/B * (soft ? -1/(1+exp(s_i)) : (margin - s_i > 0 ? -1 : 0)),  s = an - ap,
//   d/d(an) = dL/ds * (1 - hf) (+ g_an_i),  d/d(ap) = -dL/ds * (1 + hf) (+ g_ap_i)   [w.r.t. the RAW distances]
// and d dist_ij / d x_i = (x_i - x_j) / dist_ij, d dist_ij / d x_j = (x_j - x_i) / dist_ij (0 where the clamp is active).
template <typename T>
__global__ void __launch_bounds__(256) triplet_bwd_kernel(const T* __restrict__ x, int64_t ld, int B, int D, float margin, int soft,
                                                          float hard_factor, const float* __restrict__ dist_ap,
                                                          const float* __restrict__ dist_an, const int* __restrict__ p_idx,
                                                          const int* __restrict__ n_idx, const float* __restrict__ g,
                                                          const float* __restrict__ g_ap, const float* __restrict__ g_an,
                                                          T* __restrict__ dx, int64_t ldd) {
  pdl_enter();
  extern __shared__ float sm[];   // cp [B], cn [B]  (coefficient / raw distance, 0 where clamped or missing)
  float* cp = sm;
  float* cn = sm + B;
  __shared__ int sp[1024], sn[1024];   // (B <= 1024, checked by the host)
  const int j = blockIdx.x;
  const float gl = g ? *g / B : 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float ap = dist_ap[i], an = dist_an[i];
    const float s = an - ap;
    const float dls = gl * (soft ? -1.f / (1.f + expf(s)) : (margin - s > 0.f ? -1.f : 0.f));
    float c_an = dls + (g_an ? g_an[i] : 0.f);
    float c_ap = -dls + (g_ap ? g_ap[i] : 0.f);
    c_an *= (1.f - hard_factor);
    c_ap *= (1.f + hard_factor);
    const float rap = ap / (1.f + hard_factor), ran = (1.f - hard_factor) != 0.f ? an / (1.f - hard_factor) : 0.f;   // raw distances
    const int ip = p_idx[i], in_ = n_idx[i];
    sp[i] = ip; sn[i] = in_;
    cp[i] = (ip >= 0 && rap > 1.0000001e-6f) ? c_ap / rap : 0.f;     // dist == sqrt(1e-12): the clamp was active, no gradient
    cn[i] = (in_ >= 0 && ran > 1.0000001e-6f) ? c_an / ran : 0.f;
  }
  __syncthreads();
  // the pairs (i, p_i) / (i, n_i) with i != j that contain j, in ascending i (fixed order: deterministic sums)
  __shared__ int l_other[2 * 1024];
  __shared__ float l_coef[2 * 1024];
  __shared__ int l_n;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int i = 0; i < B; ++i) {
      if (sp[i] == j && cp[i] != 0.f) { l_other[n] = i; l_coef[n] = cp[i]; ++n; }
      if (sn[i] == j && cn[i] != 0.f) { l_other[n] = i; l_coef[n] = cn[i]; ++n; }
    }
    l_n = n;
  }
  __syncthreads();
  const int n = l_n;
  // dx_j = sum over the pairs that contain j of coef * (x_j - x_other)
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float xj = ldf(x + (int64_t)j * ld, c);
    float acc = 0.f;
    if (sp[j] >= 0 && cp[j] != 0.f) acc = fmaf(cp[j], xj - ldf(x + (int64_t)sp[j] * ld, c), acc);
    if (sn[j] >= 0 && cn[j] != 0.f) acc = fmaf(cn[j], xj - ldf(x + (int64_t)sn[j] * ld, c), acc);
    for (int q = 0; q < n; ++q) acc = fmaf(l_coef[q], xj - ldf(x + (int64_t)l_other[q] * ld, c), acc);
    dx[(int64_t)j * ldd + c] = from_f32<T>(acc);
  }
}

// ---- BNNeck + classifier (make_model.py:128-131,194-195,212-214: nn.BatchNorm1d -> nn.Linear(bias=False)) ------------
// One CTA = 32 feature columns x 8 row groups.  Training: batch statistics (two passes: mean, then centred variance),
// running statistics updated like nn.BatchNorm1d (momentum, unbiased variance); eval: running statistics.
// y32 [B, D] fp32 (operand of the classifier GEMM, saved for the backward), out [B, D] in the input dtype.
template <typename T>
__global__ void __launch_bounds__(256) bn1d_fwd_kernel(const T* __restrict__ x, int64_t ld, int B, int D, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float* __restrict__ run_mean,
                                                       float* __restrict__ run_var, float momentum, float eps, int training,
                                                       float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                                       float* __restrict__ y32, T* __restrict__ out, int64_t ldo) {
  pdl_enter();
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool ok = c < D;
  float mean = 0.f, rstd = 0.f;
  if (training) {
    float s = 0.f;
    if (ok) for (int r = ry; r < B; r += 8) s += ldf(x + (int64_t)r * ld, c);
    red[ry][cx] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += red[q][cx];
    mean = s / B;
    __syncthreads();
    float v = 0.f;
    if (ok) for (int r = ry; r < B; r += 8) {
      const float t = ldf(x + (int64_t)r * ld, c) - mean;
      v = fmaf(t, t, v);
    }
    red[ry][cx] = v;
    __syncthreads();
    v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += red[q][cx];
    const float var = v / B;
    rstd = rsqrtf(var + eps);
    if (ok && ry == 0) {
      if (run_mean) run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
      if (run_var) run_var[c] = (1.f - momentum) * run_var[c] + momentum * (B > 1 ? var * B / (B - 1) : var);
    }
  } else if (ok) {
    mean = run_mean[c];
    rstd = rsqrtf(run_var[c] + eps);
  }
  if (!ok) return;
  if (ry == 0) { save_mean[c] = mean; save_rstd[c] = rstd; }
  const float g = gamma[c], b = beta[c];
  for (int r = ry; r < B; r += 8) {
    const float y = (ldf(x + (int64_t)r * ld, c) - mean) * rstd * g + b;
    y32[(int64_t)r * D + c] = y;
    out[(int64_t)r * ldo + c] = from_f32<T>(y);
  }
}

// dx = gamma * rstd * (dy - mean_b(dy) - xhat * mean_b(dy * xhat))  (training; eval: gamma * rstd * dy),
// dgamma = sum_b dy * xhat, dbeta = sum_b dy;  dy = dY32 (from the classifier) + optional dout (cotangent of the BN output)
template <typename T>
__global__ void __launch_bounds__(256) bn1d_bwd_kernel(const T* __restrict__ x, int64_t ld, int B, int D, const float* __restrict__ gamma,
                                                       const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                                                       int training, const float* __restrict__ dY32, const T* __restrict__ dout,
                                                       int64_t lddo, T* __restrict__ dx, int64_t ldx, float* __restrict__ dgamma,
                                                       float* __restrict__ dbeta) {
  pdl_enter();
  __shared__ float red[2][8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool ok = c < D;
  const float mean = ok ? save_mean[c] : 0.f, rstd = ok ? save_rstd[c] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  if (ok) for (int r = ry; r < B; r += 8) {
    float dy = dY32 ? dY32[(int64_t)r * D + c] : 0.f;
    if (dout) dy += ldf(dout + (int64_t)r * lddo, c);
    const float xh = (ldf(x + (int64_t)r * ld, c) - mean) * rstd;
    s1 += dy;
    s2 = fmaf(dy, xh, s2);
  }
  red[0][ry][cx] = s1;
  red[1][ry][cx] = s2;
  __syncthreads();
  s1 = s2 = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) { s1 += red[0][q][cx]; s2 += red[1][q][cx]; }
  if (!ok) return;
  if (ry == 0) { dgamma[c] = s2; dbeta[c] = s1; }
  const float g = gamma[c] * rstd, m1 = training ? s1 / B : 0.f, m2 = training ? s2 / B : 0.f;
  for (int r = ry; r < B; r += 8) {
    float dy = dY32 ? dY32[(int64_t)r * D + c] : 0.f;
    if (dout) dy += ldf(dout + (int64_t)r * lddo, c);
    const float xh = (ldf(x + (int64_t)r * ld, c) - mean) * rstd;
    dx[(int64_t)r * ldx + c] = from_f32<T>(g * (dy - m1 - xh * m2));
  }
}

// rows of a [R, N] matrix with leading dimension ld: T -> fp32 (dense) or fp32 (dense) -> T
template <typename T>
__global__ void __launch_bounds__(256) rows_to_f32_kernel(const T* __restrict__ src, int64_t ld, int R, int N, float* __restrict__ dst) {
  pdl_enter();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < (int64_t)R * N) dst[i] = ldf(src + (i / N) * ld, i % N);
}
template <typename T>
__global__ void __launch_bounds__(256) rows_from_f32_kernel(const float* __restrict__ src, int R, int N, T* __restrict__ dst, int64_t ld) {
  pdl_enter();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < (int64_t)R * N) dst[(i / N) * ld + i % N] = from_f32<T>(src[i]);
}

template <typename T>
int bnneck_fwd(const T* x, int64_t ld, int B, int D, int C, const float* gamma, const float* beta, float* rm, float* rv, float momentum,
               float eps, int training, const float* W, T* out, int64_t ldo, T* logits, int64_t ldl, float* save_mean, float* save_rstd,
               float* y32, float* ws, cudaStream_t s) {
  SIG_LAUNCH((bn1d_fwd_kernel<T>), (unsigned)ceil_div(D, 32), 256, 0, s, x, ld, B, D, gamma, beta, rm, rv, momentum, eps, training, save_mean,
             save_rstd, y32, out, ldo);
  SIG_CHECK_LAUNCH();
  // logits = y W^T  (fp32 SIMT GEMM: 2 B C D = 0.1 GFLOP at the training shape, latency-bound)
  if (sizeof(T) == 4 && ldl == C) {
    SIG_TRY(launch_gemm(gemm_nt(y32, D, W, D, reinterpret_cast<float*>(logits), C, nullptr, B, C, D), s));
  } else {
    SIG_TRY(launch_gemm(gemm_nt(y32, D, W, D, ws, C, nullptr, B, C, D), s));
    SIG_LAUNCH((rows_from_f32_kernel<T>), (unsigned)ceil_div((int64_t)B * C, 256), 256, 0, s, ws, B, C, logits, ldl);
    SIG_CHECK_LAUNCH();
  }
  return 0;
}

template <typename T>
int bnneck_bwd(const T* x, int64_t ld, int B, int D, int C, const float* gamma, const float* W, const float* save_mean,
               const float* save_rstd, int training, const float* y32, const T* dlogits, int64_t ldl, const T* dout, int64_t lddo, T* dx,
               int64_t ldx, float* dgamma, float* dbeta, float* dW, float* ws, cudaStream_t s) {
  float* dl32 = ws;                        // [B, C]
  float* dY32 = ws + (size_t)B * C;        // [B, D]
  const float* dy = nullptr;
  if (dlogits) {
    SIG_LAUNCH((rows_to_f32_kernel<T>), (unsigned)ceil_div((int64_t)B * C, 256), 256, 0, s, dlogits, ldl, B, C, dl32);
    SIG_CHECK_LAUNCH();
    SIG_TRY(launch_gemm(gemm_tn(dl32, C, y32, D, dW, D, C, D, B), s));      // dW = dlogits^T y
    SIG_TRY(launch_gemm(gemm_nn(dl32, C, W, D, dY32, D, B, D, C), s));      // dy = dlogits W
    dy = dY32;
  } else {
    cudaMemsetAsync(dW, 0, (size_t)C * D * sizeof(float), s);
  }
  SIG_LAUNCH((bn1d_bwd_kernel<T>), (unsigned)ceil_div(D, 32), 256, 0, s, x, ld, B, D, gamma, save_mean, save_rstd, training, dy, dout, lddo, dx,
             ldx, dgamma, dbeta);
  SIG_CHECK_LAUNCH();
  return 0;
}

}  // namespace
}  // namespace sig

extern "C" {

size_t sig_loss_ws_bytes(int B) { return (size_t)(B > 0 ? B : 0) * 2 * sizeof(float); }

int sig_xent_ls_fwd(const void* logits, int dtype, int64_t ld, const int64_t* targets, int B, int C, float eps, float* loss,
                    float* lse, void* ws, size_t ws_bytes, int device, void* stream) {
  SIG_ENTER(device);
  using namespace sig;
  if (!logits || !targets || !loss || !lse || !ws) return SIG_ERR_NULL;
  if (B < 1 || C < 1 || ld < C) return SIG_ERR_SHAPE;
  if (dtype != SIG_F32 && dtype != SIG_BF16) return SIG_ERR_DTYPE;
  if (ws_bytes < sig_loss_ws_bytes(B)) return SIG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  float* rowloss = static_cast<float*>(ws);
  if (dtype == SIG_BF16)
    SIG_LAUNCH((xent_ls_fwd_kernel<__nv_bfloat16>), B, 256, 0, s, static_cast<const __nv_bfloat16*>(logits), ld, targets, C, eps, rowloss, lse);
  else
    SIG_LAUNCH((xent_ls_fwd_kernel<float>), B, 256, 0, s, static_cast<const float*>(logits), ld, targets, C, eps, rowloss, lse);
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((sum_scale_kernel), 1, 256, 0, s, rowloss, B, 1.f / B, loss);
  SIG_CHECK_LAUNCH();
  return 0;
}

int sig_xent_ls_bwd(const void* logits, int dtype, int64_t ld, const int64_t* targets, int B, int C, float eps, const float* lse,
                    const float* dloss, void* dlogits, int64_t ldd, int device, void* stream) {
  SIG_ENTER(device);
  using namespace sig;
  if (!logits || !targets || !lse || !dloss || !dlogits) return SIG_ERR_NULL;
  if (B < 1 || C < 1 || ld < C || ldd < C) return SIG_ERR_SHAPE;
  if (dtype != SIG_F32 && dtype != SIG_BF16) return SIG_ERR_DTYPE;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == SIG_BF16)
    SIG_LAUNCH((xent_ls_bwd_kernel<__nv_bfloat16>), B, 256, 0, s, static_cast<const __nv_bfloat16*>(logits), ld, targets, B, C, eps, lse, dloss,
               static_cast<__nv_bfloat16*>(dlogits), ldd);
  else
    SIG_LAUNCH((xent_ls_bwd_kernel<float>), B, 256, 0, s, static_cast<const float*>(logits), ld, targets, B, C, eps, lse, dloss,
               static_cast<float*>(dlogits), ldd);
  SIG_CHECK_LAUNCH();
  return 0;
}

int sig_triplet_fwd(const void* feat, int dtype, int64_t ld, const int64_t* labels, int B, int D, float margin, int soft_margin,
                    float hard_factor, float* loss, float* dist_ap, float* dist_an, int* p_idx, int* n_idx, void* ws, size_t ws_bytes,
                    int device, void* stream) {
  SIG_ENTER(device);
  using namespace sig;
  if (!feat || !labels || !loss || !dist_ap || !dist_an || !p_idx || !n_idx || !ws) return SIG_ERR_NULL;
  if (B < 1 || B > 1024 || D < 1 || D > 8192 || ld < D) return SIG_ERR_SHAPE;
  if (dtype != SIG_F32 && dtype != SIG_BF16) return SIG_ERR_DTYPE;
  if (ws_bytes < sig_loss_ws_bytes(B)) return SIG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  float* rowloss = static_cast<float*>(ws);
  const size_t smem = (size_t)D * sizeof(float);
  if (dtype == SIG_BF16)
    SIG_LAUNCH((triplet_fwd_kernel<__nv_bfloat16>), B, 256, smem, s, static_cast<const __nv_bfloat16*>(feat), ld, labels, B, D, margin, soft_margin,
               hard_factor, dist_ap, dist_an, p_idx, n_idx, rowloss);
  else
    SIG_LAUNCH((triplet_fwd_kernel<float>), B, 256, smem, s, static_cast<const float*>(feat), ld, labels, B, D, margin, soft_margin, hard_factor,
               dist_ap, dist_an, p_idx, n_idx, rowloss);
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((sum_scale_kernel), 1, 256, 0, s, rowloss, B, 1.f / B, loss);
  SIG_CHECK_LAUNCH();
  return 0;
}

int sig_triplet_bwd(const void* feat, int dtype, int64_t ld, int B, int D, float margin, int soft_margin, float hard_factor,
                    const float* dist_ap, const float* dist_an, const int* p_idx, const int* n_idx, const float* dloss,
                    const float* d_dist_ap, const float* d_dist_an, void* dfeat, int64_t ldd, int device, void* stream) {
  SIG_ENTER(device);
  using namespace sig;
  if (!feat || !dist_ap || !dist_an || !p_idx || !n_idx || !dfeat) return SIG_ERR_NULL;
  if (B < 1 || B > 1024 || D < 1 || ld < D || ldd < D) return SIG_ERR_SHAPE;
  if (dtype != SIG_F32 && dtype != SIG_BF16) return SIG_ERR_DTYPE;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)2 * B * sizeof(float);
  if (dtype == SIG_BF16)
    SIG_LAUNCH((triplet_bwd_kernel<__nv_bfloat16>), B, 256, smem, s, static_cast<const __nv_bfloat16*>(feat), ld, B, D, margin, soft_margin,
               hard_factor, dist_ap, dist_an, p_idx, n_idx, dloss, d_dist_ap, d_dist_an, static_cast<__nv_bfloat16*>(dfeat), ldd);
  else
    SIG_LAUNCH((triplet_bwd_kernel<float>), B, 256, smem, s, static_cast<const float*>(feat), ld, B, D, margin, soft_margin, hard_factor, dist_ap,
               dist_an, p_idx, n_idx, dloss, d_dist_ap, d_dist_an, static_cast<float*>(dfeat), ldd);
  SIG_CHECK_LAUNCH();
  return 0;
}

size_t sig_bnneck_ws_bytes(int B, int D, int C) {
  if (B < 1 || D < 1 || C < 1) return 0;
  return ((size_t)B * C + (size_t)B * D) * sizeof(float);
}

int sig_bnneck_cls_fwd(const void* feat, int dtype, int64_t ld, int B, int D, int C, const float* bn_weight, const float* bn_bias,
                       float* running_mean, float* running_var, float momentum, float eps, int training, const float* cls_weight,
                       void* bn_out, int64_t ldo, void* logits, int64_t ldl, float* save_mean, float* save_rstd, float* y32, void* ws,
                       size_t ws_bytes, int device, void* stream) {
  SIG_ENTER(device);
  using namespace sig;
  if (!feat || !bn_weight || !bn_bias || !cls_weight || !bn_out || !logits || !save_mean || !save_rstd || !y32 || !ws) return SIG_ERR_NULL;
  if (!training && (!running_mean || !running_var)) return SIG_ERR_NULL;
  if (B < 1 || D < 1 || C < 1 || ld < D || ldo < D || ldl < C) return SIG_ERR_SHAPE;
  if (dtype != SIG_F32 && dtype != SIG_BF16) return SIG_ERR_DTYPE;
  if (ws_bytes < sig_bnneck_ws_bytes(B, D, C)) return SIG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == SIG_BF16)
    return bnneck_fwd<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(feat), ld, B, D, C, bn_weight, bn_bias, running_mean, running_var, momentum,
                                     eps, training, cls_weight, static_cast<__nv_bfloat16*>(bn_out), ldo, static_cast<__nv_bfloat16*>(logits), ldl,
                                     save_mean, save_rstd, y32, static_cast<float*>(ws), s);
  return bnneck_fwd<float>(static_cast<const float*>(feat), ld, B, D, C, bn_weight, bn_bias, running_mean, running_var, momentum, eps, training,
                           cls_weight, static_cast<float*>(bn_out), ldo, static_cast<float*>(logits), ldl, save_mean, save_rstd, y32,
                           static_cast<float*>(ws), s);
}

int sig_bnneck_cls_bwd(const void* feat, int dtype, int64_t ld, int B, int D, int C, const float* bn_weight, const float* cls_weight,
                       const float* save_mean, const float* save_rstd, int training, const float* y32, const void* dlogits, int64_t ldl,
                       const void* d_bn_out, int64_t lddo, void* dfeat, int64_t ldx, float* d_bn_weight, float* d_bn_bias,
                       float* d_cls_weight, void* ws, size_t ws_bytes, int device, void* stream) {
  SIG_ENTER(device);
  using namespace sig;
  if (!feat || !bn_weight || !cls_weight || !save_mean || !save_rstd || !y32 || !dfeat || !d_bn_weight || !d_bn_bias || !d_cls_weight || !ws)
    return SIG_ERR_NULL;
  if (B < 1 || D < 1 || C < 1 || ld < D || ldx < D || (dlogits && ldl < C) || (d_bn_out && lddo < D)) return SIG_ERR_SHAPE;
  if (dtype != SIG_F32 && dtype != SIG_BF16) return SIG_ERR_DTYPE;
  if (ws_bytes < sig_bnneck_ws_bytes(B, D, C)) return SIG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == SIG_BF16)
    return bnneck_bwd<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(feat), ld, B, D, C, bn_weight, cls_weight, save_mean, save_rstd, training,
                                     y32, static_cast<const __nv_bfloat16*>(dlogits), ldl, static_cast<const __nv_bfloat16*>(d_bn_out), lddo,
                                     static_cast<__nv_bfloat16*>(dfeat), ldx, d_bn_weight, d_bn_bias, d_cls_weight, static_cast<float*>(ws), s);
  return bnneck_bwd<float>(static_cast<const float*>(feat), ld, B, D, C, bn_weight, cls_weight, save_mean, save_rstd, training, y32,
                           static_cast<const float*>(dlogits), ldl, static_cast<const float*>(d_bn_out), lddo, static_cast<float*>(dfeat), ldx,
                           d_bn_weight, d_bn_bias, d_cls_weight, static_cast<float*>(ws), s);
}

}  // extern "C"

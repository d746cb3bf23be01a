// Token producer: the tail of the CLIP vision tower that hands the fusion head its token maps,
//   x = self.ln_post(x); xproj = x @ self.proj            (modeling/clip/model.py:485-487, LayerNorm :154-160)
//   global_feat = x[:, 0]; x_cash = x[:, 1:]              (modeling/meta_arch.py:108-110)
// forward and backward behind the C ABI (include/signal_b200.h: sig_tokens_fwd / sig_tokens_bwd).
//
// Data flow (half inputs = the reference under autocast: LayerNorm computes in fp32 on x.float() and returns the input
// dtype, the matmul runs on half operands with fp32 accumulation and returns half):
//   ln_rows_fwd_kernel   one warp per token row, row read once from its (b, l) strides (the tower's [L,B,W] layout is
//                        consumed in place), two-pass statistics in registers, xn = bf16(LN(x)) stored row-contiguous in
//                        [B, 1+L, W] order + mu / rstd                                                     HBM: 2 R W s
//   tcgen05 GEMM         tokens[R, D] = xn[R, W] . proj[W, D]  (tc_gemm.cu; xn K-major, proj MN-major, bf16 out straight
//                        into the [B, 1+L, D] map the head's TMA descriptors read)                         tensor: 2 R W D
//   patch_mean_kernel    (by-product, optional) mean over the L patch rows of the ROUNDED tokens, fp32 [B, D]: the GAM
//                        mean pool (useB.py:84-86); the 17 MB it reads were written by the GEMM just before (L2)
//   backward             dxn = dtok . proj^T (bf16, like autocast's matmul backward), dproj = xn^T dtok (split-K, fp32),
//                        ln_rows_bwd_kernel: dx at x's strides + per-CTA partial d(gamma), d(beta), fixed-order reduction.
// fp32 inputs take the exact path: same kernels with fp32 xn and the fp32 SIMT GEMM (simt_ops.cuh).
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "simt_ops.cuh"
#include "tc_gemm.h"

namespace sig {

// (same scope as common.cuh's load8 / store8 overloads, so that the templated kernels below see all three dtypes)
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
  uint4 raw;
  __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(fminf(fmaxf(v[2 * i], -65504.f), 65504.f), fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f));
  *reinterpret_cast<uint4*>(p) = raw;
}

namespace {


// row r = b * L1 + l of the [B, L1, W] map: x + b * sb + l * sl
template <int NCH, typename InT, typename XnT>
__global__ void __launch_bounds__(256) ln_rows_fwd_kernel(const InT* __restrict__ x, int64_t sb, int64_t sl, int R, int L1, int W,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                         XnT* __restrict__ xn, float* __restrict__ mu, float* __restrict__ rstd) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const int b = row / L1, l = row - b * L1;
  const InT* xr = x + b * sb + l * sl;
  float v[NCH][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int e = (c * 32 + lane) * 8;
    if (e < W) {
      load8(xr + e, v[c]);
#pragma unroll
      for (int t = 0; t < 8; ++t) s += v[c][t];
    }
  }
  const float mean = warp_sum(s) / (float)W;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int e = (c * 32 + lane) * 8;
    if (e < W) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float dlt = v[c][t] - mean;
        q = fmaf(dlt, dlt, q);
      }
    }
  }
  const float rs = rsqrtf(warp_sum(q) / (float)W + eps);
  if (lane == 0) {
    mu[row] = mean;
    rstd[row] = rs;
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int e = (c * 32 + lane) * 8;
    if (e < W) {
      float g[8], bt[8], o[8];
      load8(gamma + e, g);
      load8(beta + e, bt);
#pragma unroll
      for (int t = 0; t < 8; ++t) o[t] = fmaf((v[c][t] - mean) * rs, g[t], bt[t]);
      store8(xn + (int64_t)row * W + e, o);
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dxn * gamma,  xhat = (x - mu) * rstd
// part [gridDim.x][2][W]: this CTA's sums of dxn * xhat (d gamma) and dxn (d beta) over its rows.
// Register diet (the first version held xhat, g, gamma and both accumulators as fp32 arrays: 148 registers at W = 768, one
// 8-warp CTA per SM, 2.0 TB/s): the row is kept as the RAW 16-byte loads and unpacked twice (statistics pass, output pass),
// gamma is re-read from L1 per chunk, only the two accumulators stay resident -- two to three CTAs per SM.
template <int NCH, typename InT, typename GT>
__global__ void __launch_bounds__(256, 2) ln_rows_bwd_kernel(const InT* __restrict__ x, int64_t sb, int64_t sl, int R, int L1, int W,
                                                            const float* __restrict__ gamma, const float* __restrict__ mu,
                                                            const float* __restrict__ rstd, const GT* __restrict__ dxn,
                                                            InT* __restrict__ dx, int64_t dsb, int64_t dsl, float* __restrict__ part) {
  pdl_enter();
  extern __shared__ __align__(16) float red[];   // [8 warps][2][W]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float dg[NCH][8], db[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int t = 0; t < 8; ++t) dg[c][t] = db[c][t] = 0.f;
  for (int row = blockIdx.x * 8 + w; row < R; row += gridDim.x * 8) {
    const int b = row / L1, l = row - b * L1;
    const InT* xr = x + b * sb + l * sl;
    const GT* dr_in = dxn + (int64_t)row * W;
    const float mean = mu[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int e = (c * 32 + lane) * 8;
      if (e < W) {
        float xv[8], dv[8], gm[8];
        load8(xr + e, xv);
        load8(dr_in + e, dv);
        load8(gamma + e, gm);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float xh = (xv[t] - mean) * rs, g = dv[t] * gm[t];
          s1 += g;
          s2 = fmaf(g, xh, s2);
          dg[c][t] = fmaf(dv[t], xh, dg[c][t]);
          db[c][t] += dv[t];
        }
      }
    }
    const float c1 = warp_sum(s1) / (float)W, c2 = warp_sum(s2) / (float)W;
    InT* dr = dx + b * dsb + l * dsl;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int e = (c * 32 + lane) * 8;
      if (e < W) {
        float xv[8], dv[8], gm[8], o[8];
        load8(xr + e, xv);        // (second read of the row: L1 hits, 3 KB per warp)
        load8(dr_in + e, dv);
        load8(gamma + e, gm);
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = rs * (dv[t] * gm[t] - c1 - (xv[t] - mean) * rs * c2);
        store8(dr + e, o);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int e = (c * 32 + lane) * 8;
    if (e < W) {
      store8(red + (size_t)w * 2 * W + e, dg[c]);
      store8(red + (size_t)w * 2 * W + W + e, db[c]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * W; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) t += red[(size_t)ww * 2 * W + i];   // fixed order
    part[(size_t)blockIdx.x * 2 * W + i] = t;
  }
}

// mean[b][c] = (1/L) sum_{l=1..L} tok[b][l][c]   (the patch rows of a [B, 1+L, D] map); grid B, 256 threads
template <typename T>
__global__ void __launch_bounds__(256) patch_mean_kernel(const T* __restrict__ tok, int64_t sb, int64_t sl, int L, int D, float* __restrict__ mean) {
  pdl_enter();
  __shared__ __align__(16) float red[8][1024];
  const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const T* base = tok + b * sb + sl;   // row 1
  for (int c0 = 0; c0 < D; c0 += 256) {
    const int e = c0 + lane * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (e < D) {
      for (int l = w; l < L; l += 8) {
        float v[8];
        load8(base + (int64_t)l * sl + e, v);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] += v[t];
      }
      store8(&red[w][e], acc);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) t += red[ww][i];
    mean[(int64_t)b * D + i] = t / (float)L;
  }
}

// contiguous rows [R, N] fp32 -> T at (b, l) strides
template <typename T>
__global__ void __launch_bounds__(256) rows_store_kernel(const float* __restrict__ src, int R, int L1, int N, T* __restrict__ dst, int64_t sb, int64_t sl) {
  pdl_enter();
  const int vec = N / 8;
  const int64_t total = (int64_t)R * vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    const int row = (int)(i / vec);
    const int b = row / L1, l = row - b * L1;
    float t[8];
    load8(src + (int64_t)row * N + 8 * v, t);
    store8(dst + b * sb + l * sl + 8 * v, t);
  }
}

// (b, l)-strided rows of T -> contiguous bf16 [R, N]
template <typename T>
__global__ void __launch_bounds__(256) rows_gather_bf16_kernel(const T* __restrict__ src, int64_t sb, int64_t sl, int R, int L1, int N,
                                                              __nv_bfloat16* __restrict__ dst) {
  pdl_enter();
  const int vec = N / 8;
  const int64_t total = (int64_t)R * vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    const int row = (int)(i / vec);
    const int b = row / L1, l = row - b * L1;
    float t[8];
    load8(src + b * sb + l * sl + 8 * v, t);
    store8(dst + (int64_t)row * N + 8 * v, t);
  }
}

#define SIG_TOK_NCH_SWITCH(W, STMT)                      \
  do {                                                    \
    switch ((int)ceil_div((W), 256)) {                    \
      case 1: { constexpr int NCH = 1; STMT; } break;     \
      case 2: { constexpr int NCH = 2; STMT; } break;     \
      case 3: { constexpr int NCH = 3; STMT; } break;     \
      default: { constexpr int NCH = 4; STMT; } break;    \
    }                                                     \
  } while (0)

template <typename InT, typename XnT>
int launch_ln_fwd(const InT* x, int64_t sb, int64_t sl, int R, int L1, int W, const float* g, const float* b, float eps, XnT* xn, float* mu,
                  float* rstd, cudaStream_t s) {
  const unsigned blocks = (unsigned)ceil_div(R, 8);
  SIG_TOK_NCH_SWITCH(W, SIG_LAUNCH((ln_rows_fwd_kernel<NCH, InT, XnT>), blocks, 256, 0, s, x, sb, sl, R, L1, W, g, b, eps, xn, mu, rstd));
  SIG_CHECK_LAUNCH();
  return 0;
}
template <typename InT, typename GT>
int launch_ln_bwd(int ctas, const InT* x, int64_t sb, int64_t sl, int R, int L1, int W, const float* g, const float* mu, const float* rstd,
                  const GT* dxn, InT* dx, int64_t dsb, int64_t dsl, float* part, cudaStream_t s) {
  const size_t red_smem = (size_t)8 * 2 * W * sizeof(float);
  SIG_TOK_NCH_SWITCH(W, {
    ensure_dyn_smem(ln_rows_bwd_kernel<NCH, InT, GT>, 8 * 2 * NCH * 256 * (int)sizeof(float));   // set once per device: the largest W of this NCH
    SIG_LAUNCH((ln_rows_bwd_kernel<NCH, InT, GT>), ctas, 256, red_smem, s, x, sb, sl, R, L1, W, g, mu, rstd, dxn, dx, dsb, dsl, part);
  });
  SIG_CHECK_LAUNCH();
  return 0;
}

inline size_t al256(size_t n) { return (n + 255) & ~(size_t)255; }

// caller-owned buffers.  saved (fwd -> bwd): [xn R*W (bf16 | fp32)] [mu R] [rstd R] [proj bf16 W*D (half inputs)]
// scratch fwd: fp32 tokens staging [R*D] (fp16 output only).  scratch bwd: [dxn R*W (bf16|fp32)] [dtok bf16 R*D (half)]
// [partials ctas*2*W]
struct TokLayout {
  size_t xn, mu, rstd, projb, saved_bytes;
  size_t stage, fwd_scratch_bytes;
  size_t dxn, dtokb, part, bwd_scratch_bytes;
  int ctas;
};
TokLayout tok_layout(int R, int W, int D, int dtype) {
  TokLayout t{};
  const size_t es = dtype == SIG_F32 ? 4 : 2;
  size_t o = 0;
  t.xn = o; o += al256((size_t)R * W * es);
  t.mu = o; o += al256((size_t)R * 4);
  t.rstd = o; o += al256((size_t)R * 4);
  t.projb = o; o += dtype == SIG_F32 ? 0 : al256((size_t)W * D * 2);
  t.saved_bytes = o;
  t.stage = 0;
  t.fwd_scratch_bytes = dtype == SIG_F16 ? al256((size_t)R * D * 4) : 0;
  t.ctas = 296;   // 2 CTAs of 8 warps per SM on a 148-SM part (107 registers at W = 768); any value is correct
  o = 0;
  t.dxn = o; o += al256((size_t)R * W * es);
  t.dtokb = o; o += dtype == SIG_F32 ? 0 : al256((size_t)R * D * 2);
  t.part = o; o += al256((size_t)t.ctas * 2 * W * 4);
  t.bwd_scratch_bytes = o;
  return t;
}

// XT: dtype of x / dx; TT: dtype of the tokens / their gradient.  TT == float selects the exact fp32 path (XT == float);
// half tokens select the tensor-core path, with x either in the same half type or fp32 (an autocast caller whose
// residual stream stayed fp32).  `tdt` = enum of TT (it fixes the buffer layout).
template <typename XT, typename TT>
int tokens_fwd_t(const XT* x, int tdt, int64_t sb, int64_t sl, int B, int L1, int W, int D, const float* ln_w, const float* ln_b,
                 float eps, const float* proj, TT* tokens, float* patch_mean, unsigned char* saved, unsigned char* scratch,
                 cudaStream_t s) {
  const int R = B * L1;
  const TokLayout t = tok_layout(R, W, D, tdt);
  float* mu = reinterpret_cast<float*>(saved + t.mu);
  float* rstd = reinterpret_cast<float*>(saved + t.rstd);
  if constexpr (std::is_same<TT, float>::value) {
    SIG_PHASE("tokens_ln_fwd");
    float* xn = reinterpret_cast<float*>(saved + t.xn);
    SIG_TRY(launch_ln_fwd(x, sb, sl, R, L1, W, ln_w, ln_b, eps, xn, mu, rstd, s));
    SIG_TRY(launch_gemm(gemm_nn(xn, W, proj, D, tokens, D, R, D, W), s));
  } else {
    __nv_bfloat16* xn = reinterpret_cast<__nv_bfloat16*>(saved + t.xn);
    __nv_bfloat16* projb = reinterpret_cast<__nv_bfloat16*>(saved + t.projb);
    {
      SIG_PHASE("tokens_ln_fwd");
      SIG_TRY(cast_f32_to_bf16(proj, projb, (int64_t)W * D, s));
      SIG_TRY(launch_ln_fwd(x, sb, sl, R, L1, W, ln_w, ln_b, eps, xn, mu, rstd, s));
    }
    SIG_PHASE("tokens_proj_fwd");
    TcGemmDesc g = tc_desc();
    g.A = tc_k2d(xn, R, W, W);
    g.B = tc_mn2d(projb, W, D, D);
    g.M = R; g.N = D; g.K = W;
    g.ldc = D;
    g.bn = (D % 256 == 0) ? 256 : 128;
    // (256 x 256 units, mt = 2, measured slower here: 64.6 vs 51.3 us at R = 49 536 -- 388 units are 2.6 waves of 148 CTAs)
    if constexpr (std::is_same<TT, __nv_bfloat16>::value) {
      g.C[0] = tokens; g.out_bf16 = 1;
      SIG_TRY(tc_gemm(g, s));
    } else {   // fp16 tokens: fp32 staging, one rounding to fp16
      float* stage = reinterpret_cast<float*>(scratch + t.stage);
      g.C[0] = stage;
      SIG_TRY(tc_gemm(g, s));
      SIG_LAUNCH((rows_store_kernel<TT>), (unsigned)(device_num_sms() * 4), 256, 0, s, stage, R, L1, D, tokens, (int64_t)L1 * D, (int64_t)D);
      SIG_CHECK_LAUNCH();
    }
  }
  if (patch_mean) {
    SIG_PHASE("tokens_patch_mean");
    SIG_LAUNCH((patch_mean_kernel<TT>), B, 256, 0, s, tokens, (int64_t)L1 * D, (int64_t)D, L1 - 1, D, patch_mean);
    SIG_CHECK_LAUNCH();
  }
  return 0;
}

template <typename XT, typename TT>
int tokens_bwd_t(const XT* x, int tdt, int64_t sb, int64_t sl, int B, int L1, int W, int D, const float* ln_w, const float* proj,
                 const TT* dtok, int64_t dtsb, int64_t dtsl, const unsigned char* saved, XT* dx, int64_t dxsb, int64_t dxsl,
                 float* d_ln_w, float* d_ln_b, float* d_proj, unsigned char* scratch, cudaStream_t s) {
  const int R = B * L1;
  const TokLayout t = tok_layout(R, W, D, tdt);
  const float* mu = reinterpret_cast<const float*>(saved + t.mu);
  const float* rstd = reinterpret_cast<const float*>(saved + t.rstd);
  float* part = reinterpret_cast<float*>(scratch + t.part);
  if constexpr (std::is_same<TT, float>::value) {
    const float* xn = reinterpret_cast<const float*>(saved + t.xn);
    float* dxn = reinterpret_cast<float*>(scratch + t.dxn);
    if (dtsb != (int64_t)L1 * dtsl) return SIG_ERR_SHAPE;
    {
      SIG_PHASE("tokens_proj_bwd");
      // dxn[R, W] = dtok[R, D] . proj[W, D]^T
      SIG_TRY(launch_gemm(gemm_nt(dtok, dtsl, proj, D, dxn, W, nullptr, R, W, D), s));
      // dproj[W, D] = xn[R, W]^T dtok[R, D]   (split-K into a zeroed buffer)
      cudaMemsetAsync(d_proj, 0, (size_t)W * D * sizeof(float), s);
      Gemm g = gemm_tn(xn, W, dtok, dtsl, d_proj, D, W, D, R);
      g.ksplit = R >= 2048 ? 16 : 1;
      SIG_TRY(launch_gemm(g, s));
    }
    SIG_PHASE("tokens_ln_bwd");
    SIG_TRY(launch_ln_bwd(t.ctas, x, sb, sl, R, L1, W, ln_w, mu, rstd, (const float*)dxn, dx, dxsb, dxsl, part, s));
  } else {
    const __nv_bfloat16* xn = reinterpret_cast<const __nv_bfloat16*>(saved + t.xn);
    const __nv_bfloat16* projb = reinterpret_cast<const __nv_bfloat16*>(saved + t.projb);
    __nv_bfloat16* dxn = reinterpret_cast<__nv_bfloat16*>(scratch + t.dxn);
    const __nv_bfloat16* dtb;
    int64_t ldd;
    if (std::is_same<TT, __nv_bfloat16>::value && dtsb == (int64_t)L1 * dtsl && (dtsl % 8) == 0 && ((uintptr_t)dtok % 16) == 0) {
      dtb = reinterpret_cast<const __nv_bfloat16*>(dtok);   // uniform row pitch: consumed in place
      ldd = dtsl;
    } else {
      __nv_bfloat16* tmp = reinterpret_cast<__nv_bfloat16*>(scratch + t.dtokb);
      SIG_LAUNCH((rows_gather_bf16_kernel<TT>), (unsigned)(device_num_sms() * 4), 256, 0, s, dtok, dtsb, dtsl, R, L1, D, tmp);
      SIG_CHECK_LAUNCH();
      dtb = tmp;
      ldd = D;
    }
    {
      SIG_PHASE("tokens_proj_bwd");
      TcGemmDesc g = tc_desc();   // dxn[R, W] = dtok[R, D] . proj[W, D]^T
      g.A = tc_k2d(dtb, R, D, ldd);
      g.B = tc_k2d(projb, W, D, D);
      g.M = R; g.N = W; g.K = D;
      g.C[0] = dxn; g.ldc = W; g.out_bf16 = 1;
      g.bn = (W % 256 == 0) ? 256 : 128;
      SIG_TRY(tc_gemm(g, s));
      cudaMemsetAsync(d_proj, 0, (size_t)W * D * sizeof(float), s);
      TcGemmDesc h = tc_desc();   // dproj[W, D] = xn^T dtok, K = R token rows
      h.A = tc_mn2d(xn, R, W, W);
      h.B = tc_mn2d(dtb, R, D, ldd);
      h.M = W; h.N = D; h.K = R;
      h.C[0] = d_proj; h.ldc = D;
      const int units = (int)(ceil_div(W, 128) * ceil_div(D, 128));
      int ks = (tc_num_sms() + units - 1) / units;
      const int kblocks = (int)ceil_div(R, 64);
      if (ks > kblocks) ks = kblocks;
      h.ksplit = ks < 1 ? 1 : ks;
      SIG_TRY(tc_gemm(h, s));
    }
    SIG_PHASE("tokens_ln_bwd");
    SIG_TRY(launch_ln_bwd(t.ctas, x, sb, sl, R, L1, W, ln_w, mu, rstd, (const __nv_bfloat16*)dxn, dx, dxsb, dxsl, part, s));
  }
  SIG_TRY(launch_colsum(part, 2 * (int64_t)W, t.ctas, W, d_ln_w, 1.f, s));
  SIG_TRY(launch_colsum(part + W, 2 * (int64_t)W, t.ctas, W, d_ln_b, 1.f, s));
  return 0;
}


bool tok_shape_ok(int B, int L1, int W, int D) {
  return B >= 1 && L1 >= 2 && W >= 8 && D >= 8 && W <= 1024 && D <= 1024 && (W % 8) == 0 && (D % 8) == 0 && (int64_t)B * L1 < (1 << 30);
}
// x dtype / token dtype pairs: equal, or fp32 x with half tokens (autocast)
bool tok_dtypes_ok(int xdt, int tdt) {
  if (xdt < SIG_F32 || xdt > SIG_F16 || tdt < SIG_F32 || tdt > SIG_F16) return false;
  return xdt == tdt || xdt == SIG_F32;
}
bool a16(const void* p) { return ((uintptr_t)p % 16) == 0; }
int64_t vec_elems(int dt) { return dt == SIG_F32 ? 4 : 8; }   // elements per 16 bytes

}  // namespace
}  // namespace sig

#define SIG_TOK_DISPATCH(FN, ...)                                                                    \
  do {                                                                                               \
    if (dtype == SIG_F32 && tok_dtype == SIG_F32) return FN<float, float>(__VA_ARGS__);              \
    if (dtype == SIG_BF16) return FN<__nv_bfloat16, __nv_bfloat16>(__VA_ARGS__);                     \
    if (dtype == SIG_F16) return FN<__half, __half>(__VA_ARGS__);                                    \
    if (tok_dtype == SIG_BF16) return FN<float, __nv_bfloat16>(__VA_ARGS__);                         \
    return FN<float, __half>(__VA_ARGS__);                                                           \
  } while (0)

template <typename XT, typename TT>
static int tokens_fwd_entry(const void* x, int tdt, int64_t sb, int64_t sl, int B, int L1, int W, int D, const float* ln_w,
                            const float* ln_b, float eps, const float* proj, void* tokens, float* patch_mean, void* saved, void* scratch,
                            cudaStream_t s) {
  return sig::tokens_fwd_t<XT, TT>(static_cast<const XT*>(x), tdt, sb, sl, B, L1, W, D, ln_w, ln_b, eps, proj, static_cast<TT*>(tokens),
                                   patch_mean, static_cast<unsigned char*>(saved), static_cast<unsigned char*>(scratch), s);
}
template <typename XT, typename TT>
static int tokens_bwd_entry(const void* x, int tdt, int64_t sb, int64_t sl, int B, int L1, int W, int D, const float* ln_w,
                            const float* proj, const void* dtok, int64_t dtsb, int64_t dtsl, const void* saved, void* dx, int64_t dxsb,
                            int64_t dxsl, float* d_ln_w, float* d_ln_b, float* d_proj, void* scratch, cudaStream_t s) {
  return sig::tokens_bwd_t<XT, TT>(static_cast<const XT*>(x), tdt, sb, sl, B, L1, W, D, ln_w, proj, static_cast<const TT*>(dtok), dtsb,
                                   dtsl, static_cast<const unsigned char*>(saved), static_cast<XT*>(dx), dxsb, dxsl, d_ln_w, d_ln_b,
                                   d_proj, static_cast<unsigned char*>(scratch), s);
}

extern "C" {

size_t sig_tokens_ws_bytes(int which, int B, int L1, int W, int D, int tok_dtype) {
  using namespace sig;
  if (!tok_shape_ok(B, L1, W, D) || tok_dtype < SIG_F32 || tok_dtype > SIG_F16) return 0;
  const TokLayout t = tok_layout(B * L1, W, D, tok_dtype);
  if (which == 0) return t.saved_bytes;
  if (which == 1) return t.fwd_scratch_bytes;
  if (which == 2) return t.bwd_scratch_bytes;
  return 0;
}

int sig_tokens_fwd(const void* x, int dtype, int64_t x_stride_b, int64_t x_stride_l, int B, int L1, int W, int D, const float* ln_w,
                   const float* ln_b, float eps, const float* proj, void* tokens, int tok_dtype, float* patch_mean, void* saved,
                   size_t saved_bytes, void* scratch, size_t scratch_bytes, int device, void* stream) {
  using namespace sig;
  SIG_ENTER(device);
  if (!x || !ln_w || !ln_b || !proj || !tokens || !saved) return SIG_ERR_NULL;
  if (!tok_dtypes_ok(dtype, tok_dtype)) return SIG_ERR_DTYPE;
  if (!tok_shape_ok(B, L1, W, D)) return SIG_ERR_SHAPE;
  const int64_t ev = vec_elems(dtype);
  if (!a16(x) || !a16(tokens) || !a16(ln_w) || !a16(ln_b) || !a16(proj) || !a16(saved) || (x_stride_b % ev) || (x_stride_l % ev) ||
      (patch_mean && !a16(patch_mean)))
    return SIG_ERR_ALIGN;
  const TokLayout t = tok_layout(B * L1, W, D, tok_dtype);
  if (saved_bytes < t.saved_bytes || scratch_bytes < t.fwd_scratch_bytes || (t.fwd_scratch_bytes && !scratch)) return SIG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  SIG_TOK_DISPATCH(tokens_fwd_entry, x, tok_dtype, x_stride_b, x_stride_l, B, L1, W, D, ln_w, ln_b, eps, proj, tokens, patch_mean, saved,
                   scratch, s);
}

int sig_tokens_bwd(const void* x, int dtype, int64_t x_stride_b, int64_t x_stride_l, int B, int L1, int W, int D, const float* ln_w,
                   const float* proj, const void* dtokens, int tok_dtype, int64_t dt_stride_b, int64_t dt_stride_l, const void* saved,
                   size_t saved_bytes, void* dx, int64_t dx_stride_b, int64_t dx_stride_l, float* d_ln_w, float* d_ln_b, float* d_proj,
                   void* scratch, size_t scratch_bytes, int device, void* stream) {
  using namespace sig;
  SIG_ENTER(device);
  if (!x || !ln_w || !proj || !dtokens || !saved || !dx || !d_ln_w || !d_ln_b || !d_proj || !scratch) return SIG_ERR_NULL;
  if (!tok_dtypes_ok(dtype, tok_dtype)) return SIG_ERR_DTYPE;
  if (!tok_shape_ok(B, L1, W, D)) return SIG_ERR_SHAPE;
  const int64_t ev = vec_elems(dtype), evt = vec_elems(tok_dtype);
  if (!a16(x) || !a16(dtokens) || !a16(dx) || !a16(ln_w) || !a16(proj) || !a16(saved) || !a16(scratch) || !a16(d_proj) || (x_stride_b % ev) ||
      (x_stride_l % ev) || (dt_stride_b % evt) || (dt_stride_l % evt) || (dx_stride_b % ev) || (dx_stride_l % ev))
    return SIG_ERR_ALIGN;
  const TokLayout t = tok_layout(B * L1, W, D, tok_dtype);
  if (saved_bytes < t.saved_bytes || scratch_bytes < t.bwd_scratch_bytes) return SIG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  SIG_TOK_DISPATCH(tokens_bwd_entry, x, tok_dtype, x_stride_b, x_stride_l, B, L1, W, D, ln_w, proj, dtokens, dt_stride_b, dt_stride_l, saved,
                   dx, dx_stride_b, dx_stride_l, d_ln_w, d_ln_b, d_proj, scratch, s);
}

}  // extern "C"

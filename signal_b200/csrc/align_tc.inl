// bf16 fast path of AlignmentM (included by align.cu): tokens are consumed in place (strided bf16),
// the dense 1x1 convolutions of DA_sample run on the tcgen05 GEMM core (tc_gemm.cu) with the two
// linear maps proj_q -> conv_offset[0] folded into one (DAS.py:129,136: no non-linearity between
// them), everything else is HBM-bound SIMT code with 16-byte accesses.

// ---- GAM: mean pool straight from the strided token views.  grid (B, 3), 8 * d/8 threads -----------
// Eight row groups per CTA; a thread owns 8 channels (16 B) of 16 rows and requests ALL of them before the first
// add: one memory round trip per CTA, ~200 KB in flight per SM, and CTAs retire/start continuously (2.6 waves),
// which keeps HBM busy -- a looped version with the same bytes per thread ran its rounds in lock step at 40 %.
// dyn smem: [8][d] floats.  Algorithmic traffic: the tokens, read once.
constexpr int kPoolGroups = 8;
template <typename T>
static __global__ void __launch_bounds__(768) pool_tok_kernel(TokPtrs3 tp, int B, int L, int d, float* __restrict__ mean) {
  pdl_enter();
  extern __shared__ __align__(16) float pool_red[];   // [kPoolGroups][d]
  const int b = blockIdx.x, m = blockIdx.y;
  const int tpr = d / 8;                       // threads per token row
  const int grp = threadIdx.x / tpr, c = (threadIdx.x % tpr) * 8;
  constexpr int kRows = kMaxL / kPoolGroups;   // 16 rows per thread
  if (grp < kPoolGroups) {
    const T* x = static_cast<const T*>(tp.patch[m]) + b * tp.psb[m] + c;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if constexpr (sizeof(T) == 2) {
      uint4 raw[kRows];                        // 64 registers of raw bf16: converted only while summing
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        const int l = grp + u * kPoolGroups;
        raw[u] = l < L ? *reinterpret_cast<const uint4*>(x + l * tp.psl[m]) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        const uint32_t wv[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          acc[2 * t] += __uint_as_float(wv[t] << 16);
          acc[2 * t + 1] += __uint_as_float(wv[t] & 0xffff0000u);
        }
      }
    } else {
      for (int u = 0; u < kRows; ++u) {
        const int l = grp + u * kPoolGroups;
        if (l < L) {
          float v[8];
          load8(x + l * tp.psl[m], v);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += v[i];
        }
      }
    }
    store8(pool_red + grp * d + c, acc);
  }
  __syncthreads();
  const float inv = 1.f / L;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int g2 = 0; g2 < kPoolGroups; ++g2) t += pool_red[g2 * d + i];
    mean[((int64_t)m * B + b) * d + i] = t * inv;
  }
}

// ---- LAM: depthwise 4x4/s4 conv + GELU + 1x1 -> offset logit, from the bf16 pre-activation H ------
// Streaming kernels, no shared-memory staging: thread t owns channels 2t, 2t+1 and reads its 16
// window taps of a sample point as 16 independent 4-byte loads (a warp covers one 128-byte line per
// tap, every sector fully used), so each thread keeps 16-32 loads in flight and the loop over the
// CTA's points needs no block-level synchronisation.  Algorithmic traffic: H read once (fwd);
// H read + dH written once (bwd).
constexpr int kDwFwdMaxPts = 16;   // sample points per CTA, forward (upper bound: size of the reduction scratch)
// Points per CTA such that the grid (chunks x 3 modalities) is one full wave of 2 CTAs per SM.
static int dw_pts_per_cta(int64_t BP, int max_pts) {
  const int slots = 2 * tc_num_sms() / 3;
  int pts = (int)ceil_div(BP, slots > 0 ? slots : 1);
  if (pts < 1) pts = 1;
  return max_pts > 0 && pts > max_pts ? max_pts : pts;
}

__device__ __forceinline__ void dw_load_window(const __nv_bfloat16* __restrict__ base, int w, int d, bool active, uint32_t (&hv)[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k)
    hv[k] = active ? ld_global_u32_early(base + (int64_t)((k >> 2) * w + (k & 3)) * d) : 0u;
}

// grid (ceil(B*P / pts), 3), d/2 threads (rounded up to a warp).  U saved (fp32) for backward.
// <D, W> = compile-time channel count / grid width (0 = runtime) so tap offsets become immediates.
template <int D, int W>
static __global__ void __launch_bounds__(384, 2) lam_dw_fwd_tc_kernel(const __nv_bfloat16* __restrict__ H, int64_t hms,
                                                                      sig_align_params prm, Geo g, int B, int L, int d_rt,
                                                                      int pts, float* __restrict__ U, float* __restrict__ o) {
  pdl_launch_dependents();
  const int d = D ? D : d_rt;
  if (W) g.w = W;
  __shared__ float red[kDwFwdMaxPts][16];
  const int m = blockIdx.y;
  const int c = threadIdx.x * 2;
  const bool active = c < d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // parameters are not produced by the preceding kernels: fetch them before the dependency wait
  float wk0[16], wk1[16];
  float2 w4 = make_float2(0.f, 0.f), bd = make_float2(0.f, 0.f);
  if (active) {
    const float4* wp = reinterpret_cast<const float4*>(prm.off2_w[m] + (int64_t)c * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = wp[q], b4 = wp[4 + q];
      wk0[4 * q] = a.x; wk0[4 * q + 1] = a.y; wk0[4 * q + 2] = a.z; wk0[4 * q + 3] = a.w;
      wk1[4 * q] = b4.x; wk1[4 * q + 1] = b4.y; wk1[4 * q + 2] = b4.z; wk1[4 * q + 3] = b4.w;
    }
    w4 = *reinterpret_cast<const float2*>(prm.off4_w[m] + c);
    bd = *reinterpret_cast<const float2*>(prm.off2_b[m] + c);
  }
  pdl_wait();
  const __nv_bfloat16* Hm = H + m * hms;
  const int bp0 = blockIdx.x * pts, npts = min(pts, B * g.P - bp0);
#pragma unroll 2
  for (int i = 0; i < npts; ++i) {
    const int bp = bp0 + i, b = bp / g.P, p = bp % g.P;
    const int py = p / g.Wk, px = p % g.Wk;
    uint32_t hv[16];
    dw_load_window(Hm + ((int64_t)b * L + (4 * py) * g.w + 4 * px) * d + c, g.w, d, active, hv);
    float u0 = bd.x, u1 = bd.y;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      u0 = fmaf(gelu_tanh_f(bf16lo_to_f32(hv[k])), wk0[k], u0);
      u1 = fmaf(gelu_tanh_f(bf16hi_to_f32(hv[k])), wk1[k], u1);
    }
    float part = 0.f;
    if (active) {
      *reinterpret_cast<float2*>(U + ((int64_t)m * B * g.P + bp) * d + c) = make_float2(u0, u1);
      part = gelu_fast_f(u0) * w4.x + gelu_fast_f(u1) * w4.y;
    }
    part = warp_sum(part);
    if (lane == 0) red[i][warp] = part;
  }
  __syncthreads();
  if ((int)threadIdx.x < npts) {   // fixed summation order: deterministic
    float t = 0.f;
    for (int q = 0; q < nw; ++q) t += red[threadIdx.x][q];
    o[(int64_t)m * B * g.P + bp0 + threadIdx.x] = t;
  }
}

// Backward of the offset-net tail, one pass over H:
//   dU = dO * w4 * gelu'(U);  dH[b,pos,c] = dU * wdw[c,k] * gelu'(H[b,pos,c])   (bf16 out)
// and per-CTA partial sums of the parameter gradients
//   dwdw[c,k] += dU * gelu(H[pos(p,k)]),  dbdw += dU,  dw4 += dO * gelu(U),  dbf += dH
// grid (ceil(B*P / pts), 3), d/2 threads (2 channels each); the depthwise taps live in shared
// memory as [k][c] (conflict-free 8-byte reads) so two CTAs fit the register file of an SM.
// part: [3][nchunk][19][d]
template <int D, int W>
static __global__ void __launch_bounds__(384, 2) lam_dw_bwd_tc_kernel(const __nv_bfloat16* __restrict__ H, int64_t hms,
                                                                      const float* __restrict__ U, const float* __restrict__ dO,
                                                                      sig_align_params prm, Geo g, int B, int L, int d_rt,
                                                                      int pts, __nv_bfloat16* __restrict__ dH, float* __restrict__ part) {
  pdl_launch_dependents();
  const int d = D ? D : d_rt;
  if (W) g.w = W;
  extern __shared__ __align__(16) unsigned char dw_smem[];
  float* wT = reinterpret_cast<float*>(dw_smem);   // [16][d]
  const int m = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int c = threadIdx.x * 2;
  const bool active = c < d;
  float2 w4 = make_float2(0.f, 0.f);
  if (active) {
    w4 = *reinterpret_cast<const float2*>(prm.off4_w[m] + c);
    const float4* wp = reinterpret_cast<const float4*>(prm.off2_w[m] + (int64_t)c * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = wp[q], b4 = wp[4 + q];
      *reinterpret_cast<float2*>(wT + (4 * q) * d + c) = make_float2(a.x, b4.x);
      *reinterpret_cast<float2*>(wT + (4 * q + 1) * d + c) = make_float2(a.y, b4.y);
      *reinterpret_cast<float2*>(wT + (4 * q + 2) * d + c) = make_float2(a.z, b4.z);
      *reinterpret_cast<float2*>(wT + (4 * q + 3) * d + c) = make_float2(a.w, b4.w);
    }
  }
  // (each thread reads back only the taps it wrote itself: no barrier needed)
  pdl_wait();
  const __nv_bfloat16* Hm = H + m * hms;
  __nv_bfloat16* dHm = dH + m * hms;
  const int bp0 = chunk * pts, bp_end = min(B * g.P, bp0 + pts);
  float a0[19], a1[19];
#pragma unroll
  for (int i = 0; i < 19; ++i) a0[i] = a1[i] = 0.f;
  for (int bp = bp0; bp < bp_end; ++bp) {
    const int b = bp / g.P, p = bp % g.P;
    const int py = p / g.Wk, px = p % g.Wk;
    const int64_t off = ((int64_t)b * L + (4 * py) * g.w + 4 * px) * d + c;
    uint32_t hv[16];
    dw_load_window(Hm + off, g.w, d, active, hv);
    if (active) {
      const float go = dO[(int64_t)m * B * g.P + bp];
      const float2 u = *reinterpret_cast<const float2*>(U + ((int64_t)m * B * g.P + bp) * d + c);
      float gu0, dgu0, gu1, dgu1;
      gelu_fast(u.x, gu0, dgu0);
      gelu_fast(u.y, gu1, dgu1);
      const float du0 = go * w4.x * dgu0, du1 = go * w4.y * dgu1;
      a0[16] += du0; a1[16] += du1;
      a0[17] = fmaf(go, gu0, a0[17]); a1[17] = fmaf(go, gu1, a1[17]);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float2 wk = *reinterpret_cast<const float2*>(wT + k * d + c);
        float gh0, dgh0, gh1, dgh1;   // gelu and gelu' share one tanh per element
        gelu_tanh(bf16lo_to_f32(hv[k]), gh0, dgh0);
        gelu_tanh(bf16hi_to_f32(hv[k]), gh1, dgh1);
        const float dh0 = du0 * wk.x * dgh0, dh1 = du1 * wk.y * dgh1;
        *reinterpret_cast<__nv_bfloat162*>(dHm + off + (int64_t)((k >> 2) * g.w + (k & 3)) * d) = __floats2bfloat162_rn(dh0, dh1);
        a0[k] = fmaf(du0, gh0, a0[k]);
        a1[k] = fmaf(du1, gh1, a1[k]);
        a0[18] += dh0; a1[18] += dh1;
      }
    }
  }
  if (active) {
    float* dst = part + (((int64_t)m * nchunk + chunk) * 19) * d + c;
#pragma unroll
    for (int i = 0; i < 19; ++i) *reinterpret_cast<float2*>(dst + (int64_t)i * d) = make_float2(a0[i], a1[i]);
  }
}

#include "lam_dw_ring.inl"

// deterministic reduction of the partials into the parameter gradients.  grid (ceil(d/64), 19, 3), 256 threads:
// 64 channels x 4 chunk lanes
static __global__ void __launch_bounds__(256) lam_dw_param_reduce_kernel(const float* __restrict__ part, int nchunk,
                                                                         sig_align_param_grads gr, float* __restrict__ dbf, int d,
                                                                         const float* __restrict__ wscale) {
  pdl_enter();
  __shared__ float sm[4][64];
  const int m = blockIdx.z, i = blockIdx.y;
  const int cl = threadIdx.x & 63, r = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + cl;
  float acc = 0.f;
  if (c < d)
    for (int ch = r; ch < nchunk; ch += 4) acc += part[(((int64_t)m * nchunk + ch) * 19 + i) * d + c];
  sm[r][cl] = acc;
  __syncthreads();
  if (r == 0 && c < d) {
    const float t = (sm[0][cl] + sm[1][cl] + sm[2][cl] + sm[3][cl]) * (wscale ? *wscale : 1.f);   // (eager chain: unit weight)
    if (i < 16) gr.off2_w[m][c * 16 + i] = t;
    else if (i == 16) gr.off2_b[m][c] = t;
    else if (i == 17) gr.off4_w[m][c] = t;
    else dbf[(int64_t)m * d + c] = t;
  }
}

// bilinear sampling straight from the strided token view.  grid (B*P, 3)
template <typename T>
static __global__ void __launch_bounds__(128) lam_sample_fwd_tok_kernel(TokPtrs3 tp, const float* __restrict__ o, Geo g, int B, int d,
                                                                        float* __restrict__ S) {
  pdl_enter();
  const int m = blockIdx.y, bp = blockIdx.x, b = bp / g.P, p = bp % g.P;
  const Taps t = make_taps(o[(int64_t)m * B * g.P + bp], p, g);
  const T* x = static_cast<const T*>(tp.patch[m]) + b * tp.psb[m];
  const int c = threadIdx.x * 8;
  if (c >= d) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (t.l[k] >= 0) {
      float v[8];
      load8(x + t.l[k] * tp.psl[m] + c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(t.wgt[k], v[i], acc[i]);
    }
  store8(S + ((int64_t)m * B * g.P + bp) * d + c, acc);
}

// grid-gradient reduction -> d(offset logit).  grid (B*P, 3)
template <typename T>
static __global__ void __launch_bounds__(128) lam_sample_bwd_tok_kernel(TokPtrs3 tp, const float* __restrict__ o,
                                                                        const float* __restrict__ dS, Geo g, int B, int d,
                                                                        float* __restrict__ dO) {
  pdl_enter();
  __shared__ float scratch[33];
  const int m = blockIdx.y, bp = blockIdx.x, b = bp / g.P, p = bp % g.P;
  const Taps t = make_taps(o[(int64_t)m * B * g.P + bp], p, g);
  const T* x = static_cast<const T*>(tp.patch[m]) + b * tp.psb[m];
  const int c = threadIdx.x * 8;
  float giy = 0.f, gix = 0.f;
  if (c < d) {
    float v[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (t.l[k] >= 0) load8(x + t.l[k] * tp.psl[m] + c, v[k]);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[k][i] = 0.f;
      }
    }
    float ds[8];
    load8(dS + ((int64_t)m * B * g.P + bp) * d + c, ds);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      giy += ds[i] * ((v[2][i] - v[0][i]) * (1.f - t.wx) + (v[3][i] - v[1][i]) * t.wx);
      gix += ds[i] * ((v[1][i] - v[0][i]) * (1.f - t.wy) + (v[3][i] - v[2][i]) * t.wy);
    }
  }
  giy = block_sum(giy, scratch);
  gix = block_sum(gix, scratch);
  if (threadIdx.x == 0) dO[(int64_t)m * B * g.P + bp] = giy * t.gy + gix * t.gx;
}

// dst = bf16(scale * src), n a multiple of 4 (d * d elements); scale is a device scalar
template <typename T>
static __global__ void __launch_bounds__(256) scale_to_bf16_kernel(const T* __restrict__ src, const float* __restrict__ scale,
                                                                   __nv_bfloat16* __restrict__ dst, int64_t n) {
  pdl_enter();
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float k = *scale;
  float v[4];
  if constexpr (sizeof(T) == 4) {
    const float4 x = *reinterpret_cast<const float4*>(src + i);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  } else {
    const uint2 x = *reinterpret_cast<const uint2*>(src + i);
    v[0] = bf16lo_to_f32(x.x); v[1] = bf16hi_to_f32(x.x); v[2] = bf16lo_to_f32(x.y); v[3] = bf16hi_to_f32(x.y);
  }
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0] * k, v[1] * k), b = __floats2bfloat162_rn(v[2] * k, v[3] * k);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(dst + i) = o;
}

// sparse part of d(patches): every bilinear tap adds w * dS[p,:] to the token row it read.
// grid (B, 3), d/8 threads; taps are applied one after the other, so no two threads touch the same element.
template <typename T>
static __global__ void __launch_bounds__(128) lam_sparse_add_kernel(GradPtrs3 gp, const float* __restrict__ o,
                                                                    const float* __restrict__ dS, Geo g, int B, int d,
                                                                    const float* __restrict__ wscale) {
  pdl_enter();
  const float ws = wscale ? *wscale : 1.f;   // (dS of the eager chain carries a unit loss weight)
  const int m = blockIdx.y, b = blockIdx.x;
  const int c = threadIdx.x * 8;
  if (c >= d) return;
  T* dx = static_cast<T*>(gp.dpatch[m]) + b * gp.psb[m];
  for (int p = 0; p < g.P; ++p) {
    const Taps t = make_taps(o[((int64_t)m * B + b) * g.P + p], p, g);
    float ds[8];
    load8(dS + (((int64_t)m * B + b) * g.P + p) * d + c, ds);
#pragma unroll
    for (int i = 0; i < 8; ++i) ds[i] *= ws;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (t.l[k] >= 0 && t.wgt[k] != 0.f) {
        float v[8];
        T* dst = dx + t.l[k] * gp.psl[m] + c;
        load8(dst, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(t.wgt[k], ds[i], v[i]);
        store8(dst, v);
      }
  }
}

// dW0[m][o][c] += db'[m][o] * bq[m][c]   (the bias of proj_q reaches conv_offset[0]'s weight gradient: Q = X Wq^T + bq)
// grid (ceil(d*d/256), 3)
static __global__ void rank1_add3_kernel(sig_align_param_grads gr, const float* __restrict__ dbf, sig_align_params prm, int d) {
  pdl_enter();
  const int m = blockIdx.y;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < (int64_t)d * d) gr.off0_w[m][i] += dbf[(int64_t)m * d + i / d] * prm.proj_q_b[m][i % d];
}

// b'[m] = W0[m] bq[m] + b0[m]   grid (ceil(d/8), 3), one warp per output
static __global__ void __launch_bounds__(256) lam_fold_bias_kernel(sig_align_params prm, int d, float* __restrict__ bfold) {
  pdl_enter();
  const int m = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= d) return;
  const float* W = prm.off0_w[m] + (int64_t)i * d;
  const float* x = prm.proj_q_b[m];
  float a = 0.f;
  for (int k = lane * 4; k < d; k += 128) {
    const float4 w = *reinterpret_cast<const float4*>(W + k);
    const float4 v = *reinterpret_cast<const float4*>(x + k);
    a += w.x * v.x + w.y * v.y + w.z * v.z + w.w * v.w;
  }
  a = warp_sum(a);
  if (lane == 0) bfold[(int64_t)m * d + i] = a + prm.off0_b[m][i];
}

// db0[m] = db'[m];  dbq[m] = W0[m]^T db'[m]   grid (ceil(d/32), 3), 1024 threads = 32 outputs x 32 k-lanes
static __global__ void __launch_bounds__(1024) lam_unfold_bias_kernel(sig_align_params prm, sig_align_param_grads gr,
                                                                      const float* __restrict__ dbf, int d) {
  pdl_enter();
  __shared__ float sm[32][33];
  const int m = blockIdx.y;
  const int il = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + il;
  const float* W = prm.off0_w[m];
  const float* x = dbf + (int64_t)m * d;
  float a0 = 0.f, a1 = 0.f;
  if (i < d) {
    int k = r;
    for (; k + 32 < d; k += 64) {
      a0 = fmaf(W[(int64_t)k * d + i], x[k], a0);
      a1 = fmaf(W[(int64_t)(k + 32) * d + i], x[k + 32], a1);
    }
    for (; k < d; k += 32) a0 = fmaf(W[(int64_t)k * d + i], x[k], a0);
  }
  sm[r][il] = a0 + a1;
  __syncthreads();
  if (r == 0 && i < d) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) t += sm[q][il];
    gr.proj_q_b[m][i] = t;
    gr.off0_b[m][i] = x[i];
  }
}

// zero the CLS gradient rows (packed [B,1+L,d] destination).  grid (B, 3)
template <typename T>
static __global__ void zero_cls_kernel(GradPtrs3 gp, int d) {
  pdl_enter();
  const int m = blockIdx.y, b = blockIdx.x;
  if (!gp.dcls[m]) return;
  T* dst = static_cast<T*>(gp.dcls[m]) + b * gp.csb[m];
  const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) store8(dst + c, z);
}

// ---- ctx layout of the bf16 path -------------------------------------------------------------------
struct AlignTcCtx {
  // GAM (fp32, small)
  float *mean, *f, *nrm, *self4, *lv, *la, *V, *rowstat, *colstat, *Wlv, *Wla, *rowA, *colC, *dtau, *df, *dmean;
  float *gstat, *lossp, *dtaup;
  __nv_bfloat16 *fb, *fA, *fB, *WW;
  // LAM
  __nv_bfloat16 *W0b, *Wqb, *Wfb, *dWfb, *Wfs;   // [3][d*d]
  float *bfold, *dWf, *dbf;                // [3][d], [3][d*d], [3][d]
  __nv_bfloat16 *H, *dH;                   // [3][B*L*d]
  float *U, *dwpart, *o, *dO, *S, *dS, *part;
  size_t bytes;
};

static AlignTcCtx align_tc_ctx(void* base, int B, int L, int d) {
  Arena a(base);
  AlignTcCtx c;
  const size_t BL = (size_t)B * L, P = (size_t)(L / 16), dd = (size_t)d * d;
  c.mean = a.take<float>((size_t)3 * B * d);
  c.f = a.take<float>((size_t)3 * B * d);
  c.nrm = a.take<float>((size_t)3 * B);
  c.self4 = a.take<float>((size_t)4 * B);
  c.lv = a.take<float>((size_t)B * B);
  c.la = a.take<float>((size_t)B * B);
  c.V = a.take<float>((size_t)B * B);
  c.rowstat = a.take<float>((size_t)2 * B);
  c.colstat = a.take<float>((size_t)2 * B);
  c.Wlv = a.take<float>((size_t)B * B);
  c.Wla = a.take<float>((size_t)B * B);
  c.rowA = a.take<float>((size_t)B);
  c.colC = a.take<float>((size_t)3 * B);
  c.dtau = a.take<float>(4);
  c.df = a.take<float>((size_t)3 * B * d);
  c.dmean = a.take<float>((size_t)3 * B * d);
  c.gstat = a.take<float>((size_t)2 * B);
  c.lossp = a.take<float>((size_t)2 * B);
  c.dtaup = a.take<float>((size_t)B);
  c.fb = a.take<__nv_bfloat16>((size_t)3 * B * d);
  c.fA = a.take<__nv_bfloat16>((size_t)B * 3 * d);
  c.fB = a.take<__nv_bfloat16>((size_t)2 * B * 3 * d);
  c.WW = a.take<__nv_bfloat16>((size_t)B * 2 * ((B + 7) / 8 * 8));
  c.W0b = a.take<__nv_bfloat16>(3 * dd);
  c.Wqb = a.take<__nv_bfloat16>(3 * dd);
  c.Wfb = a.take<__nv_bfloat16>(3 * dd);
  c.dWfb = a.take<__nv_bfloat16>(3 * dd);
  c.Wfs = a.take<__nv_bfloat16>(3 * dd);
  c.bfold = a.take<float>((size_t)3 * d);
  c.dWf = a.take<float>(3 * dd);
  c.dbf = a.take<float>((size_t)3 * d);
  c.H = a.take<__nv_bfloat16>(3 * BL * d);
  c.dH = a.take<__nv_bfloat16>(3 * BL * d);
  c.U = a.take<float>(3 * (size_t)B * P * d);
  c.dwpart = a.take<float>(3 * (size_t)ceil_div((int64_t)B * P, dw_pts_per_cta((int64_t)B * P, 0)) * 19 * d);
  c.o = a.take<float>(3 * (size_t)B * P);
  c.dO = a.take<float>(3 * (size_t)B * P);
  c.S = a.take<float>(3 * (size_t)B * P * d);
  c.dS = a.take<float>(3 * (size_t)B * P * d);
  c.part = a.take<float>((size_t)B * P);
  c.bytes = a.off;
  return c;
}

static TokPtrs3 tok_ptrs3(const sig_tokens* t) {
  TokPtrs3 tp;
  for (int m = 0; m < 3; ++m) {
    tp.patch[m] = t->patch[m];
    tp.psb[m] = t->patch_stride_b[m];
    tp.psl[m] = t->patch_stride_l[m];
  }
  return tp;
}
static GradPtrs3 grad_ptrs3(const sig_token_grads* g) {
  GradPtrs3 gp;
  for (int m = 0; m < 3; ++m) {
    gp.dpatch[m] = g->dpatch[m]; gp.dcls[m] = g->dcls[m];
    gp.psb[m] = g->patch_stride_b[m]; gp.psl[m] = g->patch_stride_l[m]; gp.csb[m] = g->cls_stride_b[m];
  }
  return gp;
}

static TcOperand tok_operand(const sig_tokens* t, int mode) {
  TcOperand o{};
  for (int m = 0; m < 3; ++m) o.ptr[m] = t->patch[m];
  o.mode = mode;
  o.stride_b = t->patch_stride_b[0];
  o.stride_l = t->patch_stride_l[0];
  o.rows = t->B;
  o.cols = t->d;
  return o;
}
static TcOperand batched(TcOperand o, const void* base, size_t stride_elems) { return tc_batched(o, base, stride_elems, 3); }

// ONE predicate for sizing (sig_ctx_bytes knows dtype, L, d, flags -- not the strides) and for dispatch, so the ctx
// buffer a caller sized always fits the path that runs.
static bool tc_shape_ok(int dtype, int L, int d, unsigned flags) {
  if (flags & SIG_FLAG_FORCE_SIMT) return false;
  return dtype == SIG_BF16 && L == 128 && d <= 768;   // (the LAM depthwise kernels run d/2 <= 384 threads)
}
static bool tc_path_ok(const sig_tokens* t, unsigned flags) { return tc_shape_ok(t->dtype, t->L, t->d, flags); }
// The tensor-core path reads the three modalities through one 3-D tensor-map geometry: their patch strides must agree
// (SIG_ERR_SHAPE otherwise -- the Python wrapper then hands over contiguous copies; never a silent change of path).
static bool tc_strides_ok(const sig_tokens* t) {
  for (int m = 1; m < 3; ++m)
    if (t->patch_stride_b[m] != t->patch_stride_b[0] || t->patch_stride_l[m] != t->patch_stride_l[0]) return false;
  return true;
}

// ---- GAM on the bf16 path -----------------------------------------------------------------------
// The B x B Gram entries lv = f_r f_n^T, la = f_r f_t^T feed a 3x3 determinant that cancels heavily
// when the modalities align, so they must not be rounded to bf16.  They run on the tensor cores as a
// 3-term split-bf16 product (x = hi + lo: hi*hi + hi*lo + lo*hi, fp32 accumulate, error ~2^-16):
// operands are written as [hi | hi | lo] (A side) and [hi | lo | hi] (B side) along K.
// grid B, 256 threads.  Also writes f (fp32), fb (plain bf16) and the per-sample Gram entries.
static __global__ void __launch_bounds__(256) gam_norm_split_kernel(const float* __restrict__ mean, int B, int d, float* __restrict__ f,
                                                                    __nv_bfloat16* __restrict__ fb, __nv_bfloat16* __restrict__ fA,
                                                                    __nv_bfloat16* __restrict__ fB, float* __restrict__ nrm,
                                                                    float* __restrict__ self4) {
  pdl_enter();
  __shared__ float scratch[33];
  const int b = blockIdx.x;
  float inv[3];
  for (int m = 0; m < 3; ++m) {
    const float* x = mean + ((int64_t)m * B + b) * d;
    float s = 0.f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) s += x[c] * x[c];
    s = sqrtf(block_sum(s, scratch));
    const float dn = fmaxf(s, 1e-12f);
    inv[m] = 1.f / dn;
    if (threadIdx.x == 0) nrm[m * B + b] = dn;
  }
  float ll = 0.f, vv = 0.f, aa = 0.f, va = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float v[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      v[m] = mean[((int64_t)m * B + b) * d + c] * inv[m];
      f[((int64_t)m * B + b) * d + c] = v[m];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v[m]);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v[m] - __bfloat162float(hi));
      fb[((int64_t)m * B + b) * d + c] = hi;
      if (m == 0) {
        __nv_bfloat16* a = fA + (int64_t)b * 3 * d + c;
        a[0] = hi; a[d] = hi; a[2 * d] = lo;
      } else {
        __nv_bfloat16* q = fB + ((int64_t)(m - 1) * B + b) * 3 * d + c;
        q[0] = hi; q[d] = lo; q[2 * d] = hi;
      }
    }
    ll += v[0] * v[0]; vv += v[1] * v[1]; aa += v[2] * v[2]; va += v[1] * v[2];
  }
  ll = block_sum(ll, scratch); vv = block_sum(vv, scratch);
  aa = block_sum(aa, scratch); va = block_sum(va, scratch);
  if (threadIdx.x == 0) {
    self4[0 * B + b] = ll; self4[1 * B + b] = vv; self4[2 * B + b] = aa; self4[3 * B + b] = va;
  }
}

// Stage 1 of the contrastive loss: block r < B handles row r of Z = -V/tau (log-sum-exp, mean, loss
// term), block r >= B handles column r - B.  grid 2B, 128 threads.  stat[r] = lse, lossp[r] = loss term.
static __global__ void __launch_bounds__(128) gam_stats_kernel(const float* __restrict__ self4, const float* __restrict__ lv,
                                                               const float* __restrict__ la, const float* __restrict__ tau_p, int B,
                                                               float* __restrict__ stat, float* __restrict__ lossp) {
  pdl_enter();
  __shared__ float scratch[33];
  const float* ll = self4;
  const float* vv = self4 + B;
  const float* aa = self4 + 2 * B;
  const float* va = self4 + 3 * B;
  const float itau = 1.f / *tau_p;
  const int r = blockIdx.x;
  const bool is_row = r < B;
  const int i0 = is_row ? r : r - B;
  float mx = -INFINITY;
  for (int t = threadIdx.x; t < B; t += blockDim.x) {
    const int i = is_row ? i0 : t, j = is_row ? t : i0;
    const int64_t idx = (int64_t)i * B + j;
    mx = fmaxf(mx, -sqrtf(fabsf(gram_det(ll[i], vv[j], aa[j], va[j], lv[idx], la[idx]))) * itau);
  }
  mx = block_max(mx, scratch);
  float se = 0.f, sz = 0.f, zd = 0.f;
  for (int t = threadIdx.x; t < B; t += blockDim.x) {
    const int i = is_row ? i0 : t, j = is_row ? t : i0;
    const int64_t idx = (int64_t)i * B + j;
    const float z = -sqrtf(fabsf(gram_det(ll[i], vv[j], aa[j], va[j], lv[idx], la[idx]))) * itau;
    se += expf(z - mx);
    sz += z;
    if (i == j) zd = z;
  }
  se = block_sum(se, scratch);
  sz = block_sum(sz, scratch);
  zd = block_sum(zd, scratch);
  if (threadIdx.x == 0) {
    const float lse = mx + logf(se);
    stat[r] = lse;
    lossp[r] = (1.f - kLabelSmooth) * (lse - zd) + kLabelSmooth * (lse - sz / B);
  }
}

// Stage 2: block r < B: row r of the coefficient matrices Wlv, Wla (fp32 + bf16 [B, 2*Bp] side by side, Bp = B rounded up to 8),
// rowA[r] and the row's share of d(tau); block r >= B: column sums colC[0..2][r - B].  grid 2B, 128 threads.
static __global__ void __launch_bounds__(128) gam_coef_kernel(const float* __restrict__ self4, const float* __restrict__ lv,
                                                              const float* __restrict__ la, const float* __restrict__ tau_p,
                                                              const float* __restrict__ stat, int B, int Bp, float* __restrict__ Wlv,
                                                              float* __restrict__ Wla, __nv_bfloat16* __restrict__ WW,
                                                              float* __restrict__ rowA, float* __restrict__ colC,
                                                              float* __restrict__ dtaup) {
  pdl_enter();
  __shared__ float scratch[33];
  const float* ll = self4;
  const float* vv = self4 + B;
  const float* aa = self4 + 2 * B;
  const float* va = self4 + 3 * B;
  const float itau = 1.f / *tau_p;
  const float tgt_off = kLabelSmooth / B, tgt_on = 1.f - kLabelSmooth + kLabelSmooth / B;
  const int r = blockIdx.x;
  const bool is_row = r < B;
  const int i0 = is_row ? r : r - B;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, tp = 0.f;
  for (int t = threadIdx.x; t < B; t += blockDim.x) {
    const int i = is_row ? i0 : t, j = is_row ? t : i0;
    const int64_t idx = (int64_t)i * B + j;
    const float det = gram_det(ll[i], vv[j], aa[j], va[j], lv[idx], la[idx]);
    const float v = sqrtf(fabsf(det)), z = -v * itau;
    const float tg = i == j ? tgt_on : tgt_off;
    const float pr = expf(z - stat[i]) - tg, pc = expf(z - stat[B + j]) - tg;
    const float dZ = (0.5f / B) * (pr + pc);
    const float dd = ddet_of(-dZ * itau, det, v);
    if (is_row) {
      const float wl = dd * (-2.f * (lv[idx] * aa[j] - va[j] * la[idx]));
      const float wa = dd * (2.f * (lv[idx] * va[j] - vv[j] * la[idx]));
      Wlv[idx] = wl;
      Wla[idx] = wa;
      WW[(int64_t)i * 2 * Bp + j] = __float2bfloat16_rn(wl);
      WW[(int64_t)i * 2 * Bp + Bp + j] = __float2bfloat16_rn(wa);
      s0 += dd * (vv[j] * aa[j] - va[j] * va[j]);
      // d(tau): subtract the diagonal volumes (sum_j pr = 0, sum_i pc = 0) to avoid fp32 cancellation
      const float vii = sqrtf(fabsf(gram_det(ll[i], vv[i], aa[i], va[i], lv[(int64_t)i * B + i], la[(int64_t)i * B + i])));
      const float vjj = sqrtf(fabsf(gram_det(ll[j], vv[j], aa[j], va[j], lv[(int64_t)j * B + j], la[(int64_t)j * B + j])));
      tp += (0.5f / B) * (pr * (v - vii) + pc * (v - vjj)) * itau * itau;
    } else {
      s0 += dd * (ll[i] * aa[j] - la[idx] * la[idx]);
      s1 += dd * (-2.f * (ll[i] * va[j] - lv[idx] * la[idx]));
      s2 += dd * (ll[i] * vv[j] - lv[idx] * lv[idx]);
    }
  }
  s0 = block_sum(s0, scratch);
  if (is_row) {
    tp = block_sum(tp, scratch);
    if (threadIdx.x == 0) {
      rowA[i0] = s0;
      dtaup[i0] = tp;
    }
  } else {
    s1 = block_sum(s1, scratch);
    s2 = block_sum(s2, scratch);
    if (threadIdx.x == 0) {
      colC[0 * B + i0] = s0; colC[1 * B + i0] = s1; colC[2 * B + i0] = s2;
    }
  }
}

// loss = (0.5/B) * sum lossp[0..2B) ; dtau = sum dtaup[0..B)
static __global__ void __launch_bounds__(256) gam_final_kernel(const float* __restrict__ lossp, const float* __restrict__ dtaup, int B,
                                                               float* __restrict__ loss, float* __restrict__ dtau) {
  pdl_enter();
  __shared__ float scratch[33];
  float a = 0.f, t = 0.f;
  for (int i = threadIdx.x; i < 2 * B; i += blockDim.x) a += lossp[i];
  for (int i = threadIdx.x; i < B; i += blockDim.x) t += dtaup[i];
  a = block_sum(a, scratch);
  t = block_sum(t, scratch);
  if (threadIdx.x == 0) {
    *loss = a * (0.5f / B);
    *dtau = t;
  }
}

// The part of LAM's backward that is LINEAR in the loss weight and needs nothing from the caller's backward call:
// d(MSE) -> d(samples) -> d(offset logits) -> depthwise/GELU tail (dH and the partial parameter sums).  `gscale` is the
// device scalar d(loss)/d(lam) or NULL for a unit weight -- with SIG_FLAG_EAGER_BWD the forward call runs this chain with
// a unit weight right behind the forward kernels (where the GPU is otherwise waiting for SIM's chain of small kernels),
// and the backward call applies the weight where the results are consumed.
static int lam_backward_chain_tc(const AlignTcCtx& c, const TokPtrs3& tp, const sig_align_params* p, const Geo& g, int B, int L, int d,
                                 int h, int w, const float* gscale, cudaStream_t s) {
  const size_t BL = (size_t)B * L;
  const unsigned cthreads = (unsigned)ceil_div(d / 8, 32) * 32;
  const bool dw_ring = dw_ring_ok(L, h, w, d);
  {
    SIG_PHASE("lam_sample_bwd");
    SIG_LAUNCH((lam_mse_bwd_kernel), B * g.P, 256, 0, s, c.S, (int64_t)B * g.P * d, d, 2.f / (3.f * (float)B * g.P * d), gscale, c.dS);
    SIG_CHECK_LAUNCH();
    SIG_LAUNCH((lam_sample_bwd_tok_kernel<__nv_bfloat16>), dim3(B * g.P, 3), cthreads, 0, s, tp, c.o, c.dS, g, B, d, c.dO);
    SIG_CHECK_LAUNCH();
  }
  if (dw_ring) {
    SIG_PHASE("lam_dwconv_bwd");
    if (d == 768 && w == 8) SIG_TRY((launch_dw_bwd_ring<768, 8>(c.H, c.U, c.dO, *p, B, h, c.dH, c.dwpart, s)));
    else if (d == 768) SIG_TRY((launch_dw_bwd_ring<768, 16>(c.H, c.U, c.dO, *p, B, h, c.dH, c.dwpart, s)));
    else if (w == 8) SIG_TRY((launch_dw_bwd_ring<512, 8>(c.H, c.U, c.dO, *p, B, h, c.dH, c.dwpart, s)));
    else SIG_TRY((launch_dw_bwd_ring<512, 16>(c.H, c.U, c.dO, *p, B, h, c.dH, c.dwpart, s)));
  } else {
    SIG_PHASE("lam_dwconv_bwd");
    const int pts = dw_pts_per_cta((int64_t)B * g.P, 0);
    const int nchunk = (int)ceil_div((int64_t)B * g.P, pts);
    const size_t dw_smem_bytes = (size_t)16 * d * sizeof(float);
    const unsigned thr = (unsigned)ceil_div(d / 2, 32) * 32;
#define SIG_DW_BWD(D, W)                                                                                                    \
  SIG_LAUNCH((lam_dw_bwd_tc_kernel<D, W>), dim3(nchunk, 3), thr, dw_smem_bytes, s, c.H, (int64_t)BL * d, c.U, c.dO, *p, g, B, L, d, \
             pts, c.dH, c.dwpart)
    if (d == 768 && w == 8) SIG_DW_BWD(768, 8);
    else if (d == 768 && w == 16) SIG_DW_BWD(768, 16);
    else if (d == 512 && w == 8) SIG_DW_BWD(512, 8);
    else if (d == 512 && w == 16) SIG_DW_BWD(512, 16);
    else SIG_DW_BWD(0, 0);
#undef SIG_DW_BWD
    SIG_CHECK_LAUNCH();
  }
  return 0;
}

// GAM mean pool of the three bf16 patch maps -> mean [3][B][d] fp32
static int pool_tokens_bf16(const sig_tokens* tok, float* mean, cudaStream_t s) {
  const int B = tok->B, L = tok->L, d = tok->d;
  const bool rows_contig = tok->patch_stride_l[0] == d && tok->patch_stride_l[1] == d && tok->patch_stride_l[2] == d;
  SIG_PHASE("gam_pool");
  if (tok_ring_enabled() && L == kMaxL && (d == 512 || d == 768) && rows_contig) {
    // streaming ring: 4 x 48 KB in flight per SM (tok_ring.cuh)
    TokSrc3 src;
    for (int m = 0; m < 3; ++m) { src.patch[m] = tok->patch[m]; src.psb[m] = tok->patch_stride_b[m]; }
    const int n_items = 3 * B * (kMaxL / 32);
    const int ctas = n_items < tc_num_sms() ? n_items : tc_num_sms();
    cudaMemsetAsync(mean, 0, (size_t)3 * B * d * sizeof(float), s);   // groups split between two CTAs are added atomically
    if (d == 768) {
      ensure_dyn_smem(pool_ring_kernel<768>, (int)pool_ring_smem<768>());
      SIG_LAUNCH((pool_ring_kernel<768>), ctas, TokRing<768>::kThreads, pool_ring_smem<768>(), s, src, B, n_items, mean);
    } else {
      ensure_dyn_smem(pool_ring_kernel<512>, (int)pool_ring_smem<512>());
      SIG_LAUNCH((pool_ring_kernel<512>), ctas, TokRing<512>::kThreads, pool_ring_smem<512>(), s, src, B, n_items, mean);
    }
  } else {
    const TokPtrs3 tp = tok_ptrs3(tok);
    SIG_LAUNCH((pool_tok_kernel<__nv_bfloat16>), dim3(B, 3), (unsigned)ceil_div(kPoolGroups * (d / 8), 32) * 32, (size_t)kPoolGroups * d * sizeof(float), s, tp, B, L, d, mean);
  }
  SIG_CHECK_LAUNCH();
  return 0;
}

static int align_forward_tc(const sig_tokens* tok, const sig_align_params* p, int h, int w, int do_lam, float* losses, void* ctx,
                            bool eager, bool have_mean, cudaStream_t s) {
  const int B = tok->B, L = tok->L, d = tok->d;
  AlignTcCtx c = align_tc_ctx(ctx, B, L, d);
  const TokPtrs3 tp = tok_ptrs3(tok);
  // GAM (pool -> Gram grid -> loss) is independent of LAM: run it on the side stream
  const Fork fk = do_lam ? get_fork(FORK_ALIGN_FWD) : Fork();
  cudaStream_t smain = s;
  if (fk.ok()) {
    fk.fork(smain);
    s = fk.side;
  }
  {
    SIG_PHASE("gam_fwd");
    if (have_mean) {
      // SIG_FLAG_PATCH_MEAN: the caller fills c.mean (sig_align_patch_mean_slot) with the mean pool of the patch rows --
      // the by-product of sig_tokens_fwd, or of SIM's score pass under FusionHead (sig_sim_params.pool_out), whose event
      // this stream waits for -- one pass over the tokens less
      if (p->patch_mean_event) cudaStreamWaitEvent(s, (cudaEvent_t)p->patch_mean_event, 0);
    } else {
      SIG_TRY(pool_tokens_bf16(tok, c.mean, s));
    }
    SIG_LAUNCH((gam_norm_split_kernel), B, 256, 0, s, c.mean, B, d, c.f, c.fb, c.fA, c.fB, c.nrm, c.self4);
    SIG_CHECK_LAUNCH();
    {  // lv = f_r f_n^T, la = f_r f_t^T  (split-bf16, K = 3d)
      TcGemmDesc t = tc_desc();
      t.A = tc_k2d(c.fA, B, 3 * d, 3 * d);
      t.B = tc_batched(tc_k2d(nullptr, B, 3 * d, 3 * d), c.fB, (size_t)B * 3 * d, 2);
      t.M = B; t.N = B; t.K = 3 * d; t.batch = 2;
      t.C[0] = c.lv; t.C[1] = c.la; t.ldc = B;
      SIG_TRY(tc_gemm(t, s));
    }
    SIG_LAUNCH((gam_stats_kernel), 2 * B, 128, 0, s, c.self4, c.lv, c.la, p->contra_temp, B, c.gstat, c.lossp);
    SIG_CHECK_LAUNCH();
    SIG_LAUNCH((gam_coef_kernel), 2 * B, 128, 0, s, c.self4, c.lv, c.la, p->contra_temp, c.gstat, B, (B + 7) / 8 * 8, c.Wlv, c.Wla, c.WW, c.rowA, c.colC, c.dtaup);
    SIG_CHECK_LAUNCH();
    SIG_LAUNCH((gam_final_kernel), 1, 256, 0, s, c.lossp, c.dtaup, B, losses, c.dtau);
    SIG_CHECK_LAUNCH();
  }
  s = smain;
  if (!do_lam) return 0;
  const Geo g = make_geo(h, w);
  const size_t dd = (size_t)d * d, BL = (size_t)B * L;
  {
    SIG_PHASE("lam_fold_weights");
    {
      CastJob jobs[6];
      for (int m = 0; m < 3; ++m) {
        jobs[2 * m] = {p->off0_w[m], c.W0b + m * dd, (int64_t)dd};
        jobs[2 * m + 1] = {p->proj_q_w[m], c.Wqb + m * dd, (int64_t)dd};
      }
      SIG_TRY(cast_f32_to_bf16_multi(jobs, 6, s));
    }
    SIG_LAUNCH((lam_fold_bias_kernel), dim3((unsigned)ceil_div(d, 8), 3), 256, 0, s, *p, d, c.bfold);   // b' = W0 bq + b0
    SIG_CHECK_LAUNCH();
    // W' = W0 Wq  (A = W0 [d_out, d_mid] K-major; B[n = d_in, k = d_mid] = Wq[k][n] -> MN-major)
    TcGemmDesc t = tc_desc();
    t.A = batched(tc_k2d(nullptr, d, d, d), c.W0b, dd);
    t.B = batched(tc_mn2d(nullptr, d, d, d), c.Wqb, dd);
    t.M = d; t.N = d; t.K = d; t.batch = 3;
    for (int m = 0; m < 3; ++m) t.C[m] = c.Wfb + m * dd;
    t.ldc = d; t.out_bf16 = 1;
    SIG_TRY(tc_gemm(t, s));
  }
  {
    SIG_PHASE("lam_offsetnet_fwd");
    // H = X W'^T + b'   (tokens consumed in place through a 3-D tensor map)
    TcGemmDesc t = tc_desc();
    t.A = tok_operand(tok, TC_KTOK);
    t.B = batched(tc_k2d(nullptr, d, d, d), c.Wfb, dd);
    t.M = (int)BL; t.N = d; t.K = d; t.batch = 3;
    for (int m = 0; m < 3; ++m) {
      t.C[m] = c.H + m * BL * d;
      t.bias[m] = c.bfold + (size_t)m * d;
    }
    t.ldc = d; t.out_bf16 = 1;
    t.bn = (d % 256 == 0) ? 256 : 128;
    t.mt = 1;   // (mt = 2, 256 x 256 units, measured slower: the single TMEM buffer serialises the epilogue)
    t.pair = t.bn == 256 && tc_pair_enabled();   // 256 x 256 units on CTA pairs (cta_group::2): 5 % faster
    SIG_TRY(tc_gemm(t, s));
  }
  if (dw_ring_ok(L, h, w, d)) {
    SIG_PHASE("lam_dwconv_fwd");
    if (d == 768 && w == 8) SIG_TRY((launch_dw_fwd_ring<768, 8>(c.H, *p, B, h, c.U, c.o, s)));
    else if (d == 768) SIG_TRY((launch_dw_fwd_ring<768, 16>(c.H, *p, B, h, c.U, c.o, s)));
    else if (w == 8) SIG_TRY((launch_dw_fwd_ring<512, 8>(c.H, *p, B, h, c.U, c.o, s)));
    else SIG_TRY((launch_dw_fwd_ring<512, 16>(c.H, *p, B, h, c.U, c.o, s)));
  } else {
    SIG_PHASE("lam_dwconv_fwd");
    const int pts = dw_pts_per_cta((int64_t)B * g.P, kDwFwdMaxPts);
    const dim3 grid((unsigned)ceil_div((int64_t)B * g.P, pts), 3);
    const unsigned thr = (unsigned)ceil_div(d / 2, 32) * 32;
#define SIG_DW_FWD(D, W) SIG_LAUNCH((lam_dw_fwd_tc_kernel<D, W>), grid, thr, 0, s, c.H, (int64_t)BL * d, *p, g, B, L, d, pts, c.U, c.o)
    if (d == 768 && w == 8) SIG_DW_FWD(768, 8);
    else if (d == 768 && w == 16) SIG_DW_FWD(768, 16);
    else if (d == 512 && w == 8) SIG_DW_FWD(512, 8);
    else if (d == 512 && w == 16) SIG_DW_FWD(512, 16);
    else SIG_DW_FWD(0, 0);
#undef SIG_DW_FWD
    SIG_CHECK_LAUNCH();
  }
  {
    SIG_PHASE("lam_sample_fwd");
    SIG_LAUNCH((lam_sample_fwd_tok_kernel<__nv_bfloat16>), dim3(B * g.P, 3), (unsigned)ceil_div(d / 8, 32) * 32, 0, s, tp, c.o, g, B, d, c.S);
    SIG_CHECK_LAUNCH();
    SIG_LAUNCH((lam_mse_kernel), B * g.P, 256, 0, s, c.S, (int64_t)B * g.P * d, d, c.part);
    SIG_CHECK_LAUNCH();
    SIG_LAUNCH((sum_kernel), 1, 256, 0, s, c.part, B * g.P, 1.f / (3.f * (float)B * g.P * d), losses + 1);
    SIG_CHECK_LAUNCH();
  }
  if (eager) SIG_TRY(lam_backward_chain_tc(c, tp, p, g, B, L, d, h, w, nullptr, s));
  if (fk.ok()) fk.join(smain);
  return 0;
}

static int align_backward_tc(const sig_tokens* tok, const sig_align_params* p, int h, int w, int do_lam, const float* dlosses,
                             const sig_token_grads* dtok, const sig_align_param_grads* dp, void* ctx, bool eager, cudaStream_t s) {
  const int B = tok->B, L = tok->L, d = tok->d;
  AlignTcCtx c = align_tc_ctx(ctx, B, L, d);
  const TokPtrs3 tp = tok_ptrs3(tok);
  const GradPtrs3 gp = grad_ptrs3(dtok);
  const size_t dd = (size_t)d * d, BL = (size_t)B * L;
  const Fork fkb = do_lam ? get_fork(FORK_ALIGN_BWD) : Fork();
  cudaStream_t smain = s;
  if (fkb.ok()) {
    fkb.fork(smain);
    s = fkb.side;
  }
  {
    SIG_PHASE("gam_bwd");
    const float* fr = c.f;
    const float* fn = c.f + (size_t)B * d;
    const float* ft = c.f + (size_t)2 * B * d;
    float* dfr = c.df;
    float* dfn = c.df + (size_t)B * d;
    float* dft = c.df + (size_t)2 * B * d;
    (void)fr; (void)fn; (void)ft;
    const int Bp = (B + 7) / 8 * 8;
    for (int half = 0; half < 2; ++half) {  // df_r = Wlv f_n + Wla f_t
      TcGemmDesc t = tc_desc();
      t.A = tc_k2d(c.WW + half * Bp, B, B, 2 * Bp);
      t.B = tc_mn2d(c.fb + (size_t)(1 + half) * B * d, B, d, d);
      t.M = B; t.N = d; t.K = B;
      t.C[0] = dfr; t.ldc = d; t.accumulate = half;
      SIG_TRY(tc_gemm(t, s));
    }
    {  // df_n = Wlv^T f_r, df_t = Wla^T f_r
      TcGemmDesc t = tc_desc();
      t.A = tc_mn2d(c.WW, B, B, 2 * Bp);
      t.A.ptr[1] = c.WW + Bp;
      t.B = tc_mn2d(c.fb, B, d, d);
      t.M = B; t.N = d; t.K = B; t.batch = 2;
      t.C[0] = dfn; t.C[1] = dft; t.ldc = d;
      SIG_TRY(tc_gemm(t, s));
    }
    SIG_LAUNCH((gam_finish_kernel), dim3(B, 3), 256, 0, s, c.f, c.nrm, c.rowA, c.colC, c.df, B, L, d, c.dmean);
    SIG_CHECK_LAUNCH();
    SIG_LAUNCH((scale_scalar_kernel), 1, 1, 0, s, c.dtau, dlosses, dp->contra_temp);
    SIG_CHECK_LAUNCH();
  }
  s = smain;
  if (!do_lam && dtok->wait_event) cudaStreamWaitEvent(s, (cudaEvent_t)dtok->wait_event, 0);
  if (dtok->zero_cls && !dtok->accumulate) {
    SIG_LAUNCH((zero_cls_kernel<__nv_bfloat16>), dim3(B, 3), 64, 0, s, gp, d);
    SIG_CHECK_LAUNCH();
  }
  if (!do_lam) {
    // GAM only: broadcast rows through the generic writer
    SIG_PHASE("align_write");
    const Geo g = make_geo(8, 8);
    for (int m = 0; m < 3; ++m) {
      SIG_LAUNCH((align_write_kernel<__nv_bfloat16>), (unsigned)BL, 128, 0, s, nullptr, c.dmean + (size_t)m * B * d, dlosses, nullptr, nullptr, g,
                                                                    B, L, d, static_cast<__nv_bfloat16*>(dtok->dpatch[m]),
                                                                    dtok->patch_stride_b[m], dtok->patch_stride_l[m], nullptr, 0,
                                                                    dtok->accumulate);
      SIG_CHECK_LAUNCH();
    }
    if (dtok->done_event) cudaEventRecord((cudaEvent_t)dtok->done_event, s);
    if (dp->done_event) cudaEventRecord((cudaEvent_t)dp->done_event, s);
    return 0;
  }
  const Geo g = make_geo(h, w);
  const unsigned cthreads = (unsigned)ceil_div(d / 8, 32) * 32;
  const bool dw_ring = dw_ring_ok(L, h, w, d);
  const float* lam_w = eager ? dlosses + 1 : nullptr;   // eager: the chain ran in the forward call with a unit weight
  if (!eager) SIG_TRY(lam_backward_chain_tc(c, tp, p, g, B, L, d, h, w, dlosses + 1, s));
  if (eager) {   // Wfs = d(loss)/d(lam) * W'  (B operand of the dX GEMM)
    SIG_LAUNCH((scale_to_bf16_kernel<__nv_bfloat16>), (unsigned)ceil_div((int64_t)(3 * dd) / 4, 256), 256, 0, s, c.Wfb, lam_w, c.Wfs, (int64_t)(3 * dd));
    SIG_CHECK_LAUNCH();
  }
  // Everything below that is not on the dH -> dX / dW' path runs on a second side stream: the partial-sum
  // reduction and the bias un-fold (they only need the partials), later the sparse bilinear part of
  // d(patches) (after the dX GEMM) and one of the two un-fold GEMMs.
  const Fork fk2 = get_fork(FORK_ALIGN_BWD2);
  cudaStream_t s2 = s;
  if (fk2.ok()) {
    fk2.fork(s);
    s2 = fk2.side;
  }
  {
    cudaStream_t s = s2;
    SIG_PHASE("lam_dwconv_param_grads");
    const int pts = dw_pts_per_cta((int64_t)B * g.P, 0);
    const int nchunk = dw_ring ? dw_ring_ctas_per_mod(B, h) : (int)ceil_div((int64_t)B * g.P, pts);
    SIG_LAUNCH((lam_dw_param_reduce_kernel), dim3((unsigned)ceil_div(d, 64), 19, 3), 256, 0, s, c.dwpart, nchunk, *dp, c.dbf, d, lam_w);
    SIG_CHECK_LAUNCH();
    SIG_LAUNCH((lam_unfold_bias_kernel), dim3((unsigned)ceil_div(d, 32), 3), 1024, 0, s, *p, *dp, c.dbf, d);   // db0 = db', dbq = W0^T db'
    SIG_CHECK_LAUNCH();
  }
  {
    // dW' = dH^T X  (both operands MN-major: K = all B*L positions), split-K with fp32 atomics.  First: it
    // only needs dH, and its un-fold chain then runs on the side stream under the dX GEMM.
    cudaMemsetAsync(c.dWf, 0, 3 * dd * sizeof(float), s);
    SIG_PHASE("lam_offsetnet_bwd_dw");
    TcGemmDesc t = tc_desc();
    t.A = batched(tc_mn2d(nullptr, (int64_t)BL, d, d), c.dH, BL * d);
    t.B = tok_operand(tok, TC_MNTOK);
    t.M = d; t.N = d; t.K = (int)BL; t.batch = 3;
    for (int m = 0; m < 3; ++m) t.C[m] = c.dWf + m * dd;
    t.ldc = d;
    t.bn = (d % 256 == 0) ? 256 : 128;
    t.mt = 1;
    t.pair = t.bn == 256 && tc_pair_enabled();
    const int tiles = (int)(ceil_div(d, t.pair ? 256 : 128 * t.mt) * ceil_div(d, t.bn)) * 3;
    const int workers = t.pair ? tc_num_sms() / 2 : tc_num_sms();
    int ks = (2 * workers + tiles - 1) / tiles;   // about two split-K units per CTA (pair)
    if (ks < 1) ks = 1;
    t.ksplit = ks;
    SIG_TRY(tc_gemm(t, s));
  }
  {
    if (fk2.ok()) fk2.fork(s);   // after the dW' GEMM
    cudaStream_t s = s2;
    SIG_PHASE("lam_unfold_grads");
    if (eager) {
      SIG_LAUNCH((scale_to_bf16_kernel<float>), (unsigned)ceil_div((int64_t)(3 * dd) / 4, 256), 256, 0, s, c.dWf, lam_w, c.dWfb, (int64_t)(3 * dd));
      SIG_CHECK_LAUNCH();
    } else {
      SIG_TRY(cast_f32_to_bf16(c.dWf, c.dWfb, 3 * dd, s));
    }
    {  // dWq[d_mid, d_in] = W0^T dW'  : A[m = d_mid, k = d_out] = W0[k][m] (MN-major), B[n = d_in, k = d_out] = dW'[k][n] (MN-major)
      TcGemmDesc t = tc_desc();
      t.A = batched(tc_mn2d(nullptr, d, d, d), c.W0b, dd);
      t.B = batched(tc_mn2d(nullptr, d, d, d), c.dWfb, dd);
      t.M = d; t.N = d; t.K = d; t.batch = 3;
      for (int m = 0; m < 3; ++m) t.C[m] = dp->proj_q_w[m];
      t.ldc = d;
      SIG_TRY(tc_gemm(t, s));
    }
    {  // dW0[d_out, d_mid] = dW' Wq^T : A = dW' [d_out, d_in] K-major, B = Wq [d_mid, d_in] K-major
      TcGemmDesc t = tc_desc();
      t.A = batched(tc_k2d(nullptr, d, d, d), c.dWfb, dd);
      t.B = batched(tc_k2d(nullptr, d, d, d), c.Wqb, dd);
      t.M = d; t.N = d; t.K = d; t.batch = 3;
      for (int m = 0; m < 3; ++m) t.C[m] = dp->off0_w[m];
      t.ldc = d;
      SIG_TRY(tc_gemm(t, s));
    }
    // (dbf was reduced earlier on this same side stream)
    SIG_LAUNCH((rank1_add3_kernel), dim3((unsigned)ceil_div((int64_t)dd, 256), 3), 256, 0, s, *dp, c.dbf, *p, d);
    SIG_CHECK_LAUNCH();
  }
  if (fkb.ok()) fkb.join(smain);   // the dX epilogue adds the GAM rows (dmean)
  if (dp->done_event) {
    // all parameter gradients are enqueued (LAM's on s2, contra_temp's joined into s just above): the caller's
    // exchange of this arena can run under the dX GEMM
    if (fk2.ok()) {
      cudaEventRecord(fk2.ev_join, s);         // (ev_join is re-recorded by the final join below)
      cudaStreamWaitEvent(s2, fk2.ev_join, 0);
    }
    cudaEventRecord((cudaEvent_t)dp->done_event, s2);
  }
  if (dtok->wait_event) cudaStreamWaitEvent(s, (cudaEvent_t)dtok->wait_event, 0);   // first write to the shared gradient map
  {
    SIG_PHASE("lam_offsetnet_bwd_dx");
    // d(patches) (+)= dH W' + g_gam * dmean  -> in the token dtype, at the token strides
    TcGemmDesc t = tc_desc();
    t.A = batched(tc_k2d(nullptr, (int64_t)BL, d, d), c.dH, BL * d);
    t.B = batched(tc_mn2d(nullptr, d, d, d), eager ? c.Wfs : c.Wfb, dd);   // eager: dH carries a unit weight, W' the real one
    t.M = (int)BL; t.N = d; t.K = d; t.batch = 3;
    for (int m = 0; m < 3; ++m) {
      t.C[m] = dtok->dpatch[m];
      t.rowvec[m] = c.dmean + (size_t)m * B * d;
    }
    t.rowvec_scale = dlosses;
    t.out_bf16 = 1; t.c_tok = 1;
    t.c_stride_b = dtok->patch_stride_b[0]; t.c_stride_l = dtok->patch_stride_l[0];
    t.accumulate = dtok->accumulate;
    if (dtok->fuse_pds && dtok->fuse_dxqt) {   // + SIM's token gradient as extra k-block(s) (sig_token_grads.fuse_*)
      t.xa = dtok->fuse_pds; t.xb = dtok->fuse_dxqt; t.xB = B;
    }
    t.bn = (d % 256 == 0) ? 256 : 128;
    t.mt = 1;   // (mt = 2, 256 x 256 units, measured slower: the single TMEM buffer serialises the epilogue)
    t.pair = t.bn == 256 && tc_pair_enabled() && (!t.xa || (B % 2) == 0);   // (a pair covers two samples)
    // the tail of the step: nothing of SIM's chain is left to make room for -- all SMs, one persistent wave
    const ScopedSmBudget all_sms(1 << 20);
    const ScopedSmWaves one_wave(1);
    SIG_TRY(tc_gemm(t, s));
  }
  {
    SIG_PHASE("lam_sparse_dx");
    SIG_LAUNCH((lam_sparse_add_kernel<__nv_bfloat16>), dim3(B, 3), cthreads, 0, s, gp, c.o, c.dS, g, B, d, lam_w);
    SIG_CHECK_LAUNCH();
    if (dtok->done_event) cudaEventRecord((cudaEvent_t)dtok->done_event, s);
  }
  if (fk2.ok()) fk2.join(s);
  return 0;
}

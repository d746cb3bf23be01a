// AlignmentM -- GAM (useB.py:76-126, utils/volume.py:14-62) and LAM (useB.py:128-167,
// DAS.py:107-165), fp32 SIMT path.  Formulas: SURVEY.md Appendix B4 (GAM closed form and its
// gradient) and B5 (deformable bilinear sampling and its grid gradient).
#include "align.h"
#include "common.cuh"
#include "sim.h"
#include "simt_ops.cuh"
#include "stream_ring.cuh"
#include "tok_ring.cuh"
#include "tc_gemm.h"

namespace sig {

struct TokPtrs3 {
  const void* patch[3];
  int64_t psb[3], psl[3];
};
struct GradPtrs3 {
  void* dpatch[3];
  void* dcls[3];
  int64_t psb[3], psl[3], csb[3];
};

// =============================================================================================
// GAM
// =============================================================================================
// mean over the L tokens: grid (B, 3), thread per channel.  useB.py:92-94
static __global__ void __launch_bounds__(256) pool_kernel(const float* __restrict__ Xf, int B, int L, int d, float* __restrict__ mean,
                                                          double* __restrict__ meand) {
  pdl_enter();
  const int b = blockIdx.x, m = blockIdx.y;
  const float* x = Xf + ((int64_t)m * B + b) * L * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    double a = 0.0;      // exact path: the sum of L fp32 values in fp64 is exact to the last bit of the result
    for (int l = 0; l < L; ++l) a += (double)x[(int64_t)l * d + c];
    const double mu = a / L;
    mean[((int64_t)m * B + b) * d + c] = (float)mu;
    if (meand) meand[((int64_t)m * B + b) * d + c] = mu;
  }
}

// F.normalize (useB.py:98-100) for the three modalities of sample b, then the per-sample
// Gram entries ll, vv, aa, va (volume.py:35,42-44).  grid B.
// block-wide sum in double (the exact path evaluates the Gram volume in fp64: the determinant cancels when the
// modalities align and d(loss)/d(tau) = sum_ij dZ_ij V_ij / tau^2 cancels to ~1e-3 of its terms on iid tokens)
__device__ __forceinline__ double block_sum_f64(double v, double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < nw ? scratch[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

static __global__ void __launch_bounds__(256) gam_norm_kernel(const float* __restrict__ mean, int B, int d, float* __restrict__ f,
                                                              float* __restrict__ nrm, float* __restrict__ self4,
                                                              const double* __restrict__ meand, double* __restrict__ fd,
                                                              double* __restrict__ self4d) {
  pdl_enter();
  __shared__ float scratch[33];
  __shared__ double scratchd[33];
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
    const float r = mean[((int64_t)0 * B + b) * d + c] * inv[0];
    const float n = mean[((int64_t)1 * B + b) * d + c] * inv[1];
    const float t = mean[((int64_t)2 * B + b) * d + c] * inv[2];
    f[((int64_t)0 * B + b) * d + c] = r;
    f[((int64_t)1 * B + b) * d + c] = n;
    f[((int64_t)2 * B + b) * d + c] = t;
    ll += r * r; vv += n * n; aa += t * t; va += n * t;
  }
  ll = block_sum(ll, scratch); vv = block_sum(vv, scratch);
  aa = block_sum(aa, scratch); va = block_sum(va, scratch);
  if (threadIdx.x == 0) {
    self4[0 * B + b] = ll; self4[1 * B + b] = vv; self4[2 * B + b] = aa; self4[3 * B + b] = va;
  }
  if (self4d) {   // exact path: the normalised features and their Gram entries once more in fp64, from the fp64 means
    double dinv[3];
    for (int m = 0; m < 3; ++m) {
      double sq = 0.0;
      for (int c = threadIdx.x; c < d; c += blockDim.x) { const double x = meand[((int64_t)m * B + b) * d + c]; sq += x * x; }
      dinv[m] = 1.0 / fmax(sqrt(block_sum_f64(sq, scratchd)), 1e-12);
    }
    double dll = 0.0, dvv = 0.0, daa = 0.0, dva = 0.0;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      const double r = meand[((int64_t)0 * B + b) * d + c] * dinv[0], n = meand[((int64_t)1 * B + b) * d + c] * dinv[1],
                   t = meand[((int64_t)2 * B + b) * d + c] * dinv[2];
      fd[((int64_t)0 * B + b) * d + c] = r; fd[((int64_t)1 * B + b) * d + c] = n; fd[((int64_t)2 * B + b) * d + c] = t;
      dll += r * r; dvv += n * n; daa += t * t; dva += n * t;
    }
    dll = block_sum_f64(dll, scratchd); dvv = block_sum_f64(dvv, scratchd);
    daa = block_sum_f64(daa, scratchd); dva = block_sum_f64(dva, scratchd);
    if (threadIdx.x == 0) {
      self4d[0 * B + b] = dll; self4d[1 * B + b] = dvv; self4d[2 * B + b] = daa; self4d[3 * B + b] = dva;
    }
  }
}

// lv[i,j] = f_r[i] . f_n[j], la[i,j] = f_r[i] . f_t[j] accumulated in fp64 (exact path).  grid (ceil(B/32), B), 1024 threads:
// warp w of CTA (jb, i) owns the pair (i, 32 jb + w).
static __global__ void __launch_bounds__(1024) gam_gram_f64_kernel(const double* __restrict__ f, int B, int d, double* __restrict__ lvd,
                                                                   double* __restrict__ lad) {
  pdl_enter();
  const int i = blockIdx.y, j = blockIdx.x * 32 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= B) return;
  const double* fr = f + (int64_t)i * d;
  const double* fn = f + ((int64_t)B + j) * d;
  const double* ft = f + ((int64_t)2 * B + j) * d;
  double a = 0.0, b = 0.0;
  for (int c = lane; c < d; c += 32) {
    const double r = fr[c];
    a += r * fn[c];
    b += r * ft[c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    lvd[(int64_t)i * B + j] = a;
    lad[(int64_t)i * B + j] = b;
  }
}

__device__ __forceinline__ float gram_det(float ll, float vv, float aa, float va, float lv, float la) {
  return ll * (vv * aa - va * va) - lv * (lv * aa - va * la) + la * (lv * va - vv * la);
}

// volume grid V[i,j] = sqrt|det|  (volume.py:57-60).  One thread per pair.
static __global__ void volume_kernel(const float* __restrict__ ll, const float* __restrict__ vv, const float* __restrict__ aa,
                                     const float* __restrict__ va, const float* __restrict__ lv, const float* __restrict__ la,
                                     int B1, int B2, float* __restrict__ V) {
  pdl_enter();
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B1 * B2) return;
  const int i = (int)(idx / B2), j = (int)(idx % B2);
  V[idx] = sqrtf(fabsf(gram_det(ll[i], vv[j], aa[j], va[j], lv[idx], la[idx])));
}

// d det for pair (i,j) given dV
__device__ __forceinline__ float ddet_of(float dV, float det, float V) {
  // d sqrt|det| = sign(det) / (2 sqrt|det|); the reference yields NaN at det == 0, we give 0
  if (V <= 0.f) return 0.f;
  return dV * (det > 0.f ? 0.5f : -0.5f) / V;
}

__device__ __forceinline__ double gram_det_f64(double ll, double vv, double aa, double va, double lv, double la) {
  return ll * (vv * aa - va * va) - lv * (lv * aa - va * la) + la * (lv * va - vv * la);
}

// Single CTA (1024 threads): symmetric label-smoothed CE on Z = -V/tau (useB.py:107-124) and
// d(loss)/d(everything the Gram entries depend on), for unit upstream gradient:
//   Wlv[i,j] = ddet*c_lv, Wla[i,j] = ddet*c_la, rowA[i] = sum_j ddet*c_ll,
//   colC[0..2][j] = sum_i ddet*{c_vv, c_va, c_aa},  dtau = sum_ij dZ_ij V_ij / tau^2.
// Exact path: the Gram entries arrive in fp64 (gam_norm_kernel, gam_gram_f64_kernel) and the determinant, the volume,
// the softmax statistics and the sums are evaluated in fp64 (B x B = 16 K pairs: nothing); results are stored in fp32.
static __global__ void __launch_bounds__(1024) gam_loss_kernel(const double* __restrict__ self4, const double* __restrict__ lv,
                                                               const double* __restrict__ la, const float* __restrict__ tau_p, int B,
                                                               double* __restrict__ V, double* __restrict__ rowstat /*[B]*/,
                                                               double* __restrict__ colstat /*[B]*/, float* __restrict__ Vf,
                                                               float* __restrict__ Wlv, float* __restrict__ Wla,
                                                               float* __restrict__ rowA, float* __restrict__ colC,
                                                               float* __restrict__ loss_out, float* __restrict__ dtau_out) {
  pdl_enter();
  __shared__ double scratch[33];
  const double* ll = self4;
  const double* vv = self4 + B;
  const double* aa = self4 + 2 * B;
  const double* va = self4 + 3 * B;
  const double tau = (double)*tau_p, itau = 1.0 / tau;
  const double ls = (double)kLabelSmooth;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  const int64_t n2 = (int64_t)B * B;
  for (int64_t idx = tid; idx < n2; idx += blockDim.x) {
    const int i = (int)(idx / B), j = (int)(idx % B);
    const double v = sqrt(fabs(gram_det_f64(ll[i], vv[j], aa[j], va[j], lv[idx], la[idx])));
    V[idx] = v;
    if (Vf) Vf[idx] = (float)v;
  }
  __syncthreads();
  auto wsum = [](double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
  };
  auto wmax = [](double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
  };
  // log-sum-exp and mean of Z along rows (d2a) and columns (a2d)
  double lpart = 0.0;
  for (int r = w; r < 2 * B; r += nw) {
    const bool is_row = r < B;
    const int i = is_row ? r : r - B;
    double mx = -INFINITY;
    for (int j = lane; j < B; j += 32) mx = fmax(mx, -(is_row ? V[(int64_t)i * B + j] : V[(int64_t)j * B + i]) * itau);
    mx = wmax(mx);
    double se = 0.0, sz = 0.0;
    for (int j = lane; j < B; j += 32) {
      const double z = -(is_row ? V[(int64_t)i * B + j] : V[(int64_t)j * B + i]) * itau;
      se += exp(z - mx);
      sz += z;
    }
    se = wsum(se);
    sz = wsum(sz);
    const double lse = mx + log(se);
    if (lane == 0) {
      (is_row ? rowstat : colstat)[i] = lse;
      const double zii = -V[(int64_t)i * B + i] * itau;
      lpart += (1.0 - ls) * (lse - zii) + ls * (lse - sz / B);
    }
  }
  const double loss = block_sum_f64(lpart, scratch) * (0.5 / B);
  if (tid == 0) *loss_out = (float)loss;
  __syncthreads();
  // row sweep: Wlv, Wla, rowA, dtau
  double tpart = 0.0;
  const double tgt_off = ls / B, tgt_on = 1.0 - ls + ls / B;
  auto ddet_of64 = [](double dV, double det, double v) { return v > 0.0 ? dV * (det > 0.0 ? 0.5 : -0.5) / v : 0.0; };
  for (int i = w; i < B; i += nw) {
    double ra = 0.0;
    for (int j = lane; j < B; j += 32) {
      const int64_t idx = (int64_t)i * B + j;
      const double v = V[idx], z = -v * itau;
      const double tg = i == j ? tgt_on : tgt_off;
      const double pr = exp(z - rowstat[i]) - tg, pc = exp(z - colstat[j]) - tg;
      const double dZ = (0.5 / B) * (pr + pc);
      // sum_j (P_row - T)_ij = 0 and sum_i (P_col - T)_ij = 0: subtracting the diagonal volume changes nothing in
      // exact arithmetic and removes most of the cancellation
      tpart += (0.5 / B) * (pr * (v - V[(int64_t)i * B + i]) + pc * (v - V[(int64_t)j * B + j])) * itau * itau;
      const double det = gram_det_f64(ll[i], vv[j], aa[j], va[j], lv[idx], la[idx]);
      const double dd = ddet_of64(-dZ * itau, det, v);
      Wlv[idx] = (float)(dd * (-2.0 * (lv[idx] * aa[j] - va[j] * la[idx])));
      Wla[idx] = (float)(dd * (2.0 * (lv[idx] * va[j] - vv[j] * la[idx])));
      ra += dd * (vv[j] * aa[j] - va[j] * va[j]);
    }
    ra = wsum(ra);
    if (lane == 0) rowA[i] = (float)ra;
  }
  const double dtau = block_sum_f64(tpart, scratch);
  if (tid == 0) *dtau_out = (float)dtau;
  // column sweep: colC
  for (int j = w; j < B; j += nw) {
    double cvv = 0.0, cva = 0.0, caa = 0.0;
    for (int i = lane; i < B; i += 32) {
      const int64_t idx = (int64_t)i * B + j;
      const double v = V[idx], z = -v * itau;
      const double tg = i == j ? tgt_on : tgt_off;
      const double dZ = (0.5 / B) * ((exp(z - rowstat[i]) - tg) + (exp(z - colstat[j]) - tg));
      const double det = gram_det_f64(ll[i], vv[j], aa[j], va[j], lv[idx], la[idx]);
      const double dd = ddet_of64(-dZ * itau, det, v);
      cvv += dd * (ll[i] * aa[j] - la[idx] * la[idx]);
      cva += dd * (-2.0 * (ll[i] * va[j] - lv[idx] * la[idx]));
      caa += dd * (ll[i] * vv[j] - lv[idx] * lv[idx]);
    }
    cvv = wsum(cvv); cva = wsum(cva); caa = wsum(caa);
    if (lane == 0) {
      colC[0 * B + j] = (float)cvv; colC[1 * B + j] = (float)cva; colC[2 * B + j] = (float)caa;
    }
  }
}

// add the diagonal Gram terms to df, then go back through F.normalize and the mean pool:
//   df_r += 2 rowA f_r ; df_n += 2 Cvv f_n + Cva f_t ; df_t += 2 Caa f_t + Cva f_n
//   dm = (df - (df.f) f) / max(|m|, eps) ; dmean = dm / L   (row added to every token's gradient)
// grid (B, 3)
static __global__ void __launch_bounds__(256) gam_finish_kernel(const float* __restrict__ f, const float* __restrict__ nrm,
                                                                const float* __restrict__ rowA, const float* __restrict__ colC,
                                                                const float* __restrict__ df, int B, int L, int d,
                                                                float* __restrict__ dmean) {
  pdl_enter();
  __shared__ float scratch[33];
  const int b = blockIdx.x, m = blockIdx.y;
  const float* fr = f + ((int64_t)0 * B + b) * d;
  const float* fn = f + ((int64_t)1 * B + b) * d;
  const float* ft = f + ((int64_t)2 * B + b) * d;
  const float* fm = f + ((int64_t)m * B + b) * d;
  const float* g = df + ((int64_t)m * B + b) * d;
  const float a = rowA[b], cvv = colC[b], cva = colC[B + b], caa = colC[2 * B + b];
  float dot = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float v = g[c];
    if (m == 0) v += 2.f * a * fr[c];
    else if (m == 1) v += 2.f * cvv * fn[c] + cva * ft[c];
    else v += 2.f * caa * ft[c] + cva * fn[c];
    dot += v * fm[c];
  }
  dot = block_sum(dot, scratch);
  const float s = 1.f / (nrm[m * B + b] * L);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float v = g[c];
    if (m == 0) v += 2.f * a * fr[c];
    else if (m == 1) v += 2.f * cvv * fn[c] + cva * ft[c];
    else v += 2.f * caa * ft[c] + cva * fn[c];
    dmean[((int64_t)m * B + b) * d + c] = (v - dot * fm[c]) * s;
  }
}

// =============================================================================================
// LAM
// =============================================================================================
struct Geo {
  int h, w, Hk, Wk, P;
};
__host__ __device__ inline Geo make_geo(int h, int w) {
  Geo g;
  g.h = h; g.w = w; g.Hk = h / 4; g.Wk = w / 4; g.P = g.Hk * g.Wk;
  return g;
}

// depthwise 4x4 stride-4 conv + GELU + 1x1 -> one offset logit per sample point (DAS.py:60-65).
// grid (B*P), thread per channel.  G = gelu(H) [B*L, d]; U saved for backward.
static __global__ void __launch_bounds__(256) lam_dw_fwd_kernel(const float* __restrict__ G, const float* __restrict__ wdw,
                                                                const float* __restrict__ bdw, const float* __restrict__ w4, Geo g,
                                                                int L, int d, float* __restrict__ U, float* __restrict__ o) {
  pdl_enter();
  __shared__ float scratch[33];
  const int bp = blockIdx.x, b = bp / g.P, p = bp % g.P;
  const int py = p / g.Wk, px = p % g.Wk;
  float part = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float u = bdw[c];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int l = (4 * py + (k >> 2)) * g.w + 4 * px + (k & 3);
      u = fmaf(G[((int64_t)b * L + l) * d + c], wdw[c * 16 + k], u);
    }
    U[(int64_t)bp * d + c] = u;
    part += gelu_f(u) * w4[c];
  }
  part = block_sum(part, scratch);
  if (threadIdx.x == 0) o[bp] = part;
}

// sample position and bilinear taps of point p given its offset logit (DAS.py:140-163)
struct Taps {
  int l[4];     // token index of the 4 taps (or -1 when out of bounds)
  float wgt[4]; // bilinear weights
  float wy, wx; // fractional parts
  float gy, gx; // d raw / d o  chain factors (0 when the clamp is active)
};
__device__ inline Taps make_taps(float o, int p, const Geo& g) {
  Taps t;
  const int py = p / g.Wk, px = p % g.Wk;
  const float th = tanhf(o);
  const float ry = 2.f / (g.Hk - 1.f), rx = 2.f / (g.Wk - 1.f);  // offset_range * factor(2)
  const float ref_y = (py + 0.5f) / (g.Hk - 1.f) * 2.f - 1.f;
  const float ref_x = (px + 0.5f) / (g.Wk - 1.f) * 2.f - 1.f;
  const float raw_y = th * ry + ref_y, raw_x = th * rx + ref_x;
  const float cy = fminf(fmaxf(raw_y, -1.f), 1.f), cx = fminf(fmaxf(raw_x, -1.f), 1.f);
  const float dth = 1.f - th * th;
  t.gy = (raw_y >= -1.f && raw_y <= 1.f) ? ry * dth * 0.5f * (g.h - 1) : 0.f;
  t.gx = (raw_x >= -1.f && raw_x <= 1.f) ? rx * dth * 0.5f * (g.w - 1) : 0.f;
  const float iy = (cy + 1.f) * 0.5f * (g.h - 1), ix = (cx + 1.f) * 0.5f * (g.w - 1);
  const float y0f = floorf(iy), x0f = floorf(ix);
  const int y0 = (int)y0f, x0 = (int)x0f;
  t.wy = iy - y0f;
  t.wx = ix - x0f;
  const float wy1 = t.wy, wy0 = 1.f - t.wy, wx1 = t.wx, wx0 = 1.f - t.wx;
  const int ys[4] = {y0, y0, y0 + 1, y0 + 1}, xs[4] = {x0, x0 + 1, x0, x0 + 1};
  const float ws[4] = {wy0 * wx0, wy0 * wx1, wy1 * wx0, wy1 * wx1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool inb = ys[k] >= 0 && ys[k] < g.h && xs[k] >= 0 && xs[k] < g.w;
    t.l[k] = inb ? ys[k] * g.w + xs[k] : -1;
    t.wgt[k] = ws[k];
  }
  return t;
}

// S[b,p,:] = bilinear sample of X[b] at the predicted position.  grid (B*P)
static __global__ void __launch_bounds__(256) lam_sample_fwd_kernel(const float* __restrict__ X, const float* __restrict__ o, Geo g,
                                                                    int L, int d, float* __restrict__ S) {
  pdl_enter();
  const int bp = blockIdx.x, b = bp / g.P, p = bp % g.P;
  const Taps t = make_taps(o[bp], p, g);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (t.l[k] >= 0) v = fmaf(t.wgt[k], X[((int64_t)b * L + t.l[k]) * d + c], v);
    S[(int64_t)bp * d + c] = v;
  }
}

// pairwise MSE partial sums (useB.py:161-165): grid (B*P), part[bp] = sum_c [(n-r)^2+(t-r)^2+(t-n)^2]
static __global__ void __launch_bounds__(256) lam_mse_kernel(const float* __restrict__ S, int64_t mstride, int d, float* __restrict__ part) {
  pdl_enter();
  __shared__ float scratch[33];
  const int64_t bp = blockIdx.x;
  float a = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float r = S[bp * d + c], n = S[mstride + bp * d + c], t = S[2 * mstride + bp * d + c];
    a += (n - r) * (n - r) + (t - r) * (t - r) + (t - n) * (t - n);
  }
  a = block_sum(a, scratch);
  if (threadIdx.x == 0) part[bp] = a;
}

// out[0] = scale * sum(part[0..n))   (single CTA, deterministic)
static __global__ void __launch_bounds__(256) sum_kernel(const float* __restrict__ part, int n, float scale, float* __restrict__ out) {
  pdl_enter();
  __shared__ float scratch[33];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += part[i];
  a = block_sum(a, scratch);
  if (threadIdx.x == 0) *out = a * scale;
}

// dS_m = g * 2/(3N) * (2 S_m - S_m' - S_m'')   (N = B*P*d), all three modalities.  grid (B*P)
static __global__ void __launch_bounds__(256) lam_mse_bwd_kernel(const float* __restrict__ S, int64_t mstride, int d, float scale,
                                                                 const float* __restrict__ gptr, float* __restrict__ dS) {
  pdl_enter();
  const int64_t bp = blockIdx.x;
  const float k = gptr ? scale * (*gptr) : scale;   // (NULL: unit loss weight)
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float r = S[bp * d + c], n = S[mstride + bp * d + c], t = S[2 * mstride + bp * d + c];
    dS[bp * d + c] = k * (2.f * r - n - t);
    dS[mstride + bp * d + c] = k * (2.f * n - r - t);
    dS[2 * mstride + bp * d + c] = k * (2.f * t - r - n);
  }
}

// grid gradient -> d(offset logit): grid (B*P).  GridSampler backward restricted to in-bounds taps.
static __global__ void __launch_bounds__(256) lam_sample_bwd_kernel(const float* __restrict__ X, const float* __restrict__ o,
                                                                    const float* __restrict__ dS, Geo g, int L, int d,
                                                                    float* __restrict__ dout_o) {
  pdl_enter();
  __shared__ float scratch[33];
  const int bp = blockIdx.x, b = bp / g.P, p = bp % g.P;
  const Taps t = make_taps(o[bp], p, g);
  float giy = 0.f, gix = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = t.l[k] >= 0 ? X[((int64_t)b * L + t.l[k]) * d + c] : 0.f;
    const float ds = dS[(int64_t)bp * d + c];
    giy += ds * ((v[2] - v[0]) * (1.f - t.wx) + (v[3] - v[1]) * t.wx);
    gix += ds * ((v[1] - v[0]) * (1.f - t.wy) + (v[3] - v[2]) * t.wy);
  }
  giy = block_sum(giy, scratch);
  gix = block_sum(gix, scratch);
  if (threadIdx.x == 0) dout_o[bp] = giy * t.gy + gix * t.gx;
}

// dU = do * w4 * gelu'(U); dH[b,pos,c] = dU * wdw[c,k] * gelu'(H).  grid (B*P), thread per channel
static __global__ void __launch_bounds__(256) lam_dw_bwd_kernel(const float* __restrict__ H, const float* __restrict__ U,
                                                                const float* __restrict__ dout_o, const float* __restrict__ wdw,
                                                                const float* __restrict__ w4, Geo g, int L, int d,
                                                                float* __restrict__ dU, float* __restrict__ dH) {
  pdl_enter();
  const int bp = blockIdx.x, b = bp / g.P, p = bp % g.P;
  const int py = p / g.Wk, px = p % g.Wk;
  const float go = dout_o[bp];
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float du = go * w4[c] * gelu_grad_f(U[(int64_t)bp * d + c]);
    dU[(int64_t)bp * d + c] = du;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int l = (4 * py + (k >> 2)) * g.w + 4 * px + (k & 3);
      const int64_t idx = ((int64_t)b * L + l) * d + c;
      dH[idx] = du * wdw[c * 16 + k] * gelu_grad_f(H[idx]);
    }
  }
}

// parameter gradients of the depthwise conv and the 1x1 -> 1 conv (deterministic):
// grid (ceil(d/32)), 32 channels x 8 row lanes over (b,p).
static __global__ void __launch_bounds__(256) lam_dw_param_kernel(const float* __restrict__ Gact, const float* __restrict__ U,
                                                                  const float* __restrict__ dU, const float* __restrict__ dout_o, Geo g,
                                                                  int B, int L, int d, float* __restrict__ dwdw,
                                                                  float* __restrict__ dbdw, float* __restrict__ dw4) {
  pdl_enter();
  __shared__ float sm[8][32][19];
  const int cl = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float acc[18];
#pragma unroll
  for (int i = 0; i < 18; ++i) acc[i] = 0.f;
  if (c < d) {
    for (int bp = r; bp < B * g.P; bp += 8) {
      const int b = bp / g.P, p = bp % g.P;
      const int py = p / g.Wk, px = p % g.Wk;
      const float du = dU[(int64_t)bp * d + c];
      acc[16] += du;
      acc[17] += dout_o[bp] * gelu_f(U[(int64_t)bp * d + c]);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int l = (4 * py + (k >> 2)) * g.w + 4 * px + (k & 3);
        acc[k] = fmaf(du, Gact[((int64_t)b * L + l) * d + c], acc[k]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 18; ++i) sm[r][cl][i] = acc[i];
  __syncthreads();
  if (r == 0 && c < d) {
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += sm[q][cl][i];
      if (i < 16) dwdw[c * 16 + i] = t;
      else if (i == 16) dbdw[c] = t;
      else dw4[c] = t;
    }
  }
}

// Final writer for one modality: dx[b,l,:] = dense[b,l,:] + gscale * gam_row[b,:] + sum over the
// (<= 4P) bilinear taps that hit token l of w_tap * dS[b,p,:].  One writer, no atomics (gather form).
// grid (B*L); extra rows [B*L, B*L+B) zero the CLS gradient rows when requested.
template <typename T>
static __global__ void __launch_bounds__(128) align_write_kernel(const float* __restrict__ dense, const float* __restrict__ gam_row,
                                                                 const float* __restrict__ gam_g, const float* __restrict__ o,
                                                                 const float* __restrict__ dS, Geo g, int B, int L, int d, T* dpatch,
                                                                 int64_t psb, int64_t psl, T* dcls, int64_t csb, int accumulate) {
  pdl_enter();
  __shared__ float tw[64];
  __shared__ int tp[64];
  __shared__ int ntap;
  const int64_t row = blockIdx.x;
  if (row >= (int64_t)B * L) {
    if (!dcls || accumulate) return;
    const int b = (int)(row - (int64_t)B * L);
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) store8(dcls + b * csb + c, z);
    return;
  }
  const int b = (int)(row / L), l = (int)(row % L);
  if (threadIdx.x == 0) {
    int n = 0;
    if (o && dS) {
      for (int p = 0; p < g.P; ++p) {
        const Taps t = make_taps(o[b * g.P + p], p, g);
        for (int k = 0; k < 4; ++k)
          if (t.l[k] == l && n < 64) {
            tw[n] = t.wgt[k];
            tp[n] = p;
            ++n;
          }
      }
    }
    ntap = n;
  }
  __syncthreads();
  const float gs = gam_row ? *gam_g : 0.f;
  T* dst = dpatch + b * psb + l * psl;
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (dense) load8(dense + row * d + c, v);
    if (gam_row) {
      float r[8];
      load8(gam_row + (int64_t)b * d + c, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(gs, r[i], v[i]);
    }
    for (int n = 0; n < ntap; ++n) {
      float s[8];
      load8(dS + ((int64_t)b * g.P + tp[n]) * d + c, s);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(tw[n], s[i], v[i]);
    }
    if (accumulate) {
      float old[8];
      load8(dst + c, old);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += old[i];
    }
    store8(dst + c, v);
  }
}

// d(contra_temp) = g_gam * dtau_unit
static __global__ void scale_scalar_kernel(const float* a, const float* g, float* out) {
  pdl_enter(); *out = (*a) * (*g); }

// =============================================================================================
// ctx layouts
// =============================================================================================
struct LamMod {   // per-modality LAM buffers
  float *Q, *H, *G, *U, *o, *dO, *dU;
};
struct AlignCtx {
  float* Xf;
  // GAM
  float *mean, *f, *nrm, *self4, *lv, *la, *V, *rowstat, *colstat, *Wlv, *Wla, *rowA, *colC, *dtau, *df, *dmean;
  double *self4d, *lvd, *lad, *Vd, *statd, *meand, *fd;   // fp64 evaluation of the Gram volume (exact path)
  // LAM
  LamMod mod[3];
  float *S, *dS, *part, *dH, *dQ, *dXf;
  size_t bytes;
};

static AlignCtx align_ctx(void* base, int B, int L, int d, int nmod) {
  Arena a(base);
  AlignCtx c;
  const size_t BL = (size_t)B * L, P = (size_t)(L / 16);
  c.Xf = a.take<float>((size_t)nmod * BL * d);
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
  c.self4d = a.take<double>((size_t)4 * B);
  c.lvd = a.take<double>((size_t)B * B);
  c.lad = a.take<double>((size_t)B * B);
  c.Vd = a.take<double>((size_t)B * B);
  c.statd = a.take<double>((size_t)2 * B);
  c.meand = a.take<double>((size_t)3 * B * d);
  c.fd = a.take<double>((size_t)3 * B * d);
  for (int m = 0; m < nmod; ++m) {
    c.mod[m].Q = a.take<float>(BL * d);
    c.mod[m].H = a.take<float>(BL * d);
    c.mod[m].G = a.take<float>(BL * d);
    c.mod[m].U = a.take<float>((size_t)B * P * d);
    c.mod[m].o = a.take<float>((size_t)B * P);
    c.mod[m].dO = a.take<float>((size_t)B * P);
    c.mod[m].dU = a.take<float>((size_t)B * P * d);
  }
  c.S = a.take<float>((size_t)nmod * B * P * d);
  c.dS = a.take<float>((size_t)nmod * B * P * d);
  c.part = a.take<float>((size_t)B * P);
  c.dH = a.take<float>(BL * d);
  c.dQ = a.take<float>(BL * d);
  c.dXf = a.take<float>((size_t)nmod * BL * d);
  c.bytes = a.off;
  return c;
}

size_t align_ctx_bytes(int B, int L, int d) { return align_ctx(nullptr, B, L, d, 3).bytes; }
size_t das_ctx_bytes(int B, int L, int d) { return align_ctx(nullptr, B, L, d, 1).bytes; }

#include "align_tc.inl"

// GAM mean pool for a caller outside this file (SIM's score pass delivers it under FusionHead; this is its fall-back for
// token layouts the ring kernel does not take)
int align_pool_tokens(const sig_tokens* tok, float* mean, cudaStream_t s) {
  if (tok->dtype != SIG_BF16) return SIG_ERR_DTYPE;
  return pool_tokens_bf16(tok, mean, s);
}

// SIG_FLAG_PATCH_MEAN: where the caller deposits the [3][B][d] fp32 patch means (tensor-core path only; the exact fp32
// path pools in fp64)
int align_patch_mean_slot(void* ctx, int B, int L, int d, int dtype, unsigned flags, float** slot) {
  if (!ctx || !slot) return SIG_ERR_NULL;
  if (!tc_shape_ok(dtype, L, d, flags)) return SIG_ERR_SHAPE;
  *slot = align_tc_ctx(ctx, B, L, d).mean;
  return 0;
}

size_t align_ctx_bytes_for(int B, int L, int d, int dtype, unsigned flags) {
  if (tc_shape_ok(dtype, L, d, flags)) return align_tc_ctx(nullptr, B, L, d).bytes;
  return align_ctx_bytes(B, L, d);
}

// =============================================================================================
// orchestration
// =============================================================================================
static int lam_offsets_fwd(const float* X, const sig_align_params* p, int m, const LamMod& lm, const Geo& g, int B, int L, int d,
                           cudaStream_t s) {
  const int BL = B * L;
  SIG_PHASE("lam_offsetnet_fwd");
  // q = proj_q(x) (DAS.py:129); H = conv_offset[0](q) (DAS.py:58); G = GELU(H)
  SIG_TRY(launch_gemm(gemm_nt(X, d, p->proj_q_w[m], d, lm.Q, d, p->proj_q_b[m], BL, d, d), s));
  {
    Gemm gg = gemm_nt(lm.Q, d, p->off0_w[m], d, lm.G, d, p->off0_b[m], BL, d, d);
    gg.act = 1; gg.pre = lm.H;
    SIG_TRY(launch_gemm(gg, s));
  }
  SIG_LAUNCH((lam_dw_fwd_kernel), B * g.P, 256, 0, s, lm.G, p->off2_w[m], p->off2_b[m], p->off4_w[m], g, L, d, lm.U, lm.o);
  SIG_CHECK_LAUNCH();
  return 0;
}

// given dO (gradient of the offset logits) produce the dense dX and all parameter gradients of modality m
static int lam_offsets_bwd(const float* X, const sig_align_params* p, const sig_align_param_grads* dp, int m, const LamMod& lm,
                           const AlignCtx& c, float* dXdense, const Geo& g, int B, int L, int d, cudaStream_t s) {
  const int BL = B * L;
  SIG_PHASE("lam_offsetnet_bwd");
  SIG_LAUNCH((lam_dw_bwd_kernel), B * g.P, 256, 0, s, lm.H, lm.U, lm.dO, p->off2_w[m], p->off4_w[m], g, L, d, lm.dU, c.dH);
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((lam_dw_param_kernel), (unsigned)ceil_div(d, 32), 256, 0, s, lm.G, lm.U, lm.dU, lm.dO, g, B, L, d, dp->off2_w[m], dp->off2_b[m],
                                                                dp->off4_w[m]);
  SIG_CHECK_LAUNCH();
  SIG_TRY(launch_colsum(c.dH, d, BL, d, dp->off0_b[m], 1.f, s));
  {  // dW0 = dH^T Q
    cudaMemsetAsync(dp->off0_w[m], 0, (size_t)d * d * sizeof(float), s);
    Gemm gg = gemm_tn(c.dH, d, lm.Q, d, dp->off0_w[m], d, d, d, BL);
    gg.ksplit = 8;
    SIG_TRY(launch_gemm(gg, s));
  }
  SIG_TRY(launch_gemm(gemm_nn(c.dH, d, p->off0_w[m], d, c.dQ, d, BL, d, d), s));
  SIG_TRY(launch_colsum(c.dQ, d, BL, d, dp->proj_q_b[m], 1.f, s));
  {  // dWq = dQ^T X
    cudaMemsetAsync(dp->proj_q_w[m], 0, (size_t)d * d * sizeof(float), s);
    Gemm gg = gemm_tn(c.dQ, d, X, d, dp->proj_q_w[m], d, d, d, BL);
    gg.ksplit = 8;
    SIG_TRY(launch_gemm(gg, s));
  }
  SIG_TRY(launch_gemm(gemm_nn(c.dQ, d, p->proj_q_w[m], d, dXdense, d, BL, d, d), s));
  return 0;
}

static int check_align_params(const sig_align_params* p, bool lam, int m0, int m1) {
  if (!p) return SIG_ERR_NULL;
  if (lam)
    for (int m = m0; m < m1; ++m)
      if (!p->proj_q_w[m] || !p->proj_q_b[m] || !p->off0_w[m] || !p->off0_b[m] || !p->off2_w[m] || !p->off2_b[m] || !p->off4_w[m])
        return SIG_ERR_NULL;
  return 0;
}
static int check_align_grads(const sig_align_param_grads* p, bool lam, int m0, int m1) {
  if (!p) return SIG_ERR_NULL;
  if (lam)
    for (int m = m0; m < m1; ++m)
      if (!p->proj_q_w[m] || !p->proj_q_b[m] || !p->off0_w[m] || !p->off0_b[m] || !p->off2_w[m] || !p->off2_b[m] || !p->off4_w[m])
        return SIG_ERR_NULL;
  return 0;
}
static int check_grid(int h, int w, int L) {
  if (h < 8 || w < 8 || (h % 4) || (w % 4) || h * w != L) return SIG_ERR_SHAPE;  // Hk, Wk >= 2 (DAS.py:144 divides by Hk-1)
  return 0;
}

int align_forward(const sig_tokens* tok, const sig_align_params* p, int h, int w, int do_lam, float* losses, void* ctx,
                  size_t ctx_bytes, unsigned flags, cudaStream_t s) {
  SIG_TRY(check_tokens(tok, false));
  SIG_TRY(check_align_params(p, do_lam, 0, 3));
  if (!p->contra_temp || !losses || !ctx) return SIG_ERR_NULL;
  const int B = tok->B, L = tok->L, d = tok->d;
  if (do_lam) SIG_TRY(check_grid(h, w, L));
  if (tc_path_ok(tok, flags)) {
    if (!tc_strides_ok(tok)) return SIG_ERR_SHAPE;
    if (ctx_bytes < align_tc_ctx(nullptr, B, L, d).bytes) return SIG_ERR_WORKSPACE;
    // SIG_FLAG_SHARE_SMS (FusionHead): SIM's chain runs next to this call on another stream -- leave it SMs (prof.h)
    const ScopedSmBudget sm_scope((flags & SIG_FLAG_SHARE_SMS) ? align_sm_budget() : 0);
    const ScopedSmWaves wave_scope((flags & SIG_FLAG_SHARE_SMS) ? align_sm_waves() : 1);
    return align_forward_tc(tok, p, h, w, do_lam, losses, ctx, do_lam && (flags & SIG_FLAG_EAGER_BWD), (flags & SIG_FLAG_PATCH_MEAN) != 0, s);
  }
  if (ctx_bytes < align_ctx_bytes(B, L, d)) return SIG_ERR_WORKSPACE;
  AlignCtx c = align_ctx(ctx, B, L, d, 3);
  sig_tokens t2 = *tok;
  for (int m = 0; m < 3; ++m) t2.cls[m] = nullptr;
  SIG_TRY(convert_tokens(&t2, c.Xf, nullptr, s));
  // ---- GAM
  {
  SIG_PHASE("gam_fwd");
  SIG_LAUNCH((pool_kernel), dim3(B, 3), 256, 0, s, c.Xf, B, L, d, c.mean, c.meand);
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((gam_norm_kernel), B, 256, 0, s, c.mean, B, d, c.f, c.nrm, c.self4, c.meand, c.fd, c.self4d);
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((gam_gram_f64_kernel), dim3((unsigned)ceil_div(B, 32), B), 1024, 0, s, c.fd, B, d, c.lvd, c.lad);
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((gam_loss_kernel), 1, 1024, 0, s, c.self4d, c.lvd, c.lad, p->contra_temp, B, c.Vd, c.statd, c.statd + B, (float*)nullptr,
                                     c.Wlv, c.Wla, c.rowA, c.colC, losses, c.dtau);
  SIG_CHECK_LAUNCH();
  }
  // ---- LAM
  if (do_lam) {
    const Geo g = make_geo(h, w);
    const size_t BL = (size_t)B * L;
    const int64_t ms = (int64_t)B * g.P * d;
    for (int m = 0; m < 3; ++m) {
      SIG_TRY(lam_offsets_fwd(c.Xf + m * BL * d, p, m, c.mod[m], g, B, L, d, s));
      SIG_PHASE("lam_sample_fwd");
      SIG_LAUNCH((lam_sample_fwd_kernel), B * g.P, 256, 0, s, c.Xf + m * BL * d, c.mod[m].o, g, L, d, c.S + m * ms);
      SIG_CHECK_LAUNCH();
    }
    SIG_PHASE("lam_sample_fwd");
    SIG_LAUNCH((lam_mse_kernel), B * g.P, 256, 0, s, c.S, ms, d, c.part);
    SIG_CHECK_LAUNCH();
    SIG_LAUNCH((sum_kernel), 1, 256, 0, s, c.part, B * g.P, 1.f / (3.f * (float)B * g.P * d), losses + 1);
    SIG_CHECK_LAUNCH();
  }
  return 0;
}

int align_backward(const sig_tokens* tok, const sig_align_params* p, int h, int w, int do_lam, const float* dlosses,
                   const sig_token_grads* dtok, const sig_align_param_grads* dp, void* ctx, size_t ctx_bytes, unsigned flags,
                   cudaStream_t s) {
  SIG_TRY(check_tokens(tok, false));
  SIG_TRY(check_align_params(p, do_lam, 0, 3));
  SIG_TRY(check_align_grads(dp, do_lam, 0, 3));
  SIG_TRY(check_token_grads(dtok, tok->dtype, false));
  if (!dlosses || !ctx || !dp->contra_temp) return SIG_ERR_NULL;
  const int B = tok->B, L = tok->L, d = tok->d;
  if (do_lam) SIG_TRY(check_grid(h, w, L));
  if (tc_path_ok(tok, flags)) {
    if (!tc_strides_ok(tok)) return SIG_ERR_SHAPE;
    if (ctx_bytes < align_tc_ctx(nullptr, B, L, d).bytes) return SIG_ERR_WORKSPACE;
    for (int m = 1; m < 3; ++m)
      if (dtok->patch_stride_b[m] != dtok->patch_stride_b[0] || dtok->patch_stride_l[m] != dtok->patch_stride_l[0])
        return SIG_ERR_SHAPE;
    if (dtok->fuse_pds && !do_lam) return SIG_ERR_SHAPE;   // the fused operands ride on the LAM dX GEMM
    const ScopedSmBudget sm_scope((flags & SIG_FLAG_SHARE_SMS) ? align_sm_budget() : 0);
    const ScopedSmWaves wave_scope((flags & SIG_FLAG_SHARE_SMS) ? align_sm_waves() : 1);
    return align_backward_tc(tok, p, h, w, do_lam, dlosses, dtok, dp, ctx, do_lam && (flags & SIG_FLAG_EAGER_BWD), s);
  }
  if (dtok->fuse_pds) return SIG_ERR_SHAPE;   // only the tensor-core path can take SIM's operands
  if (ctx_bytes < align_ctx_bytes(B, L, d)) return SIG_ERR_WORKSPACE;
  AlignCtx c = align_ctx(ctx, B, L, d, 3);
  const size_t BL = (size_t)B * L;
  // ---- GAM: d(features), then back through normalise + mean pool (unit upstream; scaled in the writer)
  const float* fr = c.f;
  const float* fn = c.f + (size_t)B * d;
  const float* ft = c.f + (size_t)2 * B * d;
  float* dfr = c.df;
  float* dfn = c.df + (size_t)B * d;
  float* dft = c.df + (size_t)2 * B * d;
  {
  SIG_PHASE("gam_bwd");
  SIG_TRY(launch_gemm(gemm_nn(c.Wlv, B, fn, d, dfr, d, B, d, B), s));
  {
    Gemm gg = gemm_nn(c.Wla, B, ft, d, dfr, d, B, d, B);
    gg.accumulate = 1;
    SIG_TRY(launch_gemm(gg, s));
  }
  SIG_TRY(launch_gemm(gemm_tn(c.Wlv, B, fr, d, dfn, d, B, d, B), s));
  SIG_TRY(launch_gemm(gemm_tn(c.Wla, B, fr, d, dft, d, B, d, B), s));
  SIG_LAUNCH((gam_finish_kernel), dim3(B, 3), 256, 0, s, c.f, c.nrm, c.rowA, c.colC, c.df, B, L, d, c.dmean);
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((scale_scalar_kernel), 1, 1, 0, s, c.dtau, dlosses, dp->contra_temp);
  SIG_CHECK_LAUNCH();
  }
  // ---- LAM
  Geo g = make_geo(do_lam ? h : 8, do_lam ? w : 8);
  const int64_t ms = (int64_t)B * g.P * d;
  if (do_lam) {
    {
    SIG_PHASE("lam_sample_bwd");
    SIG_LAUNCH((lam_mse_bwd_kernel), B * g.P, 256, 0, s, c.S, ms, d, 2.f / (3.f * (float)B * g.P * d), dlosses + 1, c.dS);
    SIG_CHECK_LAUNCH();
    }
    for (int m = 0; m < 3; ++m) {
      const float* X = c.Xf + m * BL * d;
      {
      SIG_PHASE("lam_sample_bwd");
      SIG_LAUNCH((lam_sample_bwd_kernel), B * g.P, 256, 0, s, X, c.mod[m].o, c.dS + m * ms, g, L, d, c.mod[m].dO);
      SIG_CHECK_LAUNCH();
      }
      SIG_TRY(lam_offsets_bwd(X, p, dp, m, c.mod[m], c, c.dXf + m * BL * d, g, B, L, d, s));
    }
  }
  // ---- single writer per modality
  SIG_PHASE("align_write");
  if (dtok->wait_event) cudaStreamWaitEvent(s, (cudaEvent_t)dtok->wait_event, 0);
  const bool zero_cls = dtok->zero_cls != 0;
  for (int m = 0; m < 3; ++m) {
    const float* dense = do_lam ? c.dXf + m * BL * d : nullptr;
    const float* o = do_lam ? c.mod[m].o : nullptr;
    const float* dS = do_lam ? c.dS + m * ms : nullptr;
    const float* grow = c.dmean + (size_t)m * B * d;
    const unsigned rows = (unsigned)(BL + (zero_cls && dtok->dcls[m] ? B : 0));
    if (tok->dtype == SIG_BF16)
      SIG_LAUNCH((align_write_kernel<__nv_bfloat16>), rows, 128, 0, s, dense, grow, dlosses, o, dS, g, B, L, d,
                                                            static_cast<__nv_bfloat16*>(dtok->dpatch[m]), dtok->patch_stride_b[m],
                                                            dtok->patch_stride_l[m], static_cast<__nv_bfloat16*>(dtok->dcls[m]),
                                                            dtok->cls_stride_b[m], dtok->accumulate);
    else
      SIG_LAUNCH((align_write_kernel<float>), rows, 128, 0, s, dense, grow, dlosses, o, dS, g, B, L, d, static_cast<float*>(dtok->dpatch[m]),
                                                    dtok->patch_stride_b[m], dtok->patch_stride_l[m],
                                                    static_cast<float*>(dtok->dcls[m]), dtok->cls_stride_b[m], dtok->accumulate);
    SIG_CHECK_LAUNCH();
  }
  if (dtok->done_event) cudaEventRecord((cudaEvent_t)dtok->done_event, s);
  if (dp->done_event) cudaEventRecord((cudaEvent_t)dp->done_event, s);
  return 0;
}

// ---- single DA_sample ---------------------------------------------------------------------------
static int one_view_tokens(const void* x, int64_t sb, int64_t sl, int dtype, int B, int L, int d, sig_tokens* t) {
  if (!x) return SIG_ERR_NULL;
  for (int m = 0; m < 3; ++m) {
    t->patch[m] = x; t->cls[m] = nullptr;
    t->patch_stride_b[m] = sb; t->patch_stride_l[m] = sl; t->cls_stride_b[m] = 0;
  }
  t->dtype = dtype; t->B = B; t->L = L; t->d = d;
  return check_tokens(t, false);
}

template <typename T>
static __global__ void convert_one_kernel(const T* __restrict__ x, int64_t sb, int64_t sl, int L, int d, float* __restrict__ Xf) {
  pdl_enter();
  const int64_t row = blockIdx.x;
  const int b = (int)(row / L), l = (int)(row % L);
  const T* src = x + b * sb + l * sl;
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    float v[8];
    load8(src + c, v);
    store8(Xf + row * d + c, v);
  }
}

int das_forward(const void* x, int64_t sb, int64_t sl, int dtype, int B, int h, int w, int d, const sig_align_params* p, int m,
                float* sampled, void* ctx, size_t ctx_bytes, unsigned flags, cudaStream_t s) {
  const int L = h * w;
  sig_tokens t;
  SIG_TRY(one_view_tokens(x, sb, sl, dtype, B, L, d, &t));
  if (m < 0 || m > 2) return SIG_ERR_SHAPE;
  SIG_TRY(check_align_params(p, true, m, m + 1));
  SIG_TRY(check_grid(h, w, L));
  if (!sampled || !ctx) return SIG_ERR_NULL;
  if (ctx_bytes < das_ctx_bytes(B, L, d)) return SIG_ERR_WORKSPACE;
  (void)flags;
  AlignCtx c = align_ctx(ctx, B, L, d, 1);
  const int threads = d / 8 >= 128 ? 128 : 64;
  if (dtype == SIG_BF16)
    SIG_LAUNCH((convert_one_kernel<__nv_bfloat16>), B * L, threads, 0, s, static_cast<const __nv_bfloat16*>(x), sb, sl, L, d, c.Xf);
  else
    SIG_LAUNCH((convert_one_kernel<float>), B * L, threads, 0, s, static_cast<const float*>(x), sb, sl, L, d, c.Xf);
  SIG_CHECK_LAUNCH();
  const Geo g = make_geo(h, w);
  SIG_TRY(lam_offsets_fwd(c.Xf, p, m, c.mod[0], g, B, L, d, s));
  SIG_LAUNCH((lam_sample_fwd_kernel), B * g.P, 256, 0, s, c.Xf, c.mod[0].o, g, L, d, sampled);
  SIG_CHECK_LAUNCH();
  return 0;
}

int das_backward(const void* x, int64_t sb, int64_t sl, int dtype, int B, int h, int w, int d, const sig_align_params* p, int m,
                 const float* dsampled, void* dx, const sig_align_param_grads* dp, void* ctx, size_t ctx_bytes, unsigned flags,
                 cudaStream_t s) {
  const int L = h * w;
  sig_tokens t;
  SIG_TRY(one_view_tokens(x, sb, sl, dtype, B, L, d, &t));
  if (m < 0 || m > 2) return SIG_ERR_SHAPE;
  SIG_TRY(check_align_params(p, true, m, m + 1));
  SIG_TRY(check_align_grads(dp, true, m, m + 1));
  SIG_TRY(check_grid(h, w, L));
  if (!dsampled || !dx || !ctx) return SIG_ERR_NULL;
  if (!aligned16(dx)) return SIG_ERR_ALIGN;
  if (ctx_bytes < das_ctx_bytes(B, L, d)) return SIG_ERR_WORKSPACE;
  (void)flags;
  AlignCtx c = align_ctx(ctx, B, L, d, 1);
  const Geo g = make_geo(h, w);
  SIG_LAUNCH((lam_sample_bwd_kernel), B * g.P, 256, 0, s, c.Xf, c.mod[0].o, dsampled, g, L, d, c.mod[0].dO);
  SIG_CHECK_LAUNCH();
  SIG_TRY(lam_offsets_bwd(c.Xf, p, dp, m, c.mod[0], c, c.dXf, g, B, L, d, s));
  if (dtype == SIG_BF16)
    SIG_LAUNCH((align_write_kernel<__nv_bfloat16>), B * L, 128, 0, s, c.dXf, nullptr, nullptr, c.mod[0].o, dsampled, g, B, L, d,
                                                           static_cast<__nv_bfloat16*>(dx), sb, sl, nullptr, 0, 0);
  else
    SIG_LAUNCH((align_write_kernel<float>), B * L, 128, 0, s, c.dXf, nullptr, nullptr, c.mod[0].o, dsampled, g, B, L, d, static_cast<float*>(dx),
                                                   sb, sl, nullptr, 0, 0);
  SIG_CHECK_LAUNCH();
  return 0;
}

// ---- volume_computation3 ---------------------------------------------------------------------------
// self dots: out[i] = a[i,:] . b[i,:]   grid rows
static __global__ void __launch_bounds__(128) rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, int d,
                                                            float* __restrict__ out) {
  pdl_enter();
  __shared__ float scratch[33];
  const int64_t i = blockIdx.x;
  float s = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) s += a[i * d + c] * b[i * d + c];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[i] = s;
}

// gradient of V = sqrt|det| w.r.t. the Gram entries, per pair; writes coefficient grids
//   Wlv, Wla [B1,B2] and E_ll, E_vv, E_va, E_aa [B1,B2] (reduced by column/row sums afterwards)
static __global__ void volume_bwd_pair_kernel(const float* __restrict__ ll, const float* __restrict__ vv, const float* __restrict__ aa,
                                              const float* __restrict__ va, const float* __restrict__ lv, const float* __restrict__ la,
                                              const float* __restrict__ dvol, int B1, int B2, float* __restrict__ Wlv,
                                              float* __restrict__ Wla, float* __restrict__ E) {
  pdl_enter();
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)B1 * B2;
  if (idx >= n) return;
  const int i = (int)(idx / B2), j = (int)(idx % B2);
  const float det = gram_det(ll[i], vv[j], aa[j], va[j], lv[idx], la[idx]);
  const float V = sqrtf(fabsf(det));
  const float dd = ddet_of(dvol[idx], det, V);
  Wlv[idx] = dd * (-2.f * (lv[idx] * aa[j] - va[j] * la[idx]));
  Wla[idx] = dd * (2.f * (lv[idx] * va[j] - vv[j] * la[idx]));
  E[idx] = dd * (vv[j] * aa[j] - va[j] * va[j]);                    // c_ll
  E[n + idx] = dd * (ll[i] * aa[j] - la[idx] * la[idx]);            // c_vv
  E[2 * n + idx] = dd * (-2.f * (ll[i] * va[j] - lv[idx] * la[idx]));  // c_va
  E[3 * n + idx] = dd * (ll[i] * vv[j] - lv[idx] * lv[idx]);        // c_aa
}

// out[r,:] += 2*s1[r]*x[r,:] + s2[r]*y[r,:]
static __global__ void __launch_bounds__(128) axpy_rows_kernel(const float* __restrict__ s1, const float* __restrict__ x,
                                                               const float* __restrict__ s2, const float* __restrict__ y, int d,
                                                               float* __restrict__ out) {
  pdl_enter();
  const int64_t r = blockIdx.x;
  const float a = 2.f * s1[r], b = s2 ? s2[r] : 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) out[r * d + c] += a * x[r * d + c] + (y ? b * y[r * d + c] : 0.f);
}

// rowsum: out[i] = sum_j X[i*N + j]
static __global__ void __launch_bounds__(128) rowsum_kernel(const float* __restrict__ X, int N, float* __restrict__ out) {
  pdl_enter();
  __shared__ float scratch[33];
  const int64_t i = blockIdx.x;
  float s = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) s += X[i * N + j];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[i] = s;
}

size_t volume_ws_floats(int B1, int B2) { return (size_t)8 * B1 * B2 + 8 * (size_t)(B1 + B2) + 1024; }

int volume3_forward(const float* l, const float* v, const float* a, int B1, int B2, int d, float* vol, float* ws, cudaStream_t s) {
  // ws: ll[B1] vv[B2] aa[B2] va[B2] lv[B1*B2] la[B1*B2]
  float* ll = ws;
  float* vv = ll + B1;
  float* aa = vv + B2;
  float* va = aa + B2;
  float* lv = va + B2;
  float* la = lv + (size_t)B1 * B2;
  SIG_LAUNCH((rowdot_kernel), B1, 128, 0, s, l, l, d, ll);
  SIG_LAUNCH((rowdot_kernel), B2, 128, 0, s, v, v, d, vv);
  SIG_LAUNCH((rowdot_kernel), B2, 128, 0, s, a, a, d, aa);
  SIG_LAUNCH((rowdot_kernel), B2, 128, 0, s, v, a, d, va);
  SIG_CHECK_LAUNCH();
  SIG_TRY(launch_gemm(gemm_nt(l, d, v, d, lv, B2, nullptr, B1, B2, d), s));
  SIG_TRY(launch_gemm(gemm_nt(l, d, a, d, la, B2, nullptr, B1, B2, d), s));
  const int64_t n = (int64_t)B1 * B2;
  SIG_LAUNCH((volume_kernel), (unsigned)ceil_div(n, 256), 256, 0, s, ll, vv, aa, va, lv, la, B1, B2, vol);
  SIG_CHECK_LAUNCH();
  return 0;
}

int volume3_backward(const float* l, const float* v, const float* a, int B1, int B2, int d, const float* dvol, float* dl, float* dv,
                     float* da, float* ws, cudaStream_t s) {
  float* ll = ws;
  float* vv = ll + B1;
  float* aa = vv + B2;
  float* va = aa + B2;
  const size_t n = (size_t)B1 * B2;
  float* lv = va + B2;
  float* la = lv + n;
  float* Wlv = la + n;
  float* Wla = Wlv + n;
  float* E = Wla + n;          // 4n
  float* rs = E + 4 * n;       // rowA[B1]
  float* cs = rs + B1;         // colC[3][B2]
  SIG_LAUNCH((rowdot_kernel), B1, 128, 0, s, l, l, d, ll);
  SIG_LAUNCH((rowdot_kernel), B2, 128, 0, s, v, v, d, vv);
  SIG_LAUNCH((rowdot_kernel), B2, 128, 0, s, a, a, d, aa);
  SIG_LAUNCH((rowdot_kernel), B2, 128, 0, s, v, a, d, va);
  SIG_CHECK_LAUNCH();
  SIG_TRY(launch_gemm(gemm_nt(l, d, v, d, lv, B2, nullptr, B1, B2, d), s));
  SIG_TRY(launch_gemm(gemm_nt(l, d, a, d, la, B2, nullptr, B1, B2, d), s));
  SIG_LAUNCH((volume_bwd_pair_kernel), (unsigned)ceil_div((int64_t)n, 256), 256, 0, s, ll, vv, aa, va, lv, la, dvol, B1, B2, Wlv, Wla, E);
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((rowsum_kernel), B1, 128, 0, s, E, B2, rs);
  SIG_CHECK_LAUNCH();
  SIG_TRY(launch_colsum(E + n, B2, B1, B2, cs, 1.f, s));
  SIG_TRY(launch_colsum(E + 2 * n, B2, B1, B2, cs + B2, 1.f, s));
  SIG_TRY(launch_colsum(E + 3 * n, B2, B1, B2, cs + 2 * B2, 1.f, s));
  // dl = Wlv v + Wla a + 2 rowA l
  SIG_TRY(launch_gemm(gemm_nn(Wlv, B2, v, d, dl, d, B1, d, B2), s));
  {
    Gemm gg = gemm_nn(Wla, B2, a, d, dl, d, B1, d, B2);
    gg.accumulate = 1;
    SIG_TRY(launch_gemm(gg, s));
  }
  SIG_LAUNCH((axpy_rows_kernel), B1, 128, 0, s, rs, l, nullptr, nullptr, d, dl);
  // dv = Wlv^T l + 2 Cvv v + Cva a ; da = Wla^T l + 2 Caa a + Cva v
  SIG_TRY(launch_gemm(gemm_tn(Wlv, B2, l, d, dv, d, B2, d, B1), s));
  SIG_LAUNCH((axpy_rows_kernel), B2, 128, 0, s, cs, v, cs + B2, a, d, dv);
  SIG_TRY(launch_gemm(gemm_tn(Wla, B2, l, d, da, d, B2, d, B1), s));
  SIG_LAUNCH((axpy_rows_kernel), B2, 128, 0, s, cs + 2 * B2, a, cs + B2, v, d, da);
  SIG_CHECK_LAUNCH();
  return 0;
}

}  // namespace sig

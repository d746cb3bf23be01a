// volume_computation4 / volume_computation5 (utils/volume.py:65-116, :119-182): V[i,j] = sqrt|det G(i,j)| of the n x n Gram
// matrix of (language_i, video_j, audio_j, subtitles_j[, depth_j]), forward and backward, behind the C ABI
// (include/signal_b200.h: sig_volume_n_fwd / _bwd; n = 3 is accepted too and cross-checks sig_volume3_*).
// The reference never calls these two (only volume_computation3, useB.py:14); they complete the utils/volume.py surface.
//
//   G_00 = l_i.l_i      G_0k = l_i.m_k_j  (n-1 B1 x B2 cross grids = fp32 GEMMs)      G_kl = m_k_j.m_l_j  (per-j row dots)
//   forward : per pair, det by Gaussian elimination with partial pivoting in fp64 (the reference: torch.det, fp32 LU)
//   backward: d det / d G = cofactor matrix C (minors by the same elimination), ddet = dV sign(det) / (2 V);
//             d l_i   = sum_j 2 ddet [C_00 l_i + sum_k C_0k m_k_j]      = 2 rowA_i l_i + sum_k (W_k m_k)_i
//             d m_k_j = sum_i 2 ddet [C_k0 l_i + sum_l C_kl m_l_j]      = (W_k^T l)_j + sum_l 2 S_kl[j] m_l_j
//             with W_k = 2 ddet C_0k [B1,B2], rowA = rowsum(ddet C_00), S_kl = colsum(ddet C_kl): the same structure as
//             volume3_backward (align.cu), generalised over n.  All sums in a fixed order: deterministic.
#include "common.cuh"
#include "simt_ops.cuh"

namespace sig {
namespace {

constexpr int kVolMaxN = 5;

struct VolFeats {
  const float* f[kVolMaxN];   // [0] language [B1,d]; [1..n-1] [B2,d]
};
struct VolGrads {
  float* g[kVolMaxN];
};

// out[r] = a[r,:] . b[r,:]
__global__ void __launch_bounds__(128) voln_rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, int d,
                                                          float* __restrict__ out) {
  pdl_enter();
  __shared__ float scratch[33];
  const int64_t i = blockIdx.x;
  float s = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) s += a[i * d + c] * b[i * d + c];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[i] = s;
}

// determinant of the leading m x m block of A (destroyed): Gaussian elimination, partial pivoting
template <int N>
__device__ __forceinline__ double det_inplace(double (&A)[N][N], int m) {
  double det = 1.0;
#pragma unroll
  for (int c = 0; c < N; ++c) {
    if (c >= m) break;
    int piv = c;
    double best = fabs(A[c][c]);
#pragma unroll
    for (int r = 0; r < N; ++r)
      if (r > c && r < m && fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (best == 0.0) return 0.0;
    if (piv != c) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        // (piv is a run-time row index: swap through selects so that A stays in registers)
#pragma unroll
        for (int r = 0; r < N; ++r)
          if (r == piv) { const double t = A[c][k]; A[c][k] = A[r][k]; A[r][k] = t; }
      }
      det = -det;
    }
    det *= A[c][c];
    const double inv = 1.0 / A[c][c];
#pragma unroll
    for (int r = 0; r < N; ++r) {
      if (r > c && r < m) {
        const double f = A[r][c] * inv;
#pragma unroll
        for (int k = 0; k < N; ++k)
          if (k > c && k < m) A[r][k] -= f * A[c][k];
      }
    }
  }
  return det;
}

// G(i,j): self [B1] = l.l; mm [(n-1)^2][B2] (entry (k-1)*(n-1)+(l-1) for k <= l); cross [n-1][B1*B2]
template <int N>
__device__ __forceinline__ void build_gram(double (&G)[N][N], const float* __restrict__ ll, const float* __restrict__ mm,
                                           const float* __restrict__ cross, int i, int j, int64_t idx, int64_t n, int B2) {
  G[0][0] = ll[i];
#pragma unroll
  for (int k = 1; k < N; ++k) {
    const double c = cross[(int64_t)(k - 1) * n + idx];
    G[0][k] = c;
    G[k][0] = c;
#pragma unroll
    for (int l = k; l < N; ++l) {
      const double v = mm[(int64_t)((k - 1) * (N - 1) + (l - 1)) * B2 + j];
      G[k][l] = v;
      G[l][k] = v;
    }
  }
}

template <int N>
__global__ void voln_fwd_pair_kernel(const float* __restrict__ ll, const float* __restrict__ mm, const float* __restrict__ cross, int B1,
                                     int B2, float* __restrict__ vol) {
  pdl_enter();
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)B1 * B2;
  if (idx >= n) return;
  const int i = (int)(idx / B2), j = (int)(idx % B2);
  double G[N][N];
  build_gram<N>(G, ll, mm, cross, i, j, idx, n, B2);
  vol[idx] = (float)sqrt(fabs(det_inplace<N>(G, N)));
}

// W [n-1][B1*B2] = 2 ddet C_0k; E00 [B1*B2] = ddet C_00; E [(n-1)^2][B1*B2] = ddet C_kl for 1 <= k <= l
template <int N>
__global__ void voln_bwd_pair_kernel(const float* __restrict__ ll, const float* __restrict__ mm, const float* __restrict__ cross,
                                     const float* __restrict__ dvol, int B1, int B2, float* __restrict__ W, float* __restrict__ E00,
                                     float* __restrict__ E) {
  pdl_enter();
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)B1 * B2;
  if (idx >= n) return;
  const int i = (int)(idx / B2), j = (int)(idx % B2);
  double G[N][N];
  build_gram<N>(G, ll, mm, cross, i, j, idx, n, B2);
  double A[N][N];
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c) A[r][c] = G[r][c];
  const double det = det_inplace<N>(A, N);
  const double V = sqrt(fabs(det));
  // d sqrt|det| = sign(det) / (2 sqrt|det|); the reference yields NaN / inf at det == 0, we give 0 (like sig_volume3_bwd)
  const double dd = V > 0.0 ? (double)dvol[idx] * (det > 0.0 ? 0.5 : -0.5) / V : 0.0;
  // cofactors C_ab for a <= b (G is symmetric): (-1)^(a+b) det(G without row a and column b)
#pragma unroll
  for (int a = 0; a < N; ++a) {
#pragma unroll
    for (int b = a; b < N; ++b) {
      double M[N][N];
#pragma unroll
      for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c) {
          const int rs = r < a ? r : r + 1, cs = c < b ? c : c + 1;   // compile-time after unrolling
          M[r][c] = (r < N - 1 && c < N - 1) ? G[rs < N ? rs : 0][cs < N ? cs : 0] : 0.0;
        }
      double cof = det_inplace<N>(M, N - 1);
      if ((a + b) & 1) cof = -cof;
      const float v = (float)(dd * cof);
      if (a == 0 && b == 0) E00[idx] = v;
      else if (a == 0) W[(int64_t)(b - 1) * n + idx] = 2.f * v;
      else E[(int64_t)((a - 1) * (N - 1) + (b - 1)) * n + idx] = v;
    }
  }
}

// out[r,:] (+)= 2 * s[r] * x[r,:]
__global__ void __launch_bounds__(128) voln_axpy_rows_kernel(const float* __restrict__ s1, const float* __restrict__ x, int d,
                                                             float* __restrict__ out) {
  pdl_enter();
  const int64_t r = blockIdx.x;
  const float a = 2.f * s1[r];
  for (int c = threadIdx.x; c < d; c += blockDim.x) out[r * d + c] += a * x[r * d + c];
}

__global__ void __launch_bounds__(128) voln_rowsum_kernel(const float* __restrict__ X, int N, float* __restrict__ out) {
  pdl_enter();
  __shared__ float scratch[33];
  const int64_t i = blockIdx.x;
  float s = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) s += X[i * N + j];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[i] = s;
}

struct VolWs {
  float *ll, *mm, *cross, *W, *E00, *E, *rowA, *S;
  size_t floats;
};
VolWs vol_ws(float* base, int n, int B1, int B2) {
  VolWs w{};
  const size_t np = (size_t)B1 * B2, q = (size_t)(n - 1) * (n - 1);
  size_t o = 0;
  auto take = [&](size_t k) { float* p = base ? base + o : nullptr; o += (k + 63) & ~(size_t)63; return p; };
  w.ll = take(B1);
  w.mm = take(q * B2);
  w.cross = take((size_t)(n - 1) * np);
  w.W = take((size_t)(n - 1) * np);
  w.E00 = take(np);
  w.E = take(q * np);
  w.rowA = take(B1);
  w.S = take(q * B2);
  w.floats = o;
  return w;
}

int voln_gram_inputs(int n, const VolFeats& f, int B1, int B2, int d, const VolWs& w, cudaStream_t s) {
  SIG_LAUNCH((voln_rowdot_kernel), B1, 128, 0, s, f.f[0], f.f[0], d, w.ll);
  for (int k = 1; k < n; ++k)
    for (int l = k; l < n; ++l)
      SIG_LAUNCH((voln_rowdot_kernel), B2, 128, 0, s, f.f[k], f.f[l], d, w.mm + (size_t)((k - 1) * (n - 1) + (l - 1)) * B2);
  SIG_CHECK_LAUNCH();
  const size_t np = (size_t)B1 * B2;
  for (int k = 1; k < n; ++k) SIG_TRY(launch_gemm(gemm_nt(f.f[0], d, f.f[k], d, w.cross + (size_t)(k - 1) * np, B2, nullptr, B1, B2, d), s));
  return 0;
}

int voln_forward(int n, const VolFeats& f, int B1, int B2, int d, float* vol, float* ws, cudaStream_t s) {
  const VolWs w = vol_ws(ws, n, B1, B2);
  SIG_TRY(voln_gram_inputs(n, f, B1, B2, d, w, s));
  const int64_t np = (int64_t)B1 * B2;
  const unsigned grid = (unsigned)ceil_div(np, 128);
  if (n == 3) SIG_LAUNCH((voln_fwd_pair_kernel<3>), grid, 128, 0, s, w.ll, w.mm, w.cross, B1, B2, vol);
  else if (n == 4) SIG_LAUNCH((voln_fwd_pair_kernel<4>), grid, 128, 0, s, w.ll, w.mm, w.cross, B1, B2, vol);
  else SIG_LAUNCH((voln_fwd_pair_kernel<5>), grid, 128, 0, s, w.ll, w.mm, w.cross, B1, B2, vol);
  SIG_CHECK_LAUNCH();
  return 0;
}

int voln_backward(int n, const VolFeats& f, int B1, int B2, int d, const float* dvol, const VolGrads& g, float* ws, cudaStream_t s) {
  const VolWs w = vol_ws(ws, n, B1, B2);
  SIG_TRY(voln_gram_inputs(n, f, B1, B2, d, w, s));
  const int64_t np = (int64_t)B1 * B2;
  const unsigned grid = (unsigned)ceil_div(np, 128);
  if (n == 3) SIG_LAUNCH((voln_bwd_pair_kernel<3>), grid, 128, 0, s, w.ll, w.mm, w.cross, dvol, B1, B2, w.W, w.E00, w.E);
  else if (n == 4) SIG_LAUNCH((voln_bwd_pair_kernel<4>), grid, 128, 0, s, w.ll, w.mm, w.cross, dvol, B1, B2, w.W, w.E00, w.E);
  else SIG_LAUNCH((voln_bwd_pair_kernel<5>), grid, 128, 0, s, w.ll, w.mm, w.cross, dvol, B1, B2, w.W, w.E00, w.E);
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((voln_rowsum_kernel), B1, 128, 0, s, w.E00, B2, w.rowA);
  SIG_CHECK_LAUNCH();
  for (int k = 1; k < n; ++k)
    for (int l = k; l < n; ++l) {
      const size_t e = (size_t)((k - 1) * (n - 1) + (l - 1));
      SIG_TRY(launch_colsum(w.E + e * np, B2, B1, B2, w.S + e * B2, 1.f, s));
    }
  // d l = sum_k W_k m_k + 2 rowA l
  for (int k = 1; k < n; ++k) {
    Gemm gg = gemm_nn(w.W + (size_t)(k - 1) * np, B2, f.f[k], d, g.g[0], d, B1, d, B2);
    gg.accumulate = k > 1;
    SIG_TRY(launch_gemm(gg, s));
  }
  SIG_LAUNCH((voln_axpy_rows_kernel), B1, 128, 0, s, w.rowA, f.f[0], d, g.g[0]);
  SIG_CHECK_LAUNCH();
  // d m_k = W_k^T l + sum_l 2 S_kl m_l
  for (int k = 1; k < n; ++k) {
    SIG_TRY(launch_gemm(gemm_tn(w.W + (size_t)(k - 1) * np, B2, f.f[0], d, g.g[k], d, B2, d, B1), s));
    for (int l = 1; l < n; ++l) {
      const int a = k < l ? k : l, b = k < l ? l : k;
      SIG_LAUNCH((voln_axpy_rows_kernel), B2, 128, 0, s, w.S + (size_t)((a - 1) * (n - 1) + (b - 1)) * B2, f.f[l], d, g.g[k]);
    }
    SIG_CHECK_LAUNCH();
  }
  return 0;
}


}  // namespace
}  // namespace sig

extern "C" {

size_t sig_volume_n_ws_bytes(int n, int B1, int B2) {
  if (n < 3 || n > sig::kVolMaxN || B1 < 1 || B2 < 1) return 0;
  return sig::vol_ws(nullptr, n, B1, B2).floats * sizeof(float);
}

int sig_volume_n_fwd(int n, const float* const* feats, int B1, int B2, int d, float* vol, void* ws, size_t ws_bytes, int device,
                     void* stream) {
  using namespace sig;
  SIG_ENTER(device);
  if (n < 3 || n > kVolMaxN || B1 < 1 || B2 < 1 || d < 1) return SIG_ERR_SHAPE;
  if (!feats || !vol || !ws) return SIG_ERR_NULL;
  VolFeats f{};
  for (int k = 0; k < n; ++k) {
    if (!feats[k]) return SIG_ERR_NULL;
    f.f[k] = feats[k];
  }
  if (ws_bytes < sig_volume_n_ws_bytes(n, B1, B2)) return SIG_ERR_WORKSPACE;
  return voln_forward(n, f, B1, B2, d, vol, static_cast<float*>(ws), (cudaStream_t)stream);
}

int sig_volume_n_bwd(int n, const float* const* feats, int B1, int B2, int d, const float* dvol, float* const* dfeats, void* ws,
                     size_t ws_bytes, int device, void* stream) {
  using namespace sig;
  SIG_ENTER(device);
  if (n < 3 || n > kVolMaxN || B1 < 1 || B2 < 1 || d < 1) return SIG_ERR_SHAPE;
  if (!feats || !dfeats || !dvol || !ws) return SIG_ERR_NULL;
  VolFeats f{};
  VolGrads g{};
  for (int k = 0; k < n; ++k) {
    if (!feats[k] || !dfeats[k]) return SIG_ERR_NULL;
    f.f[k] = feats[k];
    g.g[k] = dfeats[k];
  }
  if (ws_bytes < sig_volume_n_ws_bytes(n, B1, B2)) return SIG_ERR_WORKSPACE;
  return voln_backward(n, f, B1, B2, d, dvol, g, static_cast<float*>(ws), (cudaStream_t)stream);
}

}  // extern "C"

// fp32 SIMT building blocks (exact-precision path, also the cross-check for the
// tcgen05 kernels): strided GEMM, column sums, LayerNorm fwd/bwd, GELU backward.
#pragma once
#include "common.cuh"

namespace sig {

// C(m,n) = alpha * sum_k A(m,k) B(n,k) [+ bias(n)] [+ C(m,n)]  ;  optional GELU.
// A(m,k) = A[m*am + k*ak], B(n,k) = B[n*bn + k*bk], C(m,n) = C[m*cm + n].
// batch: blockIdx.z offsets (az, bz, cz, biasz).  ksplit > 1: blockIdx.z splits K and
// the result is atomically added into a pre-zeroed C (no bias/act in that mode).
struct Gemm {
  const float* A; int64_t am, ak;
  const float* B; int64_t bn, bk;
  float* C; int64_t cm;
  const float* bias;
  float* pre;  // optional pre-activation copy (same layout as C)
  int M, N, K;
  int batch; int64_t az, bz, cz, biasz;
  float alpha;
  int act;         // 0 none, 1 GELU
  int accumulate;  // C += result
  int ksplit;
};

inline Gemm gemm_nt(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
                    const float* bias, int M, int N, int K) {
  // C[M,N] = A[M,K] W[N,K]^T + bias
  Gemm g{};
  g.A = A; g.am = lda; g.ak = 1; g.B = W; g.bn = ldw; g.bk = 1; g.C = C; g.cm = ldc; g.bias = bias;
  g.M = M; g.N = N; g.K = K; g.batch = 1; g.alpha = 1.f; g.ksplit = 1;
  return g;
}
inline Gemm gemm_nn(const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc,
                    int M, int N, int K) {
  // C[M,N] = A[M,K] Bm[K,N]
  Gemm g{};
  g.A = A; g.am = lda; g.ak = 1; g.B = Bm; g.bn = 1; g.bk = ldb; g.C = C; g.cm = ldc;
  g.M = M; g.N = N; g.K = K; g.batch = 1; g.alpha = 1.f; g.ksplit = 1;
  return g;
}
inline Gemm gemm_tn(const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc,
                    int M, int N, int K) {
  // C[M,N] = A[K,M]^T Bm[K,N]
  Gemm g{};
  g.A = A; g.am = 1; g.ak = lda; g.B = Bm; g.bn = 1; g.bk = ldb; g.C = C; g.cm = ldc;
  g.M = M; g.N = N; g.K = K; g.batch = 1; g.alpha = 1.f; g.ksplit = 1;
  return g;
}

constexpr int kGemmBM = 64, kGemmBN = 64, kGemmBK = 16;

static __global__ void __launch_bounds__(256) gemm_simt_kernel(Gemm g) {
  pdl_enter();
  __shared__ float As[kGemmBK][kGemmBM + 4];
  __shared__ float Bs[kGemmBK][kGemmBN + 4];
  const int tid = threadIdx.x;
  int z = blockIdx.z;
  int kbeg = 0, kend = g.K;
  if (g.ksplit > 1) {
    const int per = (int)ceil_div(ceil_div(g.K, g.ksplit), kGemmBK) * kGemmBK;
    kbeg = z * per;
    kend = min(g.K, kbeg + per);
    z = 0;
    if (kbeg >= kend) return;
  }
  const float* A = g.A + z * g.az;
  const float* B = g.B + z * g.bz;
  float* C = g.C + z * g.cz;
  const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += kGemmBK) {
    if (g.ak == 1) {
      const int m = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gm = m0 + m, gk = k0 + kk + i;
        As[kk + i][m] = (gm < g.M && gk < kend) ? A[gm * g.am + gk] : 0.f;
      }
    } else {
      const int kk = tid >> 4, m = (tid & 15) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gm = m0 + m + i, gk = k0 + kk;
        As[kk][m + i] = (gm < g.M && gk < kend) ? A[gm * g.am + gk * g.ak] : 0.f;
      }
    }
    if (g.bk == 1) {
      const int n = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gn = n0 + n, gk = k0 + kk + i;
        Bs[kk + i][n] = (gn < g.N && gk < kend) ? B[gn * g.bn + gk] : 0.f;
      }
    } else {
      const int kk = tid >> 4, n = (tid & 15) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gn = n0 + n + i, gk = k0 + kk;
        Bs[kk][n + i] = (gn < g.N && gk < kend) ? B[gn * g.bn + gk * g.bk] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGemmBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const float* bias = g.bias ? g.bias + z * g.biasz : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float v = acc[i][j] * g.alpha;
      float* dst = C + gm * g.cm + gn;
      if (g.ksplit > 1) {
        atomicAdd(dst, v);
        continue;
      }
      if (bias) v += bias[gn];
      if (g.accumulate) v += *dst;
      if (g.pre) g.pre[z * g.cz + gm * g.cm + gn] = v;
      if (g.act == 1) v = gelu_f(v);
      *dst = v;
    }
  }
}

inline int launch_gemm(const Gemm& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
  dim3 grid((unsigned)ceil_div(g.N, kGemmBN), (unsigned)ceil_div(g.M, kGemmBM),
            (unsigned)(g.ksplit > 1 ? g.ksplit : g.batch));
  SIG_LAUNCH((gemm_simt_kernel), grid, 256, 0, s, g);
  SIG_CHECK_LAUNCH();
  return 0;
}

// out[n] = scale * sum_m X[m*ldx + n] (deterministic; 32 columns x 32 row lanes per CTA, 4 loads in flight per thread)
static __global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ X, int64_t ldx, int M, int N,
                                                             float* __restrict__ out, float scale) {
  pdl_enter();
  __shared__ float sm[32][33];
  const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + c;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (n < N) {
    int m = r;
    for (; m + 96 < M; m += 128) {
      a0 += X[(int64_t)m * ldx + n];
      a1 += X[(int64_t)(m + 32) * ldx + n];
      a2 += X[(int64_t)(m + 64) * ldx + n];
      a3 += X[(int64_t)(m + 96) * ldx + n];
    }
    for (; m < M; m += 32) a0 += X[(int64_t)m * ldx + n];
  }
  sm[r][c] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (r == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += sm[i][c];
    out[n] = t * scale;
  }
}
// Up to four column sums of same-shaped matrices in one launch (grid.y = job).
struct ColsumJobs {
  const float* X[4];
  float* out[4];
};
static __global__ void __launch_bounds__(1024) colsum_multi_kernel(ColsumJobs jobs, int64_t ldx, int M, int N) {
  pdl_enter();
  __shared__ float sm[32][33];
  const float* __restrict__ X = jobs.X[blockIdx.y];
  const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + c;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (n < N) {
    int m = r;
    for (; m + 96 < M; m += 128) {
      a0 += X[(int64_t)m * ldx + n];
      a1 += X[(int64_t)(m + 32) * ldx + n];
      a2 += X[(int64_t)(m + 64) * ldx + n];
      a3 += X[(int64_t)(m + 96) * ldx + n];
    }
    for (; m < M; m += 32) a0 += X[(int64_t)m * ldx + n];
  }
  sm[r][c] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (r == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += sm[i][c];
    jobs.out[blockIdx.y][n] = t;
  }
}
inline int launch_colsum3(const float* X0, float* o0, const float* X1, float* o1, const float* X2, float* o2, int64_t ldx, int M, int N,
                          cudaStream_t s) {
  ColsumJobs j{};
  j.X[0] = X0; j.out[0] = o0; j.X[1] = X1; j.out[1] = o1; j.X[2] = X2; j.out[2] = o2;
  SIG_LAUNCH((colsum_multi_kernel), dim3((unsigned)ceil_div(N, 32), 3), 1024, 0, s, j, ldx, M, N);
  SIG_CHECK_LAUNCH();
  return 0;
}
inline int launch_colsum(const float* X, int64_t ldx, int M, int N, float* out, float scale, cudaStream_t s) {
  SIG_LAUNCH((colsum_kernel), (unsigned)ceil_div(N, 32), 1024, 0, s, X, ldx, M, N, out, scale);
  SIG_CHECK_LAUNCH();
  return 0;
}

// y = LN(x [+ res]) * gamma + beta over rows of width d.  Saves xhat-ready stats.
// OutT: float or bf16 (the final SIM output is produced in the token dtype).
template <typename OutT>
static __global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   int d, float* __restrict__ sum_out, float* __restrict__ mean,
                                                                   float* __restrict__ rstd, OutT* __restrict__ y,
                                                                   __nv_bfloat16* __restrict__ y_shadow) {
  pdl_enter();
  __shared__ float scratch[33];
  const int64_t row = blockIdx.x;
  const float* xr = x + row * d;
  const float* rr = res ? res + row * d : nullptr;
  float s = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) s += xr[c] + (rr ? rr[c] : 0.f);
  const float mu = block_sum(s, scratch) / d;
  float v = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float t = xr[c] + (rr ? rr[c] : 0.f) - mu;
    v += t * t;
  }
  const float rs = rsqrtf(block_sum(v, scratch) / d + kLnEps);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float t = xr[c] + (rr ? rr[c] : 0.f);
    if (sum_out) sum_out[row * d + c] = t;
    const float yv = (t - mu) * rs * gamma[c] + beta[c];
    y[row * d + c] = from_f32<OutT>(yv);
    if (y_shadow) y_shadow[row * d + c] = __float2bfloat16_rn(yv);   // operand of the next tcgen05 GEMM
  }
  if (threadIdx.x == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; also writes
// xhat*dy and dy in fp32 (dgamma/dbeta follow as column sums).  [+ extra] is added to dx.
template <typename InT>
static __global__ void __launch_bounds__(256) layernorm_bwd_kernel(const InT* __restrict__ dy, const float* __restrict__ x,
                                                                   const float* __restrict__ gamma, const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, const float* __restrict__ extra,
                                                                   int d, float* __restrict__ dx, float* __restrict__ dyx,
                                                                   float* __restrict__ dyf, __nv_bfloat16* __restrict__ dx_shadow) {
  pdl_enter();
  __shared__ float scratch[33];
  const int64_t row = blockIdx.x;
  const float mu = mean[row], rs = rstd[row];
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float g = to_f32<InT>(dy[row * d + c]) * gamma[c];
    const float xh = (x[row * d + c] - mu) * rs;
    s1 += g;
    s2 += g * xh;
  }
  s1 = block_sum(s1, scratch) / d;
  s2 = block_sum(s2, scratch) / d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float dyv = to_f32<InT>(dy[row * d + c]);
    const float g = dyv * gamma[c];
    const float xh = (x[row * d + c] - mu) * rs;
    float v = rs * (g - s1 - xh * s2);
    if (extra) v += extra[row * d + c];
    dx[row * d + c] = v;
    if (dx_shadow) dx_shadow[row * d + c] = __float2bfloat16_rn(v);
    dyx[row * d + c] = dyv * xh;
    dyf[row * d + c] = dyv;
  }
}

// da = dh * gelu'(a)
static __global__ void gelu_bwd_kernel(const float* dh, const float* __restrict__ a, float* da, int64_t n) {
  pdl_enter();  // in-place safe
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) da[i] = dh[i] * gelu_grad_f(a[i]);
}

// y[i] = a[i] + b[i]
static __global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n) {
  pdl_enter();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}

// y[i] = (add ? add[i] : 0) + sum_k W(i,k) x[k];  W(i,k) = W[i*ws_i + k*ws_k].  One warp per output.  grid ceil(n/8), 256 threads
static __global__ void __launch_bounds__(256) gemv_kernel(const float* __restrict__ W, int64_t ws_i, int64_t ws_k,
                                                          const float* __restrict__ x, const float* __restrict__ add, int n, int kdim,
                                                          float* __restrict__ y) {
  pdl_enter();
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  float a = 0.f;
  for (int k = lane; k < kdim; k += 32) a = fmaf(W[i * ws_i + k * ws_k], x[k], a);
  a = warp_sum(a);
  if (lane == 0) y[i] = a + (add ? add[i] : 0.f);
}

// y[i] = sum_k W[k*ld + i] x[k]  (W^T x with coalesced row reads).  grid ceil(n/32), 1024 threads = 32 outputs x 32 k-lanes
static __global__ void __launch_bounds__(1024) gemv_t_kernel(const float* __restrict__ W, int64_t ld, const float* __restrict__ x, int n,
                                                             int kdim, float* __restrict__ y) {
  pdl_enter();
  __shared__ float sm[32][33];
  const int il = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + il;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (i < n) {
    int k = r;
    for (; k + 96 < kdim; k += 128) {
      a0 = fmaf(W[(int64_t)k * ld + i], x[k], a0);
      a1 = fmaf(W[(int64_t)(k + 32) * ld + i], x[k + 32], a1);
      a2 = fmaf(W[(int64_t)(k + 64) * ld + i], x[k + 64], a2);
      a3 = fmaf(W[(int64_t)(k + 96) * ld + i], x[k + 96], a3);
    }
    for (; k < kdim; k += 32) a0 = fmaf(W[(int64_t)k * ld + i], x[k], a0);
  }
  sm[r][il] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (r == 0 && i < n) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) t += sm[q][il];
    y[i] = t;
  }
}

}  // namespace sig

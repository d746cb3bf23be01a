// Inference tail of the Signal model, behind the C ABI (include/signal_b200.h):
//   sig_infer_features     make_model.py:284-290  feat = cat([RGB_global, NI_global, TI_global, vars_total]) (+ L2 normalise,
//                                                 utils/metrics.py:266-268)
//   sig_euclidean_distmat  utils/metrics.py:494-501   dist = |q|^2 + |g|^2 - 2 q g^T  (squared distances, fp32)
//   sig_rank_eval          utils/metrics.py:111-170   market1501 CMC / mAP: gallery entries with the query's pid AND camid are
//                                                     discarded, ranking by ascending distance
// The reference moves every feature to the host (metrics.py:245) and ranks with numpy; here the features, the distance
// matrix and the per-query statistics stay on the device and nothing synchronises with the host.
//
// Ranking without a sort: a query has few matches (the gallery images of its identity), so the rank of each match is
// COUNTED -- rank_j = #{valid k : d_k < d_j or (d_k == d_j and k < j)}, i.e. a stable ascending argsort (numpy's default
// argsort leaves the order of equal distances unspecified; the oracle and this kernel both take the lowest index first).
// AP = (1/m) sum_matches (#matches ranked at or before j) / (rank_j + 1), CMC[r] = 1 iff the best match has rank <= r.
#include "common.cuh"
#include "simt_ops.cuh"

namespace sig {
namespace {

// one CTA per sample: out[b] = [cls_r | cls_n | cls_t | sim_out[b]] (normalised in fp32 when `normalize`)
template <typename T>
__global__ void __launch_bounds__(256) infer_features_kernel(const T* __restrict__ c0, const T* __restrict__ c1, const T* __restrict__ c2,
                                                             int64_t s0, int64_t s1, int64_t s2, const T* __restrict__ sim,
                                                             int64_t ld_sim, int d, int normalize, float* __restrict__ out) {
  pdl_enter();
  __shared__ float scratch[33];
  const int64_t b = blockIdx.x;
  const int D = 6 * d;
  const T* src[3] = {c0 + b * s0, c1 + b * s1, c2 + b * s2};
  float ss = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float v = c < 3 * d ? to_f32(src[c / d][c % d]) : to_f32(sim[b * ld_sim + (c - 3 * d)]);
    out[b * D + c] = v;
    ss += v * v;
  }
  if (!normalize) return;
  const float inv = 1.f / fmaxf(sqrtf(block_sum(ss, scratch)), 1e-12f);   // F.normalize eps
  for (int c = threadIdx.x; c < D; c += blockDim.x) out[b * D + c] *= inv;
}

// sq[i] = sum_c x[i,c]^2   grid rows
__global__ void __launch_bounds__(256) rownorm2_kernel(const float* __restrict__ x, int64_t ld, int D, float* __restrict__ sq) {
  pdl_enter();
  __shared__ float scratch[33];
  const int64_t i = blockIdx.x;
  float s = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) s = fmaf(x[i * ld + c], x[i * ld + c], s);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) sq[i] = s;
}

// dist[i,j] = qq[i] + gg[j] - 2 * dot[i,j]   (dot already holds q g^T)
__global__ void dist_finish_kernel(float* __restrict__ dist, const float* __restrict__ qq, const float* __restrict__ gg, int nq, int ng) {
  pdl_enter();
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)nq * ng) return;
  const int i = (int)(idx / ng), j = (int)(idx % ng);
  dist[idx] = (qq[i] + gg[j]) + (-2.f) * dist[idx];    // same order of operations as the reference's addmm_(beta=1, alpha=-2)
}

constexpr int kMaxMatches = 2048;

// one CTA per query.  stats[q] = {valid (0/1), AP, first-match rank}
__global__ void __launch_bounds__(512) rank_eval_kernel(const float* __restrict__ dist, int64_t ld, const int64_t* __restrict__ q_pids,
                                                        const int64_t* __restrict__ g_pids, const int64_t* __restrict__ q_cams,
                                                        const int64_t* __restrict__ g_cams, int ng, double* __restrict__ stats,
                                                        int* __restrict__ overflow) {
  pdl_enter();
  __shared__ int match_idx[kMaxMatches];
  __shared__ int match_rank[kMaxMatches];
  __shared__ int n_match;
  const int q = blockIdx.x, tid = threadIdx.x;
  const float* drow = dist + (int64_t)q * ld;
  const int64_t pid = q_pids[q], cam = q_cams[q];
  if (tid == 0) n_match = 0;
  __syncthreads();
  // matches (valid = not [same pid and same camera]); in ascending gallery index
  for (int base = 0; base < ng; base += blockDim.x) {
    const int j = base + tid;
    const bool m = j < ng && g_pids[j] == pid && g_cams[j] != cam;
    // ordered compaction: ballot per warp, then warp offsets
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    __shared__ int warp_cnt[16];
    if ((tid & 31) == 0) warp_cnt[tid >> 5] = __popc(bal);
    __syncthreads();
    int off = n_match;
    for (int w = 0; w < (tid >> 5); ++w) off += warp_cnt[w];
    if (m) {
      const int pos = off + __popc(bal & ((1u << (tid & 31)) - 1u));
      if (pos < kMaxMatches) match_idx[pos] = j;
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_cnt[w];
      n_match += tot;
    }
    __syncthreads();
  }
  int nm = n_match;
  if (nm > kMaxMatches) {
    if (tid == 0) atomicExch(overflow, 1);
    nm = kMaxMatches;
  }
  if (nm == 0) {   // this identity does not appear in the gallery: the query is skipped (metrics.py:141-143)
    if (tid == 0) { stats[3 * q] = 0.0; stats[3 * q + 1] = 0.0; stats[3 * q + 2] = 0.0; }
    return;
  }
  // rank of every match among the valid gallery entries: one warp per match, lanes over the gallery
  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int mi = warp; mi < nm; mi += nwarps) {
    const int j = match_idx[mi];
    const float dj = drow[j];
    int cnt = 0;
    for (int k = lane; k < ng; k += 32) {
      const float dk = drow[k];
      const bool junk = g_pids[k] == pid && g_cams[k] == cam;
      cnt += (!junk && (dk < dj || (dk == dj && k < j))) ? 1 : 0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) match_rank[mi] = cnt;
  }
  __syncthreads();
  // AP = mean over matches of (#matches with rank <= rank_j) / (rank_j + 1); ranks are distinct
  double ap = 0.0;
  int best = 0x7fffffff;
  for (int mi = tid; mi < nm; mi += blockDim.x) {
    const int r = match_rank[mi];
    int ahead = 0;
    for (int k = 0; k < nm; ++k) ahead += match_rank[k] <= r ? 1 : 0;
    ap += (double)ahead / (double)(r + 1);
    best = min(best, r);
  }
  __shared__ double red[16];
  __shared__ int redi[16];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ap += __shfl_xor_sync(0xffffffffu, ap, o);
    best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  }
  if (lane == 0) { red[warp] = ap; redi[warp] = best; }
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    int bb = 0x7fffffff;
    for (int w = 0; w < nwarps; ++w) { t += red[w]; bb = min(bb, redi[w]); }
    stats[3 * q] = 1.0;
    stats[3 * q + 1] = t / (double)nm;
    stats[3 * q + 2] = (double)bb;
  }
}

// cmc[r] = mean over valid queries of [first-match rank <= r]  (float32 like the reference), out[max_rank] = mAP, out[max_rank+1] = #valid
__global__ void __launch_bounds__(256) rank_reduce_kernel(const double* __restrict__ stats, int nq, int max_rank, float* __restrict__ cmc,
                                                          double* __restrict__ map_out) {
  pdl_enter();
  __shared__ double sd[256];
  const int tid = threadIdx.x;
  double nv = 0.0, ap = 0.0;
  for (int q = tid; q < nq; q += blockDim.x) { nv += stats[3 * q]; ap += stats[3 * q] * stats[3 * q + 1]; }
  sd[tid] = nv;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (tid < o) sd[tid] += sd[tid + o]; __syncthreads(); }
  const double nvalid = sd[0];
  __syncthreads();
  sd[tid] = ap;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (tid < o) sd[tid] += sd[tid + o]; __syncthreads(); }
  if (tid == 0) { map_out[0] = nvalid > 0 ? sd[0] / nvalid : 0.0; map_out[1] = nvalid; }
  for (int r = tid; r < max_rank; r += blockDim.x) {
    float hits = 0.f;      // the reference sums float32 0/1 rows
    for (int q = 0; q < nq; ++q) hits += (stats[3 * q] != 0.0 && stats[3 * q + 2] <= (double)r) ? 1.f : 0.f;
    cmc[r] = nvalid > 0 ? hits / (float)nvalid : 0.f;
  }
}

}  // namespace

int infer_features(const void* const cls[3], const int64_t cls_stride_b[3], const void* sim_out, int64_t ld_sim, int dtype, int B, int d,
                   int normalize, float* out, cudaStream_t s) {
  if (dtype == SIG_BF16) {
    using T = __nv_bfloat16;
    SIG_LAUNCH((infer_features_kernel<T>), B, 256, 0, s, (const T*)cls[0], (const T*)cls[1], (const T*)cls[2], cls_stride_b[0],
               cls_stride_b[1], cls_stride_b[2], (const T*)sim_out, ld_sim, d, normalize, out);
  } else {
    SIG_LAUNCH((infer_features_kernel<float>), B, 256, 0, s, (const float*)cls[0], (const float*)cls[1], (const float*)cls[2],
               cls_stride_b[0], cls_stride_b[1], cls_stride_b[2], (const float*)sim_out, ld_sim, d, normalize, out);
  }
  SIG_CHECK_LAUNCH();
  return 0;
}

int euclidean_distmat(const float* qf, const float* gf, int nq, int ng, int D, float* dist, float* ws, cudaStream_t s) {
  float* qq = ws;
  float* gg = ws + nq;
  SIG_LAUNCH((rownorm2_kernel), nq, 256, 0, s, qf, (int64_t)D, D, qq);
  SIG_LAUNCH((rownorm2_kernel), ng, 256, 0, s, gf, (int64_t)D, D, gg);
  SIG_CHECK_LAUNCH();
  SIG_TRY(launch_gemm(gemm_nt(qf, D, gf, D, dist, ng, nullptr, nq, ng, D), s));
  const int64_t n = (int64_t)nq * ng;
  SIG_LAUNCH((dist_finish_kernel), (unsigned)ceil_div(n, 256), 256, 0, s, dist, qq, gg, nq, ng);
  SIG_CHECK_LAUNCH();
  return 0;
}

int rank_eval(const float* dist, int64_t ld, const int64_t* q_pids, const int64_t* g_pids, const int64_t* q_cams, const int64_t* g_cams,
              int nq, int ng, int max_rank, float* cmc, double* map_out, double* stats, int* overflow, cudaStream_t s) {
  cudaMemsetAsync(overflow, 0, sizeof(int), s);
  SIG_LAUNCH((rank_eval_kernel), nq, 512, 0, s, dist, ld, q_pids, g_pids, q_cams, g_cams, ng, stats, overflow);
  SIG_CHECK_LAUNCH();
  SIG_LAUNCH((rank_reduce_kernel), 1, 256, 0, s, stats, nq, max_rank, cmc, map_out);
  SIG_CHECK_LAUNCH();
  return 0;
}

}  // namespace sig

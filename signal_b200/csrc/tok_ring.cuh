// One-pass token kernels on the streaming ring (stream_ring.cuh): the selection scores of SIM (useA.py:50-221, four
// dot products per patch token) and the mean pool of GAM (useB.py:84-86).  Both read the three bf16 patch maps
// exactly once -- T*s bytes per sample -- and do a handful of FP32 operations per element, so the only thing that
// matters is bytes in flight: item = 32 consecutive token rows of one (modality, sample) = 48 KB at d = 768, four
// stages per SM.  Requires row-contiguous patches (stride_l == d), L == 128, d in {512, 768}; other layouts keep the
// register-staged kernels (sim_scores_tok_kernel, pool_tok_kernel).
#pragma once
#include "common.cuh"
#include "stream_ring.cuh"

namespace sig {

struct TokSrc3 {
  const void* patch[3];
  int64_t psb[3];   // elements between samples
};

template <int D>
struct TokRing {
  static constexpr int kRows = 32;                          // token rows per item
  static constexpr int kItemsPerGroup = kMaxL / kRows;      // a group = one (modality, sample): 128 rows
  static constexpr int kStageBytes = kRows * D * 2;
  static constexpr int kStages = (192 * 1024) / kStageBytes;
  static constexpr int kConsumers = 256;                    // 8 warps x 4 rows
  static constexpr int kThreads = kConsumers + 32;
  static constexpr int kChunks = D / 256;                   // 16-byte chunks per lane and row
  static constexpr int kQBytes = 4 * D * 4;                 // the four query vectors of a group (scores only)
};

// bf16 pair -> two floats (exact)
__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    v[2 * t] = __uint_as_float(w[t] << 16);
    v[2 * t + 1] = __uint_as_float(w[t] & 0xffff0000u);
  }
}

// ---- SIM selection scores ---------------------------------------------------------------------------------------
// sel_logits [B][3 queries][3 modalities][L] = (qt_r . x + c_r) / sqrt(d);  intra_raw [B][3][L] = cls_m . x
// grid = min(#SMs, items); a CTA owns the items [i0, i1) of the sequence ((b * 3 + m) * 4 + chunk).
// The producer also stages each group's four query vectors (qtsel rows 3b..3b+2 and cls row 3b+m, fp32) through a
// two-deep buffer; a lane keeps its 4 x 8 x kChunks query values in registers for the whole group.  The per-lane
// accumulation order (chunks lane, lane+32, .. ; 8 channels in order; then a butterfly) is that of
// sim_scores_tok_kernel, so both kernels return the same bits.
// kPool (FusionHead): the same pass also delivers GAM's mean pool of the patch rows (useB.py:84-86) --
// every token value is in a register here anyway, so AlignM's own pass over the 75 MB of tokens (pool_ring_kernel) goes
// away.  pool_mean [3][B][D] (zero-filled by the caller) receives sum / L per group exactly like pool_ring_kernel: the
// eight warps' partial sums meet in a [8][256] shared-memory tile in a fixed order, a group that straddles two CTAs gets
// two atomic adds onto zero (order-independent), so the result is deterministic.
template <int D, bool kPool>
static __global__ void __launch_bounds__(TokRing<D>::kThreads, 1)
sim_scores_ring_kernel(TokSrc3 src, const float* __restrict__ clsf, const float* __restrict__ qtsel, const float* __restrict__ csel,
                       int B, int n_items, float* __restrict__ sel_logits, float* __restrict__ intra_raw, float* __restrict__ pool_mean) {
  using R = TokRing<D>;
  constexpr int L = kMaxL;
  pdl_launch_dependents();
  extern __shared__ __align__(128) unsigned char tr_smem[];
  unsigned char* stages = tr_smem;
  float* qbuf = reinterpret_cast<float*>(tr_smem + (size_t)R::kStages * R::kStageBytes);   // [2][4][D]
  auto* bars = reinterpret_cast<ring::Bars<R::kStages>*>(tr_smem + (size_t)R::kStages * R::kStageBytes + 2 * R::kQBytes);
  uint64_t* qfull = reinterpret_cast<uint64_t*>(bars + 1);   // [2]
  uint64_t* qempty = qfull + 2;                               // [2]
  float* pred = reinterpret_cast<float*>(qempty + 2);         // [8 warps][256] (kPool only)
  const int i0 = (int)((int64_t)blockIdx.x * n_items / gridDim.x), i1 = (int)((int64_t)(blockIdx.x + 1) * n_items / gridDim.x);
  if (threadIdx.x == 0) {
    for (int q = 0; q < 2; ++q) {
      ptx::mbar_init(&qfull[q], 1);
      ptx::mbar_init(&qempty[q], R::kConsumers / 32);
    }
  }
  ring::init(bars, R::kConsumers / 32);
  if ((int)threadIdx.x >= R::kConsumers) {
    if ((int)threadIdx.x == R::kConsumers) {
      pdl_wait();   // the query vectors come from the kernels launched just before
      int gcount = 0;
      for (int it = i0, k = 0; it < i1; ++it, ++k) {
        const int g = it / R::kItemsPerGroup, chunk = it % R::kItemsPerGroup;
        const int b = g / 3, m = g % 3;
        if (k == 0 || chunk == 0) {
          const int qs = gcount & 1;
          if (gcount >= 2) ptx::mbar_wait(&qempty[qs], (uint32_t)((gcount >> 1) - 1) & 1u);
          ptx::mbar_expect_tx(&qfull[qs], R::kQBytes);
          ring::bulk_g2s(qbuf + (size_t)qs * 4 * D, qtsel + (int64_t)b * 3 * D, 3 * D * 4, &qfull[qs]);
          ring::bulk_g2s(qbuf + (size_t)qs * 4 * D + 3 * D, clsf + ((int64_t)b * 3 + m) * D, D * 4, &qfull[qs]);
          ++gcount;
        }
        const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(src.patch[m]) + b * src.psb[m] + (int64_t)chunk * R::kRows * D;
        ring::produce(bars, k, stages + (size_t)(k % R::kStages) * R::kStageBytes, x, R::kStageBytes, ptx::kPolTokens);
      }
    }
    return;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float inv = rsqrtf((float)D);
  float q[4][R::kChunks][8];
  float pacc[kPool ? R::kChunks : 1][8];
#pragma unroll
  for (int ch = 0; ch < (kPool ? R::kChunks : 1); ++ch)
#pragma unroll
    for (int t = 0; t < 8; ++t) pacc[ch][t] = 0.f;
  float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f;
  int gcount = 0;
  pdl_wait();
  for (int it = i0, k = 0; it < i1; ++it, ++k) {
    const int g = it / R::kItemsPerGroup, chunk = it % R::kItemsPerGroup;
    const int b = g / 3, m = g % 3;
    if (k == 0 || chunk == 0) {   // new group: its query vectors -> registers
      const int qs = gcount & 1;
      ptx::mbar_wait(&qfull[qs], (uint32_t)(gcount >> 1) & 1u);
      const float* qv = qbuf + (size_t)qs * 4 * D;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int ch = 0; ch < R::kChunks; ++ch) {
          const float4 lo = *reinterpret_cast<const float4*>(qv + r * D + (lane + 32 * ch) * 8);
          const float4 hi = *reinterpret_cast<const float4*>(qv + r * D + (lane + 32 * ch) * 8 + 4);
          q[r][ch][0] = lo.x; q[r][ch][1] = lo.y; q[r][ch][2] = lo.z; q[r][ch][3] = lo.w;
          q[r][ch][4] = hi.x; q[r][ch][5] = hi.y; q[r][ch][6] = hi.z; q[r][ch][7] = hi.w;
        }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&qempty[qs]);
      cs0 = csel[b * 3 + 0]; cs1 = csel[b * 3 + 1]; cs2 = csel[b * 3 + 2];
      ++gcount;
    }
    ring::consumer_wait(bars, k);
    const unsigned char* st = stages + (size_t)(k % R::kStages) * R::kStageBytes;
    float a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i][0] = a[i][1] = a[i][2] = a[i][3] = 0.f;
#pragma unroll
    for (int ch = 0; ch < R::kChunks; ++ch) {
      uint4 raw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        raw[i] = *reinterpret_cast<const uint4*>(st + ((size_t)(w * 4 + i) * D + (lane + 32 * ch) * 8) * 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float xv[8];
        unpack8(raw[i], xv);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int t = 0; t < 8; ++t) a[i][r] = fmaf(xv[t], q[r][ch][t], a[i][r]);
        if constexpr (kPool) {
#pragma unroll
          for (int t = 0; t < 8; ++t) pacc[ch][t] += xv[t];
        }
      }
    }
    ring::consumer_release(bars, k);
    if constexpr (kPool) {
      if (chunk == R::kItemsPerGroup - 1 || it == i1 - 1) {   // end of the group (or of this CTA's part of it): warp-uniform
        const bool whole = (it - i0) >= R::kItemsPerGroup - 1 && chunk == R::kItemsPerGroup - 1;   // all four items were ours
        float* dst = pool_mean + ((int64_t)m * B + b) * D;
#pragma unroll
        for (int ch = 0; ch < R::kChunks; ++ch) {
          float* pw = pred + w * 256 + lane * 8;
          *reinterpret_cast<float4*>(pw) = make_float4(pacc[ch][0], pacc[ch][1], pacc[ch][2], pacc[ch][3]);
          *reinterpret_cast<float4*>(pw + 4) = make_float4(pacc[ch][4], pacc[ch][5], pacc[ch][6], pacc[ch][7]);
#pragma unroll
          for (int t = 0; t < 8; ++t) pacc[ch][t] = 0.f;
          ring::consumer_sync(R::kConsumers);
          {
            const int i = threadIdx.x;      // 256 consumer threads: channel ch * 256 + i (lane j of a warp holds 8 j .. 8 j + 7)
            float t = 0.f;
#pragma unroll
            for (int qw = 0; qw < R::kConsumers / 32; ++qw) t += pred[qw * 256 + i];
            t *= 1.f / kMaxL;
            if (whole) dst[ch * 256 + i] = t;
            else atomicAdd(dst + ch * 256 + i, t);
          }
          ring::consumer_sync(R::kConsumers);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int r = 0; r < 4; ++r) a[i][r] = warp_sum(a[i][r]);
    if (lane < 4) {
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (lane == i) { v0 = a[i][0]; v1 = a[i][1]; v2 = a[i][2]; v3 = a[i][3]; }
      const int l = chunk * R::kRows + w * 4 + lane;
      const int64_t base = (int64_t)b * 3 * 3 * L + (int64_t)m * L + l;
      sel_logits[base] = (v0 + cs0) * inv;
      sel_logits[base + 3 * L] = (v1 + cs1) * inv;
      sel_logits[base + 6 * L] = (v2 + cs2) * inv;
      intra_raw[((int64_t)b * 3 + m) * L + l] = v3;
    }
  }
}

// ---- column-split variant: 8 * (D / 256) consumer warps --------------------------------------------------------------
// ncu (profiles/r1e_ring_kernels_full.md): with 8 consumer warps the kernel above issues 0.49 instructions per cycle
// and scheduler and a quarter of them are the 16 butterfly reductions per warp.  Here warp (rg, cb) owns rows 4rg..4rg+3
// and ONE 16-byte chunk per lane (columns 256 cb + 8 lane ..): 32 query registers instead of 96, three times the warps,
// a 16-shuffle reduce-scatter instead of 80 shuffles, and the D/256 partial dot products of a row group meet in shared
// memory behind a 32*CB-thread named barrier.  The summation order differs from sim_scores_tok_kernel (column blocks are
// added last), so the scores agree to fp32 rounding, not bit for bit.
template <int D>
struct ScoreSplit {
  using R = TokRing<D>;
  static constexpr int kCB = D / 256;                       // column blocks = warps per row group
  static constexpr int kConsumers = 256 * kCB;
  static constexpr int kThreads = kConsumers + 32;
  static constexpr int kPartFloats = 2 * 8 * kCB * 16;      // [item parity][row group][column block][16]
  static constexpr size_t kSmemBytes = (size_t)R::kStages * R::kStageBytes + 2 * R::kQBytes + sizeof(ring::Bars<R::kStages>) +
                                       4 * sizeof(uint64_t) + kPartFloats * sizeof(float) + 128;
};

// 16 per-lane partial sums -> their totals over the warp: lanes 2k and 2k+1 return the total of v[k]
__device__ __forceinline__ float warp_reduce_scatter16(float (&v)[16], int lane) {
  float a8[8], a4[4], a2[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float send = b4 ? v[k] : v[k + 8];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
    a8[k] = (b4 ? v[k + 8] : v[k]) + recv;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float send = b3 ? a8[k] : a8[k + 4];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
    a4[k] = (b3 ? a8[k + 4] : a8[k]) + recv;
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float send = b2 ? a4[k] : a4[k + 2];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
    a2[k] = (b2 ? a4[k + 2] : a4[k]) + recv;
  }
  const float send = b1 ? a2[0] : a2[1];
  const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
  float t = (b1 ? a2[1] : a2[0]) + recv;
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  return t;   // index k = 8*b4 + 4*b3 + 2*b2 + b1 = lane >> 1
}

template <int D>
static __global__ void __launch_bounds__(ScoreSplit<D>::kThreads, 1)
sim_scores_split_kernel(TokSrc3 src, const float* __restrict__ clsf, const float* __restrict__ qtsel, const float* __restrict__ csel,
                        int B, int n_items, float* __restrict__ sel_logits, float* __restrict__ intra_raw) {
  using R = TokRing<D>;
  using S = ScoreSplit<D>;
  constexpr int L = kMaxL;
  pdl_launch_dependents();
  extern __shared__ __align__(128) unsigned char tr_smem[];
  unsigned char* stages = tr_smem;
  float* qbuf = reinterpret_cast<float*>(tr_smem + (size_t)R::kStages * R::kStageBytes);   // [2][4][D]
  auto* bars = reinterpret_cast<ring::Bars<R::kStages>*>(tr_smem + (size_t)R::kStages * R::kStageBytes + 2 * R::kQBytes);
  uint64_t* qfull = reinterpret_cast<uint64_t*>(bars + 1);   // [2]
  uint64_t* qempty = qfull + 2;                               // [2]
  float* part = reinterpret_cast<float*>(qempty + 2);         // [2][8][kCB][16]
  const int i0 = (int)((int64_t)blockIdx.x * n_items / gridDim.x), i1 = (int)((int64_t)(blockIdx.x + 1) * n_items / gridDim.x);
  if (threadIdx.x == 0) {
    for (int q = 0; q < 2; ++q) {
      ptx::mbar_init(&qfull[q], 1);
      ptx::mbar_init(&qempty[q], S::kConsumers / 32);
    }
  }
  ring::init(bars, S::kConsumers / 32);
  if ((int)threadIdx.x >= S::kConsumers) {
    if ((int)threadIdx.x == S::kConsumers) {
      pdl_wait();   // the query vectors come from the kernels launched just before
      int gcount = 0;
      for (int it = i0, k = 0; it < i1; ++it, ++k) {
        const int g = it / R::kItemsPerGroup, chunk = it % R::kItemsPerGroup;
        const int b = g / 3, m = g % 3;
        if (k == 0 || chunk == 0) {
          const int qs = gcount & 1;
          if (gcount >= 2) ptx::mbar_wait(&qempty[qs], (uint32_t)((gcount >> 1) - 1) & 1u);
          ptx::mbar_expect_tx(&qfull[qs], R::kQBytes);
          ring::bulk_g2s(qbuf + (size_t)qs * 4 * D, qtsel + (int64_t)b * 3 * D, 3 * D * 4, &qfull[qs]);
          ring::bulk_g2s(qbuf + (size_t)qs * 4 * D + 3 * D, clsf + ((int64_t)b * 3 + m) * D, D * 4, &qfull[qs]);
          ++gcount;
        }
        const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(src.patch[m]) + b * src.psb[m] + (int64_t)chunk * R::kRows * D;
        ring::produce(bars, k, stages + (size_t)(k % R::kStages) * R::kStageBytes, x, R::kStageBytes, ptx::kPolTokens);
      }
    }
    return;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int rg = w & 7, cb = w >> 3;                 // row group (4 rows), column block (256 columns)
  const int col = (cb * 32 + lane) * 8;              // first of this lane's 8 columns
  const float inv = rsqrtf((float)D);
  float q[4][8];
  float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f;
  int gcount = 0;
  pdl_wait();
  for (int it = i0, k = 0; it < i1; ++it, ++k) {
    const int g = it / R::kItemsPerGroup, chunk = it % R::kItemsPerGroup;
    const int b = g / 3, m = g % 3;
    if (k == 0 || chunk == 0) {   // new group: this lane's slice of the query vectors -> registers
      const int qs = gcount & 1;
      ptx::mbar_wait(&qfull[qs], (uint32_t)(gcount >> 1) & 1u);
      const float* qv = qbuf + (size_t)qs * 4 * D;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 lo = *reinterpret_cast<const float4*>(qv + r * D + col);
        const float4 hi = *reinterpret_cast<const float4*>(qv + r * D + col + 4);
        q[r][0] = lo.x; q[r][1] = lo.y; q[r][2] = lo.z; q[r][3] = lo.w;
        q[r][4] = hi.x; q[r][5] = hi.y; q[r][6] = hi.z; q[r][7] = hi.w;
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&qempty[qs]);
      cs0 = csel[b * 3 + 0]; cs1 = csel[b * 3 + 1]; cs2 = csel[b * 3 + 2];
      ++gcount;
    }
    ring::consumer_wait(bars, k);
    const unsigned char* st = stages + (size_t)(k % R::kStages) * R::kStageBytes;
    uint4 raw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) raw[i] = *reinterpret_cast<const uint4*>(st + ((size_t)(rg * 4 + i) * D + col) * 2);
    ring::consumer_release(bars, k);      // (the stage's bytes are in registers)
    float a[16];                           // a[4 * i + r]: row i, query r
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float xv[8];
      unpack8(raw[i], xv);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float t = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) t = fmaf(xv[e], q[r][e], t);
        a[4 * i + r] = t;
      }
    }
    const float tot = warp_reduce_scatter16(a, lane);          // lanes 2k, 2k+1: total of a[k] over this column block
    float* pk = part + (size_t)(k & 1) * 8 * S::kCB * 16;
    if (!(lane & 1)) pk[(rg * S::kCB + cb) * 16 + (lane >> 1)] = tot;
    asm volatile("bar.sync %0, %1;" ::"r"(2 + rg), "r"(32 * S::kCB) : "memory");   // the kCB warps of this row group
    if (cb == 0 && lane < 16) {
      float v = 0.f;
#pragma unroll
      for (int c2 = 0; c2 < S::kCB; ++c2) v += pk[(rg * S::kCB + c2) * 16 + lane];
      const int i = lane >> 2, r = lane & 3;
      const int l = chunk * R::kRows + rg * 4 + i;
      if (r < 3) {
        const float cs = r == 0 ? cs0 : (r == 1 ? cs1 : cs2);
        sel_logits[(int64_t)b * 3 * 3 * L + (int64_t)r * 3 * L + (int64_t)m * L + l] = (v + cs) * inv;
      } else {
        intra_raw[((int64_t)b * 3 + m) * L + l] = v;
      }
    }
  }
}

template <int D>
static size_t sim_scores_ring_smem(bool pool = false) {
  using R = TokRing<D>;
  return (size_t)R::kStages * R::kStageBytes + 2 * R::kQBytes + sizeof(ring::Bars<R::kStages>) + 4 * sizeof(uint64_t) + 128 +
         (pool ? (R::kConsumers / 32) * 256 * sizeof(float) : 0);
}

// ---- GAM mean pool ----------------------------------------------------------------------------------------------
// mean [3][B][D] (zero-filled by the caller) += sum over the CTA's rows of the group / L.  Item order (m * B + b) * 4 +
// chunk.  A group that straddles two CTAs receives two atomic adds onto zero: order-independent, so the result is
// deterministic; within a CTA the eight warps' partial sums are added in a fixed order.
template <int D>
static __global__ void __launch_bounds__(TokRing<D>::kThreads, 1)
pool_ring_kernel(TokSrc3 src, int B, int n_items, float* __restrict__ mean) {
  using R = TokRing<D>;
  constexpr int kWarps = R::kConsumers / 32;
  pdl_launch_dependents();
  extern __shared__ __align__(128) unsigned char tr_smem[];
  unsigned char* stages = tr_smem;
  float* red = reinterpret_cast<float*>(tr_smem + (size_t)R::kStages * R::kStageBytes);   // [8][D]
  auto* bars = reinterpret_cast<ring::Bars<R::kStages>*>(tr_smem + (size_t)R::kStages * R::kStageBytes + kWarps * D * 4);
  const int i0 = (int)((int64_t)blockIdx.x * n_items / gridDim.x), i1 = (int)((int64_t)(blockIdx.x + 1) * n_items / gridDim.x);
  ring::init(bars, kWarps);
  if ((int)threadIdx.x >= R::kConsumers) {
    if ((int)threadIdx.x == R::kConsumers) {
      // (the tokens are inputs of the call: the ring fills under the previous kernel's tail)
      for (int it = i0, k = 0; it < i1; ++it, ++k) {
        const int g = it / R::kItemsPerGroup, chunk = it % R::kItemsPerGroup;
        const int m = g / B, b = g % B;
        const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(src.patch[m]) + b * src.psb[m] + (int64_t)chunk * R::kRows * D;
        ring::produce(bars, k, stages + (size_t)(k % R::kStages) * R::kStageBytes, x, R::kStageBytes, ptx::kPolTokens);
      }
    }
    return;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float acc[R::kChunks][8];
#pragma unroll
  for (int ch = 0; ch < R::kChunks; ++ch)
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[ch][t] = 0.f;
  pdl_wait();   // `mean` was zero-filled by the memset node in front of this kernel
  for (int it = i0, k = 0; it < i1; ++it, ++k) {
    const int g = it / R::kItemsPerGroup, chunk = it % R::kItemsPerGroup;
    ring::consumer_wait(bars, k);
    const unsigned char* st = stages + (size_t)(k % R::kStages) * R::kStageBytes;
#pragma unroll
    for (int ch = 0; ch < R::kChunks; ++ch) {
      uint4 raw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        raw[i] = *reinterpret_cast<const uint4*>(st + ((size_t)(w * 4 + i) * D + (lane + 32 * ch) * 8) * 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float xv[8];
        unpack8(raw[i], xv);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[ch][t] += xv[t];
      }
    }
    ring::consumer_release(bars, k);
    if (chunk == R::kItemsPerGroup - 1 || it == i1 - 1) {   // end of the group (or of this CTA's part of it)
#pragma unroll
      for (int ch = 0; ch < R::kChunks; ++ch) {
        float* dst = red + w * D + (lane + 32 * ch) * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(acc[ch][0], acc[ch][1], acc[ch][2], acc[ch][3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[ch][4], acc[ch][5], acc[ch][6], acc[ch][7]);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[ch][t] = 0.f;
      }
      ring::consumer_sync(R::kConsumers);
      const bool whole = (it - i0) >= R::kItemsPerGroup - 1 && chunk == R::kItemsPerGroup - 1;   // all four items were ours
      for (int i = threadIdx.x; i < D; i += R::kConsumers) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < kWarps; ++q) t += red[q * D + i];
        t *= 1.f / kMaxL;
        if (whole) mean[(int64_t)g * D + i] = t;
        else atomicAdd(mean + (int64_t)g * D + i, t);
      }
      ring::consumer_sync(R::kConsumers);
    }
  }
}

template <int D>
static size_t pool_ring_smem() {
  using R = TokRing<D>;
  return (size_t)R::kStages * R::kStageBytes + (R::kConsumers / 32) * D * 4 + sizeof(ring::Bars<R::kStages>) + 128;
}

// SIG_SCORES_SPLIT=1 selects the column-split score kernel.  Parity-verified on a B200 (tests/test_gpu_parity.py, 29 passed)
// but NOT faster there: 30.5 us vs 28.5 us event-timed at B = 128, d = 768 -- so the issue rate of the 8-warp kernel is not
// what bounds it either; off by default until an ncu capture explains it.
inline bool scores_split_enabled() {
  static const bool on = [] {
    const char* e = getenv("SIG_SCORES_SPLIT");
    return e && e[0] == '1';
  }();
  return on;
}

inline bool tok_ring_enabled() {
  static const bool on = [] {
    const char* e = getenv("SIG_TOK_RING");
    return !(e && e[0] == '0');
  }();
  return on;
}

}  // namespace sig

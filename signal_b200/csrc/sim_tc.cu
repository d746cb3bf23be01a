// SIM cross-modal attention over the (masked) tokens on tcgen05/TMEM tiles fed by TMA -- bf16 path.
//
// With the K/V projections folded into 24 effective queries per sample (e = q*8 + h)
// the token-side work of ModalInteractive (useA.py:383-388) is four thin GEMMs whose wide operand is
// the token map itself, read in place through 3-D tensor maps:
//   fwd  A: logits  S[j,e]   = mask_j * (x_j . qt_e) + c_e            M = tokens, N = 32, K = d
//        B: xbar[e,:]        = sum_j P~[e,j] x_j                      M = d (channels), N = 32, K = 384 tokens
//   bwd  C: dP[j,e]          = x_j . dxbar_e  ->  dS~ = P~ (dP - delta)
//        D: dx_j             = sum_e P~[j,e] dxbar_e + dS~[j,e] qt_e  M = tokens, N = d, K = 64
//        E: dqt[e,:]         = sum_j dS~[e,j] x_j                     (same kernel as B)
// P~ = softmax * mask.  All four are HBM-bound (one pass over the tokens each); the MMAs are N = 32
// (A, B, C, E) or K = 64 (D) and ride along for free.
#include "sim_tc.h"

#include "tc_pipeline.cuh"

namespace sig {

namespace {

using tc::BK;
using tc::BM;

// ------------------------------------------------------------------------------------------------
// A / C : per-token logits against the sample's 32 query rows.  units = (sample, modality)
// ------------------------------------------------------------------------------------------------
struct RowsParams {
  CUtensorMap ta[3];   // token maps, box [1 x 128 x 64]
  CUtensorMap tb;      // DXQT_all [B*64, d], box [32 x 64]
  int B, d, L;
  int b_row_off;       // 32 -> query rows (fwd), 0 -> dxbar rows (bwd)
  int mode;            // 0: logits (A), 1: dS (C)
  const float* maskf;  // [3][B][L]
  const float* catt;   // [B][24]
  float* S32;          // [B][384][32]
  const __nv_bfloat16* Ptok;  // [B][384][32]
  const float* delta;  // [B][32]
  __nv_bfloat16* PdS;  // [B][384][64]
  __nv_bfloat16* dST;  // [B][32][384]
};

struct RowsProblem {
  using Params = RowsParams;
  static constexpr int kAMn = 0, kBMn = 0;
  __device__ static void prefetch(const Params& p) {
    for (int z = 0; z < 3; ++z) ptx::prefetch_tmap(&p.ta[z]);
    ptx::prefetch_tmap(&p.tb);
  }
  __device__ static int num_units(const Params& p) { return 3 * p.B; }
  __device__ static void krange(const Params& p, int, int& kb0, int& kb1) { kb0 = 0; kb1 = p.d / BK; }
  __device__ static void load(const Params& p, int unit, int kb, uint8_t* sa, uint8_t* sb, uint64_t* bar) {
    const int b = unit / 3, z = unit % 3;
    tc::load_kmajor_tok(&p.ta[z], sa, bar, kb * BK, b, 1);
    tc::load_kmajor_2d(&p.tb, sb, bar, kb * BK, b * 64 + p.b_row_off);
  }
  __device__ static void epilogue(const Params& p, int unit, int /*mt*/, uint32_t tmem_acc, int q, int lane) {
    const int b = unit / 3, z = unit % 3;
    const int l = q * 32 + lane;
    const int j = z * 128 + l;
    uint32_t r[32];
    ptx::tmem_ld32(tmem_acc, r);
    ptx::tmem_ld_wait();
    if (l >= p.L) return;
    if (p.mode == 0) {
      const float mk = p.maskf[((int64_t)z * p.B + b) * p.L + l];
      float* dst = p.S32 + ((int64_t)b * 384 + j) * 32;
      const float* cq = p.catt + (int64_t)b * 24;
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        float4 v;
        v.x = e + 0 < 24 ? fmaf(mk, __uint_as_float(r[e + 0]), cq[e + 0]) : 0.f;
        v.y = e + 1 < 24 ? fmaf(mk, __uint_as_float(r[e + 1]), cq[e + 1]) : 0.f;
        v.z = e + 2 < 24 ? fmaf(mk, __uint_as_float(r[e + 2]), cq[e + 2]) : 0.f;
        v.w = e + 3 < 24 ? fmaf(mk, __uint_as_float(r[e + 3]), cq[e + 3]) : 0.f;
        *reinterpret_cast<float4*>(dst + e) = v;
      }
    } else {
      const __nv_bfloat16* pt = p.Ptok + ((int64_t)b * 384 + j) * 32;
      __nv_bfloat16* out = p.PdS + ((int64_t)b * 384 + j) * 64;
      const float* dl = p.delta + (int64_t)b * 32;
      float ds[32];
#pragma unroll
      for (int e = 0; e < 32; e += 8) {
        float pv[8];
        load8(pt + e, pv);
        store8(out + e, pv);                      // P~ half of the [P~ | dS~] operand
#pragma unroll
        for (int i = 0; i < 8; ++i) ds[e + i] = e + i < 24 ? pv[i] * (__uint_as_float(r[e + i]) - dl[e + i]) : 0.f;
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = ds[e + i];
        store8(out + 32 + e, t);
      }
      __nv_bfloat16* dt = p.dST + (int64_t)b * 32 * 384 + j;
#pragma unroll
      for (int e = 0; e < 32; ++e) dt[(int64_t)e * 384] = __float2bfloat16_rn(ds[e]);
    }
  }
};

// ------------------------------------------------------------------------------------------------
// B / E : out[b][e][c] = sum_j W[b][e][j] x_j[c],  W = P~T or dS~T ([B*32, 384] K-major).
// units = (sample, 128-channel tile); K runs over the 3 x 128 tokens of the sample.
// ------------------------------------------------------------------------------------------------
struct ColsParams {
  CUtensorMap ta[3];   // token maps, box [1 x 64 x 64] (MN-major operand: 64 token rows x 64 channels)
  CUtensorMap tb;      // W_all [B*32, 384], box [32 x 64]
  int B, d;
  float* out;          // [B][24][d]
};

struct ColsProblem {
  using Params = ColsParams;
  static constexpr int kAMn = 1, kBMn = 0;
  __device__ static void prefetch(const Params& p) {
    for (int z = 0; z < 3; ++z) ptx::prefetch_tmap(&p.ta[z]);
    ptx::prefetch_tmap(&p.tb);
  }
  __device__ static int num_units(const Params& p) { return p.B * ((p.d + 127) / 128); }
  __device__ static void krange(const Params&, int, int& kb0, int& kb1) { kb0 = 0; kb1 = 6; }
  __device__ static void load(const Params& p, int unit, int kb, uint8_t* sa, uint8_t* sb, uint64_t* bar) {
    const int mt = (p.d + 127) / 128;
    const int b = unit / mt, c0 = (unit % mt) * 128;
    tc::load_mnmajor_tok(&p.ta[kb >> 1], sa, bar, c0, (kb & 1) * 64, b, 128);
    tc::load_kmajor_2d(&p.tb, sb, bar, kb * BK, b * 32);
  }
  __device__ static void epilogue(const Params& p, int unit, int /*mt*/, uint32_t tmem_acc, int q, int lane) {
    const int mt = (p.d + 127) / 128;
    const int b = unit / mt, c = (unit % mt) * 128 + q * 32 + lane;
    uint32_t r[32];
    ptx::tmem_ld32(tmem_acc, r);
    ptx::tmem_ld_wait();
    if (c >= p.d) return;
    float* dst = p.out + (int64_t)b * 24 * p.d + c;
#pragma unroll
    for (int e = 0; e < 24; ++e) dst[(int64_t)e * p.d] = __uint_as_float(r[e]);
  }
};

// ------------------------------------------------------------------------------------------------
// D : dx[(b,z,l), :] = [P~ | dS~][b, z*128+l, 0:64] . [dxbar ; qt][b]   -> bf16 at the token strides
// units = (sample, modality, n-tile); one k-block.
// ------------------------------------------------------------------------------------------------
struct DxParams {
  CUtensorMap ta;      // PdS_all [B*384, 64], box [128 x 64]
  CUtensorMap tb;      // DXQT_all [B*64, d] read as MN-major: box [64 x 64]
  int B, d, L;
  void* dpatch[3];
  long long psb[3], psl[3];
  int accumulate;
};

template <int BN>
struct DxProblem {
  using Params = DxParams;
  static constexpr int kAMn = 0, kBMn = 1;
  __device__ static void prefetch(const Params& p) {
    ptx::prefetch_tmap(&p.ta);
    ptx::prefetch_tmap(&p.tb);
  }
  __device__ static int num_units(const Params& p) { return 3 * p.B * ((p.d + BN - 1) / BN); }
  __device__ static void krange(const Params&, int, int& kb0, int& kb1) { kb0 = 0; kb1 = 1; }
  __device__ static void load(const Params& p, int unit, int, uint8_t* sa, uint8_t* sb, uint64_t* bar) {
    const int nt = (p.d + BN - 1) / BN;
    const int n0 = (unit % nt) * BN, bz = unit / nt, b = bz / 3, z = bz % 3;
    tc::load_kmajor_2d(&p.ta, sa, bar, 0, b * 384 + z * 128);
    tc::load_mnmajor_2d(&p.tb, sb, bar, n0, b * 64, BN);
  }
  __device__ static void epilogue(const Params& p, int unit, int /*mt*/, uint32_t tmem_acc, int q, int lane) {
    const int nt = (p.d + BN - 1) / BN;
    const int n0 = (unit % nt) * BN, bz = unit / nt, b = bz / 3, z = bz % 3;
    const int l = q * 32 + lane;
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.dpatch[z]) + b * p.psb[z] + l * p.psl[z];
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      ptx::tmem_ld32(tmem_acc + c * 32, r);
      ptx::tmem_ld_wait();
      const int col0 = n0 + c * 32;
      if (l < p.L && col0 < p.d) {
#pragma unroll
        for (int jj = 0; jj < 32; jj += 8) {
          float t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = __uint_as_float(r[jj + i]);
          if (p.accumulate) {
            float o[8];
            load8(dst + col0 + jj, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] += o[i];
          }
          store8(dst + col0 + jj, t);
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// SIMT glue
// ------------------------------------------------------------------------------------------------
// DXQT[b][32 + e][:] = bf16(qt[b][e][:]) (e < 24), zero rows elsewhere in [32, 64).  grid (B, 32)
static __global__ void build_qt_kernel(const float* __restrict__ qt, __nv_bfloat16* __restrict__ DXQT, int d) {
  pdl_enter();
  const int b = blockIdx.x, e = blockIdx.y;
  __nv_bfloat16* dst = DXQT + ((int64_t)b * 64 + 32 + e) * d;
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (e < 24) load8(qt + ((int64_t)b * 24 + e) * d + c, v);
    store8(dst + c, v);
  }
}

// DXQT[b][e][:] = bf16(dxbar[b][e][:]) (e < 24), zero rows in [24, 32); delta[b][e] = dxbar_e . xbar_e.  grid (B, 32)
static __global__ void __launch_bounds__(128) build_dx_kernel(const float* __restrict__ dxbar, const float* __restrict__ xbar,
                                                              __nv_bfloat16* __restrict__ DXQT, float* __restrict__ delta, int d) {
  pdl_enter();
  __shared__ float scratch[33];
  const int b = blockIdx.x, e = blockIdx.y;
  __nv_bfloat16* dst = DXQT + ((int64_t)b * 64 + e) * d;
  float acc = 0.f;
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (e < 24) {
      float x[8];
      load8(dxbar + ((int64_t)b * 24 + e) * d + c, v);
      load8(xbar + ((int64_t)b * 24 + e) * d + c, x);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(v[i], x[i], acc);
    }
    store8(dst + c, v);
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) delta[b * 32 + e] = acc;
}

// softmax over the 384 tokens for each of the 24 effective queries of sample b; writes P~ = P * mask in
// both layouts (token-major [384][32] and query-major [32][384], bf16).  grid B, 256 threads, dyn smem 384*33 floats
static __global__ void __launch_bounds__(256) sim_softmax_kernel(const float* __restrict__ S32, const float* __restrict__ maskf, int B,
                                                                 int L, __nv_bfloat16* __restrict__ Ptok, __nv_bfloat16* __restrict__ PT) {
  pdl_enter();
  extern __shared__ float sm[];   // [384][33]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* src = S32 + (int64_t)b * 384 * 32;
  for (int i = tid; i < 384 * 32; i += blockDim.x) sm[(i >> 5) * 33 + (i & 31)] = src[i];
  __syncthreads();
  for (int e = w; e < 32; e += 8) {
    __nv_bfloat16* prow = PT + ((int64_t)b * 32 + e) * 384;
    if (e >= 24) {
      for (int j = lane; j < 384; j += 32) {
        prow[j] = __float2bfloat16_rn(0.f);
        sm[j * 33 + e] = 0.f;
      }
      continue;
    }
    float mx = -INFINITY;
    for (int j = lane; j < 384; j += 32)
      if (j % 128 < L) mx = fmaxf(mx, sm[j * 33 + e]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < 384; j += 32) {
      const float ev = (j % 128 < L) ? expf(sm[j * 33 + e] - mx) : 0.f;
      sm[j * 33 + e] = ev;
      sum += ev;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < 384; j += 32) {
      const int z = j >> 7, l = j & 127;
      const float mk = l < L ? maskf[((int64_t)z * B + b) * L + l] : 0.f;
      const float pv = sm[j * 33 + e] * inv * mk;
      sm[j * 33 + e] = pv;
      prow[j] = __float2bfloat16_rn(pv);
    }
  }
  __syncthreads();
  for (int j = tid; j < 384; j += blockDim.x) {
    __nv_bfloat16* dst = Ptok + ((int64_t)b * 384 + j) * 32;
#pragma unroll
    for (int e = 0; e < 32; e += 8) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = sm[j * 33 + e + i];
      store8(dst + e, t);
    }
  }
}

int make_tok_maps(const sig_tokens* tok, int box_rows, CUtensorMap* out) {
  for (int z = 0; z < 3; ++z)
    SIG_TRY(tc::make_map_tok(tok->patch[z], tok->B, tok->d, tok->patch_stride_b[z], tok->patch_stride_l[z], box_rows, &out[z]));
  return 0;
}

}  // namespace

size_t sim_tc_softmax_smem() { return (size_t)384 * 33 * sizeof(float); }

int sim_tc_tokens_fwd(const sig_tokens* tok, const SimTcBufs& k, cudaStream_t s) {
  const int B = tok->B, d = tok->d, L = tok->L;
  SIG_LAUNCH((build_qt_kernel), dim3(B, 32), 96, 0, s, k.qtatt, k.DXQT, d);
  SIG_CHECK_LAUNCH();
  {
    RowsParams p{};
    SIG_TRY(make_tok_maps(tok, 128, p.ta));
    SIG_TRY(tc::make_map_2d(k.DXQT, (int64_t)B * 64, d, d, 32, &p.tb));
    p.B = B; p.d = d; p.L = L; p.b_row_off = 32; p.mode = 0;
    p.maskf = k.maskf; p.catt = k.catt; p.S32 = k.S32;
    SIG_TRY((tc::launch<32, RowsProblem>(p, 3 * B, s, d / 64)));
  }
  {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(sim_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sim_tc_softmax_smem());
      attr = true;
    }
    SIG_LAUNCH((sim_softmax_kernel), B, 256, sim_tc_softmax_smem(), s, k.S32, k.maskf, B, L, k.Ptok, k.PT);
    SIG_CHECK_LAUNCH();
  }
  {
    ColsParams p{};
    SIG_TRY(make_tok_maps(tok, 64, p.ta));
    SIG_TRY(tc::make_map_2d(k.PT, (int64_t)B * 32, 384, 384, 32, &p.tb));
    p.B = B; p.d = d; p.out = k.xbar;
    SIG_TRY((tc::launch<32, ColsProblem>(p, B * (int)ceil_div(d, 128), s, 6)));
  }
  return 0;
}

int sim_tc_tokens_bwd(const sig_tokens* tok, const SimTcBufs& k, const sig_token_grads* dtok, cudaStream_t s) {
  const int B = tok->B, d = tok->d, L = tok->L;
  SIG_LAUNCH((build_dx_kernel), dim3(B, 32), 96, 0, s, k.dxbar, k.xbar, k.DXQT, k.delta, d);
  SIG_CHECK_LAUNCH();
  {
    RowsParams p{};
    SIG_TRY(make_tok_maps(tok, 128, p.ta));
    SIG_TRY(tc::make_map_2d(k.DXQT, (int64_t)B * 64, d, d, 32, &p.tb));
    p.B = B; p.d = d; p.L = L; p.b_row_off = 0; p.mode = 1;
    p.Ptok = k.Ptok; p.delta = k.delta; p.PdS = k.PdS; p.dST = k.dST;
    SIG_TRY((tc::launch<32, RowsProblem>(p, 3 * B, s, d / 64)));
  }
  if (dtok->wait_event) cudaStreamWaitEvent(s, (cudaEvent_t)dtok->wait_event, 0);
  {
    DxParams p{};
    SIG_TRY(tc::make_map_2d(k.PdS, (int64_t)B * 384, 64, 64, 128, &p.ta));
    SIG_TRY(tc::make_map_2d(k.DXQT, (int64_t)B * 64, d, d, 64, &p.tb));
    p.B = B; p.d = d; p.L = L;
    for (int z = 0; z < 3; ++z) {
      p.dpatch[z] = dtok->dpatch[z];
      p.psb[z] = dtok->patch_stride_b[z];
      p.psl[z] = dtok->patch_stride_l[z];
    }
    p.accumulate = dtok->accumulate;
    if (d % 256 == 0) SIG_TRY((tc::launch<256, DxProblem<256>>(p, 3 * B * (d / 256), s, 1)));
    else SIG_TRY((tc::launch<128, DxProblem<128>>(p, 3 * B * (int)ceil_div(d, 128), s, 1)));
  }
  {
    ColsParams p{};
    SIG_TRY(make_tok_maps(tok, 64, p.ta));
    SIG_TRY(tc::make_map_2d(k.dST, (int64_t)B * 32, 384, 384, 32, &p.tb));
    p.B = B; p.d = d; p.out = k.dqt;
    SIG_TRY((tc::launch<32, ColsProblem>(p, B * (int)ceil_div(d, 128), s, 6)));
  }
  return 0;
}

}  // namespace sig

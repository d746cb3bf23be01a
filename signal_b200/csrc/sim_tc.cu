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

#include "prof.h"
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
  struct Unit {
    int b, z, kb0, kb1;
  };
  __device__ static Unit unit_info(const Params& p, int unit) {
    Unit u;
    u.b = unit / 3; u.z = unit - 3 * u.b; u.kb0 = 0; u.kb1 = p.d / BK;
    return u;
  }
  __device__ static void load_a(const Params& p, const Unit& u, int kb, uint8_t* sa, uint64_t* bar) {
    tc::load_kmajor_tok(&p.ta[u.z], sa, bar, kb * BK, u.b, 1);
  }
  __device__ static void load_b(const Params& p, const Unit& u, int kb, uint8_t* sb, uint64_t* bar) {
    tc::load_kmajor_2d(&p.tb, sb, bar, kb * BK, u.b * 64 + p.b_row_off);
  }
  __device__ static void epilogue(const Params& p, const Unit& u, int /*mt*/, uint32_t tmem_acc, int q, int half, int lane, float* /*scratch*/,
                                  uint64_t* acc_bar, uint32_t acc_phase, long long* /*stamps*/) {
    ptx::mbar_wait(acc_bar, acc_phase);   // (also the idle half: its arrival on tmem_empty must not run a phase ahead)
    ptx::tc_fence_after();
    if (half) return;   // a single 32-column chunk: the second set of epilogue warps has nothing to do
    const int b = u.b, z = u.z;
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
  __nv_bfloat16* out_b;   // same layout: bf16 shadow feeding the next GEMM (may be null)
};

struct ColsProblem {
  using Params = ColsParams;
  static constexpr int kAMn = 1, kBMn = 0;
  __device__ static void prefetch(const Params& p) {
    for (int z = 0; z < 3; ++z) ptx::prefetch_tmap(&p.ta[z]);
    ptx::prefetch_tmap(&p.tb);
  }
  __device__ static int num_units(const Params& p) { return p.B * ((p.d + 127) / 128); }
  struct Unit {
    int b, c0, kb0, kb1;
  };
  __device__ static Unit unit_info(const Params& p, int unit) {
    Unit u;
    const int mt = (p.d + 127) / 128;
    u.b = unit / mt; u.c0 = (unit - u.b * mt) * 128; u.kb0 = 0; u.kb1 = 6;
    return u;
  }
  __device__ static void load_a(const Params& p, const Unit& u, int kb, uint8_t* sa, uint64_t* bar) {
    tc::load_mnmajor_tok(&p.ta[kb >> 1], sa, bar, u.c0, (kb & 1) * 64, u.b, 128);
  }
  __device__ static void load_b(const Params& p, const Unit& u, int kb, uint8_t* sb, uint64_t* bar) {
    tc::load_kmajor_2d(&p.tb, sb, bar, kb * BK, u.b * 32);
  }
  __device__ static void epilogue(const Params& p, const Unit& u, int /*mt*/, uint32_t tmem_acc, int q, int half, int lane, float* /*scratch*/,
                                  uint64_t* acc_bar, uint32_t acc_phase, long long* /*stamps*/) {
    ptx::mbar_wait(acc_bar, acc_phase);   // (also the idle half: its arrival on tmem_empty must not run a phase ahead)
    ptx::tc_fence_after();
    if (half) return;   // a single 32-column chunk: the second set of epilogue warps has nothing to do
    const int b = u.b, c = u.c0 + q * 32 + lane;
    uint32_t r[32];
    ptx::tmem_ld32(tmem_acc, r);
    ptx::tmem_ld_wait();
    if (c >= p.d) return;
    float* dst = p.out + (int64_t)b * 24 * p.d + c;
#pragma unroll
    for (int e = 0; e < 24; ++e) dst[(int64_t)e * p.d] = __uint_as_float(r[e]);
    if (p.out_b) {
      __nv_bfloat16* db = p.out_b + (int64_t)b * 24 * p.d + c;
#pragma unroll
      for (int e = 0; e < 24; ++e) db[(int64_t)e * p.d] = __float2bfloat16_rn(__uint_as_float(r[e]));
    }
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
  struct Unit {
    int b, z, n0, kb0, kb1;
  };
  __device__ static Unit unit_info(const Params& p, int unit) {
    Unit u;
    const int nt = (p.d + BN - 1) / BN;
    const int bz = unit / nt;
    u.n0 = (unit - bz * nt) * BN; u.b = bz / 3; u.z = bz - 3 * u.b; u.kb0 = 0; u.kb1 = 1;
    return u;
  }
  __device__ static void load_a(const Params& p, const Unit& u, int, uint8_t* sa, uint64_t* bar) {
    tc::load_kmajor_2d(&p.ta, sa, bar, 0, u.b * 384 + u.z * 128);
  }
  __device__ static void load_b(const Params& p, const Unit& u, int, uint8_t* sb, uint64_t* bar) {
    tc::load_mnmajor_2d(&p.tb, sb, bar, u.n0, u.b * 64, BN);
  }
  // one k-block per unit: the epilogue IS the kernel.  Each 32 x 32 block goes through the transpose
  // tile so that the read-modify-write of the gradient map is 4 rows x 64 contiguous bytes per warp
  // instruction, with all 8 old-value loads of a chunk issued before the first add.
  __device__ static void epilogue(const Params& p, const Unit& u, int /*mt*/, uint32_t tmem_acc, int q, int half, int lane, float* scratch,
                                  uint64_t* acc_bar, uint32_t acc_phase, long long* /*stamps*/) {
    const int n0 = u.n0, z = u.z;
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    const int l0 = q * 32 + sub;                                   // this lane's token row in step i is l0 + 4 i
    const long long step = 4 * p.psl[z];
    __nv_bfloat16* const base = static_cast<__nv_bfloat16*>(p.dpatch[z]) + u.b * p.psb[z] + l0 * p.psl[z];
    const int nrows = p.L - l0;
    const int accumulate = p.accumulate;
    bool waited = false;
    constexpr int kChunks = BN / 32 / 2;
#pragma unroll 1
    for (int c = half * kChunks; c < (half + 1) * kChunks; ++c) {
      const int col = n0 + c * 32 + c4;
      if (n0 + c * 32 >= p.d) break;   // warp-uniform
      const bool colok = col + 3 < p.d;   // d is a multiple of 8 on this path
      __nv_bfloat16* const dst = base + col;
      uint2 old[8];                      // old values (zeros when overwriting): requested first, they overlap the TMEM read
#pragma unroll
      for (int i = 0; i < 8; ++i)
        old[i] = (accumulate && 4 * i < nrows && colok) ? *reinterpret_cast<const uint2*>(dst + i * step) : make_uint2(0u, 0u);
      if (!waited) {                     // first chunk: its old-value loads are already in flight
        ptx::mbar_wait(acc_bar, acc_phase);
        ptx::tc_fence_after();
        waited = true;
      }
      uint32_t r[32];
      ptx::tmem_ld32(tmem_acc + c * 32, r);
      ptx::tmem_ld_wait();
      {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        __syncwarp();
        tc::epi_put_row(scratch, lane, v);
        __syncwarp();
      }
      if (p.L == 128 && n0 + c * 32 + 32 <= p.d) {   // warp-uniform: whole 32 x 32 block in range -> branch-free stores
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t4 = tc::epi_get(scratch, lane, i);
          float t[4] = {t4.x, t4.y, t4.z, t4.w};
          t[0] += __uint_as_float(old[i].x << 16); t[1] += __uint_as_float(old[i].x & 0xffff0000u);
          t[2] += __uint_as_float(old[i].y << 16); t[3] += __uint_as_float(old[i].y & 0xffff0000u);
          store4(dst + i * step, t);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t4 = tc::epi_get(scratch, lane, i);
          float t[4] = {t4.x, t4.y, t4.z, t4.w};
          t[0] += __uint_as_float(old[i].x << 16); t[1] += __uint_as_float(old[i].x & 0xffff0000u);
          t[2] += __uint_as_float(old[i].y << 16); t[3] += __uint_as_float(old[i].y & 0xffff0000u);
          if (4 * i < nrows && colok) store4(dst + i * step, t);
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
// both layouts (token-major [384][32] and query-major [32][384], bf16).  grid B, 1024 threads (one warp per
// query row: the kernel is one CTA per sample, i.e. latency-bound, so it wants every warp it can get), dyn smem 384*33 floats
static __global__ void __launch_bounds__(1024) sim_softmax_kernel(const float* __restrict__ S32, const float* __restrict__ maskf, int B,
                                                                 int L, __nv_bfloat16* __restrict__ Ptok, __nv_bfloat16* __restrict__ PT) {
  pdl_enter();
  extern __shared__ float sm[];   // [384][33]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* src = S32 + (int64_t)b * 384 * 32;
  for (int i = tid; i < 384 * 32; i += blockDim.x) sm[(i >> 5) * 33 + (i & 31)] = src[i];
  __syncthreads();
  for (int e = w; e < 32; e += (int)(blockDim.x >> 5)) {   // one warp per query row
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
    SIG_PHASE("sim_attn_logits_fwd");
    SIG_TRY((tc::launch<32, RowsProblem>(p, 3 * B, s, d / 64)));
  }
  {
    ensure_dyn_smem(sim_softmax_kernel, (int)sim_tc_softmax_smem());
    SIG_LAUNCH((sim_softmax_kernel), B, 1024, sim_tc_softmax_smem(), s, k.S32, k.maskf, B, L, k.Ptok, k.PT);
    SIG_CHECK_LAUNCH();
  }
  {
    ColsParams p{};
    SIG_TRY(make_tok_maps(tok, 64, p.ta));
    SIG_TRY(tc::make_map_2d(k.PT, (int64_t)B * 32, 384, 384, 32, &p.tb));
    p.B = B; p.d = d; p.out = k.xbar; p.out_b = k.xbarb;
    SIG_PHASE("sim_attn_pool_fwd");
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
    SIG_PHASE("sim_attn_dlogits_bwd");
    SIG_TRY((tc::launch<32, RowsProblem>(p, 3 * B, s, d / 64)));
  }
  if (dtok->fuse_skip_dx) {
    // AlignM's dX GEMM consumes [P~ | dS~] (PdS) and [dxbar ; qt] (DXQT) as one more k-block: they are complete here
    if (dtok->done_event) cudaEventRecord((cudaEvent_t)dtok->done_event, s);
  } else {
    if (dtok->wait_event) cudaStreamWaitEvent(s, (cudaEvent_t)dtok->wait_event, 0);
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
    SIG_PHASE("sim_attn_dx_bwd");
    if (d % 256 == 0) SIG_TRY((tc::launch<256, DxProblem<256>>(p, 3 * B * (d / 256), s, 1)));
    else SIG_TRY((tc::launch<128, DxProblem<128>>(p, 3 * B * (int)ceil_div(d, 128), s, 1)));
    // the patch rows of the shared gradient map are complete (the CLS rows belong to SIM alone)
    if (dtok->done_event) cudaEventRecord((cudaEvent_t)dtok->done_event, s);
  }
  {
    ColsParams p{};
    SIG_TRY(make_tok_maps(tok, 64, p.ta));
    SIG_TRY(tc::make_map_2d(k.dST, (int64_t)B * 32, 384, 384, 32, &p.tb));
    p.B = B; p.d = d; p.out = k.dqt; p.out_b = k.dqtb;
    SIG_PHASE("sim_attn_dq_bwd");
    SIG_TRY((tc::launch<32, ColsProblem>(p, B * (int)ceil_div(d, 128), s, 6)));
  }
  return 0;
}

}  // namespace sig

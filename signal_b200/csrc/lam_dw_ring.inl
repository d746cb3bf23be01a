// LAM offset-net tail (GELU -> depthwise 4x4/s4 -> GELU -> 1x1 -> 1, DAS.py:60-72) on the streaming ring
// (stream_ring.cuh).  Included by align_tc.inl.
//
// The 4x4 / stride-4 windows tile the h x w grid exactly, so a *band* of 4 grid rows (4*W tokens, contiguous in
// H [3][B*L][D] bf16) holds W/4 complete windows: item = band, 48 KB at D = 768, W = 8.  A CTA belongs to ONE
// modality (its 16 + 16 depthwise taps and the partial parameter sums stay in registers) and owns a contiguous
// range of that modality's B*h/4 bands; grid = 3 * 49 CTAs on 148 SMs.  Consumer thread t (of a group: the windows of
// a band can be split over two groups, see kDwFwdGroups) owns channels 2t, 2t+1 and reads its taps from the stage
// as 4-byte words (a warp covers one 128-byte row segment per tap, conflict-free).
// Algorithmic traffic: H read once (fwd); H read + dH written once (bwd).
#include <type_traits>

template <int D, int W, int TAPS_SMEM, int GROUPS>
struct DwRing {
  // grid rows per item: a whole band (4 rows) or, when the taps also live in shared memory and W = 16, half a band
  static constexpr int kRows = (TAPS_SMEM && W == 16) ? 2 : 4;
  static constexpr int kItemsPerBand = 4 / kRows;
  static constexpr int kTok = kRows * W;                   // tokens per item
  static constexpr int kWk = W / 4;                        // windows (sample points) per band
  static constexpr int kStageBytes = kTok * D * 2;
  static constexpr int kTapBytes = TAPS_SMEM ? 16 * D * 4 : 0;
  static constexpr int kStages = (192 * 1024 - kTapBytes) / kStageBytes;
  static constexpr int kGroups = GROUPS;                   // consumer groups: group g takes the windows px = g, g + GROUPS, ...
  static constexpr int kGroupThreads = D / 2;              // a thread owns channels 2t, 2t + 1 of its group's windows
  static constexpr int kConsumers = kGroups * kGroupThreads;
  static constexpr int kThreads = kConsumers + 32;         // + the producer warp
  static constexpr int kGroupWarps = kGroupThreads / 32;
  static constexpr int kRedFloats = 2 * kWk * kGroupWarps;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + kTapBytes + sizeof(ring::Bars<kStages>) + kRedFloats * sizeof(float) + 128;
  static_assert(kStages >= 2, "ring needs two stages");
  static_assert(kStages * kStageBytes >= 19 * D * 4, "the partial sums of group 1 are combined through the ring memory");
  static_assert(GROUPS == 1 || GROUPS == 2, "one or two consumer groups");
};
constexpr int kDwRingCtasPerMod = 49;   // 3 * 49 = 147 of 148 SMs
// Measured (B = 128, d = 768, 16 x 8): two consumer groups (24 warps, 72 registers) were SLOWER than one group with
// the taps in registers (fwd 33.9 vs 27.0 us, bwd 46.8 vs 46.6 us): the kernels are bound by their ~12 (fwd) / ~20 (bwd)
// instructions per element (one MUFU.TANH each), not by latency or by the ring.
constexpr int kDwFwdGroups = 1, kDwBwdGroups = 1, kDwBwdTapsSmem = 0;

template <int D, int W>
static __global__ void __launch_bounds__(DwRing<D, W, 0, kDwFwdGroups>::kThreads, 1)
lam_dw_fwd_ring_kernel(const __nv_bfloat16* __restrict__ H, sig_align_params prm, int B, int bands_per_mod, int ctas_per_mod,
                       float* __restrict__ U, float* __restrict__ o) {
  using R = DwRing<D, W, 0, kDwFwdGroups>;
  pdl_launch_dependents();
  extern __shared__ __align__(128) unsigned char dwr_smem[];
  unsigned char* stages = dwr_smem;
  auto* bars = reinterpret_cast<ring::Bars<R::kStages>*>(dwr_smem + (size_t)R::kStages * R::kStageBytes);
  float* red = reinterpret_cast<float*>(bars + 1);         // [2][kWk][kGroupWarps]
  const int m = blockIdx.x / ctas_per_mod, j = blockIdx.x % ctas_per_mod;
  const int i0 = (int)((int64_t)j * bands_per_mod / ctas_per_mod), i1 = (int)((int64_t)(j + 1) * bands_per_mod / ctas_per_mod);
  const int64_t pts_per_mod = (int64_t)bands_per_mod * R::kWk;   // = B * P
  const __nv_bfloat16* Hm = H + (int64_t)m * bands_per_mod * R::kTok * D;
  ring::init(bars, R::kConsumers / 32);
  if ((int)threadIdx.x >= R::kConsumers) {
    if ((int)threadIdx.x == R::kConsumers) {
      pdl_wait();   // H is the output of the GEMM launched just before
      for (int it = i0, k = 0; it < i1; ++it, ++k)
        ring::produce(bars, k, stages + (size_t)(k % R::kStages) * R::kStageBytes, Hm + (int64_t)it * R::kTok * D, R::kStageBytes, ptx::kPolStream);
    }
    return;
  }
  const int grp = threadIdx.x / R::kGroupThreads, gt = threadIdx.x % R::kGroupThreads;
  const int c = gt * 2, warp = gt >> 5, lane = gt & 31;
  constexpr int kWarps = R::kGroupWarps;
  float wk0[16], wk1[16];
  {
    const float4* wp = reinterpret_cast<const float4*>(prm.off2_w[m] + (int64_t)c * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = wp[q], b4 = wp[4 + q];
      wk0[4 * q] = a.x; wk0[4 * q + 1] = a.y; wk0[4 * q + 2] = a.z; wk0[4 * q + 3] = a.w;
      wk1[4 * q] = b4.x; wk1[4 * q + 1] = b4.y; wk1[4 * q + 2] = b4.z; wk1[4 * q + 3] = b4.w;
    }
  }
  const float2 w4 = *reinterpret_cast<const float2*>(prm.off4_w[m] + c);
  const float2 bd = *reinterpret_cast<const float2*>(prm.off2_b[m] + c);
  pdl_wait();
  for (int it = i0, k = 0; it < i1; ++it, ++k) {
    ring::consumer_wait(bars, k);
    const unsigned char* st = stages + (size_t)(k % R::kStages) * R::kStageBytes;
    float* rd = red + (k & 1) * R::kWk * kWarps;
#pragma unroll
    for (int pi = 0; pi < R::kWk / R::kGroups; ++pi) {
      const int px = grp + pi * R::kGroups;
      float u0 = bd.x, u1 = bd.y;
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const int t = (kk >> 2) * W + 4 * px + (kk & 3);
        const uint32_t hv = *reinterpret_cast<const uint32_t*>(st + ((size_t)t * D + c) * 2);
        u0 = fmaf(gelu_tanh_f(bf16lo_to_f32(hv)), wk0[kk], u0);
        u1 = fmaf(gelu_tanh_f(bf16hi_to_f32(hv)), wk1[kk], u1);
      }
      const int64_t gp = (int64_t)it * R::kWk + px;       // = b * P + p
      *reinterpret_cast<float2*>(U + ((int64_t)m * pts_per_mod + gp) * D + c) = make_float2(u0, u1);
      float part = gelu_fast_f(u0) * w4.x + gelu_fast_f(u1) * w4.y;
      part = warp_sum(part);
      if (lane == 0) rd[px * kWarps + warp] = part;
    }
    ring::consumer_release(bars, k);
    ring::consumer_sync(R::kConsumers);
    if ((int)threadIdx.x < R::kWk) {   // fixed summation order: deterministic
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < kWarps; ++q) t += rd[threadIdx.x * kWarps + q];
      o[(int64_t)m * pts_per_mod + (int64_t)it * R::kWk + threadIdx.x] = t;
    }
  }
}

// Backward, one pass over H:  dU = dO * w4 * gelu'(U);  dH = dU * wdw[c,k] * gelu'(H)  (bf16), and the CTA's partial
// parameter sums  dwdw[c,k] += dU * gelu(H),  dbdw += dU,  dw4 += dO * gelu(U),  dbf += dH   -> part [3][ctas_per_mod][19][D]
template <int D, int W>
static __global__ void __launch_bounds__(DwRing<D, W, kDwBwdTapsSmem, kDwBwdGroups>::kThreads, 1)
lam_dw_bwd_ring_kernel(const __nv_bfloat16* __restrict__ H, const float* __restrict__ U, const float* __restrict__ dO,
                       sig_align_params prm, int B, int bands_per_mod, int ctas_per_mod, __nv_bfloat16* __restrict__ dH,
                       float* __restrict__ part) {
  using R = DwRing<D, W, kDwBwdTapsSmem, kDwBwdGroups>;
  pdl_launch_dependents();
  extern __shared__ __align__(128) unsigned char dwr_smem[];
  unsigned char* stages = dwr_smem;
  float* wT = reinterpret_cast<float*>(dwr_smem + (size_t)R::kStages * R::kStageBytes);   // [16][D] depthwise taps
  auto* bars = reinterpret_cast<ring::Bars<R::kStages>*>(dwr_smem + (size_t)R::kStages * R::kStageBytes + R::kTapBytes);
  const int m = blockIdx.x / ctas_per_mod, j = blockIdx.x % ctas_per_mod;
  const int i0 = (int)((int64_t)j * bands_per_mod / ctas_per_mod), i1 = (int)((int64_t)(j + 1) * bands_per_mod / ctas_per_mod);
  // (here bands_per_mod counts ITEMS: bands, or half bands when W = 16)
  const int64_t pts_per_mod = (int64_t)(bands_per_mod / R::kItemsPerBand) * R::kWk;
  const __nv_bfloat16* Hm = H + (int64_t)m * bands_per_mod * R::kTok * D;
  __nv_bfloat16* dHm = dH + (int64_t)m * bands_per_mod * R::kTok * D;
  ring::init(bars, R::kConsumers / 32);
  if ((int)threadIdx.x >= R::kConsumers) {
    if ((int)threadIdx.x == R::kConsumers) {
      // (H was written by the forward call: the ring starts filling under the previous kernel's tail)
      for (int it = i0, k = 0; it < i1; ++it, ++k)
        ring::produce(bars, k, stages + (size_t)(k % R::kStages) * R::kStageBytes, Hm + (int64_t)it * R::kTok * D, R::kStageBytes, ptx::kPolStream);
    }
    return;
  }
  const int grp = threadIdx.x / R::kGroupThreads, gt = threadIdx.x % R::kGroupThreads;
  const int c = gt * 2;
  float wk0[16], wk1[16];
  {
    const float4* wp = reinterpret_cast<const float4*>(prm.off2_w[m] + (int64_t)c * 16);
    if (R::kTapBytes == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 a = wp[q], b4 = wp[4 + q];
        wk0[4 * q] = a.x; wk0[4 * q + 1] = a.y; wk0[4 * q + 2] = a.z; wk0[4 * q + 3] = a.w;
        wk1[4 * q] = b4.x; wk1[4 * q + 1] = b4.y; wk1[4 * q + 2] = b4.z; wk1[4 * q + 3] = b4.w;
      }
    } else if (grp == 0) {   // taps as [k][c] in shared memory (conflict-free 8-byte reads)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 a = wp[q], b4 = wp[4 + q];
        *reinterpret_cast<float2*>(wT + (4 * q) * D + c) = make_float2(a.x, b4.x);
        *reinterpret_cast<float2*>(wT + (4 * q + 1) * D + c) = make_float2(a.y, b4.y);
        *reinterpret_cast<float2*>(wT + (4 * q + 2) * D + c) = make_float2(a.z, b4.z);
        *reinterpret_cast<float2*>(wT + (4 * q + 3) * D + c) = make_float2(a.w, b4.w);
      }
    }
  }
  const float2 w4 = *reinterpret_cast<const float2*>(prm.off4_w[m] + c);
  float a0[19], a1[19];
#pragma unroll
  for (int i = 0; i < 19; ++i) a0[i] = a1[i] = 0.f;
  if (R::kTapBytes) ring::consumer_sync(R::kConsumers);   // group 1 reads the taps group 0 staged
  pdl_wait();   // U is older, dO comes from the kernel launched just before
  for (int it = i0, k = 0; it < i1; ++it, ++k) {
    const int band = it / R::kItemsPerBand, part_of_band = it % R::kItemsPerBand;
    float go[R::kWk / R::kGroups];
    float2 uu[R::kWk / R::kGroups];
#pragma unroll
    for (int pi = 0; pi < R::kWk / R::kGroups; ++pi) {   // requested before the wait for the stage
      const int64_t gp = (int64_t)m * pts_per_mod + (int64_t)band * R::kWk + grp + pi * R::kGroups;
      go[pi] = dO[gp];
      uu[pi] = *reinterpret_cast<const float2*>(U + gp * D + c);
    }
    ring::consumer_wait(bars, k);
    const unsigned char* st = stages + (size_t)(k % R::kStages) * R::kStageBytes;
    __nv_bfloat16* dst = dHm + (int64_t)it * R::kTok * D + c;
    // `first` item of a band also takes the window-level sums; ky0 = first kernel row this item holds
    auto body = [&](auto ky0_tag) {
      constexpr int ky0 = decltype(ky0_tag)::value;
#pragma unroll
      for (int pi = 0; pi < R::kWk / R::kGroups; ++pi) {
        const int px = grp + pi * R::kGroups;
        float gu0, dgu0, gu1, dgu1;
        gelu_fast(uu[pi].x, gu0, dgu0);
        gelu_fast(uu[pi].y, gu1, dgu1);
        const float du0 = go[pi] * w4.x * dgu0, du1 = go[pi] * w4.y * dgu1;
        if (ky0 == 0) {
          a0[16] += du0; a1[16] += du1;
          a0[17] = fmaf(go[pi], gu0, a0[17]); a1[17] = fmaf(go[pi], gu1, a1[17]);
        }
#pragma unroll
        for (int r = 0; r < R::kRows; ++r)
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) {
            const int kk = (ky0 + r) * 4 + kx;
            const int t = r * W + 4 * px + kx;
            const uint32_t hv = *reinterpret_cast<const uint32_t*>(st + ((size_t)t * D + c) * 2);
            const float2 wk = R::kTapBytes ? *reinterpret_cast<const float2*>(wT + kk * D + c) : make_float2(wk0[kk], wk1[kk]);
            float gh0, dgh0, gh1, dgh1;   // gelu and gelu' share one tanh per element
            gelu_tanh(bf16lo_to_f32(hv), gh0, dgh0);
            gelu_tanh(bf16hi_to_f32(hv), gh1, dgh1);
            const float dh0 = du0 * wk.x * dgh0, dh1 = du1 * wk.y * dgh1;
            *reinterpret_cast<__nv_bfloat162*>(dst + (size_t)t * D) = __floats2bfloat162_rn(dh0, dh1);
            a0[kk] = fmaf(du0, gh0, a0[kk]);
            a1[kk] = fmaf(du1, gh1, a1[kk]);
            a0[18] += dh0; a1[18] += dh1;
          }
      }
    };
    if (R::kItemsPerBand == 1 || part_of_band == 0) body(std::integral_constant<int, 0>{});
    else body(std::integral_constant<int, 2>{});
    ring::consumer_release(bars, k);
  }
  float* pd = part + (((int64_t)m * ctas_per_mod + j) * 19) * D + c;
  if (R::kGroups == 1) {
#pragma unroll
    for (int i = 0; i < 19; ++i) *reinterpret_cast<float2*>(pd + (int64_t)i * D) = make_float2(a0[i], a1[i]);
    return;
  }
  // the two groups' partial sums meet in the (now idle) ring memory: group 1 writes, group 0 adds and stores
  ring::consumer_sync(R::kConsumers);
  float* xch = reinterpret_cast<float*>(stages);   // [19][D]
  if (grp == 1) {
#pragma unroll
    for (int i = 0; i < 19; ++i) *reinterpret_cast<float2*>(xch + i * D + c) = make_float2(a0[i], a1[i]);
  }
  ring::consumer_sync(R::kConsumers);
  if (grp == 0) {
#pragma unroll
    for (int i = 0; i < 19; ++i) {
      const float2 o2 = *reinterpret_cast<const float2*>(xch + i * D + c);
      *reinterpret_cast<float2*>(pd + (int64_t)i * D) = make_float2(a0[i] + o2.x, a1[i] + o2.y);
    }
  }
}

static bool dw_ring_enabled() {
  static const bool on = [] {
    const char* e = getenv("SIG_DW_RING");
    return !(e && e[0] == '0');
  }();
  return on;
}
// the ring kernels cover the shipped shapes: L = 128 as h x w = 16 x 8 or 8 x 16, d = 512 or 768
static bool dw_ring_ok(int L, int h, int w, int d) {
  return dw_ring_enabled() && L == 128 && h * w == 128 && (w == 8 || w == 16) && (d == 512 || d == 768);
}
static int dw_ring_ctas_per_mod(int B, int h) {
  const int bands = B * (h / 4);
  // (never more CTAs than kDwRingCtasPerMod: the per-CTA partial-sum buffers of the backward are sized for it)
  const int cap = sm_budget() / 3 < kDwRingCtasPerMod ? sm_budget() / 3 : kDwRingCtasPerMod;
  return bands < cap ? bands : cap;
}

template <int D, int W>
static int launch_dw_fwd_ring(const __nv_bfloat16* H, const sig_align_params& p, int B, int h, float* U, float* o, cudaStream_t s) {
  using R = DwRing<D, W, 0, kDwFwdGroups>;
  ensure_dyn_smem(lam_dw_fwd_ring_kernel<D, W>, (int)R::kSmemBytes);
  const int bands = B * (h / 4), cpm = dw_ring_ctas_per_mod(B, h);
  SIG_LAUNCH((lam_dw_fwd_ring_kernel<D, W>), 3 * cpm, R::kThreads, R::kSmemBytes, s, H, p, B, bands, cpm, U, o);
  SIG_CHECK_LAUNCH();
  return 0;
}

template <int D, int W>
static int launch_dw_bwd_ring(const __nv_bfloat16* H, const float* U, const float* dO, const sig_align_params& p, int B, int h,
                              __nv_bfloat16* dH, float* part, cudaStream_t s) {
  using R = DwRing<D, W, kDwBwdTapsSmem, kDwBwdGroups>;
  ensure_dyn_smem(lam_dw_bwd_ring_kernel<D, W>, (int)R::kSmemBytes);
  const int items = B * (h / 4) * R::kItemsPerBand, cpm = dw_ring_ctas_per_mod(B, h);
  SIG_LAUNCH((lam_dw_bwd_ring_kernel<D, W>), 3 * cpm, R::kThreads, R::kSmemBytes, s, H, U, dO, p, B, items, cpm, dH, part);
  SIG_CHECK_LAUNCH();
  return 0;
}

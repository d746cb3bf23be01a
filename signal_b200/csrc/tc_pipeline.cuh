// Warp-specialised tcgen05 pipeline skeleton shared by the tensor-core kernels.
//
//   warp 0, 3 : TMA producers (A / B operand; cp.async.bulk.tensor -> 128B-swizzled shared-memory ring, mbarrier complete_tx)
//   warp 1 : MMA issuer    (one thread, tcgen05.mma kind::f16, fp32 accumulators in TMEM, tcgen05.commit)
//   warp 2 : TMEM allocator
//   warps 4-11 : epilogue  (tcgen05.ld of the accumulator, problem-specific math, global stores); warp w reads TMEM lane
//                quadrant w % 4, warps 4-7 the first half of the accumulator columns, warps 8-11 the second half (the
//                epilogue is instruction-latency bound: twice the warps, twice the rate)
//
// Persistent: CTA i processes work units i, i+grid, ...; the smem ring and the two TMEM accumulator
// buffers run continuously across units, so the epilogue of unit n overlaps the MMAs of unit n+1.
//
// A Problem supplies: Params (kernel argument, holds the CUtensorMaps), kAMn/kBMn (operand
// major-ness), num_units(), Unit + unit_info() (a work unit decoded once: its k-block range kb0..kb1 and
// whatever the rest needs), load_a()/load_b() (issue the TMA copies of one operand's k-block) and epilogue().
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace sig {
namespace tc {

constexpr int BM = 128, BK = 64;
constexpr int kEpiWarps = 8;                 // two warps per TMEM lane quadrant: each takes half of the accumulator columns
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kATileBytes = BM * BK * 2;
// Epilogue transpose scratch: one [32 rows][kEpiLd floats] tile per epilogue warp.  tcgen05.ld hands a
// thread one accumulator ROW (32 consecutive columns); stored from there a warp store touches 32
// different lines.  Staged through this tile a warp instead writes 4 rows x 128 contiguous bytes per
// instruction (kEpiLd = 36 keeps 16-byte alignment and is bank-conflict free both ways).
constexpr int kEpiLd = 36;
constexpr int kEpiBytes = kEpiWarps * 32 * kEpiLd * 4;

// MT = number of 128-row M tiles a CTA computes against ONE B tile (MT = 2: a 256 x BN output per
// unit; the B operand is fetched once for both, which raises the FLOPs per byte pulled from L2 --
// these GEMMs are bound by the ~8.7 TB/s L2 -> SM operand stream, not by the tensor pipe).
template <int BN, int MT = 1>
struct Cfg {
  static constexpr int kABytes = MT * kATileBytes;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (220 * 1024 - kEpiBytes) / kStageBytes > 8 ? 8 : (220 * 1024 - kEpiBytes) / kStageBytes;
  static constexpr int kAccCols = MT * BN;                              // fp32 accumulator columns of one unit
  static constexpr int kAccBufs = 2 * kAccCols <= 512 ? 2 : 1;          // double buffered when TMEM allows
  static constexpr int kTmemCols = kAccBufs * kAccCols < 32 ? 32 : kAccBufs * kAccCols;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kEpiBytes;
  static constexpr int smem_bytes(int stages) { return stages * kStageBytes + 1024 + 256 + kEpiBytes; }
};

// Loads one operand tile of `extent` M/N-rows for k-block starting at element k0.
//   K-major  : smem rows = M/N index, 128 B (64 bf16 of K) per row
//   MN-major : 64-element M/N chunks of [64 K-rows x 128 B], chunk j at dst + j*8192
__device__ __forceinline__ void load_kmajor_2d(const CUtensorMap* tm, uint8_t* dst, uint64_t* bar, int k0, int row0) {
  ptx::tma_load_2d(dst, tm, bar, k0, row0);
}
__device__ __forceinline__ void load_kmajor_tok(const CUtensorMap* tm, uint8_t* dst, uint64_t* bar, int k0, int sample, int nsamples) {
  for (int j = 0; j < nsamples; ++j) ptx::tma_load_3d(dst + j * 128 * BK * 2, tm, bar, k0, 0, sample + j, ptx::kPolTokens);
}
__device__ __forceinline__ void load_mnmajor_2d(const CUtensorMap* tm, uint8_t* dst, uint64_t* bar, int mn0, int krow0, int extent) {
  for (int j = 0; j < extent / 64; ++j) ptx::tma_load_2d(dst + j * 64 * BK * 2, tm, bar, mn0 + 64 * j, krow0);
}
__device__ __forceinline__ void load_mnmajor_tok(const CUtensorMap* tm, uint8_t* dst, uint64_t* bar, int mn0, int l0, int sample, int extent) {
  for (int j = 0; j < extent / 64; ++j) ptx::tma_load_3d(dst + j * 64 * BK * 2, tm, bar, mn0 + 64 * j, l0, sample, ptx::kPolTokens);
}

// Row layout -> tile: lane (= accumulator row) writes its 32 fp32 values.  Call epi_sync() before reading.
__device__ __forceinline__ void epi_put_row(float* scratch, int lane, const float (&v)[32]) {
  float4* dst = reinterpret_cast<float4*>(scratch + lane * kEpiLd);
#pragma unroll
  for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
// Tile -> transposed layout: in step i (0..7) lane handles row 4*i + lane/8, columns 4*(lane%8) .. +3.
__device__ __forceinline__ float4 epi_get(const float* scratch, int lane, int i) {
  return *reinterpret_cast<const float4*>(scratch + (4 * i + (lane >> 3)) * kEpiLd + 4 * (lane & 7));
}

// Debug time stamps (clock64 of CTA 0): set SIG_TC_STAMPS=1, read with sig_debug_tc_stamps().
//  0 kernel entry  1 prologue done  2 dependency wait done  3 first TMA issued  4 first stage landed
//  5 accumulator committed  6 epilogue sees accumulator  7 epilogue done  8 CTA exit
long long* stamps_ptr();   // nullptr unless SIG_TC_STAMPS=1 (then a 16-slot device buffer, allocated once)
#define SIG_STAMP(i) do { if (stamps && blockIdx.x == 0) stamps[i] = clock64(); } while (0)

// Register budget: kThreads (384) x 160 registers = 61 440 of the SM's 65 536, which leaves exactly one 128-thread,
// 32-register CTA of the gradient-exchange kernel (xchg.cu) room on the same SM -- that kernel must run NEXT TO these
// persistent kernels, not between them.  (The compiler's free choice was 161-162: a 384-thread CTA then owns the whole
// register file and nothing can be co-resident.)
#ifndef SIG_TC_MAXNREG
#define SIG_TC_MAXNREG 160
#endif
template <int BN, int MT, class Problem>
__global__ void __maxnreg__(SIG_TC_MAXNREG) pipeline_kernel(const __grid_constant__ typename Problem::Params p, const int nstages,
                                                               long long* __restrict__ stamps) {
  if (threadIdx.x == 0) SIG_STAMP(0);
  using C = Cfg<BN, MT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);   // 1024 B aligned (128B-swizzle atoms)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nstages * C::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + nstages;
  uint64_t* tmem_full = bars + 2 * nstages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_scratch = reinterpret_cast<float*>(smem + nstages * C::kStageBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_units = Problem::num_units(p);

  pdl_launch_dependents();   // PDL: the next kernel may be scheduled; it waits for us in its own pdl_wait()
  if (warp == 0 && lane == 0) Problem::prefetch(p);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < nstages; ++i) {
      ptx::mbar_init(&full[i], 2);    // one arrive.expect_tx per producer (A, B)
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < C::kAccBufs; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 32 * kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // decoding the CTA's first unit needs only the parameters: do it before the dependency wait
  const typename Problem::Unit u_first = Problem::unit_info(p, blockIdx.x);
  if (threadIdx.x == 0) SIG_STAMP(1);
  pdl_wait();                // setup (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
  if (threadIdx.x == 0) SIG_STAMP(2);

  // The three roles below are single instruction streams (one thread each for TMA and MMA, one warp per
  // scheduler for the epilogue): their cost is instruction latency, so everything per-unit is decoded
  // ONCE (Problem::unit_info) and the ring position is carried as (stage, phase) counters, not it % n.
  if ((warp == 0 || warp == 3) && lane == 0) {
    // ================= TMA producers: warp 0 feeds the A operand, warp 3 the B operand =================
    // (two independent instruction streams: the per-k-block issue latency of one thread was the ring's limit)
    const bool is_a = warp == 0;
    const uint32_t bytes = is_a ? (uint32_t)C::kABytes : (uint32_t)C::kBBytes;
    uint32_t stage = 0, ph = 0;
    bool first = true;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      const typename Problem::Unit u = unit == (int)blockIdx.x ? u_first : Problem::unit_info(p, unit);
      for (int kb = u.kb0; kb < u.kb1; ++kb) {
        ptx::mbar_wait(&empty[stage], ph ^ 1);
        ptx::mbar_expect_tx(&full[stage], bytes);
        uint8_t* sa = smem + stage * C::kStageBytes;
        if (is_a) Problem::load_a(p, u, kb, sa, &full[stage]);
        else Problem::load_b(p, u, kb, sa + C::kABytes, &full[stage]);
        if (first && is_a) SIG_STAMP(3);
        first = false;
        if (++stage == (uint32_t)nstages) { stage = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ================= MMA issuer =================
    constexpr int a_mn = Problem::kAMn, b_mn = Problem::kBMn;
    const uint32_t idesc = ptx::make_idesc_bf16(BM, BN, a_mn, b_mn);
    // K-major: 8-row groups 1024 B apart (SBO); one UMMA_K (16 bf16) = 32 B inside the 128 B swizzle row.
    // MN-major: 64-element chunks 8192 B apart (LBO); 8 K-rows = 1024 B (SBO); one UMMA_K = 16 rows = 2048 B.
    constexpr uint32_t a_lbo = a_mn ? 64 * 128 : 16, b_lbo = b_mn ? 64 * 128 : 16;
    constexpr uint32_t a_kstep = a_mn ? 16 * 128 : 32, b_kstep = b_mn ? 16 * 128 : 32;
    // descriptors of stage 0; a later stage / k-step only moves the 14-bit start-address field (>> 4)
    const uint32_t smem_base = ptx::smem_u32(smem);
    const uint64_t da0 = ptx::make_smem_desc(smem_base, a_lbo, 1024);
    const uint64_t db0 = ptx::make_smem_desc(smem_base + C::kABytes, b_lbo, 1024);
    uint32_t stage = 0, ph = 0, acc = 0, aph = 0;
    bool first = true;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      const typename Problem::Unit u = unit == (int)blockIdx.x ? u_first : Problem::unit_info(p, unit);
      ptx::mbar_wait(&tmem_empty[acc], aph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * C::kAccCols;
      for (int kb = u.kb0; kb < u.kb1; ++kb) {
        ptx::mbar_wait(&full[stage], ph);
        ptx::tc_fence_after();
        if (first) SIG_STAMP(4);
        const uint64_t soff = (uint64_t)((stage * (uint32_t)C::kStageBytes) >> 4);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t db = db0 + soff + (uint64_t)((k * b_kstep) >> 4);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t da = da0 + soff + (uint64_t)((mt * kATileBytes + k * a_kstep) >> 4);
            ptx::umma_bf16(d_tmem + mt * BN, da, db, idesc, kb > u.kb0 || k > 0);
          }
        }
        ptx::umma_commit(&empty[stage]);   // frees the smem slot once these MMAs have read it
        if (++stage == (uint32_t)nstages) { stage = 0; ph ^= 1; }
      }
      ptx::umma_commit(&tmem_full[acc]);   // accumulator complete
      if (first) { SIG_STAMP(5); first = false; }
      if (++acc == (uint32_t)C::kAccBufs) { acc = 0; aph ^= 1; }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int q = warp & 3;
    uint32_t acc = 0, aph = 0;
    bool first = true;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      const typename Problem::Unit u = unit == (int)blockIdx.x ? u_first : Problem::unit_info(p, unit);
      if (first && threadIdx.x == 128) SIG_STAMP(6);
      // (the epilogue waits for the accumulator itself -- tmem_full[acc], phase aph -- after it has
      //  requested everything that does not depend on it)
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt)
        Problem::epilogue(p, u, mt, tmem_base + acc * C::kAccCols + mt * BN + ((uint32_t)(q * 32) << 16), q, (warp - 4) >> 2, lane,
                          epi_scratch + (warp - 4) * 32 * kEpiLd, &tmem_full[acc], aph,
                          (first && threadIdx.x == 128 && blockIdx.x == 0) ? stamps : nullptr);
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);
      if (first && threadIdx.x == 128) SIG_STAMP(7);
      first = false;
      if (++acc == (uint32_t)C::kAccBufs) { acc = 0; aph ^= 1; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
  if (threadIdx.x == 0) SIG_STAMP(8);
}

// ---- host helpers -------------------------------------------------------------------------------
// bf16 tensor maps with 128B swizzle.  2-D: row-major [rows, cols] with pitch ld (elements), box
// [box_rows x 64].  Token view: [B, 128, d] with element strides, box [1 x box_rows x 64].
int make_map_2d(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out);
int make_map_tok(const void* ptr, int64_t B, int64_t d, int64_t stride_b, int64_t stride_l, int box_rows, CUtensorMap* out);
int make_map_3d(const void* ptr, int64_t cols, int64_t rows, int64_t nb, int64_t stride_row, int64_t stride_b, int box_rows,
                CUtensorMap* out);   // [nb][rows][cols], box [1 x box_rows x 64]
int num_sms();

int sm_reserve();       // SIG_TC_RESERVE in the environment: SMs a persistent launch leaves free
int stage_override();   // SIG_TC_STAGES in the environment (tuning aid), 0 = automatic

// kblocks_per_unit: K-blocks one work unit runs through; short K loops get a shallower ring (less
// shared memory to carve out, fewer barriers to initialise) -- it only needs to cover the load latency.
template <int BN, class Problem, int MT = 1>
int launch(const typename Problem::Params& p, int units, cudaStream_t s, int kblocks_per_unit = 1 << 20) {
  ensure_dyn_smem(pipeline_kernel<BN, MT, Problem>, Cfg<BN, MT>::kSmemBytes);
  if (units <= 0) return 0;
  // Kernels with at least one unit per SM are persistent and would hold every SM for their whole run; leaving a
  // few SMs free (SIG_TC_RESERVE, default in tc_gemm.cu) lets the short kernels of a concurrent stream -- the other
  // module's dependency chain -- keep moving.  These kernels are bound by the L2 operand stream, not by SM count.
  const int cap = (num_sms() - sm_reserve()) * sm_waves();
  const int grid = units < cap ? units : cap;
  int stages = Cfg<BN, MT>::kStages;
  const int per_cta = kblocks_per_unit * (int)ceil_div(units, grid);
  if (per_cta < stages) stages = per_cta < 2 ? 2 : per_cta;
  if (stage_override() > 0 && stage_override() < stages) stages = stage_override();
  SIG_LAUNCH((pipeline_kernel<BN, MT, Problem>), grid, kThreads, (Cfg<BN, MT>::smem_bytes(stages)), s, p, stages, stamps_ptr());
  SIG_CHECK_LAUNCH();
  return 0;
}

}  // namespace tc
}  // namespace sig

// Generic bf16 GEMM on the 5th-gen tensor cores (pipeline skeleton: tc_pipeline.cuh).
//   C[z][M,N] = alpha * A[z] . B[z]^T (+ bias) (GELU) (+ per-sample row vector), fp32 accumulate;
// operands are described by TMA tensor maps so strided token views are consumed in place.
#include "tc_gemm.h"

#include <cstdlib>
#include <mutex>

#include "tc_pipeline.cuh"

namespace sig {

namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

static int encode(void* ptr, cuuint32_t rank, const cuuint64_t* gdim, const cuuint64_t* gstr, const cuuint32_t* box, CUtensorMap* out) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return SIG_ERR_ARCH;
  if (!ptr) return SIG_ERR_NULL;
  if (!aligned16(ptr)) return SIG_ERR_ALIGN;
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : SIG_ERR_SHAPE;
}

int make_map_2d(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out) {
  if ((ld * 2) % 16) return SIG_ERR_ALIGN;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  return encode(const_cast<void*>(ptr), 2, gdim, gstr, box, out);
}

int make_map_tok(const void* ptr, int64_t B, int64_t d, int64_t stride_b, int64_t stride_l, int box_rows, CUtensorMap* out) {
  if ((stride_l * 2) % 16 || (stride_b * 2) % 16) return SIG_ERR_ALIGN;
  cuuint64_t gdim[3] = {(cuuint64_t)d, 128, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)stride_l * 2, (cuuint64_t)stride_b * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  return encode(const_cast<void*>(ptr), 3, gdim, gstr, box, out);
}

long long* stamps_ptr() {
  static long long* ptr = [] {
    const char* e = getenv("SIG_TC_STAMPS");
    long long* q = nullptr;
    if (e && atoi(e) && cudaMalloc(&q, 16 * sizeof(long long)) != cudaSuccess) q = nullptr;
    return q;
  }();
  return ptr;
}
int read_stamps(long long* out16) {
  if (!stamps_ptr()) return 1;
  return cudaMemcpy(out16, stamps_ptr(), sizeof(long long) * 16, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 1;
}

int sm_reserve() {
  static const int v = [] {
    const char* e = getenv("SIG_TC_RESERVE");
    int r = e ? atoi(e) : 0;
    return r < 0 ? 0 : (r > 64 ? 64 : r);
  }();
  return v;
}

int make_map_3d(const void* ptr, int64_t cols, int64_t rows, int64_t nb, int64_t stride_row, int64_t stride_b, int box_rows,
                CUtensorMap* out) {
  if ((stride_row * 2) % 16 || (stride_b * 2) % 16) return SIG_ERR_ALIGN;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)nb};
  cuuint64_t gstr[2] = {(cuuint64_t)stride_row * 2, (cuuint64_t)stride_b * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  return encode(const_cast<void*>(ptr), 3, gdim, gstr, box, out);
}

int stage_override() {
  static const int v = [] {
    const char* e = getenv("SIG_TC_STAGES");
    return e ? atoi(e) : 0;
  }();
  return v;
}

int num_sms() { return sm_budget(); }

}  // namespace tc

namespace {

using tc::BK;
using tc::BM;

struct TcKernelParams {
  CUtensorMap ta[8], tb[8];
  int a_mode, b_mode;
  int M, N, K, batch;
  void* C[8];
  long long ldc;
  int out_bf16;
  const float* bias[8];
  float alpha;
  int act, ksplit;
  int tiles_m, tiles_n, kblocks;
  int c_tok;
  long long c_stride_b, c_stride_l;
  const float* rowvec[8];
  const float* rowvec_scale;
  int accumulate;
  void* C2[8];
  long long ldc2;
  float* pre[8];
  int extra;              // 1: one more k-block (index kblocks) from the per-sample operands tax / tbx
  CUtensorMap tax, tbx;
};

template <int BN, bool AMN, bool BMN, int MT = 1>
struct GemmProblem {
  using Params = TcKernelParams;
  static constexpr int kAMn = AMN, kBMn = BMN;

  __device__ static void prefetch(const Params& p) {
    for (int z = 0; z < p.batch; ++z) {
      ptx::prefetch_tmap(&p.ta[z]);
      ptx::prefetch_tmap(&p.tb[z]);
    }
    if (p.extra) {
      ptx::prefetch_tmap(&p.tax);
      ptx::prefetch_tmap(&p.tbx);
    }
  }
  __device__ static int num_units(const Params& p) { return p.tiles_m * p.tiles_n * p.ksplit * p.batch; }
  struct Unit {
    int z, m0, n0, kb0, kb1;
  };
  __device__ static Unit unit_info(const Params& p, int unit) {
    Unit u;
    const int upb = p.tiles_m * p.tiles_n * p.ksplit;
    const int per = (p.kblocks + p.ksplit - 1) / p.ksplit;
    u.z = unit / upb;
    int r = unit - u.z * upb;
    const int ks = r % p.ksplit;
    r /= p.ksplit;
    const int tm = r / p.tiles_n;
    u.n0 = (r - tm * p.tiles_n) * BN;
    u.m0 = tm * BM * MT;
    u.kb0 = ks * per;
    u.kb1 = min(p.kblocks, u.kb0 + per) + p.extra;   // (extra implies ksplit == 1)
    return u;
  }
  template <bool kMn>   // the operand's major-ness is a compile-time property: one runtime test (2-D vs token view) remains
  __device__ static void load_one(const CUtensorMap* tm, int mode, uint8_t* dst, uint64_t* bar, int mn0, int kb, int extent) {
    const int k0 = kb * BK;
    if constexpr (!kMn) {
      if (mode == TC_K2D) tc::load_kmajor_2d(tm, dst, bar, k0, mn0);
      else tc::load_kmajor_tok(tm, dst, bar, k0, mn0 >> 7, extent >> 7);
    } else {
      if (mode == TC_MN2D) tc::load_mnmajor_2d(tm, dst, bar, mn0, k0, extent);
      else tc::load_mnmajor_tok(tm, dst, bar, mn0, k0 & 127, k0 >> 7, extent);   // K index = sample * 128 + l
    }
  }
  __device__ static void load_a(const Params& p, const Unit& u, int kb, uint8_t* sa, uint64_t* bar) {
    if (kb == p.kblocks) {   // the per-sample extra block: rows z*128.. of sample m0/128
      ptx::tma_load_3d(sa, &p.tax, bar, 0, u.z * 128, u.m0 >> 7);
      return;
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) load_one<AMN>(&p.ta[u.z], p.a_mode, sa + mt * tc::kATileBytes, bar, u.m0 + mt * BM, kb, BM);
  }
  __device__ static void load_b(const Params& p, const Unit& u, int kb, uint8_t* sb, uint64_t* bar) {
    if (kb == p.kblocks) {
      tc::load_mnmajor_2d(&p.tbx, sb, bar, u.n0, (u.m0 >> 7) * 64, BN);
      return;
    }
    load_one<BMN>(&p.tb[u.z], p.b_mode, sb, bar, u.n0, kb, BN);
  }

  // Epilogue of one 128 x BN accumulator.  tcgen05.ld delivers a thread one accumulator ROW; each 32 x 32
  // block goes through the warp's transpose tile (tc_pipeline.cuh) so that every global access -- bias,
  // per-sample row vectors, old values for accumulate, all stores -- is 4 rows x 128 contiguous bytes
  // (fp32) or 4 x 64 B (bf16) per warp instruction.  The epilogue is ONE warp per scheduler: exposed
  // latency is its cost and a branch costs it ~25 cycles, so the common case (full 32 x 32 block, aligned
  // pitch, no activation) runs branch-free code selected once per chunk (store_fast<...>), column vectors
  // are requested before the accumulator is waited for, and the TMEM read of chunk c+1 is in flight
  // while chunk c is stored.
  template <bool kBf16, bool kAcc, bool kC2>
  __device__ __forceinline__ static void store_fast(const float* scratch, int lane, float4 cv, void* cptr, long long step,
                                                    __nv_bfloat16* c2ptr, long long c2step) {
    if constexpr (kBf16) {
      __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(cptr);
      uint2 old[8];
      if constexpr (kAcc) {
#pragma unroll
        for (int i = 0; i < 8; ++i) old[i] = *reinterpret_cast<const uint2*>(dst + i * step);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t4 = tc::epi_get(scratch, lane, i);
        float t[4] = {t4.x + cv.x, t4.y + cv.y, t4.z + cv.z, t4.w + cv.w};
        if constexpr (kC2) store4(c2ptr + i * c2step, t);
        if constexpr (kAcc) {
          t[0] += __uint_as_float(old[i].x << 16); t[1] += __uint_as_float(old[i].x & 0xffff0000u);
          t[2] += __uint_as_float(old[i].y << 16); t[3] += __uint_as_float(old[i].y & 0xffff0000u);
        }
        store4(dst + i * step, t);
      }
    } else {
      float* dst = static_cast<float*>(cptr);
      float4 old[8];
      if constexpr (kAcc) {
#pragma unroll
        for (int i = 0; i < 8; ++i) old[i] = *reinterpret_cast<const float4*>(dst + i * step);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t4 = tc::epi_get(scratch, lane, i);
        float t[4] = {t4.x + cv.x, t4.y + cv.y, t4.z + cv.z, t4.w + cv.w};
        if constexpr (kC2) store4(c2ptr + i * c2step, t);
        if constexpr (kAcc) { t[0] += old[i].x; t[1] += old[i].y; t[2] += old[i].z; t[3] += old[i].w; }
        *reinterpret_cast<float4*>(dst + i * step) = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
  }
  __device__ static void epilogue(const Params& p, const Unit& u, int mt, uint32_t tmem_acc, int q, int half, int lane, float* scratch,
                                  uint64_t* acc_bar, uint32_t acc_phase, long long* stamps) {
#define EPI_STAMP(i) do { if (stamps) stamps[i] = clock64(); } while (0)
    const int z = u.z, n0 = u.n0;
    const int m0 = u.m0 + mt * BM;
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    const int row0 = m0 + q * 32 + sub;              // this lane's row in step i is row0 + 4 i
    const bool red = p.ksplit > 1;
    // split-K: the partial tiles are added into a pre-zeroed C; the unit that owns the first k-blocks also adds the bias
    const float* bias = (red && u.kb0 != 0) ? nullptr : p.bias[z];
    const float* rvec = (p.rowvec[z] && !red) ? p.rowvec[z] + (long long)(m0 >> 7) * p.N : nullptr;   // one sample per 128-row tile
    const float rscale = rvec ? *p.rowvec_scale : 0.f;
    const bool vec_ok = p.c_tok ? ((p.c_stride_b % 8) == 0 && (p.c_stride_l % 8) == 0) : ((p.ldc % 8) == 0);
    const float alpha = p.alpha;
    const int act = p.act, out_bf16 = p.out_bf16, accumulate = p.accumulate;
    float* const pre = p.pre[z];
    __nv_bfloat16* const c2 = static_cast<__nv_bfloat16*>(p.C2[z]);
    // element offset of (row0, column 0) in C and the stride of 4 rows
    const long long off0 = p.c_tok ? (long long)(row0 >> 7) * p.c_stride_b + (long long)(row0 & 127) * p.c_stride_l
                                   : (long long)row0 * p.ldc;
    const long long step = 4 * (p.c_tok ? p.c_stride_l : p.ldc);
    const long long c2off0 = (long long)row0 * p.ldc2, c2step = 4 * p.ldc2;
    const int nrows = p.M - row0;                    // step i is in range iff 4 i < nrows
    // warp-uniform: all 32 rows of this warp exist, pitches allow 16-byte accesses, nothing element-wise to do
    const bool fast_unit = (m0 + q * 32 + 32 <= p.M) && vec_ok && !act && !pre && (!c2 || (p.ldc2 % 8) == 0);
    const int mode = (out_bf16 ? 1 : 0) | (accumulate ? 2 : 0) | (c2 ? 4 : 0);
    // column vector of a chunk for this lane's 4 columns: bias + rscale * rowvec (bias is applied before
    // the activation and the row vector after it; they only fold into one add on the fast path, which has
    // no activation)
    auto colv_at = [&](const int c) -> float4 {
      const int col = n0 + c * 32 + c4;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col + 3 < p.N) {
        if (bias) t = *reinterpret_cast<const float4*>(bias + col);
        if (rvec) {
          const float4 r4 = *reinterpret_cast<const float4*>(rvec + col);
          t.x = fmaf(rscale, r4.x, t.x); t.y = fmaf(rscale, r4.y, t.y); t.z = fmaf(rscale, r4.z, t.z); t.w = fmaf(rscale, r4.w, t.w);
        }
      }
      return t;
    };
    constexpr int kChunks = BN / 32 / 2;             // this warp's share: chunks [c_lo, c_lo + kChunks)
    const int c_lo = half * kChunks;
    float4 cv = colv_at(c_lo);                       // requested before the accumulator is waited for
    EPI_STAMP(9);
    ptx::mbar_wait(acc_bar, acc_phase);              // accumulator complete
    ptx::tc_fence_after();
    EPI_STAMP(10);
    if (u.kb1 <= u.kb0) return;
    uint32_t r[32];
    ptx::tmem_ld32(tmem_acc + c_lo * 32, r);         // the read of chunk c+1 is issued as soon as chunk c is in `v`
#pragma unroll 1
    for (int c = c_lo; c < c_lo + kChunks; ++c) {
      const int col0 = n0 + c * 32;
      if (col0 >= p.N) break;                        // warp-uniform
      const int col = col0 + c4;
      ptx::tmem_ld_wait();
      if (c == c_lo) EPI_STAMP(11);
      {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * alpha;
        if (c + 1 < c_lo + kChunks) ptx::tmem_ld32(tmem_acc + (c + 1) * 32, r);   // in flight during this chunk's stores
        __syncwarp();                                // the previous chunk's tile reads are done
        tc::epi_put_row(scratch, lane, v);
        __syncwarp();
      }
      const float4 cvn = (c + 1 < c_lo + kChunks) ? colv_at(c + 1) : cv;   // next chunk's column vector: a chunk ahead
      if (c == c_lo) EPI_STAMP(12);
      if (fast_unit && col0 + 32 <= p.N) {
        if (red) {
          float* dst = static_cast<float*>(p.C[z]) + off0 + col;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t = tc::epi_get(scratch, lane, i);
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i * step), "f"(t.x + cv.x), "f"(t.y + cv.y),
                         "f"(t.z + cv.z), "f"(t.w + cv.w)
                         : "memory");
          }
        } else {
          void* cptr = out_bf16 ? static_cast<void*>(static_cast<__nv_bfloat16*>(p.C[z]) + off0 + col)
                                : static_cast<void*>(static_cast<float*>(p.C[z]) + off0 + col);
          __nv_bfloat16* c2p = c2 + c2off0 + col;
          switch (mode) {
            case 0: store_fast<false, false, false>(scratch, lane, cv, cptr, step, c2p, c2step); break;
            case 1: store_fast<true, false, false>(scratch, lane, cv, cptr, step, c2p, c2step); break;
            case 2: store_fast<false, true, false>(scratch, lane, cv, cptr, step, c2p, c2step); break;
            case 3: store_fast<true, true, false>(scratch, lane, cv, cptr, step, c2p, c2step); break;
            case 4: store_fast<false, false, true>(scratch, lane, cv, cptr, step, c2p, c2step); break;
            case 5: store_fast<true, false, true>(scratch, lane, cv, cptr, step, c2p, c2step); break;
            case 6: store_fast<false, true, true>(scratch, lane, cv, cptr, step, c2p, c2step); break;
            default: store_fast<true, true, true>(scratch, lane, cv, cptr, step, c2p, c2step); break;
          }
        }
      } else if (col < p.N) {
        // ---- general path (activation / pre-activation copy / row or column tail / unaligned pitch)
        float bb[4] = {0.f, 0.f, 0.f, 0.f}, rr[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < 4; ++j)
          if (col + j < p.N) {
            bb[j] = bias ? bias[col + j] : 0.f;
            rr[j] = rvec ? rscale * rvec[col + j] : 0.f;
          }
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          const float4 t4 = tc::epi_get(scratch, lane, i);
          const float t[4] = {t4.x, t4.y, t4.z, t4.w};
          if (4 * i >= nrows) continue;
          const long long o = off0 + i * step + col;
          const long long orow = (long long)(row0 + 4 * i);
          for (int j = 0; j < 4; ++j) {
            if (col + j >= p.N) break;
            if (red) { atomicAdd(static_cast<float*>(p.C[z]) + o + j, t[j] + bb[j]); continue; }
            float x = t[j] + bb[j];
            if (pre) pre[orow * p.ldc + col + j] = x;
            if (act == 1) x = gelu_f(x);
            x += rr[j];
            if (c2) c2[orow * p.ldc2 + col + j] = __float2bfloat16_rn(x);
            if (out_bf16) {
              __nv_bfloat16* d = static_cast<__nv_bfloat16*>(p.C[z]) + o + j;
              *d = __float2bfloat16_rn(x + (accumulate ? __bfloat162float(*d) : 0.f));
            } else {
              float* d = static_cast<float*>(p.C[z]) + o + j;
              *d = x + (accumulate ? *d : 0.f);
            }
          }
        }
      }
      cv = cvn;
      if (c == c_lo) EPI_STAMP(13);
      if (c == c_lo + 1) EPI_STAMP(14);
    }
    EPI_STAMP(15);
#undef EPI_STAMP
    ptx::tmem_ld_wait();   // (nothing is in flight when the loop ran to the end; a column-tail break leaves one read)
  }
};

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for the large GEMMs: a cluster of two CTAs computes a 256 x BN unit.
// Each CTA stages ITS 128 rows of A and HALF of the B tile (BN/2 rows); the leader's MMA thread issues
// M = 256 instructions that read both halves, so the bytes a SM pulls per k-block drop from 48 KB to 32 KB
// at BN = 256 -- the single-CTA kernel is bound by exactly that L2 -> SM stream.  Same ring / TMEM double
// buffering / epilogue as tc::pipeline_kernel; differences: both CTAs' TMA transactions land on the LEADER's
// full barrier, MMA completion is multicast to both CTAs' barriers, and the non-leader's epilogue threads
// release the accumulator with a remote arrive on the leader's tmem_empty barrier.
// ------------------------------------------------------------------------------------------------
template <int BN, bool AMN, bool BMN>
__global__ void __maxnreg__(SIG_TC_MAXNREG) pair_gemm_kernel(const __grid_constant__ TcKernelParams p, const int nstages) {
  using P1 = GemmProblem<BN, AMN, BMN, 1>;
  using P2 = GemmProblem<BN, AMN, BMN, 2>;   // unit decode with 256-row M tiles
  constexpr int kABytes = tc::kATileBytes, kBBytes = (BN / 2) * BK * 2, kStageBytes = kABytes + kBBytes;
  constexpr int kAccCols = BN, kAccBufs = 2 * BN <= 512 ? 2 : 1, kTmemCols = kAccBufs * BN < 32 ? 32 : kAccBufs * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nstages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + nstages;
  uint64_t* tmem_full = bars + 2 * nstages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_scratch = reinterpret_cast<float*>(smem + nstages * kStageBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total_units = P2::num_units(p);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) P1::prefetch(p);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < nstages; ++i) {
      ptx::mbar_init(&full[i], 2);      // the leader's two producers (A, B) arm it for BOTH CTAs' bytes
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kAccBufs; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 2 * 32 * tc::kEpiWarps);   // the epilogue threads of both CTAs (the leader's barrier is the one waited on)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc_2sm<kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync();                  // barriers of both CTAs initialised before any remote arrive / transaction
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if ((warp == 0 || warp == 3) && lane == 0) {
    // ================= TMA producers (both CTAs): warp 0 = this CTA's 128 rows of A, warp 3 = its half of B =================
    const bool is_a = warp == 0;
    uint32_t stage = 0, ph = 0;
    for (int unit = pair; unit < total_units; unit += npairs) {
      const typename P2::Unit u = P2::unit_info(p, unit);
      const int m0 = u.m0 + (int)rank * BM;
      const int n0 = u.n0 + (int)rank * (BN / 2);
      const CUtensorMap* tm = is_a ? &p.ta[u.z] : &p.tb[u.z];
      const int mode = is_a ? p.a_mode : p.b_mode;
      const int mn0 = is_a ? m0 : n0;
      constexpr int kExtent = 128;        // rows of A / rows of this CTA's half of B (BN = 256)
      for (int kb = u.kb0; kb < u.kb1; ++kb) {
        ptx::mbar_wait(&empty[stage], ph ^ 1);
        const uint32_t lbar = ptx::mapa_u32(&full[stage], 0);
        if (leader) ptx::mbar_expect_tx(&full[stage], 2u * (is_a ? (uint32_t)kABytes : (uint32_t)kBBytes));
        uint8_t* dst = smem + stage * kStageBytes + (is_a ? 0 : kABytes);
        const int k0 = kb * BK;
        if (kb >= p.kblocks) {
          // per-sample extra blocks (TcGemmDesc::xa/xb): the pair covers samples b0 = m0/128 and b0 + 1, whose
          // B operands differ, so block e = 0/1 contracts sample b0 + e only: the OTHER CTA's A tile is all zeros
          // (a box past the last sample: TMA zero-fills it)
          const int e = kb - p.kblocks, b = (u.m0 >> 7) + e;
          if (is_a) ptx::tma_load_3d_2sm(dst, &p.tax, lbar, 0, (int)rank == e ? u.z * 128 : 0, (int)rank == e ? b : p.M);
          else {
#pragma unroll
            for (int j = 0; j < kExtent / 64; ++j) ptx::tma_load_2d_2sm(dst + j * 64 * BK * 2, &p.tbx, lbar, n0 + 64 * j, b * 64);
          }
        } else
        // (L2 hints: token operands are re-read by other kernels of the step -> evict_last; the A operand of the dX / dW'
        //  GEMMs is dH, read once here and never again -> evict_first; weights: normal)
        if (mode == TC_K2D) ptx::tma_load_2d_2sm(dst, tm, lbar, k0, mn0, is_a ? ptx::kPolStream : ptx::kPolNormal);
        else if (mode == TC_KTOK) ptx::tma_load_3d_2sm(dst, tm, lbar, k0, 0, mn0 >> 7, ptx::kPolTokens);
        else if (mode == TC_MN2D) {
#pragma unroll
          for (int j = 0; j < kExtent / 64; ++j)
            ptx::tma_load_2d_2sm(dst + j * 64 * BK * 2, tm, lbar, mn0 + 64 * j, k0, is_a ? ptx::kPolStream : ptx::kPolNormal);
        } else {
#pragma unroll
          for (int j = 0; j < kExtent / 64; ++j)
            ptx::tma_load_3d_2sm(dst + j * 64 * BK * 2, tm, lbar, mn0 + 64 * j, k0 & 127, k0 >> 7, ptx::kPolTokens);
        }
        if (++stage == (uint32_t)nstages) { stage = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ================= MMA issuer (leader CTA only) =================
    const uint32_t idesc = ptx::make_idesc_bf16(2 * BM, BN, AMN, BMN);
    constexpr uint32_t a_lbo = AMN ? 64 * 128 : 16, b_lbo = BMN ? 64 * 128 : 16;
    constexpr uint32_t a_kstep = AMN ? 16 * 128 : 32, b_kstep = BMN ? 16 * 128 : 32;
    const uint32_t smem_base = ptx::smem_u32(smem);
    const uint64_t da0 = ptx::make_smem_desc(smem_base, a_lbo, 1024);
    const uint64_t db0 = ptx::make_smem_desc(smem_base + kABytes, b_lbo, 1024);
    uint32_t stage = 0, ph = 0, acc = 0, aph = 0;
    for (int unit = pair; unit < total_units; unit += npairs) {
      const typename P2::Unit u = P2::unit_info(p, unit);
      ptx::mbar_wait(&tmem_empty[acc], aph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccCols;
      for (int kb = u.kb0; kb < u.kb1; ++kb) {
        ptx::mbar_wait(&full[stage], ph);
        ptx::tc_fence_after();
        const uint64_t soff = (uint64_t)((stage * (uint32_t)kStageBytes) >> 4);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          ptx::umma_bf16_2sm(d_tmem, da0 + soff + (uint64_t)((k * a_kstep) >> 4), db0 + soff + (uint64_t)((k * b_kstep) >> 4), idesc,
                             kb > u.kb0 || k > 0);
        ptx::umma_commit_2sm(&empty[stage]);      // frees the stage in both CTAs
        if (++stage == (uint32_t)nstages) { stage = 0; ph ^= 1; }
      }
      ptx::umma_commit_2sm(&tmem_full[acc]);      // accumulator complete, both CTAs
      if (++acc == (uint32_t)kAccBufs) { acc = 0; aph ^= 1; }
    }
  } else if (warp >= 4) {
    // ================= epilogue (both CTAs, each its own 128 accumulator rows) =================
    const int q = warp & 3;
    uint32_t acc = 0, aph = 0;
    for (int unit = pair; unit < total_units; unit += npairs) {
      typename P2::Unit u2 = P2::unit_info(p, unit);
      typename P1::Unit u;
      u.z = u2.z; u.m0 = u2.m0 + (int)rank * BM; u.n0 = u2.n0; u.kb0 = u2.kb0; u.kb1 = u2.kb1;
      P1::epilogue(p, u, 0, tmem_base + acc * kAccCols + ((uint32_t)(q * 32) << 16), q, (warp - 4) >> 2, lane,
                   epi_scratch + (warp - 4) * 32 * tc::kEpiLd, &tmem_full[acc], aph, nullptr);
      ptx::tc_fence_before();
      if (leader) ptx::mbar_arrive(&tmem_empty[acc]);
      else ptx::mbar_arrive_cluster(ptx::mapa_u32(&tmem_empty[acc], 0));
      if (++acc == (uint32_t)kAccBufs) { acc = 0; aph ^= 1; }
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();                  // nobody tears down while the peer may still touch this CTA's smem / TMEM
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm<kTmemCols>(tmem_base);
  }
}

template <bool AMN, bool BMN>
static int launch_pair(const TcKernelParams& p, int units, int kpu, cudaStream_t s) {
  constexpr int BN = 256;
  constexpr int kStageBytes = tc::kATileBytes + (BN / 2) * BK * 2;
  auto kernel = pair_gemm_kernel<BN, AMN, BMN>;
  int stages = (220 * 1024 - tc::kEpiBytes) / kStageBytes;
  if (stages > 8) stages = 8;
  const int max_smem = stages * kStageBytes + 1024 + 256 + tc::kEpiBytes;
  ensure_dyn_smem(kernel, max_smem);
  if (units <= 0) return 0;
  const int npairs_max = (tc::num_sms() - tc::sm_reserve()) / 2 * sm_waves();
  const int npairs = units < npairs_max ? units : npairs_max;
  const int per_cta = kpu * (int)ceil_div(units, npairs);
  if (per_cta < stages) stages = per_cta < 2 ? 2 : per_cta;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * npairs);
  cfg.blockDim = dim3(tc::kThreads);
  cfg.dynamicSmemBytes = stages * kStageBytes + 1024 + 256 + tc::kEpiBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kernel, p, stages);
  SIG_CHECK_LAUNCH();
  return 0;
}

// g.pair: 256 x 256 units on CTA pairs.  A: K-major token view or MN-major 2-D; B: K-major 2-D or MN-major (2-D / token view).
static int launch_gemm_pair(const TcGemmDesc& g, cudaStream_t s) {
  if (g.bn != 256 || g.mt != 1) return SIG_ERR_SHAPE;
  TcKernelParams p{};
  auto mk = [&](const TcOperand& o, int z, CUtensorMap* out) -> int {   // every box is 128 M/N-rows (K-major) or 64 x 64 (MN-major)
    if (o.mode == TC_K2D) return tc::make_map_2d(o.ptr[z], o.rows, o.cols, o.ld, 128, out);
    if (o.mode == TC_MN2D) return tc::make_map_2d(o.ptr[z], o.rows, o.cols, o.ld, 64, out);
    return tc::make_map_tok(o.ptr[z], o.rows, o.cols, o.stride_b, o.stride_l, o.mode == TC_KTOK ? 128 : 64, out);
  };
  for (int z = 0; z < g.batch; ++z) {
    SIG_TRY(mk(g.A, z, &p.ta[z]));
    SIG_TRY(mk(g.B, z, &p.tb[z]));
    p.C[z] = g.C[z]; p.bias[z] = g.bias[z]; p.rowvec[z] = g.rowvec[z]; p.C2[z] = g.C2[z]; p.pre[z] = g.pre[z];
    if (!g.C[z]) return SIG_ERR_NULL;
  }
  p.a_mode = g.A.mode; p.b_mode = g.B.mode;
  p.M = g.M; p.N = g.N; p.K = g.K; p.batch = g.batch;
  p.ldc = g.ldc; p.out_bf16 = g.out_bf16; p.alpha = g.alpha; p.act = g.act; p.ksplit = g.ksplit < 1 ? 1 : g.ksplit;
  p.c_tok = g.c_tok; p.c_stride_b = g.c_stride_b; p.c_stride_l = g.c_stride_l;
  p.rowvec_scale = g.rowvec_scale; p.accumulate = g.accumulate; p.ldc2 = g.ldc2;
  p.tiles_m = (int)ceil_div(g.M, 2 * BM);
  p.tiles_n = (int)ceil_div(g.N, 256);
  p.kblocks = (int)ceil_div(g.K, BK);
  if (p.ksplit > p.kblocks) p.ksplit = p.kblocks;
  const int units = p.tiles_m * p.tiles_n * p.ksplit * p.batch;
  const bool amn = g.A.mode >= TC_MN2D, bmn = g.B.mode >= TC_MN2D;
  if (g.xa || g.xb) {   // two extra k-blocks per unit (one per sample of the pair)
    if (!g.xa || !g.xb || g.xB < 2 || (g.xB & 1) || amn || !bmn || p.ksplit != 1 || g.M != g.xB * 128 || g.batch > 3) return SIG_ERR_SHAPE;
    SIG_TRY(tc::make_map_3d(g.xa, 64, 384, g.xB, 64, 384 * 64, 128, &p.tax));
    SIG_TRY(tc::make_map_2d(g.xb, (int64_t)g.xB * 64, g.N, g.N, 64, &p.tbx));
    p.extra = 2;
  }
  const int kpu = (p.kblocks + p.ksplit - 1) / p.ksplit + p.extra;
  if (amn && bmn) return launch_pair<true, true>(p, units, kpu, s);
  if (amn) return launch_pair<true, false>(p, units, kpu, s);
  if (bmn) return launch_pair<false, true>(p, units, kpu, s);
  return launch_pair<false, false>(p, units, kpu, s);
}

template <int BN, int MT>
int launch_gemm_bn(const TcGemmDesc& g, cudaStream_t s) {
  TcKernelParams p{};
  auto mk = [&](const TcOperand& o, int z, int extent, CUtensorMap* out) -> int {
    if (o.mode == TC_K2D) return tc::make_map_2d(o.ptr[z], o.rows, o.cols, o.ld, extent, out);
    if (o.mode == TC_MN2D) return tc::make_map_2d(o.ptr[z], o.rows, o.cols, o.ld, 64, out);
    return tc::make_map_tok(o.ptr[z], o.rows, o.cols, o.stride_b, o.stride_l, o.mode == TC_KTOK ? 128 : 64, out);
  };
  for (int z = 0; z < g.batch; ++z) {
    SIG_TRY(mk(g.A, z, BM, &p.ta[z]));
    SIG_TRY(mk(g.B, z, BN, &p.tb[z]));
    p.C[z] = g.C[z];
    p.bias[z] = g.bias[z];
    p.rowvec[z] = g.rowvec[z];
    p.C2[z] = g.C2[z];
    p.pre[z] = g.pre[z];
    if (!g.C[z]) return SIG_ERR_NULL;
  }
  p.a_mode = g.A.mode; p.b_mode = g.B.mode;
  p.M = g.M; p.N = g.N; p.K = g.K; p.batch = g.batch;
  p.ldc = g.ldc; p.out_bf16 = g.out_bf16; p.alpha = g.alpha; p.act = g.act; p.ksplit = g.ksplit < 1 ? 1 : g.ksplit;
  p.c_tok = g.c_tok; p.c_stride_b = g.c_stride_b; p.c_stride_l = g.c_stride_l;
  p.rowvec_scale = g.rowvec_scale; p.accumulate = g.accumulate; p.ldc2 = g.ldc2;
  p.tiles_m = (int)ceil_div(g.M, BM * MT);
  p.tiles_n = (int)ceil_div(g.N, BN);
  p.kblocks = (int)ceil_div(g.K, BK);
  if (p.ksplit > p.kblocks) p.ksplit = p.kblocks;
  const int units = p.tiles_m * p.tiles_n * p.ksplit * p.batch;
  const bool amn = g.A.mode >= TC_MN2D, bmn = g.B.mode >= TC_MN2D;
  if (g.xa || g.xb) {
    if (!g.xa || !g.xb || g.xB < 1 || amn || !bmn || MT != 1 || p.ksplit != 1 || g.M != g.xB * 128 || g.batch > 3) return SIG_ERR_SHAPE;
    SIG_TRY(tc::make_map_3d(g.xa, 64, 384, g.xB, 64, 384 * 64, 128, &p.tax));
    SIG_TRY(tc::make_map_2d(g.xb, (int64_t)g.xB * 64, g.N, g.N, 64, &p.tbx));
    p.extra = 1;
  }
  const int kpu = (p.kblocks + p.ksplit - 1) / p.ksplit + p.extra;
  if (amn && bmn) return tc::launch<BN, GemmProblem<BN, true, true, MT>, MT>(p, units, s, kpu);
  if (amn) return tc::launch<BN, GemmProblem<BN, true, false, MT>, MT>(p, units, s, kpu);
  if (bmn) return tc::launch<BN, GemmProblem<BN, false, true, MT>, MT>(p, units, s, kpu);
  return tc::launch<BN, GemmProblem<BN, false, false, MT>, MT>(p, units, s, kpu);
}

}  // namespace

int tc_num_sms() { return tc::num_sms(); }
bool tc_pair_enabled() {
  static const bool on = [] {
    const char* e = getenv("SIG_TC_PAIR");
    return !(e && e[0] == '0');
  }();
  return on;
}
int tc_read_stamps(long long* out16) { return tc::read_stamps(out16); }

int tc_gemm(const TcGemmDesc& g, cudaStream_t s) {
  if (g.M < 1 || g.N < 1 || g.K < 1 || g.batch < 1 || g.batch > 8) return SIG_ERR_SHAPE;
  if (g.ksplit > 1 && (g.out_bf16 || g.act)) return SIG_ERR_SHAPE;
  if (g.pair) return launch_gemm_pair(g, s);
  if (g.bn == 256 && g.mt == 2) return launch_gemm_bn<256, 2>(g, s);
  if (g.bn == 256) return launch_gemm_bn<256, 1>(g, s);
  return launch_gemm_bn<128, 1>(g, s);
}

// ---- bf16 helpers ---------------------------------------------------------------------------------
static __global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  pdl_enter();
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + i) = o;
  } else {
    for (int64_t j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
  }
}

int cast_f32_to_bf16(const float* src, __nv_bfloat16* dst, int64_t n, cudaStream_t s) {
  if (n <= 0) return 0;
  SIG_LAUNCH((cast_bf16_kernel), (unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, s, src, dst, n);
  SIG_CHECK_LAUNCH();
  return 0;
}

struct CastJobs {
  CastJob j[8];
};
static __global__ void cast_bf16_multi_kernel(CastJobs jobs) {
  pdl_enter();
  const CastJob jb = jobs.j[blockIdx.y];
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4; i < jb.n; i += (int64_t)gridDim.x * blockDim.x * 4) {
    if (i + 3 < jb.n) {
      const float4 v = *reinterpret_cast<const float4*>(jb.src + i);
      __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(jb.dst + i) = o;
    } else {
      for (int64_t q = i; q < jb.n; ++q) jb.dst[q] = __float2bfloat16_rn(jb.src[q]);
    }
  }
}

int cast_f32_to_bf16_multi(const CastJob* jobs, int n, cudaStream_t s) {
  if (n <= 0) return 0;
  if (n > 8) return SIG_ERR_SHAPE;
  CastJobs js{};
  int64_t mx = 0;
  for (int i = 0; i < n; ++i) {
    js.j[i] = jobs[i];
    mx = jobs[i].n > mx ? jobs[i].n : mx;
  }
  int64_t blocks = ceil_div(ceil_div(mx, 4), 256);
  if (blocks > 1184) blocks = 1184;
  SIG_LAUNCH((cast_bf16_multi_kernel), dim3((unsigned)blocks, (unsigned)n), 256, 0, s, js);
  SIG_CHECK_LAUNCH();
  return 0;
}

static __global__ void transpose_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols) {
  pdl_enter();
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dst[(int64_t)c * rows + r] = __float2bfloat16_rn(t[threadIdx.x][i]);
  }
}

int transpose_f32_to_bf16(const float* src, __nv_bfloat16* dst, int rows, int cols, cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
  SIG_LAUNCH((transpose_bf16_kernel), grid, dim3(32, 8), 0, s, src, dst, rows, cols);
  SIG_CHECK_LAUNCH();
  return 0;
}

}  // namespace sig

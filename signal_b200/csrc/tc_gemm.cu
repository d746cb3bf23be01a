// Generic bf16 GEMM on the 5th-gen tensor cores: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared
// memory ring -> tcgen05.mma (one issuing thread, fp32 accumulators in TMEM, double buffered) ->
// tcgen05.ld epilogue.  Persistent: one CTA per SM loops over (batch, m-tile, n-tile, k-split) units.
//
// Warp roles (256 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue (warp w reads TMEM lanes 32*(w%4) .. +31).
#include "tc_gemm.h"

#include <mutex>

#include "ptx.cuh"

namespace sig {

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kThreads = 256;

struct TcKernelParams {
  CUtensorMap ta[3], tb[3];
  int a_mode, b_mode;
  int M, N, K, batch;
  void* C[3];
  long long ldc;
  int out_bf16;
  const float* bias[3];
  float alpha;
  int act, ksplit;
  int tiles_m, tiles_n, kblocks;
  int c_tok;
  long long c_stride_b, c_stride_l;
  const float* rowvec[3];
  const float* rowvec_scale;
  int accumulate;
};

template <int BN>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN == 128 ? 6 : 4;
  static constexpr int kTmemCols = 2 * BN;   // two accumulator buffers
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN>
__device__ __forceinline__ void load_operand(const CUtensorMap* tm, int mode, uint8_t* dst, uint64_t* bar, int mn0, int kb,
                                             int extent /*128 or BN*/) {
  const int k0 = kb * BK;
  if (mode == TC_K2D) {
    ptx::tma_load_2d(dst, tm, bar, k0, mn0);
  } else if (mode == TC_KTOK) {
    // rows = (sample, position): one 128-row tile = one sample (extent 128); a 256-row tile = two samples
    for (int j = 0; j < extent / 128; ++j) ptx::tma_load_3d(dst + j * 128 * BK * 2, tm, bar, k0, 0, mn0 / 128 + j);
  } else if (mode == TC_MN2D) {
    for (int j = 0; j < extent / 64; ++j) ptx::tma_load_2d(dst + j * 64 * BK * 2, tm, bar, mn0 + 64 * j, k0);
  } else {  // TC_MNTOK: K index = global position = sample * 128 + l
    for (int j = 0; j < extent / 64; ++j)
      ptx::tma_load_3d(dst + j * 64 * BK * 2, tm, bar, mn0 + 64 * j, k0 % 128, k0 / 128);
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) tc_gemm_kernel(const __grid_constant__ TcKernelParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);   // 1024 B aligned (128B-swizzle atoms)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units_per_batch = p.tiles_m * p.tiles_n * p.ksplit;
  const int total_units = units_per_batch * p.batch;
  const int kb_per_split = (p.kblocks + p.ksplit - 1) / p.ksplit;

  if (warp == 0 && lane == 0) {
    for (int z = 0; z < p.batch; ++z) {
      ptx::prefetch_tmap(&p.ta[z]);
      ptx::prefetch_tmap(&p.tb[z]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int unit, int& z, int& m0, int& n0, int& kb0, int& kb1) {
    z = unit / units_per_batch;
    int r = unit % units_per_batch;
    const int ks = r % p.ksplit;
    r /= p.ksplit;
    n0 = (r % p.tiles_n) * BN;
    m0 = (r / p.tiles_n) * BM;
    kb0 = ks * kb_per_split;
    kb1 = min(p.kblocks, kb0 + kb_per_split);
  };

  if (warp == 0 && lane == 0) {
    // ================= TMA producer =================
    uint32_t it = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      int z, m0, n0, kb0, kb1;
      decode(unit, z, m0, n0, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const uint32_t stage = it % C::kStages, ph = (it / C::kStages) & 1;
        ptx::mbar_wait(&empty[stage], ph ^ 1);
        ptx::mbar_expect_tx(&full[stage], C::kStageBytes);
        uint8_t* sa = smem + stage * C::kStageBytes;
        load_operand<BN>(&p.ta[z], p.a_mode, sa, &full[stage], m0, kb, BM);
        load_operand<BN>(&p.tb[z], p.b_mode, sa + C::kABytes, &full[stage], n0, kb, BN);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ================= MMA issuer =================
    const int a_mn = p.a_mode >= TC_MN2D, b_mn = p.b_mode >= TC_MN2D;
    const uint32_t idesc = ptx::make_idesc_bf16(BM, BN, a_mn, b_mn);
    // K-major: rows of 128 B, 8-row groups 1024 B apart; one UMMA_K (16 bf16) = 32 B inside the swizzle row.
    // MN-major: 64-element chunks 64 K-rows * 128 B = 8192 B apart (LBO); 8 K-rows = 1024 B (SBO); UMMA_K = 16 rows = 2048 B.
    const uint32_t a_lbo = a_mn ? 64 * 128 : 16, b_lbo = b_mn ? 64 * 128 : 16;
    const uint32_t a_kstep = a_mn ? 16 * 128 : 32, b_kstep = b_mn ? 16 * 128 : 32;
    uint32_t it = 0, tcount = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++tcount) {
      int z, m0, n0, kb0, kb1;
      decode(unit, z, m0, n0, kb0, kb1);
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      ptx::mbar_wait(&tmem_empty[acc], aph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const uint32_t stage = it % C::kStages, ph = (it / C::kStages) & 1;
        ptx::mbar_wait(&full[stage], ph);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem + stage * C::kStageBytes);
        const uint32_t b_addr = a_addr + C::kABytes;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t da = ptx::make_smem_desc(a_addr + k * a_kstep, a_lbo, 1024);
          const uint64_t db = ptx::make_smem_desc(b_addr + k * b_kstep, b_lbo, 1024);
          ptx::umma_bf16(d_tmem, da, db, idesc, kb > kb0 || k > 0);
        }
        ptx::umma_commit(&empty[stage]);   // frees the smem slot when these MMAs have read it
      }
      ptx::umma_commit(&tmem_full[acc]);   // accumulator complete
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int q = warp & 3;
    uint32_t tcount = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++tcount) {
      int z, m0, n0, kb0, kb1;
      decode(unit, z, m0, n0, kb0, kb1);
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      ptx::mbar_wait(&tmem_full[acc], aph);
      ptx::tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const float* bias = p.bias[z];
      const long long row_off = p.c_tok ? (long long)(row >> 7) * p.c_stride_b + (long long)(row & 127) * p.c_stride_l
                                        : (long long)row * p.ldc;
      const float* rvec = p.rowvec[z] ? p.rowvec[z] + (long long)(row >> 7) * p.N : nullptr;
      const float rscale = p.rowvec[z] ? *p.rowvec_scale : 0.f;
      const bool vec_ok = p.c_tok ? ((p.c_stride_b % 8) == 0 && (p.c_stride_l % 8) == 0) : ((p.ldc % 8) == 0);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem_base + acc * BN + c * 32 + ((uint32_t)(q * 32) << 16), r);
        ptx::tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row < p.M && col0 < p.N && kb1 > kb0) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
          const bool full_cols = col0 + 32 <= p.N;
          if (p.ksplit > 1) {
            float* dst = static_cast<float*>(p.C[z]) + row_off + col0;
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) atomicAdd(dst + j, v[j]);
          } else {
            if (bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) v[j] += bias[col0 + j];
            }
            if (p.act == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
            }
            if (rvec) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) v[j] = fmaf(rscale, rvec[col0 + j], v[j]);
            }
            if (p.out_bf16) {
              __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.C[z]) + row_off + col0;
              if (full_cols && vec_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  float t[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) t[i] = v[j + i];
                  if (p.accumulate) {
                    float o[8];
                    load8(dst + j, o);
#pragma unroll
                    for (int i = 0; i < 8; ++i) t[i] += o[i];
                  }
                  store8(dst + j, t);
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) dst[j] = __float2bfloat16_rn(v[j] + (p.accumulate ? __bfloat162float(dst[j]) : 0.f));
              }
            } else {
              float* dst = static_cast<float*>(p.C[z]) + row_off + col0;
              if (full_cols && vec_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 t = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                  if (p.accumulate) {
                    const float4 o = *reinterpret_cast<const float4*>(dst + j);
                    t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
                  }
                  *reinterpret_cast<float4*>(dst + j) = t;
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) dst[j] = v[j] + (p.accumulate ? dst[j] : 0.f);
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// ---- host: tensor maps -----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// extent: tile extent along the operand's M/N index (128 for A, BN for B)
int make_map(const TcOperand& o, int z, int extent, CUtensorMap* out) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return SIG_ERR_ARCH;
  void* ptr = const_cast<void*>(o.ptr[z]);
  if (!ptr) return SIG_ERR_NULL;
  if (!aligned16(ptr)) return SIG_ERR_ALIGN;
  cuuint64_t gdim[3];
  cuuint64_t gstr[2];
  cuuint32_t box[3];
  cuuint32_t estr[3] = {1, 1, 1};
  cuuint32_t rank;
  if (o.mode == TC_K2D || o.mode == TC_MN2D) {
    if ((o.ld * 2) % 16) return SIG_ERR_ALIGN;
    rank = 2;
    gdim[0] = (cuuint64_t)o.cols; gdim[1] = (cuuint64_t)o.rows;
    gstr[0] = (cuuint64_t)o.ld * 2;
    box[0] = 64;
    box[1] = o.mode == TC_K2D ? (cuuint32_t)extent : 64;
  } else {
    if ((o.stride_l * 2) % 16 || (o.stride_b * 2) % 16) return SIG_ERR_ALIGN;
    rank = 3;
    gdim[0] = (cuuint64_t)o.cols; gdim[1] = 128; gdim[2] = (cuuint64_t)o.rows;
    gstr[0] = (cuuint64_t)o.stride_l * 2; gstr[1] = (cuuint64_t)o.stride_b * 2;
    box[0] = 64;
    box[1] = o.mode == TC_KTOK ? 128 : 64;
    box[2] = 1;
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : SIG_ERR_SHAPE;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN>
int launch(const TcGemmDesc& g, cudaStream_t s) {
  TcKernelParams p{};
  for (int z = 0; z < g.batch; ++z) {
    SIG_TRY(make_map(g.A, z, BM, &p.ta[z]));
    SIG_TRY(make_map(g.B, z, BN, &p.tb[z]));
    p.C[z] = g.C[z];
    p.bias[z] = g.bias[z];
    if (!g.C[z]) return SIG_ERR_NULL;
  }
  p.a_mode = g.A.mode; p.b_mode = g.B.mode;
  p.M = g.M; p.N = g.N; p.K = g.K; p.batch = g.batch;
  p.ldc = g.ldc; p.out_bf16 = g.out_bf16; p.alpha = g.alpha; p.act = g.act; p.ksplit = g.ksplit < 1 ? 1 : g.ksplit;
  p.c_tok = g.c_tok; p.c_stride_b = g.c_stride_b; p.c_stride_l = g.c_stride_l;
  for (int z = 0; z < g.batch; ++z) p.rowvec[z] = g.rowvec[z];
  p.rowvec_scale = g.rowvec_scale; p.accumulate = g.accumulate;
  p.tiles_m = (int)ceil_div(g.M, BM);
  p.tiles_n = (int)ceil_div(g.N, BN);
  p.kblocks = (int)ceil_div(g.K, BK);
  if (p.ksplit > p.kblocks) p.ksplit = p.kblocks;
  const int units = p.tiles_m * p.tiles_n * p.ksplit * p.batch;
  const int grid = units < num_sms() ? units : num_sms();
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmemBytes);
    attr_set = true;
  }
  tc_gemm_kernel<BN><<<grid, kThreads, Cfg<BN>::kSmemBytes, s>>>(p);
  SIG_CHECK_LAUNCH();
  return 0;
}

}  // namespace

int tc_gemm(const TcGemmDesc& g, cudaStream_t s) {
  if (g.M < 1 || g.N < 1 || g.K < 1 || g.batch < 1 || g.batch > 3) return SIG_ERR_SHAPE;
  if (g.K % 8) return SIG_ERR_SHAPE;
  if (g.ksplit > 1 && (g.out_bf16 || g.act)) return SIG_ERR_SHAPE;
  if ((g.A.mode == TC_KTOK || g.A.mode == TC_MNTOK || g.B.mode == TC_KTOK || g.B.mode == TC_MNTOK) && false) return SIG_ERR_SHAPE;
  if (g.bn == 256) return launch<256>(g, s);
  return launch<128>(g, s);
}

// ---- bf16 helpers ---------------------------------------------------------------------------------
static __global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
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
  cast_bf16_kernel<<<(unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, s>>>(src, dst, n);
  SIG_CHECK_LAUNCH();
  return 0;
}

static __global__ void transpose_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols) {
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
  transpose_bf16_kernel<<<grid, dim3(32, 8), 0, s>>>(src, dst, rows, cols);
  SIG_CHECK_LAUNCH();
  return 0;
}

}  // namespace sig

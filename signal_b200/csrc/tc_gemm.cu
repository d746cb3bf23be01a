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

int stage_override() {
  static const int v = [] {
    const char* e = getenv("SIG_TC_STAGES");
    return e ? atoi(e) : 0;
  }();
  return v;
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
  }
  __device__ static int num_units(const Params& p) { return p.tiles_m * p.tiles_n * p.ksplit * p.batch; }
  __device__ static void decode(const Params& p, int unit, int& z, int& m0, int& n0, int& kb0, int& kb1) {
    const int upb = p.tiles_m * p.tiles_n * p.ksplit;
    const int per = (p.kblocks + p.ksplit - 1) / p.ksplit;
    z = unit / upb;
    int r = unit % upb;
    const int ks = r % p.ksplit;
    r /= p.ksplit;
    n0 = (r % p.tiles_n) * BN;
    m0 = (r / p.tiles_n) * BM * MT;
    kb0 = ks * per;
    kb1 = min(p.kblocks, kb0 + per);
  }
  __device__ static void krange(const Params& p, int unit, int& kb0, int& kb1) {
    int z, m0, n0;
    decode(p, unit, z, m0, n0, kb0, kb1);
  }
  __device__ static void load_one(const CUtensorMap* tm, int mode, uint8_t* dst, uint64_t* bar, int mn0, int kb, int extent) {
    const int k0 = kb * BK;
    if (mode == TC_K2D) tc::load_kmajor_2d(tm, dst, bar, k0, mn0);
    else if (mode == TC_KTOK) tc::load_kmajor_tok(tm, dst, bar, k0, mn0 / 128, extent / 128);
    else if (mode == TC_MN2D) tc::load_mnmajor_2d(tm, dst, bar, mn0, k0, extent);
    else tc::load_mnmajor_tok(tm, dst, bar, mn0, k0 % 128, k0 / 128, extent);   // K index = sample * 128 + l
  }
  __device__ static void load(const Params& p, int unit, int kb, uint8_t* sa, uint8_t* sb, uint64_t* bar) {
    int z, m0, n0, kb0, kb1;
    decode(p, unit, z, m0, n0, kb0, kb1);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) load_one(&p.ta[z], p.a_mode, sa + mt * tc::kATileBytes, bar, m0 + mt * BM, kb, BM);
    load_one(&p.tb[z], p.b_mode, sb, bar, n0, kb, BN);
  }
  __device__ static void epilogue(const Params& p, int unit, int mt, uint32_t tmem_acc, int q, int lane) {
    int z, m0, n0, kb0, kb1;
    decode(p, unit, z, m0, n0, kb0, kb1);
    m0 += mt * BM;
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
        ptx::tmem_ld32(tmem_acc + c * 32, r);
        ptx::tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row < p.M && col0 < p.N && kb1 > kb0) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
          const bool full_cols = col0 + 32 <= p.N;
          if (p.ksplit > 1) {
            float* dst = static_cast<float*>(p.C[z]) + row_off + col0;
            if (full_cols && vec_ok) {   // 16-byte vector reductions (red.global.add.v4.f32)
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]),
                             "f"(v[j + 3])
                             : "memory");
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) atomicAdd(dst + j, v[j]);
            }
          } else {
            if (bias) {
              if (full_cols) {   // unpredicated 16-byte loads, all issued before the first use
                float4 b4[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) b4[j] = *reinterpret_cast<const float4*>(bias + col0 + 4 * j);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  v[4 * j] += b4[j].x; v[4 * j + 1] += b4[j].y; v[4 * j + 2] += b4[j].z; v[4 * j + 3] += b4[j].w;
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) v[j] += bias[col0 + j];
              }
            }
            if (p.pre[z]) {
              float* pd = p.pre[z] + (long long)row * p.ldc + col0;
              if (full_cols && vec_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(pd + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) pd[j] = v[j];
              }
            }
            if (p.act == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
            }
            if (rvec) {
              if (full_cols) {
                float4 r4[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) r4[j] = *reinterpret_cast<const float4*>(rvec + col0 + 4 * j);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  v[4 * j] = fmaf(rscale, r4[j].x, v[4 * j]); v[4 * j + 1] = fmaf(rscale, r4[j].y, v[4 * j + 1]);
                  v[4 * j + 2] = fmaf(rscale, r4[j].z, v[4 * j + 2]); v[4 * j + 3] = fmaf(rscale, r4[j].w, v[4 * j + 3]);
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) v[j] = fmaf(rscale, rvec[col0 + j], v[j]);
              }
            }
            if (p.C2[z]) {
              __nv_bfloat16* d2 = static_cast<__nv_bfloat16*>(p.C2[z]) + (long long)row * p.ldc2 + col0;
              if (full_cols && (p.ldc2 % 8) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  float t[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) t[i] = v[j + i];
                  store8(d2 + j, t);
                }
              } else
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) d2[j] = __float2bfloat16_rn(v[j]);
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
  }
};

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
  const int kpu = (p.kblocks + p.ksplit - 1) / p.ksplit;
  if (amn && bmn) return tc::launch<BN, GemmProblem<BN, true, true, MT>, MT>(p, units, s, kpu);
  if (amn) return tc::launch<BN, GemmProblem<BN, true, false, MT>, MT>(p, units, s, kpu);
  if (bmn) return tc::launch<BN, GemmProblem<BN, false, true, MT>, MT>(p, units, s, kpu);
  return tc::launch<BN, GemmProblem<BN, false, false, MT>, MT>(p, units, s, kpu);
}

}  // namespace

int tc_gemm(const TcGemmDesc& g, cudaStream_t s) {
  if (g.M < 1 || g.N < 1 || g.K < 1 || g.batch < 1 || g.batch > 8) return SIG_ERR_SHAPE;
  if (g.ksplit > 1 && (g.out_bf16 || g.act)) return SIG_ERR_SHAPE;
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

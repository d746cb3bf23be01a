// Generic bf16 tcgen05 GEMM (TMA-fed, TMEM accumulators, persistent, warp-specialised).
//   C[z][M,N] = alpha * A[z] . B[z]^T (+ bias[z][n]) (GELU optional), fp32 accumulate,
// operands described by TMA tensor maps so that strided token views are consumed in place.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace sig {

enum TcOpMode : int {
  TC_K2D = 0,    // row-major [rows, K] matrix, rows = M or N index                (K-major operand)
  TC_KTOK = 1,   // token view [B, L=128, d]: rows = (b, l), K = d                  (K-major operand, 3-D map)
  TC_MN2D = 2,   // row-major [K, cols] matrix, cols = M or N index                 (MN-major operand)
  TC_MNTOK = 3,  // token view [B, L=128, d]: K = (b, l) positions, cols = channels (MN-major operand, 3-D map)
};

struct TcOperand {
  const void* ptr[8];       // per batch entry (bf16)
  int mode;
  int64_t ld;               // 2-D modes: row pitch in elements
  int64_t stride_b, stride_l;  // token modes: element strides; B samples of L=128 rows
  int64_t rows, cols;       // 2-D modes: matrix extents (rows x cols as stored); token modes: rows = B, cols = d
};

struct TcGemmDesc {
  TcOperand A, B;
  int M, N, K, batch;       // batch <= 8
  void* C[8];
  int64_t ldc;
  int out_bf16;             // 0: fp32 C, 1: bf16 C
  const float* bias[8];     // optional, per n
  float alpha;
  int act;                  // 0 none, 1 GELU
  int ksplit;               // > 1: fp32 atomic accumulation into a pre-zeroed C (bias allowed: added once; no act / bf16 / C2 / pre)
  int bn;                   // 128 or 256 (N tile)
  int mt;                   // 1 or 2 (with bn == 256): 128-row M tiles per work unit sharing one B tile
  int pair;                 // 1 (with bn == 256, mt == 1): 256 x 256 units on CTA pairs (cta_group::2), see tc_gemm.cu
  // token-strided output: row = (b, l) with l in [0,128): element (row, n) at C + b*c_stride_b + l*c_stride_l + n
  int c_tok;
  int64_t c_stride_b, c_stride_l;
  // optional per-sample row vector: v += (*rowvec_scale) * rowvec[z][(row / 128) * N + n]
  const float* rowvec[8];
  const float* rowvec_scale;
  int accumulate;           // v += existing C (non split-K)
  void* C2[8];              // optional bf16 copy of the stored value (row pitch ldc2)
  int64_t ldc2;
  float* pre[8];            // optional fp32 copy of the value before the activation (row pitch ldc)
  // Optional extra 64-wide K block appended to the contraction, with PER-SAMPLE operands (M tile = one sample,
  // batch entry z = modality): C[z][(b,l), n] += sum_k xa[b][z*128 + l][k] * xb[b*64 + k][n].
  //   xa: [xB][384][64] bf16 (K-major rows), xb: [xB*64][N] bf16 (read MN-major).
  // Requires A K-major (2-D), B MN-major, ksplit == 1, mt == 1.  (SIM's token gradient inside AlignM's dX GEMM.)
  const void* xa;
  const void* xb;
  int xB;
};

inline TcGemmDesc tc_desc() {
  TcGemmDesc g{};
  g.batch = 1; g.alpha = 1.f; g.ksplit = 1; g.bn = 128; g.mt = 1;
  return g;
}
inline TcOperand tc_k2d(const void* p, int64_t rows, int64_t k, int64_t ld) {
  TcOperand o{}; for (int i = 0; i < 8; ++i) o.ptr[i] = p; o.mode = TC_K2D; o.ld = ld; o.rows = rows; o.cols = k; return o;
}
inline TcOperand tc_mn2d(const void* p, int64_t k, int64_t cols, int64_t ld) {
  TcOperand o{}; for (int i = 0; i < 8; ++i) o.ptr[i] = p; o.mode = TC_MN2D; o.ld = ld; o.rows = k; o.cols = cols; return o;
}

// per-batch base pointers: entry z = base + z * stride_elems (bf16 elements)
inline TcOperand tc_batched(TcOperand o, const void* base, size_t stride_elems, int n = 3) {
  for (int z = 0; z < n; ++z) o.ptr[z] = static_cast<const __nv_bfloat16*>(base) + z * stride_elems;
  return o;
}

// Enqueue the GEMM on `s`.  Returns 0 / SIG_ERR_* / cudaError_t.
int tc_gemm(const TcGemmDesc& g, cudaStream_t s);

// bf16 helpers used around the GEMMs
int cast_f32_to_bf16(const float* src, __nv_bfloat16* dst, int64_t n, cudaStream_t s);
// several fp32 -> bf16 casts in one launch (n <= 8 jobs)
struct CastJob {
  const float* src;
  __nv_bfloat16* dst;
  int64_t n;
};
int cast_f32_to_bf16_multi(const CastJob* jobs, int n, cudaStream_t s);
// dst[c][r] = bf16(src[r][c]) for a row-major [rows, cols] fp32 matrix
int transpose_f32_to_bf16(const float* src, __nv_bfloat16* dst, int rows, int cols, cudaStream_t s);

int tc_read_stamps(long long* out16);   // SIG_TC_STAMPS=1 debug aid (tc_pipeline.cuh)
bool tc_pair_enabled();   // SIG_TC_PAIR=0 in the environment keeps the large GEMMs on the single-CTA kernel
int tc_num_sms();   // multiprocessor count of the current device (cached)

}  // namespace sig

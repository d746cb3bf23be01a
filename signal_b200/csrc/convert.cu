// fp16 boundary of the bf16 tensor-core path (include/signal_b200.h: sig_convert_half).
//
// The reference trains under torch.cuda.amp.autocast() = fp16 (engine/processor.py:165): the backbone's last matmul
// (clip/model.py:487) hands SIM and AlignM fp16 token maps.  The kernels of this library compute on bf16 operands with
// fp32 accumulation, so fp16 callers are served by ONE streaming conversion pass per token map at the module boundary
// (fp16 -> bf16: 8 exponent bits hold every fp16 value, the mantissa is rounded from 10 to 7 bits, RNE) and one pass
// back for the token gradients (bf16 -> fp16, RNE, saturating to +-65504 like a GradScaler overflow would show).
// HBM-bound: 4 B per element moved, 16-byte vector accesses, grid = a multiple of the SM count.
#include <cuda_fp16.h>

#include "common.cuh"

namespace sig {
namespace {

template <bool kToBf16>
__device__ __forceinline__ uint32_t cvt2(uint32_t w) {
  if (kToBf16) {
    const __half2 h = *reinterpret_cast<const __half2*>(&w);
    const float2 f = __half22float2(h);
    const __nv_bfloat162 b = __floats2bfloat162_rn(f.x, f.y);
    return *reinterpret_cast<const uint32_t*>(&b);
  } else {
    const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
    const __half2 h = __floats2half2_rn(fminf(fmaxf(lo, -65504.f), 65504.f), fminf(fmaxf(hi, -65504.f), 65504.f));
    return *reinterpret_cast<const uint32_t*>(&h);
  }
}

// rows = nb * nl rows of d elements (d % 8 == 0); row (b, l) at base + b * sb + l * sl (element strides)
template <bool kToBf16>
__global__ void __launch_bounds__(256) convert_half_kernel(const uint16_t* __restrict__ src, int64_t ssb, int64_t ssl,
                                                          uint16_t* __restrict__ dst, int64_t dsb, int64_t dsl, int nb, int nl, int d) {
  pdl_enter();
  const int vec_per_row = d / 8;
  const int64_t total = (int64_t)nb * nl * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec_per_row);
    const int64_t row = i / vec_per_row;
    const int l = (int)(row % nl);
    const int64_t b = row / nl;
    const uint4 raw = *reinterpret_cast<const uint4*>(src + b * ssb + l * ssl + 8 * v);
    uint4 out;
    out.x = cvt2<kToBf16>(raw.x); out.y = cvt2<kToBf16>(raw.y); out.z = cvt2<kToBf16>(raw.z); out.w = cvt2<kToBf16>(raw.w);
    *reinterpret_cast<uint4*>(dst + b * dsb + l * dsl + 8 * v) = out;
  }
}

}  // namespace

int convert_half(const void* src, int64_t ssb, int64_t ssl, int to_bf16, void* dst, int64_t dsb, int64_t dsl, int nb, int nl, int d,
                 cudaStream_t s) {
  const int64_t total = (int64_t)nb * nl * (d / 8);
  int64_t blocks = ceil_div(total, 256 * 4);
  const int cap = device_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (to_bf16) {
    SIG_LAUNCH((convert_half_kernel<true>), (unsigned)blocks, 256, 0, s, static_cast<const uint16_t*>(src), ssb, ssl,
               static_cast<uint16_t*>(dst), dsb, dsl, nb, nl, d);
  } else {
    SIG_LAUNCH((convert_half_kernel<false>), (unsigned)blocks, 256, 0, s, static_cast<const uint16_t*>(src), ssb, ssl,
               static_cast<uint16_t*>(dst), dsb, dsl, nb, nl, d);
  }
  SIG_CHECK_LAUNCH();
  return 0;
}

}  // namespace sig

// Shared device/host helpers for libsignal_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "prof.h"
#include "signal_b200.h"

#define SIG_CHECK_LAUNCH()                         \
  do {                                             \
    ::sig::prof_count_launch();                    \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

// SIG_LAUNCH((kernel<...>), grid, block, smem_bytes, stream, args...)
#define SIG_LAUNCH(kernel, grid, block, smem, stream, ...) \
  ::sig::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, ##__VA_ARGS__)

#define SIG_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != 0) return rc__;  \
  } while (0)

namespace sig {

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and
// starts with pdl_enter(): it lets the next kernel of the stream be scheduled early (its launch latency
// and prologue overlap this kernel) and then waits until the previous kernel has completed and its
// writes are visible.  Nothing touches global memory before the wait, so stream order is preserved.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_launch_dependents();
  pdl_wait();
}

bool pdl_enabled();   // prof.cu: SIG_PDL=0 in the environment turns the launch attribute off

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) function attribute: it is set the first
// time a kernel is launched on EACH device (prof.cu keeps a mutex-guarded set of (function, device) pairs), so a
// process that touches several GPUs -- tests on cuda:1, multi-device threads -- gets it on every one of them.
bool first_launch_on_device(const void* fn);
template <class K>
inline void ensure_dyn_smem(K kernel, int bytes) {
  if (first_launch_on_device(reinterpret_cast<const void*>(kernel)))
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
int device_num_sms();   // multiprocessor count of the CURRENT device (cached per device)
// infer.cu: inference tail (feature concat + L2 norm, euclidean distance matrix, CMC / mAP)
int infer_features(const void* const cls[3], const int64_t cls_stride_b[3], const void* sim_out, int64_t ld_sim, int dtype, int B, int d,
                   int normalize, float* out, cudaStream_t s);
int euclidean_distmat(const float* qf, const float* gf, int nq, int ng, int D, float* dist, float* ws, cudaStream_t s);
int rank_eval(const float* dist, int64_t ld, const int64_t* q_pids, const int64_t* g_pids, const int64_t* q_cams, const int64_t* g_cams,
              int nq, int ng, int max_rank, float* cmc, double* map_out, double* stats, int* overflow, cudaStream_t s);
// xchg.cu: in-place all-reduce of a symmetric fp32 arena over NVLink peer memory / NVLS
size_t xchg_flag_bytes();
int xchg_allreduce_f32(const sig_xchg_peers* peers, size_t off, size_t count, float scale, int ctas, cudaStream_t s);
// convert.cu: fp16 <-> bf16 over a strided [nb, nl, d] map (element strides, unit channel stride)
int convert_half(const void* src, int64_t ssb, int64_t ssl, int to_bf16, void* dst, int64_t dsb, int64_t dsl, int nb, int nl, int d,
                 cudaStream_t s);

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr int kHeads = 8;        // useA.py:449
constexpr int kMaxL = 128;       // patch tokens per modality (make_model.py:67)
constexpr float kLnEps = 1e-5f;  // nn.LayerNorm default
constexpr float kLabelSmooth = 0.1f;  // useB.py:121-122

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- scalar conversions ---------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive channels -> fp32 registers (16 B load for bf16, 2 x 16 B for fp32).
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 raw;
  uint32_t* w = reinterpret_cast<uint32_t*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = raw;
}

// 4 consecutive bf16 channels (8 B).
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 raw = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(raw.x << 16); v[1] = __uint_as_float(raw.x & 0xffff0000u);
  v[2] = __uint_as_float(raw.y << 16); v[3] = __uint_as_float(raw.y & 0xffff0000u);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&v)[4]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 raw;
  raw.x = *reinterpret_cast<const uint32_t*>(&a);
  raw.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}

// 4-byte global load the compiler may not sink to its use (an invariant `const __restrict__` load is
// re-scheduled next to its consumer to save registers, which serialises the memory latency).
__device__ __forceinline__ uint32_t ld_global_u32_early(const void* p) {
  uint32_t v;
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// ---- cp.async (LDGSTS) staging -------------------------------------------------
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- math -------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float x) {  // nn.GELU() erf form
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {  // Phi(x) + x phi(x)
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Fast GELU value + derivative for the bf16 path: erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7),
// sharing exp(-x^2/2) between the cdf and the pdf.  No branches, two MUFU ops per element.
__device__ __forceinline__ void gelu_fast(float x, float& g, float& dg) {
  const float ay = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ay, 1.0f));
  const float e = __expf(-0.5f * x * x);
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float half_tail = 0.5f * poly * t * e;          // 0.5 * (1 - erf(|y|))
  const float cdf = x >= 0.f ? 1.0f - half_tail : half_tail;
  g = x * cdf;
  dg = fmaf(x * 0.39894228040143267794f, e, cdf);
}
__device__ __forceinline__ float gelu_fast_f(float x) {
  float g, dg;
  gelu_fast(x, g, dg);
  return g;
}

// tanh-form GELU value + derivative for the HBM-bound LAM depthwise kernels of the bf16 path, where the
// erf form above is the bottleneck (issue slots, not memory).  One MUFU.TANH and 7 FP32 ops per element.
// |gelu_tanh - gelu_erf| <= 5e-4 absolute (at |x| ~ 2), below the 4e-3 rounding step of the bf16
// pre-activation it is applied to.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(x2, 0.0356774081f, 0.7978845608f));
  return x * fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ void gelu_tanh(float x, float& g, float& dg) {
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(x2, 0.0356774081f, 0.7978845608f));
  const float cdf = fmaf(0.5f, t, 0.5f);
  g = x * cdf;
  dg = fmaf(0.5f * x * fmaf(-t, t, 1.0f), fmaf(x2, 0.1070322243f, 0.7978845608f), cdf);
}

// ---- reductions ---------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum; `scratch` holds >= 33 floats; result broadcast to all threads.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // scratch may still be read from a previous call
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < nw ? scratch[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < nw ? scratch[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// ---- entry-point prologue --------------------------------------------------------
// Every extern "C" entry runs on the device ordinal it is given and restores the caller's current device on return;
// a sticky error left by earlier, unrelated CUDA calls of the process is cleared so that it is not reported as ours.
struct DeviceGuard {
  int prev = -1;
  int rc = 0;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) {
      cudaError_t e = cudaSetDevice(dev);
      if (e != cudaSuccess) rc = (int)e;
    }
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};
#define SIG_ENTER(device)              \
  ::sig::DeviceGuard guard__(device);  \
  if (guard__.rc) return guard__.rc;   \
  cudaGetLastError();

// ---- host-side argument checks ---------------------------------------------------
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t elem_size(int dtype) { return dtype == SIG_BF16 ? 2 : 4; }

inline int check_tokens(const sig_tokens* t, bool need_cls) {
  if (!t) return SIG_ERR_NULL;
  if (t->dtype != SIG_F32 && t->dtype != SIG_BF16) return SIG_ERR_DTYPE;
  if (t->B < 1 || t->B > 4096 || t->L < 1 || t->L > kMaxL || t->d < 64 || t->d > 1024 || (t->d % 64) != 0)
    return SIG_ERR_SHAPE;
  const size_t es = elem_size(t->dtype);
  for (int m = 0; m < 3; ++m) {
    if (!t->patch[m]) return SIG_ERR_NULL;
    if (!aligned16(t->patch[m]) || (t->patch_stride_b[m] * es) % 16 || (t->patch_stride_l[m] * es) % 16)
      return SIG_ERR_ALIGN;
    if (need_cls) {
      if (!t->cls[m]) return SIG_ERR_NULL;
      if (!aligned16(t->cls[m]) || (t->cls_stride_b[m] * es) % 16) return SIG_ERR_ALIGN;
    }
  }
  return 0;
}

inline int check_token_grads(const sig_token_grads* g, int dtype, bool need_cls) {
  if (!g) return SIG_ERR_NULL;
  const size_t es = elem_size(dtype);
  for (int m = 0; m < 3; ++m) {
    if (!g->dpatch[m]) return SIG_ERR_NULL;
    if (!aligned16(g->dpatch[m]) || (g->patch_stride_b[m] * es) % 16 || (g->patch_stride_l[m] * es) % 16)
      return SIG_ERR_ALIGN;
    if (need_cls) {
      if (!g->dcls[m]) return SIG_ERR_NULL;
    }
    if (g->dcls[m] && (!aligned16(g->dcls[m]) || (g->cls_stride_b[m] * es) % 16)) return SIG_ERR_ALIGN;
  }
  return 0;
}

// Bump allocator over the caller's ctx buffer (256 B aligned sub-buffers).
struct Arena {
  char* base;
  size_t off;
  explicit Arena(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t n) {
    T* r = reinterpret_cast<T*>(base + off);
    off += (n * sizeof(T) + 255) & ~size_t(255);
    return r;
  }
};

}  // namespace sig

// Gradient exchange of the data-parallel head (the reference: DDP all-reduce, engine/processor.py:100-105)
// as ONE kernel over NVLink peer memory -- include/signal_b200.h: sig_xchg_allreduce_f32.
//
// Why not ncclAllReduce: the exchange has to run UNDER the backward, whose GEMM / ring kernels are persistent and hold
// all 148 SMs with ~200 KB of shared memory each.  NCCL's kernels want tens of KB of shared memory and up to 640 threads
// per CTA, so they only get onto the SMs between two compute kernels and then delay the next one (measured round 1:
// a collective costs 2-3x its stand-alone time there, 8-GPU efficiency 0.78).  This kernel uses NO shared memory and
// few registers, so its CTAs are co-resident with the persistent kernels on the same SMs, and it is bound by NVLink
// latency/bandwidth, not by SM time.
//
// Algorithm (two-shot, in place, every rank runs the same kernel on its own symmetric arena):
//   barrier A   every rank's piece [off, off+count) is final                (flags in peer memory, release/acquire.sys)
//   reduce      rank r owns slice r of the piece: sum over all ranks -- with NVLS one multimem.ld_reduce on the
//               multicast address (the NVSwitch adds the replicas), else N-1 peer loads -- times `scale`
//   broadcast   the owner writes the reduced slice into every rank's arena -- one multimem.st, else N-1 peer stores
//   barrier B   all slices have landed everywhere, nobody reads the old values any more
// Per rank and direction this moves count*4*(N-1)/N bytes over NVLink (P2P) or count*4/N through the switch (NVLS).
// Each CTA synchronises only with the CTA of the same index on the peers (own flag slots, monotonically increasing
// epochs kept in device memory, so CUDA-graph replays need no host state and the flags are never reset).
#include <cstdlib>

#include "common.cuh"

namespace sig {
namespace {

constexpr int kXchgMaxRanks = 8;
constexpr int kXchgMaxCtas = 1024;
// 128 threads x 32 registers = 4096 registers per CTA: exactly what a 384-thread, 160-register tcgen05 GEMM CTA
// (tc_pipeline.cuh: SIG_TC_MAXNREG) leaves free on its SM, and no shared memory -- so one CTA of this kernel runs NEXT TO
// every persistent compute CTA of the backward (more per SM next to the lighter kernels).
constexpr int kXchgThreads = 128;
constexpr int kXchgUnroll = 4;      // multimem path: 4 x 16 B per thread in flight
constexpr int kXchgUnrollP2P = 2;   // peer-load path (accumulator + incoming value per slot): 2 x 16 B
// flag region of one rank (uint32): [cta][src rank] arrival flags, then [cta] the rank-private epoch counters
// ... then 16 words of phase time stamps of CTA 0's last call (%globaltimer, ns: start, after barrier A, after the data
// loop, after barrier B) -- a measurement aid, read by tools/bench_xchg.py
constexpr int kXchgStampWord = kXchgMaxCtas * kXchgMaxRanks + kXchgMaxCtas;
constexpr int kXchgFlagWords = kXchgStampWord + 16;
#ifndef SIG_XCHG_SPIN_CLOCKS
#define SIG_XCHG_SPIN_CLOCKS 60000000000LL   /* ~30 s: a peer that never arrives traps instead of hanging the box */
#endif

struct XchgArgs {
  float* buf[kXchgMaxRanks];
  uint32_t* flags[kXchgMaxRanks];
  float* mc;
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// All CTAs of index `cta` on all ranks meet here.  Thread t < world signals peer t and waits for peer t.
__device__ __forceinline__ void xchg_barrier(const XchgArgs& a, int cta, uint32_t val) {
  __syncthreads();   // every thread's data accesses of this CTA happen-before the signalling threads' release below
  if ((int)threadIdx.x < a.world) {
    const int peer = threadIdx.x;
    // (st.release.sys is cumulative over what bar.sync ordered before it: no separate fence.sys -- it measured ~3 us)
    st_release_sys(a.flags[peer] + cta * kXchgMaxRanks + a.rank, val);
    const uint32_t* mine = a.flags[a.rank] + cta * kXchgMaxRanks + peer;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - val) < 0) {
      if (clock64() - t0 > SIG_XCHG_SPIN_CLOCKS) __trap();
    }
  }
  __syncthreads();
}

template <bool kMultimem>
__global__ void __launch_bounds__(kXchgThreads, 16) xchg_allreduce_kernel(const XchgArgs a, size_t off, size_t n4, float scale) {
  const int cta = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
  __shared__ uint32_t epoch_s;
  uint32_t* epoch_p = a.flags[a.rank] + kXchgMaxCtas * kXchgMaxRanks + cta;
  if (tid == 0) epoch_s = *epoch_p + 2;          // every call uses two flag values: e - 1 (barrier A) and e (barrier B)
  __syncthreads();
  const uint32_t e = epoch_s;
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(a.flags[a.rank] + kXchgStampWord);
  auto stamp = [&](int i) {
    if (cta == 0 && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      stamps[i] = t;
    }
  };
  stamp(0);
  xchg_barrier(a, cta, e - 1);
  stamp(1);

  const size_t per = (n4 + a.world - 1) / a.world;
  const size_t lo = min(n4, (size_t)a.rank * per), hi = min(n4, lo + per);
  constexpr int U = kMultimem ? kXchgUnroll : kXchgUnrollP2P;
  const size_t chunk = (size_t)kXchgThreads * U;
  for (size_t base = lo + (size_t)cta * chunk; base < hi; base += (size_t)G * chunk) {
    float4 acc[U];
    if (kMultimem) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t i = base + (size_t)u * kXchgThreads + tid;
        if (i < hi) acc[u] = multimem_ld_reduce(a.mc + off + 4 * i);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t i = base + (size_t)u * kXchgThreads + tid;
        if (i < hi) {
          acc[u].x *= scale; acc[u].y *= scale; acc[u].z *= scale; acc[u].w *= scale;
          multimem_st(a.mc + off + 4 * i, acc[u]);
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      // fixed summation order (rank 0, 1, ...) on every owner: the result does not depend on who reduces
      for (int p = 0; p < a.world; ++p) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t i = base + (size_t)u * kXchgThreads + tid;
          if (i < hi) v[u] = ld_peer(a.buf[p] + off + 4 * i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t i = base + (size_t)u * kXchgThreads + tid;
          if (i < hi) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t i = base + (size_t)u * kXchgThreads + tid;
        if (i < hi) {
          acc[u].x *= scale; acc[u].y *= scale; acc[u].z *= scale; acc[u].w *= scale;
          for (int p = 0; p < a.world; ++p) st_peer(a.buf[p] + off + 4 * i, acc[u]);
        }
      }
    }
  }
  stamp(2);
  xchg_barrier(a, cta, e);
  stamp(3);
  if (tid == 0) *epoch_p = e;
}

}  // namespace

size_t xchg_flag_bytes() { return (size_t)kXchgFlagWords * sizeof(uint32_t); }

int xchg_allreduce_f32(const sig_xchg_peers* pr, size_t off, size_t count, float scale, int ctas, cudaStream_t s) {
  if (!pr) return SIG_ERR_NULL;
  if (pr->world < 1 || pr->world > kXchgMaxRanks || pr->rank < 0 || pr->rank >= pr->world) return SIG_ERR_SHAPE;
  if ((off | count) % 4) return SIG_ERR_ALIGN;
  XchgArgs a{};
  for (int r = 0; r < pr->world; ++r) {
    if (!pr->buf[r] || !pr->flags[r]) return SIG_ERR_NULL;
    if (((uintptr_t)pr->buf[r] | (uintptr_t)pr->flags[r]) & 15) return SIG_ERR_ALIGN;
    a.buf[r] = static_cast<float*>(pr->buf[r]);
    a.flags[r] = static_cast<uint32_t*>(pr->flags[r]);
  }
  a.mc = static_cast<float*>(pr->multicast);
  a.rank = pr->rank;
  a.world = pr->world;
  if (count == 0) return 0;
  if (ctas <= 0) ctas = device_num_sms();   // ONE CTA per SM: that is what fits next to a persistent GEMM CTA (a second wave
                                            // would wait for room and then delay the next compute kernel -- measured)
  if (ctas > kXchgMaxCtas) ctas = kXchgMaxCtas;
  const size_t n4 = count / 4;
  // Co-residency with the compute kernels: an SM runs CTAs of two kernels at the same time only if both accept its
  // current L1 / shared-memory split.  The persistent GEMM and ring kernels carve out (nearly) all of it as shared
  // memory, so this kernel -- which uses none -- must ask for the same carve-out or it waits for whole SMs to drain.
  static const int carve = [] {
    const char* e = getenv("SIG_XCHG_CARVEOUT");
    return e ? atoi(e) : (int)cudaSharedmemCarveoutMaxShared;
  }();
  if (carve >= 0) {
    if (first_launch_on_device(reinterpret_cast<const void*>(xchg_allreduce_kernel<true>))) {
      cudaFuncSetAttribute(xchg_allreduce_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      cudaFuncSetAttribute(xchg_allreduce_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    }
  }
  // (no programmatic dependent launch: the kernel is ordered behind the gradient producers by events on other streams)
  if (a.mc) xchg_allreduce_kernel<true><<<ctas, kXchgThreads, 0, s>>>(a, off, n4, scale);
  else      xchg_allreduce_kernel<false><<<ctas, kXchgThreads, 0, s>>>(a, off, n4, scale);
  SIG_CHECK_LAUNCH();
  return 0;
}

}  // namespace sig

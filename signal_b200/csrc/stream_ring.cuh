// Streaming skeleton for the HBM-bound token passes: a persistent CTA per SM owns a contiguous range of
// equally sized, contiguous items; ONE producer thread keeps a ring of shared-memory stages full with 1-D bulk
// copies (cp.async.bulk, completion counted in bytes on an mbarrier), the consumer warps work from shared memory
// and hand each stage back through a second mbarrier.  The bytes in flight per SM are the ring size (144-192 KB),
// independent of the consumers' register budget -- tools/membw measured that ~128 KB in flight per SM are needed
// to reach 6 TB/s, which register-staged loops (16-48 KB per SM) could not provide.
#pragma once
#include "ptx.cuh"

namespace sig {
namespace ring {

// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar` (complete_tx)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol = ptx::kPolNormal) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(ptx::smem_u32(dst)), "l"(src), "r"(bytes), "r"(ptx::smem_u32(bar)), "l"(pol)
               : "memory");
}

// barrier among the first `nthreads` threads of the CTA only (the producer warp does not take part)
__device__ __forceinline__ void consumer_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

template <int STAGES>
struct Bars {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
};

// thread 0 initialises; every thread of the CTA must call this (it ends with __syncthreads)
template <int STAGES>
__device__ __forceinline__ void init(Bars<STAGES>* b, int consumer_warps) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&b->full[s], 1);
      ptx::mbar_init(&b->empty[s], (uint32_t)consumer_warps);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
}

// producer side of item number k (0-based count of this CTA's items): waits until stage k % STAGES is free,
// then starts the copy
template <int STAGES>
__device__ __forceinline__ void produce(Bars<STAGES>* b, int k, void* stage, const void* src, uint32_t bytes,
                                        uint64_t pol = ptx::kPolNormal) {
  const int s = k % STAGES;
  if (k >= STAGES) ptx::mbar_wait(&b->empty[s], (uint32_t)((k / STAGES) - 1) & 1u);
  ptx::mbar_expect_tx(&b->full[s], bytes);
  bulk_g2s(stage, src, bytes, &b->full[s], pol);
}

template <int STAGES>
__device__ __forceinline__ void consumer_wait(Bars<STAGES>* b, int k) {
  ptx::mbar_wait(&b->full[k % STAGES], (uint32_t)(k / STAGES) & 1u);
}

// one arrival per consumer warp, after all its lanes are done with the stage
template <int STAGES>
__device__ __forceinline__ void consumer_release(Bars<STAGES>* b, int k) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) ptx::mbar_arrive(&b->empty[k % STAGES]);
}

}  // namespace ring
}  // namespace sig

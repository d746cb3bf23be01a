// Optional phase profiler + launch counter (debug facility; off by default, mutex guarded).
// When enabled, each SIG_PHASE scope brackets its launches with a pair of CUDA events on the
// launching stream; sig_profile_collect() synchronises those events and reports per-phase time.
#pragma once
#include <cuda_runtime.h>

namespace sig {
void prof_count_launch();
struct ProfScope {
  int idx;
  cudaStream_t s;
  ProfScope(const char* name, cudaStream_t stream);
  ~ProfScope();
};
}  // namespace sig

#define SIG_PHASE(name) ::sig::ProfScope prof_scope__(name, s)

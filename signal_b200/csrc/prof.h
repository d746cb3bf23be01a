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
  bool capturing;
  ProfScope(const char* name, cudaStream_t stream);
  ~ProfScope();
};
}  // namespace sig

#define SIG_PHASE(name) ::sig::ProfScope prof_scope__(name, s)

namespace sig {
// Internal side streams: independent sub-chains of one entry point (e.g. the GAM grid next to the LAM
// GEMMs, weight-gradient GEMMs next to the activation-gradient chain) are enqueued on a library-owned
// non-blocking stream, forked from and joined back into the caller's stream with events, so the call
// is still ordered on `stream` as a whole and stays CUDA-graph capturable (the fork/join becomes
// parallel branches of the graph).  One stream + two events per slot, created once per device.
enum ForkSlot { FORK_SIM_FWD = 0, FORK_SIM_BWD = 1, FORK_ALIGN_FWD = 2, FORK_ALIGN_BWD = 3, FORK_ALIGN_BWD2 = 4, FORK_SLOTS = 5 };
struct Fork {
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool ok() const { return side != nullptr; }
  void fork(cudaStream_t main) const {   // side continues from main's current position
    cudaEventRecord(ev_fork, main);
    cudaStreamWaitEvent(side, ev_fork, 0);
  }
  void join(cudaStream_t main) const {   // main waits for everything enqueued on side
    cudaEventRecord(ev_join, side);
    cudaStreamWaitEvent(main, ev_join, 0);
  }
};
Fork get_fork(int slot);   // returns a disabled Fork (ok() == false) when SIG_FORK=0

int device_num_sms();
// SM budget of the calling thread's current entry point (prof.cu): persistent kernels size their grids with
// sm_budget() instead of the device's SM count.  sig_align_fwd / sig_align_bwd run with SIG_ALIGN_SMS (environment,
// default: all) so that AlignM's long persistent kernels leave SMs to the short kernels of SIM's dependency chain, which
// runs on another stream and is the critical path of the fused step.
int sm_budget();
struct ScopedSmBudget {
  int prev;
  explicit ScopedSmBudget(int n);
  ~ScopedSmBudget();
};
int align_sm_budget();   // SIG_ALIGN_SMS or 0 (= all)
// Waves: a persistent kernel holds its SMs until it is done, so a short high-priority kernel of another stream waits for
// the whole kernel.  With sm_waves() = k > 1 the calling entry point launches its persistent kernels with k x as many
// CTAs (each doing 1/k of the work): SMs are handed back k times as often and the hardware scheduler can slot the other
// stream's CTAs in between (it dispatches pending CTAs of the higher-priority stream first).
int sm_waves();
struct ScopedSmWaves {
  int prev;
  explicit ScopedSmWaves(int k);
  ~ScopedSmWaves();
};
int align_sm_waves();    // SIG_ALIGN_WAVES or 1
}  // namespace sig

#include "prof.h"

#include <atomic>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <set>
#include <utility>
#include <string>
#include <vector>

#include "signal_b200.h"

namespace sig {
namespace {
std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_enabled{0};
std::mutex g_mu;
struct Rec {
  std::string name;
  cudaEvent_t a, b;
};
std::vector<Rec> g_recs;
}  // namespace

Fork get_fork(int slot) {
  static std::mutex mu;
  static Fork forks[8][FORK_SLOTS];
  static const bool enabled = [] {
    const char* e = getenv("SIG_FORK");
    return !(e && e[0] == '0');
  }();
  Fork none;
  if (!enabled || slot < 0 || slot >= FORK_SLOTS) return none;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 8) return none;
  std::lock_guard<std::mutex> lk(mu);
  Fork& f = forks[dev][slot];
  if (!f.side) {
    cudaStream_t st = nullptr;
    cudaEvent_t a = nullptr, b = nullptr;
    // SIG_PRIO=1 (experiment, off by default: measured 6 % slower at B = 128): SIM's side streams get the highest
    // priority and AlignM's the lowest, so that SIM's chain of short kernels goes first when both have work pending
    static const bool prio = [] {
      const char* e = getenv("SIG_PRIO");
      return e && e[0] == '1';
    }();
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // lo = least (numerically largest), hi = greatest
    const int pr = !prio ? lo : ((slot == FORK_SIM_FWD || slot == FORK_SIM_BWD) ? hi : lo);
    if (cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, pr) != cudaSuccess) return none;
    if (cudaEventCreateWithFlags(&a, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b, cudaEventDisableTiming) != cudaSuccess)
      return none;
    f.side = st; f.ev_fork = a; f.ev_join = b;
  }
  return f;
}

bool first_launch_on_device(const void* fn) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> seen;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(mu);
  return seen.insert({fn, dev}).second;
}

int device_num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

static thread_local int g_sm_budget = 0;
int sm_budget() {
  const int n = device_num_sms();
  return (g_sm_budget > 0 && g_sm_budget < n) ? g_sm_budget : n;
}
ScopedSmBudget::ScopedSmBudget(int n) : prev(g_sm_budget) { if (n > 0) g_sm_budget = n; }
ScopedSmBudget::~ScopedSmBudget() { g_sm_budget = prev; }
static thread_local int g_sm_waves = 1;
int sm_waves() { return g_sm_waves; }
ScopedSmWaves::ScopedSmWaves(int k) : prev(g_sm_waves) { if (k > 0) g_sm_waves = k; }
ScopedSmWaves::~ScopedSmWaves() { g_sm_waves = prev; }
int align_sm_waves() {
  static const int v = [] {
    const char* e = getenv("SIG_ALIGN_WAVES");
    return e ? atoi(e) : 1;
  }();
  return v;
}

int align_sm_budget() {
  static const int v = [] {
    // default: AlignM's persistent kernels leave ~1/5 of the SMs to the short kernels of SIM's chain, which runs next to
    // them and is the critical path of the fused step (measured at B = 128, d = 768: 148 -> 0.605, 132 -> 0.595,
    // 116 -> 0.589, 100 -> 0.607, 84 -> 0.642 ms/step; re-measured at the end of round 2, same box back to back:
    // 132 -> 0.595, 124 -> 0.589, 116 -> 0.585 / 0.588 / 0.591, 112 -> 0.583 / 0.589, 108 -> 0.574 / 0.581 / 0.582,
    // 104 -> 0.598 / 0.600, 100 -> 0.593); the dX GEMM at the tail of the step always takes all SMs
    const char* e = getenv("SIG_ALIGN_SMS");
    return e ? atoi(e) : (device_num_sms() * 108) / 148;
  }();
  return v;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("SIG_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

void prof_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

ProfScope::ProfScope(const char* name, cudaStream_t stream) : idx(-1), s(stream), capturing(false) {
  if (!g_enabled.load(std::memory_order_relaxed)) return;
  // SIG_PROF_CAPTURE=1: also record inside stream capture (the events become event-record nodes of the graph,
  // so a replay can be laid out on a time line with sig_profile_timeline)
  static const bool in_capture = [] {
    const char* e = getenv("SIG_PROF_CAPTURE");
    return e && e[0] == '1';
  }();
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) return;
  if (st != cudaStreamCaptureStatusNone && !in_capture) return;
  Rec r;
  r.name = name;
  if (cudaEventCreate(&r.a) != cudaSuccess) return;
  if (cudaEventCreate(&r.b) != cudaSuccess) { cudaEventDestroy(r.a); return; }
  capturing = st != cudaStreamCaptureStatusNone;
  if (capturing) cudaEventRecordWithFlags(r.a, stream, cudaEventRecordExternal);   // a node the host can time after a replay
  else cudaEventRecord(r.a, stream);
  std::lock_guard<std::mutex> lk(g_mu);
  g_recs.push_back(r);
  idx = (int)g_recs.size() - 1;
}

ProfScope::~ProfScope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (idx < (int)g_recs.size()) {
    if (capturing) cudaEventRecordWithFlags(g_recs[idx].b, s, cudaEventRecordExternal);
    else cudaEventRecord(g_recs[idx].b, s);
  }
}
}  // namespace sig

extern "C" {

unsigned long long sig_debug_launch_count(void) { return sig::g_launches.load(); }

int sig_profile_enable(int on) {
  sig::g_enabled.store(on ? 1 : 0);
  return 0;
}

// A caller-side scope on the same time line (e.g. around a collective the caller issues between two library
// calls): begin returns a handle for sig_profile_scope_end, or NULL when the profiler is off.
void* sig_profile_scope_begin(const char* name, void* stream) {
  if (!sig::g_enabled.load(std::memory_order_relaxed) || !name) return nullptr;
  return new sig::ProfScope(name, static_cast<cudaStream_t>(stream));
}

void sig_profile_scope_end(void* scope) { delete static_cast<sig::ProfScope*>(scope); }

// Time line of the recorded scopes: "name start_us end_us\n" per scope, relative to the earliest start.
// Does not clear the record (a captured graph keeps using the events); returns the number of scopes written.
int sig_profile_timeline(char* buf, size_t bytes) {
  std::lock_guard<std::mutex> lk(sig::g_mu);
  if (sig::g_recs.empty() || !buf || !bytes) return 0;
  std::string out;
  int n = 0;
  cudaEvent_t ref = nullptr;
  float best = 0.f;
  for (auto& r : sig::g_recs) {   // earliest start = the one no other start precedes
    if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
    if (!ref) { ref = r.a; continue; }
    float t = 0.f;
    if (cudaEventElapsedTime(&t, ref, r.a) == cudaSuccess && t < best) { ref = r.a; best = 0.f; }
  }
  if (!ref) return 0;
  for (auto& r : sig::g_recs) {
    float t0 = 0.f, t1 = 0.f;
    if (cudaEventElapsedTime(&t0, ref, r.a) != cudaSuccess || cudaEventElapsedTime(&t1, ref, r.b) != cudaSuccess) continue;
    char line[160];
    snprintf(line, sizeof line, "%s %.1f %.1f\n", r.name.c_str(), t0 * 1e3f, t1 * 1e3f);
    if (out.size() + strlen(line) + 1 > bytes) break;
    out += line;
    ++n;
  }
  std::strncpy(buf, out.c_str(), bytes - 1);
  buf[bytes - 1] = 0;
  return n;
}

// Synchronises the recorded events, writes up to `max` aggregated phases ('\n'-separated names
// into names_buf, summed milliseconds into ms, scope counts into counts) and clears the record.
int sig_profile_collect(char* names_buf, size_t names_bytes, float* ms, int* counts, int max) {
  std::lock_guard<std::mutex> lk(sig::g_mu);
  std::vector<std::string> names;
  std::vector<float> sums;
  std::vector<int> cnt;
  for (auto& r : sig::g_recs) {
    float t = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess) cudaEventElapsedTime(&t, r.a, r.b);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    size_t i = 0;
    for (; i < names.size(); ++i)
      if (names[i] == r.name) break;
    if (i == names.size()) { names.push_back(r.name); sums.push_back(0.f); cnt.push_back(0); }
    sums[i] += t;
    cnt[i] += 1;
  }
  sig::g_recs.clear();
  std::string joined;
  int n = 0;
  for (size_t i = 0; i < names.size() && n < max; ++i, ++n) {
    if (joined.size() + names[i].size() + 2 > names_bytes) break;
    joined += names[i];
    joined += '\n';
    ms[n] = sums[i];
    counts[n] = cnt[i];
  }
  if (names_buf && names_bytes) {
    std::strncpy(names_buf, joined.c_str(), names_bytes - 1);
    names_buf[names_bytes - 1] = 0;
  }
  return n;
}
}

// Launch counter and per-stage CUDA-event timing (used by bench.py for the roofline numbers).
// Events are recorded on the launching stream around each stage; nothing is synchronised until
// drin_profile_collect(), so enabling the profiler does not serialise the step.
#include <atomic>
#include <mutex>
#include <vector>

#include "kernels.cuh"

namespace drin {

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

namespace prof {

struct Rec {
  int cat;
  cudaEvent_t e0, e1;
  double flops, bytes;
};
static bool g_on = false;
static std::vector<Rec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static std::mutex g_mu;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void enable(bool on) {
  std::lock_guard<std::mutex> l(g_mu);
  g_on = on;
}
bool enabled() { return g_on; }

Scope::Scope(cudaStream_t s, int cat, double flops, double bytes) : stream_(s), idx_(-1) {
  if (!g_on) return;
  std::lock_guard<std::mutex> l(g_mu);
  Rec r{cat, get_event(), get_event(), flops, bytes};
  cudaEventRecord(r.e0, s);
  idx_ = (int)g_recs.size();
  g_recs.push_back(r);
}
Scope::~Scope() {
  if (idx_ < 0) return;
  std::lock_guard<std::mutex> l(g_mu);
  cudaEventRecord(g_recs[idx_].e1, stream_);
}

// Sums per category since the last collect; returns the number of records.
int collect(double* ms, double* flops, double* bytes, long long* count) {
  std::lock_guard<std::mutex> l(g_mu);
  for (int c = 0; c < NCAT; ++c) { ms[c] = 0; flops[c] = 0; bytes[c] = 0; count[c] = 0; }
  for (Rec& r : g_recs) {
    cudaEventSynchronize(r.e1);
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
      ms[r.cat] += t;
      flops[r.cat] += r.flops;
      bytes[r.cat] += r.bytes;
      count[r.cat] += 1;
    }
    g_pool.push_back(r.e0);
    g_pool.push_back(r.e1);
  }
  const int n = (int)g_recs.size();
  g_recs.clear();
  return n;
}

}  // namespace prof
}  // namespace drin

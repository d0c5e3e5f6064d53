#include <atomic>
#include <mutex>
#include <vector>

#include "ar_common.cuh"
#include "prof.cuh"

namespace ar {

static std::atomic<long long> g_launches{0};
static bool g_on = false;
static std::mutex g_mu;
struct Rec { int cat; cudaEvent_t a, b; double flops; int launches; };
static std::vector<Rec> g_recs;      // live records of the current window
static std::vector<Rec> g_pool;      // recycled event pairs

void prof_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long prof_launch_count() { return g_launches.load(); }

ProfScope::ProfScope(int cat, cudaStream_t s, double flops, int launches) : slot(-1), stream(s) {
  prof_count_launch(launches);
  if (!g_on) return;
  std::lock_guard<std::mutex> lk(g_mu);
  Rec r;
  if (!g_pool.empty()) { r = g_pool.back(); g_pool.pop_back(); }
  else { cudaEventCreate(&r.a); cudaEventCreate(&r.b); }
  r.cat = cat; r.flops = flops; r.launches = launches;
  cudaEventRecord(r.a, s);
  g_recs.push_back(r);
  slot = (int)g_recs.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEventRecord(g_recs[slot].b, stream);
}

int prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_on = on != 0;
  for (auto& r : g_recs) g_pool.push_back(r);
  g_recs.clear();
  return AR_OK;
}

// Sums the window's records per category (synchronises on their stop events) and clears it.
int prof_read(double* ms, double* flops, long long* launches, int n) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < n; ++i) { ms[i] = 0; flops[i] = 0; launches[i] = 0; }
  for (auto& r : g_recs) {
    AR_CUDA_OK(cudaEventSynchronize(r.b));
    float t = 0.f;
    AR_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
    if (r.cat < n) { ms[r.cat] += t; flops[r.cat] += r.flops; launches[r.cat] += r.launches; }
    g_pool.push_back(r);
  }
  g_recs.clear();
  return AR_OK;
}

}  // namespace ar

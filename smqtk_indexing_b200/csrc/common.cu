// Library-wide state: thread-local error string, launch counter, version.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace sb {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = NUM_SMS_B200;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return NUM_SMS_B200;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
      cached_sms = n;
      cached_dev = dev;
    }
  }
  return cached_sms;
}

// ---- per-launch event timing ------------------------------------------------
namespace {
struct ProfRec {
  const char* name;
  cudaEvent_t e0, e1;
};
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_recs;        // records of the current window
std::vector<cudaEvent_t> g_prof_pool;    // recycled events
constexpr size_t PROF_MAX_RECS = 1 << 16;

cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

ProfScope::ProfScope(const char* name, cudaStream_t st) : slot_(-1), st_(st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof_recs.size() >= PROF_MAX_RECS) return;
  ProfRec r{name, prof_event(), prof_event()};
  if (!r.e0 || !r.e1) return;
  cudaEventRecord(r.e0, st);
  slot_ = (int)g_prof_recs.size();
  g_prof_recs.push_back(r);
}

ProfScope::~ProfScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if ((size_t)slot_ < g_prof_recs.size()) cudaEventRecord(g_prof_recs[slot_].e1, st_);
}

}  // namespace sb

extern "C" {
int sb_profile_enable(int on) {
  sb::g_prof_on.store(on ? 1 : 0);
  return SB_OK;
}

// Waits for the recorded launches, writes up to `cap` (name, milliseconds) pairs
// in launch order, clears the window and returns the number of records it held.
int64_t sb_profile_fetch(const char** names, float* ms, int64_t cap) {
  std::lock_guard<std::mutex> lk(sb::g_prof_mu);
  const int64_t n = (int64_t)sb::g_prof_recs.size();
  for (int64_t i = 0; i < n; ++i) {
    sb::ProfRec& r = sb::g_prof_recs[i];
    float t = -1.0f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess) cudaEventElapsedTime(&t, r.e0, r.e1);
    if (i < cap) {
      if (names) names[i] = r.name;
      if (ms) ms[i] = t;
    }
    sb::g_prof_pool.push_back(r.e0);
    sb::g_prof_pool.push_back(r.e1);
  }
  sb::g_prof_recs.clear();
  return n;
}

int sb_version(void) { return 100; }
const char* sb_last_error(void) { return sb::t_err; }
uint64_t sb_launch_count(void) { return sb::g_launches.load(); }
}

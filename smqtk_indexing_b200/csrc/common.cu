// Library-wide state: thread-local error string, launch counter, version.
#include <stdarg.h>

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

}  // namespace sb

extern "C" {
int sb_version(void) { return 100; }
const char* sb_last_error(void) { return sb::t_err; }
uint64_t sb_launch_count(void) { return sb::g_launches.load(); }
}

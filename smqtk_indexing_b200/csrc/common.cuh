// Shared helpers for the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "smqtk_b200.h"

namespace sb {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SB_ERR_CUDA;
  }
  return SB_OK;
}

#define SB_CUDA_TRY(expr)                                                        \
  do {                                                                           \
    cudaError_t e__ = (expr);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      sb::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));            \
      return SB_ERR_CUDA;                                                        \
    }                                                                            \
  } while (0)

#define SB_REQUIRE(cond, ...)                                                    \
  do {                                                                           \
    if (!(cond)) {                                                               \
      sb::set_error(__VA_ARGS__);                                                \
      return SB_ERR_INVALID_ARGUMENT;                                            \
    }                                                                            \
  } while (0)

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int NUM_SMS_B200 = 148;

// Cached device properties (SM count) for the current device.
int sm_count();

// Optional per-launch timing (sb_profile_enable): when on, a ProfScope records a
// CUDA event on `st` before and after the launch it brackets; sb_profile_fetch
// turns the pairs into milliseconds.  Off (default) it costs one relaxed load.
struct ProfScope {
  ProfScope(const char* name, cudaStream_t st);
  ~ProfScope();
  int slot_;
  cudaStream_t st_;
};

}  // namespace sb

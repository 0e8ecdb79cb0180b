// uavb_common.cuh -- error plumbing shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/uavb.h"

namespace uavb {

// thread-local message returned by uavb_last_error()
char* error_buffer();
int set_error(int code, const char* fmt, ...);

// Fails loudly when there is no usable device: the library has no CPU path.
int require_device();

#define UAVB_CUDA_OK(expr)                                                                          \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::uavb::set_error(UAVB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                                 \
  } while (0)

#define UAVB_REQUIRE(cond, msg)                                         \
  do {                                                                  \
    if (!(cond)) return ::uavb::set_error(UAVB_EINVAL, "%s", msg);      \
  } while (0)

// Library-private stream-ordered memory pool of the current device (scratch for the *_host entry points and the
// time-sliced rollout).  Its release threshold is unlimited, so scratch freed after one launch is reused by the next
// instead of being returned to the driver at every synchronisation (the default pool does that, which costs
// milliseconds per 100 MB).  Returns nullptr if pools are unavailable; callers then fall back to the default pool.
cudaMemPool_t scratch_pool();

// SM count of the current device, cached per device (thread-safe).
int sm_count_cached(int* sms);

constexpr int kMaxDevices = 64;

inline int div_up(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace uavb

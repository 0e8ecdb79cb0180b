// uavb_common.cuh -- error plumbing shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <functional>
#include <vector>

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

// Stream-ordered device buffer that frees itself; allocation failures are recorded, not thrown.
struct DevPool {
  cudaStream_t st;
  std::vector<void*> owned;
  cudaError_t err = cudaSuccess;
  explicit DevPool(cudaStream_t s) : st(s) {}
  ~DevPool() {
    for (void* p : owned) cudaFreeAsync(p, st);
  }
  template <class T> T* alloc(size_t n) {
    void* p = nullptr;
    if (err == cudaSuccess) {
      cudaMemPool_t pool = scratch_pool();
      err = pool ? cudaMallocFromPoolAsync(&p, (n ? n : 1) * sizeof(T), pool, st) : cudaMallocAsync(&p, (n ? n : 1) * sizeof(T), st);
    }
    if (err == cudaSuccess) owned.push_back(p);
    return static_cast<T*>(p);
  }
  template <class T> T* upload(const T* host, size_t n) {
    if (!host) return nullptr;
    T* d = alloc<T>(n);
    if (err == cudaSuccess) err = cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, st);
    return d;
  }
};

// Thread-local pinned (portable) host scratch of at least `bytes` bytes, grown on demand and kept for the life of the thread;
// two independent slots (0: the planner's control block, 1: the host-buffer mission call's small inputs).  nullptr on failure.
void* pinned_scratch(int slot, size_t bytes);

// The obstacle-correction loop behind uavb_minsnap_correct_f64 (minsnap_correct.cu).  after_first_round (optional) enqueues
// follow-up work behind round 1 (complete after that round's synchronisation when nothing was hit); n_wp_known (optional): the
// initial n_waypoints on the host; ctrl (optional): a zeroed device block whose first 8 ints are the loop's counters, copied whole
// to ctrl->host (pinned) at every round's read-back.
struct CtrlBlock {
  int* dev;
  int* host;
  int ints;
  bool no_wait;       // enqueue round 1 and the copy of the block, do not synchronise (speculative plan)
};
int correct_missions(double* waypoints, int* n_waypoints, const double* velocity, int B, int max_wp, double factor, double dt, const double* cuboids,
                     int n_obs, long long cuboid_stride, double* coeffs_out, double* times_out, int* status_out, int* rounds_out, cudaStream_t st,
                     const std::function<int()>* after_first_round, const int* n_wp_known, const CtrlBlock* ctrl);

// One mission of T consecutive tables planned with the correction loop into shared-mission segment arrays (uavb_plan_shared_f64).
constexpr int kMaxSharedTables = 8;
int plan_shared_tables(int T, const double* const* d_waypoints, const int* n_wp_in, const double* d_velocity, double factor, double dt,
                       const double* d_cuboids, int n_obs, int cap_seg, double* seg_coeffs, double* seg_times, int* seg_rows, int* seg_table,
                       double* seg_yaw0, int* n_seg_out, int* rows_out, int* status_out, int* rounds_out, cudaStream_t st, int* async_report = nullptr);

// SM count of the current device, cached per device (thread-safe).
int sm_count_cached(int* sms);

constexpr int kMaxDevices = 64;

inline int div_up(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace uavb

// host_api.cu -- host-buffer (end-to-end) entry point: one mission, B drones, numpy-style arrays in,
// metrics out.  Only plumbing lives here: device allocation from the stream-ordered pool, copies,
// and calls into the device-pointer entry points (K1, table geometry, K2).
#include <vector>

#include "uavb_common.cuh"

namespace uavb {

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

// Streams and events of the host-buffer calls: one set per host thread and device, created on first use and kept for the
// life of the thread (creating and destroying a stream costs ~55 + ~65 us on the B200 box, 2 % of a 5.7 ms mission call).
// Never destroyed explicitly: at thread / process exit the context owns them.
//   st        everything except the bulk uploads
//   st_up     the per-rollout inputs (the bulk of the host->device bytes): they travel while the mission is planned on `st`
//   ev_alloc  `st` -> `st_up`: the upload buffers exist;  ev_up  `st_up` -> `st`: the uploads are complete
struct HostCallSet {
  cudaStream_t st = nullptr, st_up = nullptr, st_aux = nullptr;     // st_aux: third lane of the chunked solve pipeline
  cudaEvent_t ev_alloc = nullptr, ev_up = nullptr;
};

static cudaError_t host_call_set(HostCallSet** out) {
  thread_local HostCallSet sets[kMaxDevices];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  HostCallSet& h = sets[dev];
  if (h.st == nullptr) e = cudaStreamCreateWithFlags(&h.st, cudaStreamNonBlocking);
  if (e == cudaSuccess && h.st_up == nullptr) e = cudaStreamCreateWithFlags(&h.st_up, cudaStreamNonBlocking);
  if (e == cudaSuccess && h.st_aux == nullptr) e = cudaStreamCreateWithFlags(&h.st_aux, cudaStreamNonBlocking);
  if (e == cudaSuccess && h.ev_alloc == nullptr) e = cudaEventCreateWithFlags(&h.ev_alloc, cudaEventDisableTiming);
  if (e == cudaSuccess && h.ev_up == nullptr) e = cudaEventCreateWithFlags(&h.ev_up, cudaEventDisableTiming);
  if (e != cudaSuccess) return e;
  *out = &h;
  return cudaSuccess;
}

}  // namespace uavb

using namespace uavb;

// B missions, host buffers in and out (BASELINE configs[1] end to end).  The batch is cut into chunks of 2^15 missions that run
// H2D -> K1 -> D2H on three cached streams in rotation, so the upload of chunk k+1 and the download of chunk k-1 overlap the solve
// of chunk k; device scratch for the three lanes comes from the library's stream-ordered pool (no cudaMalloc / cudaFree per
// call).  928 B per solve cross the host link, 804 of them device-to-host: with pinned host buffers the call is bound by that
// link; pageable buffers work but are staged by the driver.
extern "C" int uavb_minsnap_solve_f64_host(const double* waypoints, const double* velocity, int B, int S, double factor,
                                           double* coeffs_out, double* times_out, int* status_out) {
  UAVB_REQUIRE(waypoints && velocity && coeffs_out && times_out, "minsnap_solve_host: NULL pointer");
  UAVB_REQUIRE(B >= 0 && S >= 1 && S <= UAVB_MAX_SPLINES, "minsnap_solve_host: B >= 0 and 1 <= S <= UAVB_MAX_SPLINES required");
  int rc = require_device();
  if (rc) return rc;
  if (B == 0) return UAVB_OK;
  HostCallSet* hs = nullptr;
  UAVB_CUDA_OK(host_call_set(&hs));
  cudaStream_t lanes[3] = {hs->st, hs->st_up, hs->st_aux};
  constexpr int kChunk = 1 << 15;
  const int n_chunks = (B + kChunk - 1) / kChunk, n_lanes = n_chunks < 3 ? n_chunks : 3;
  const size_t cw = (size_t)(S + 1) * 3, cc = (size_t)24 * S, ct = (size_t)S;        // doubles per mission
  const size_t chunk = (size_t)(B < kChunk ? B : kChunk);
  int result = UAVB_OK;
  cudaError_t e = cudaSuccess;
  {
    DevPool pool(hs->st);
    double *dw[3], *dv[3], *dc[3], *dtm[3];
    int* ds[3];
    for (int l = 0; l < n_lanes; ++l) {
      dw[l] = pool.alloc<double>(chunk * cw); dv[l] = pool.alloc<double>(chunk); dc[l] = pool.alloc<double>(chunk * cc);
      dtm[l] = pool.alloc<double>(chunk * ct); ds[l] = pool.alloc<int>(chunk);
    }
    if (pool.err != cudaSuccess) {
      result = set_error(UAVB_ENOMEM, "minsnap_solve_host: %s", cudaGetErrorString(pool.err));
    } else {
      e = cudaEventRecord(hs->ev_alloc, hs->st);                      // the buffers exist once `st` reaches this point
      for (int l = 1; l < n_lanes && e == cudaSuccess; ++l) e = cudaStreamWaitEvent(lanes[l], hs->ev_alloc, 0);
      for (int k = 0; k < n_chunks && e == cudaSuccess && !result; ++k) {
        const int l = k % n_lanes;
        cudaStream_t st = lanes[l];
        const size_t first = (size_t)k * kChunk, n = (size_t)B - first < (size_t)kChunk ? (size_t)B - first : (size_t)kChunk;
        e = cudaMemcpyAsync(dw[l], waypoints + first * cw, n * cw * 8, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dv[l], velocity + first, n * 8, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) break;
        result = uavb_minsnap_solve_f64(dw[l], dv[l], (int)n, S, factor, dc[l], dtm[l], ds[l], st);
        if (result) break;
        e = cudaMemcpyAsync(coeffs_out + first * cc, dc[l], n * cc * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(times_out + first * ct, dtm[l], n * ct * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && status_out) e = cudaMemcpyAsync(status_out + first, ds[l], n * 4, cudaMemcpyDeviceToHost, st);
      }
      // the frees of `pool` are ordered on `st`: make it wait for the other lanes, then wait for everything
      for (int l = 1; l < n_lanes; ++l) {
        cudaError_t e2 = cudaEventRecord(hs->ev_up, lanes[l]);
        if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(hs->st, hs->ev_up, 0);
        if (e2 != cudaSuccess && e == cudaSuccess) e = e2;
      }
      const cudaError_t e3 = cudaStreamSynchronize(hs->st);
      if (e == cudaSuccess) e = e3;
      if (e != cudaSuccess && !result) {
        for (int l = 0; l < n_lanes; ++l) cudaStreamSynchronize(lanes[l]);
        cudaGetLastError();
        result = set_error(UAVB_ECUDA, "minsnap_solve_host: %s", cudaGetErrorString(e));
      }
    }
  }
  return result;
}

extern "C" int uavb_fly_mission_host(const uavb_mission_host* m, float* metrics_out, float* state_out, int* n_ticks_out) {
  UAVB_REQUIRE(m != nullptr && metrics_out != nullptr, "fly_mission_host: mission and metrics_out are required");
  UAVB_REQUIRE(m->B >= 0, "fly_mission_host: B must be >= 0");
  UAVB_REQUIRE(m->waypoints != nullptr && m->n_waypoints >= 2, "fly_mission_host: at least two waypoints are required");
  UAVB_REQUIRE(m->n_takeoff_waypoints == 0 || (m->n_takeoff_waypoints >= 2 && m->n_takeoff_waypoints < m->n_waypoints),
               "fly_mission_host: n_takeoff_waypoints must be 0 or in [2, n_waypoints)");
  UAVB_REQUIRE(m->frequency >= 1 && m->velocity > 0.0 && m->n_ticks >= 0, "fly_mission_host: frequency >= 1, velocity > 0, n_ticks >= 0 required");
  UAVB_REQUIRE(m->n_obs >= 0 && (m->n_obs == 0 || m->aabbs != nullptr), "fly_mission_host: n_obs > 0 needs aabbs");
  const int n_tab = m->n_takeoff_waypoints ? 2 : 1;
  const int S0 = m->n_takeoff_waypoints ? m->n_takeoff_waypoints - 1 : m->n_waypoints - 1;
  const int S1 = m->n_takeoff_waypoints ? m->n_waypoints - m->n_takeoff_waypoints : 0;
  UAVB_REQUIRE(S0 <= UAVB_MAX_SPLINES && S1 <= UAVB_MAX_SPLINES, "fly_mission_host: too many splines in one table");
  int rc = require_device();
  if (rc) return rc;
  const int n_seg = S0 + S1;
  const size_t B = (size_t)m->B;
  const double dt_outer = m->veh.dt * m->frequency;

  HostCallSet* hs = nullptr;
  UAVB_CUDA_OK(host_call_set(&hs));
  cudaStream_t st = hs->st, st_up = hs->st_up;
  int result = UAVB_OK;
  {
    DevPool pool(st);
    // per-rollout inputs: allocated in `st` order, copied on `st_up` while the plan below runs and its results travel back
    const bool have_mc = m->mc_mass || m->mc_inertia || m->mc_gains || m->mc_wind;
    bool uploads_in_flight = false;
    float* d_mc_mass = nullptr; float* d_mc_inertia = nullptr; float* d_mc_gains = nullptr; float* d_mc_wind = nullptr;
    if (have_mc && B > 0) {
      if (m->mc_mass) d_mc_mass = pool.alloc<float>(B);
      if (m->mc_inertia) d_mc_inertia = pool.alloc<float>(3 * B);
      if (m->mc_gains) d_mc_gains = pool.alloc<float>((size_t)UAVB_N_GAINS * B);
      if (m->mc_wind) d_mc_wind = pool.alloc<float>(3 * B);
      cudaError_t e = pool.err;
      if (e == cudaSuccess) e = cudaEventRecord(hs->ev_alloc, st);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(st_up, hs->ev_alloc, 0);
      if (e == cudaSuccess) uploads_in_flight = true;
      if (e == cudaSuccess && d_mc_mass) e = cudaMemcpyAsync(d_mc_mass, m->mc_mass, sizeof(float) * B, cudaMemcpyHostToDevice, st_up);
      if (e == cudaSuccess && d_mc_inertia) e = cudaMemcpyAsync(d_mc_inertia, m->mc_inertia, sizeof(float) * 3 * B, cudaMemcpyHostToDevice, st_up);
      if (e == cudaSuccess && d_mc_gains) e = cudaMemcpyAsync(d_mc_gains, m->mc_gains, sizeof(float) * UAVB_N_GAINS * B, cudaMemcpyHostToDevice, st_up);
      if (e == cudaSuccess && d_mc_wind) e = cudaMemcpyAsync(d_mc_wind, m->mc_wind, sizeof(float) * 3 * B, cudaMemcpyHostToDevice, st_up);
      if (e == cudaSuccess) e = cudaEventRecord(hs->ev_up, st_up);
      if (e != cudaSuccess) result = set_error(UAVB_ECUDA, "fly_mission_host: %s", cudaGetErrorString(e));
    }
    double* d_wp = pool.upload(m->waypoints, (size_t)m->n_waypoints * 3);
    const double vel2[2] = {m->velocity, m->velocity};
    double* d_vel = pool.upload(vel2, 2);
    double* d_coeffs = pool.alloc<double>((size_t)n_seg * 24);
    double* d_times = pool.alloc<double>(n_seg);
    int* d_status = pool.alloc<int>(2);
    int* d_rows = pool.alloc<int>(n_seg);
    double* d_yaw0 = pool.alloc<double>(2);
    int* d_total = pool.alloc<int>(2);
    const int offs[3] = {0, S0, n_seg};
    int* d_offs = pool.upload(offs, 3);
    if (pool.err != cudaSuccess) result = set_error(UAVB_ENOMEM, "fly_mission_host: %s", cudaGetErrorString(pool.err));
    // plan: one K1 launch per table (main.py:80-83), then the table geometry of both
    if (!result) result = uavb_minsnap_solve_f64(d_wp, d_vel, 1, S0, m->start_end_time_factor, d_coeffs, d_times, d_status, st);
    if (!result && n_tab == 2)
      result = uavb_minsnap_solve_f64(d_wp + 3 * (size_t)(m->n_takeoff_waypoints - 1), d_vel + 1, 1, S1, m->start_end_time_factor,
                                      d_coeffs + (size_t)S0 * 24, d_times + S0, d_status + 1, st);
    if (!result) result = uavb_minsnap_table_meta_f64(d_coeffs, d_times, d_offs, n_tab, dt_outer, d_rows, d_yaw0, d_total, st);
    std::vector<int> rows(n_seg), seg_table(n_seg, 0);
    std::vector<double> seg_yaw0(n_seg, 0.0);
    double yaw0[2] = {0.0, 0.0};
    int status[2] = {0, 0};
    if (!result) {
      cudaError_t e = cudaMemcpyAsync(rows.data(), d_rows, sizeof(int) * n_seg, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(yaw0, d_yaw0, sizeof(double) * n_tab, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(status, d_status, sizeof(int) * n_tab, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) result = set_error(UAVB_ECUDA, "fly_mission_host: %s", cudaGetErrorString(e));
      else if (status[0] || status[1])
        result = set_error(UAVB_EINVAL, "fly_mission_host: degenerate mission (zero-length spline): the reference's KKT matrix is singular");
    }
    if (!result) {
      long long total_rows = 0;
      for (int s = 0; s < n_seg; ++s) total_rows += rows[s];
      seg_table[0] = 1; seg_yaw0[0] = yaw0[0];
      if (n_tab == 2) { seg_table[S0] = 1; seg_yaw0[S0] = yaw0[1]; }
      const long long whole = total_rows * m->frequency;
      const int n_ticks = m->n_ticks ? m->n_ticks : (int)(whole < 2147483647LL ? whole : 2147483647LL);
      if (n_ticks_out) *n_ticks_out = n_ticks;

      uavb_rollout_args a = {};
      a.B = m->B; a.n_ticks = n_ticks; a.inner_per_outer = m->frequency; a.thrust_frame_lag = m->thrust_frame_lag;
      a.n_obs = m->n_obs; a.n_obs_sets = m->n_obs > 0 ? 1 : 0;
      a.veh = m->veh;
      a.mc_mass = d_mc_mass; a.mc_inertia = d_mc_inertia; a.mc_gains = d_mc_gains; a.mc_wind = d_mc_wind;
      a.seg_coeffs = d_coeffs; a.seg_rows = d_rows;
      a.seg_table = pool.upload(seg_table.data(), n_seg);
      a.seg_yaw0 = pool.upload(seg_yaw0.data(), n_seg);
      a.n_seg_shared = n_seg;
      a.dt_outer = dt_outer;
      void* d_targets = total_rows > 0 ? static_cast<void*>(pool.alloc<char>((size_t)total_rows * UAVB_TARGET_ROW_BYTES)) : nullptr;
      if (d_targets && pool.err == cudaSuccess) {
        result = uavb_rollout_targets_f64(d_coeffs, d_rows, a.seg_table, a.seg_yaw0, n_seg, dt_outer, d_targets, (int)total_rows, st);
        a.shared_targets = d_targets;
        a.n_target_rows = (int)total_rows;
      }
      a.start = m->start ? pool.upload(m->start, 3) : d_wp;
      a.goal = m->goal ? pool.upload(m->goal, 3) : d_wp + 3 * (size_t)(m->n_waypoints - 1);
      a.aabbs = m->n_obs > 0 ? pool.upload(m->aabbs, (size_t)m->n_obs * 6) : nullptr;
      float* d_metrics = pool.alloc<float>(B * UAVB_N_METRICS);
      float* d_state = state_out ? pool.alloc<float>(B * UAVB_STATE_DIM) : nullptr;
      a.metrics_out = d_metrics; a.state_out = d_state;
      if (pool.err != cudaSuccess) result = set_error(UAVB_ENOMEM, "fly_mission_host: %s", cudaGetErrorString(pool.err));
      if (!result && uploads_in_flight) {                    // the rollout reads what `st_up` wrote
        const cudaError_t e = cudaStreamWaitEvent(st, hs->ev_up, 0);
        if (e != cudaSuccess) result = set_error(UAVB_ECUDA, "fly_mission_host: %s", cudaGetErrorString(e));
      }
      if (!result) result = uavb_rollout_f32(&a, st);
      if (!result && B > 0) {
        cudaError_t e = cudaMemcpyAsync(metrics_out, d_metrics, sizeof(float) * B * UAVB_N_METRICS, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && state_out) e = cudaMemcpyAsync(state_out, d_state, sizeof(float) * B * UAVB_STATE_DIM, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) result = set_error(UAVB_ECUDA, "fly_mission_host: %s", cudaGetErrorString(e));
      }
    }
    // every path, also the failing ones: the pool's destructor (next brace) frees in `st` order, so the uploads must be over
    if (uploads_in_flight) cudaStreamSynchronize(st_up);
  }
  cudaStreamSynchronize(st);                               // also on the failing paths: nothing of this call is in flight afterwards
  return result;
}

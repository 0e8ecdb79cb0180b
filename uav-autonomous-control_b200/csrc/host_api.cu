// host_api.cu -- host-buffer (end-to-end) entry point: one mission, B drones, numpy-style arrays in,
// metrics out.  Only plumbing lives here: device allocation from the stream-ordered pool, copies,
// and calls into the device-pointer entry points (K1, table geometry, K2).
#include <vector>

#include "uavb_common.cuh"

namespace uavb {

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
    // plan: both tables (main.py:80-83) go through the reference's obstacle-correction loop (midpoints are inserted where a
    // sampled point lies inside a box, minimum_snap.py:63-95; nothing is inserted on lab_course) and are packed into the
    // segment arrays of a shared-mission rollout -- one synchronisation when nothing is hit (plan_shared_tables)
    const bool correct = m->n_obs > 0 && !m->no_correction;
    // the small inputs travel as ONE pinned block: waypoints, planner boxes (fp64), velocities, start, goal | collision boxes (fp32)
    const size_t n_wpd = (size_t)m->n_waypoints * 3, n_boxd = correct ? (size_t)m->n_obs * 6 : 0, n_aabb = (size_t)m->n_obs * 6;
    const size_t n_dbl = n_wpd + n_boxd + 2 + 3 + 3, blob_bytes = sizeof(double) * n_dbl + sizeof(float) * n_aabb;
    double* h_blob = static_cast<double*>(pinned_scratch(1, blob_bytes));
    char* d_blob = pool.alloc<char>(blob_bytes);
    if (h_blob == nullptr || pool.err != cudaSuccess) result = set_error(UAVB_ENOMEM, "fly_mission_host: scratch allocation failed");
    double* d_wp = reinterpret_cast<double*>(d_blob);
    double* d_boxes = correct ? d_wp + n_wpd : nullptr;
    double* d_vel = d_wp + n_wpd + n_boxd;
    double* d_start = d_vel + 2;
    double* d_goal = d_start + 3;
    float* d_aabbs = m->n_obs > 0 ? reinterpret_cast<float*>(d_goal + 3) : nullptr;
    if (!result) {
      for (size_t k = 0; k < n_wpd; ++k) h_blob[k] = m->waypoints[k];
      for (size_t k = 0; k < n_boxd; ++k) h_blob[n_wpd + k] = m->plan_aabbs ? m->plan_aabbs[k] : (double)m->aabbs[k];
      double* hv = h_blob + n_wpd + n_boxd;
      hv[0] = hv[1] = m->velocity;
      const double* s3 = m->start ? m->start : m->waypoints;
      const double* g3 = m->goal ? m->goal : m->waypoints + 3 * (size_t)(m->n_waypoints - 1);
      for (int k = 0; k < 3; ++k) { hv[2 + k] = s3[k]; hv[5 + k] = g3[k]; }
      float* ha = reinterpret_cast<float*>(hv + 8);
      for (size_t k = 0; k < n_aabb; ++k) ha[k] = m->aabbs[k];
      const cudaError_t e = cudaMemcpyAsync(d_blob, h_blob, blob_bytes, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) result = set_error(UAVB_ECUDA, "fly_mission_host: %s", cudaGetErrorString(e));
    }
    constexpr int kCapSeg = 2 * UAVB_MAX_SPLINES;
    double* d_coeffs = pool.alloc<double>((size_t)kCapSeg * 24);
    double* d_times = pool.alloc<double>(kCapSeg);
    int* d_rows = pool.alloc<int>(kCapSeg);
    int* d_seg_table = pool.alloc<int>(kCapSeg);
    double* d_seg_yaw0 = pool.alloc<double>(kCapSeg);
    if (pool.err != cudaSuccess) result = set_error(UAVB_ENOMEM, "fly_mission_host: %s", cudaGetErrorString(pool.err));
    const double* tab_wp[2] = {d_wp, d_wp + 3 * (size_t)(m->n_takeoff_waypoints ? m->n_takeoff_waypoints - 1 : 0)};
    const int tab_n[2] = {S0 + 1, S1 + 1};
    int status[2] = {0, 0}, tab_rows[2] = {0, 0}, n_seg = 0;
    if (!result)
      result = plan_shared_tables(n_tab, tab_wp, tab_n, d_vel, m->start_end_time_factor, dt_outer, d_boxes, correct ? m->n_obs : 0, kCapSeg, d_coeffs,
                                  d_times, d_rows, d_seg_table, d_seg_yaw0, &n_seg, tab_rows, status, nullptr, st);
    if (!result && (status[0] == UAVB_SOLVE_TOO_MANY || status[1] == UAVB_SOLVE_TOO_MANY))
      result = set_error(UAVB_EINVAL, "fly_mission_host: obstacle correction needs more than %d splines (an obstacle probably contains a waypoint; "
                         "the reference loops forever in this case)", UAVB_MAX_SPLINES);
    else if (!result && (status[0] || status[1]))
      result = set_error(UAVB_EINVAL, "fly_mission_host: degenerate mission (zero-length spline): the reference's KKT matrix is singular");
    if (!result) {
      const long long total_rows = (long long)tab_rows[0] + tab_rows[1];
      const long long whole = total_rows * m->frequency;
      const int n_ticks = m->n_ticks ? m->n_ticks : (int)(whole < 2147483647LL ? whole : 2147483647LL);
      if (n_ticks_out) *n_ticks_out = n_ticks;

      uavb_rollout_args a = {};
      a.B = m->B; a.n_ticks = n_ticks; a.inner_per_outer = m->frequency; a.thrust_frame_lag = m->thrust_frame_lag;
      a.n_obs = m->n_obs; a.n_obs_sets = m->n_obs > 0 ? 1 : 0;
      a.veh = m->veh;
      a.mc_mass = d_mc_mass; a.mc_inertia = d_mc_inertia; a.mc_gains = d_mc_gains; a.mc_wind = d_mc_wind;
      a.seg_coeffs = d_coeffs; a.seg_rows = d_rows;
      a.seg_table = d_seg_table;
      a.seg_yaw0 = d_seg_yaw0;
      a.n_seg_shared = n_seg;
      a.dt_outer = dt_outer;
      void* d_targets = total_rows > 0 ? static_cast<void*>(pool.alloc<char>((size_t)total_rows * UAVB_TARGET_ROW_BYTES)) : nullptr;
      if (d_targets && pool.err == cudaSuccess) {
        result = uavb_rollout_targets_f64(d_coeffs, d_rows, a.seg_table, a.seg_yaw0, n_seg, dt_outer, d_targets, (int)total_rows, st);
        a.shared_targets = d_targets;
        a.n_target_rows = (int)total_rows;
      }
      a.start = d_start;
      a.goal = d_goal;
      a.aabbs = d_aabbs;
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

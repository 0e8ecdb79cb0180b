// minsnap_correct.cu -- the obstacle-correction loop of MinimumSnap._generate_collision_free_trajectory
// (uav_ac/planning/minimum_snap.py:63-95) for B missions, entirely on the device.
//
// Reference, per mission:   for coord in obstacles:  plan;  loop { ids = splines that own a sampled point inside coord;
//                           if none: break;  insert the midpoint of every such spline as a new waypoint;  plan again }
// Missions are independent and a mission's waypoints change only when it is hit, so "walk the obstacles in order" is the same
// as "find the first obstacle at or after my cursor that my current plan hits": every mission keeps its own obstacle cursor and
// the batch needs no per-obstacle pass.  One round of the pipeline is
//     solve   the missions of the work lists (K1 arithmetic, one thread per mission; the lists are bucketed by spline count so
//             that missions of up to 4 / 8 splines run the fully unrolled register-resident solver and only longer ones the
//             rolled one with local arrays)
//     sweep   one CTA of four warps per listed mission (one launch for all buckets), the warps take the splines in turn: the lanes
//             evaluate the sampled positions of a spline (the rows j * dt,
//             j < ceil(T_i / dt), by the same Horner recurrence as the sampled table -- the same bits), test them against the
//             boxes from the mission's cursor on, reduce to (first obstacle hit, mask of its splines), insert the midpoints in
//             place and append the mission to the next round's list of its new bucket
// and the host reads four counters per round; only missions that were hit are planned again.  A mission that would grow past
// its waypoint capacity is reported (UAVB_SOLVE_TOO_MANY) instead of looping forever like the reference does when a box
// contains a waypoint.
#include <functional>

#include "minsnap_core.cuh"
#include "uavb_common.cuh"

namespace uavb {

constexpr int kBuckets = 4;
constexpr int kCorrectThreads = 64;
constexpr int kSweepWarps = 4;

__host__ __device__ __forceinline__ int bucket_of(int splines) { return splines <= 4 ? 0 : (splines <= 8 ? 1 : (splines <= 16 ? 2 : 3)); }

// Work lists of one round: list[k] holds count[k] mission indices of bucket k ([kBuckets][B] / [kBuckets]).
struct WorkLists {
  int* list;
  int* count;
};

__global__ void __launch_bounds__(256) correct_classify_kernel(const int* __restrict__ n_wp, int B, int max_wp, int* __restrict__ obs_idx,
                                                               int* __restrict__ status, WorkLists out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  obs_idx[b] = 0;
  const int S = n_wp[b] - 1;
  if (S < 1 || S > max_wp - 1 || S > UAVB_MAX_SPLINES) {
    status[b] = S < 1 ? UAVB_SOLVE_DEGENERATE : UAVB_SOLVE_TOO_MANY;
    return;
  }
  status[b] = UAVB_SOLVE_OK;
  const int k = bucket_of(S);
  out.list[(size_t)k * B + atomicAdd(out.count + k, 1)] = b;
}

// K1 over a work list, fixed-pitch layout: mission b owns waypoints[b][0 .. n_wp[b]) of [B][max_wp][3] and writes
// coeffs[b][s][8][3], times[b][s] of [B][max_wp-1][...].
template <int MAXS>
__global__ void __launch_bounds__(kCorrectThreads) minsnap_solve_list_kernel(const double* __restrict__ waypoints, const int* __restrict__ n_wp,
                                                                              const double* __restrict__ velocity, const int* __restrict__ list,
                                                                              const int* __restrict__ n_list, int max_wp, double factor,
                                                                              double* __restrict__ coeffs, double* __restrict__ times,
                                                                              int* __restrict__ status) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= *n_list) return;
  const int b = list[e];
  const int S = n_wp[b] - 1;
  if (S < 1 || S > MAXS) return;                       // cannot happen: the lists are bucketed by S
  const double* w = waypoints + (size_t)b * max_wp * 3;
  double* cout = coeffs + (size_t)b * (max_wp - 1) * 24;
  double* tout = times + (size_t)b * (max_wp - 1);
  const int st = minsnap_solve_one<MAXS>(
      S, velocity[b], factor, [w](int i, int ax) { return __ldg(w + 3 * i + ax); },
      [cout](int seg, int j, int ax, double val) { cout[seg * 24 + j * 3 + ax] = val; }, [tout](int seg, double t) { tout[seg] = t; });
  if (st) status[b] = st;
}

__device__ __forceinline__ unsigned long long bits_below(int i) { return i >= 64 ? ~0ull : (1ull << i) - 1ull; }

// One CTA of kSweepWarps warps per listed mission, every bucket's list in ONE launch (block x belongs to the bucket whose cumulative
// list length first exceeds x); warp w takes the splines s = w, w + kSweepWarps, ...; see the header.  cuboids: [n_obs][6] doubles
// shared by all missions (cuboid_stride 0) or one set per mission (cuboid_stride = doubles between consecutive missions' sets).
__global__ void __launch_bounds__(32 * kSweepWarps) correct_sweep_kernel(const double* __restrict__ coeffs, const double* __restrict__ times,
                                                                        double* __restrict__ waypoints, int* __restrict__ n_wp,
                                                                        int* __restrict__ obs_idx, WorkLists in, int B, int max_wp, double dt,
                                                                        const double* __restrict__ cuboids, int n_obs, long long cuboid_stride,
                                                                        int* __restrict__ status, WorkLists next) {
  __shared__ int s_best[kSweepWarps];
  __shared__ unsigned long long s_mask[kSweepWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int e = blockIdx.x, bucket = 0;
  while (bucket < kBuckets && e >= in.count[bucket]) { e -= in.count[bucket]; ++bucket; }
  if (bucket == kBuckets) return;                           // (uniform per CTA)
  const int b = in.list[(size_t)bucket * B + e];
  const int nw = n_wp[b], S = nw - 1;
  const int o0 = obs_idx[b];
  const double* boxes = cuboids + (size_t)cuboid_stride * b;
  const unsigned full = 0xffffffffu;
  int best = n_obs;                       // first obstacle (>= o0) with a sampled point inside, over this lane's rows
  unsigned long long mask = 0ull;         // splines that own such a point of obstacle `best`
  if (status[b] == UAVB_SOLVE_OK) {       // a degenerate mission has NaN coefficients: nothing to test
    for (int s = warp; s < S; s += kSweepWarps) {
      const double* c = coeffs + ((size_t)b * (max_wp - 1) + s) * 24;
      const int n = arange_len(times[(size_t)b * (max_wp - 1) + s], dt);
      if (n <= lane) continue;
      double cf[24];
#pragma unroll
      for (int k = 0; k < 24; ++k) cf[k] = __ldg(c + k);
      for (int j = lane; j < n; j += 32) {
        const double t = (double)j * dt;
        double p[3];
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {                 // the position chain of eval_row (flight_core.cuh): same operations, same bits
          double pp = cf[21 + ax];
#pragma unroll
          for (int k = 6; k >= 0; --k) pp = fma(pp, t, cf[3 * k + ax]);
          p[ax] = pp;
        }
        for (int o = o0; o < n_obs && o <= best; ++o) {
          const double* q = boxes + 6 * o;
          if (q[0] <= p[0] && p[0] <= q[1] && q[2] <= p[1] && p[1] <= q[3] && q[4] <= p[2] && p[2] <= q[5]) {   // is_collision_cuboid, inclusive (:352-357)
            if (o < best) { best = o; mask = 0ull; }
            mask |= 1ull << s;
            break;
          }
        }
      }
    }
  }
  // (first obstacle hit, mask of its splines): over the warp, then over the CTA's warps
  int first = best;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) first = min(first, __shfl_xor_sync(full, first, off));
  if (best != first) mask = 0ull;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mask |= __shfl_xor_sync(full, mask, off);
  if (lane == 0) { s_best[warp] = first; s_mask[warp] = mask; }
  __syncthreads();
  if (warp != 0) return;
  first = n_obs;
#pragma unroll
  for (int w = 0; w < kSweepWarps; ++w) first = min(first, s_best[w]);
  if (first >= n_obs) {                                   // clean against every remaining obstacle: this mission is done
    if (lane == 0) obs_idx[b] = n_obs;
    return;
  }
  mask = 0ull;
#pragma unroll
  for (int w = 0; w < kSweepWarps; ++w) mask |= s_best[w] == first ? s_mask[w] : 0ull;
  const int n_new = nw + __popcll(mask);
  if (n_new > max_wp || n_new - 1 > UAVB_MAX_SPLINES) {
    if (lane == 0) { status[b] = UAVB_SOLVE_TOO_MANY; obs_idx[b] = first; }
    return;
  }
  // insert_midpoints_at_indexes (:359-391): waypoint i moves behind the midpoints of the hit splines s < i, and spline i-1's
  // midpoint (p[i-1] + p[i]) / 2 goes right before it.  All reads, then all writes (the insertion is in place).
  double* w = waypoints + (size_t)b * max_wp * 3;
  double cur[3][3], prv[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int i = lane + 32 * r;
    if (i < nw) {
      cur[r][0] = w[3 * i]; cur[r][1] = w[3 * i + 1]; cur[r][2] = w[3 * i + 2];
      if (i > 0 && ((mask >> (i - 1)) & 1ull)) { prv[r][0] = w[3 * i - 3]; prv[r][1] = w[3 * i - 2]; prv[r][2] = w[3 * i - 1]; }
    }
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int i = lane + 32 * r;
    if (i < nw) {
      const int ni = i + __popcll(mask & bits_below(i));
      w[3 * ni] = cur[r][0]; w[3 * ni + 1] = cur[r][1]; w[3 * ni + 2] = cur[r][2];
      if (i > 0 && ((mask >> (i - 1)) & 1ull)) {
        w[3 * ni - 3] = (prv[r][0] + cur[r][0]) / 2.0; w[3 * ni - 2] = (prv[r][1] + cur[r][1]) / 2.0; w[3 * ni - 1] = (prv[r][2] + cur[r][2]) / 2.0;
      }
    }
  }
  if (lane == 0) {
    n_wp[b] = n_new;
    obs_idx[b] = first;                                   // the reference tests the same obstacle again after re-planning
    const int k = bucket_of(n_new - 1);
    next.list[(size_t)k * B + atomicAdd(next.count + k, 1)] = b;
  }
}

// Fixed pitch -> packed segments: thread (b, s) copies spline s of mission b to packed segment seg_offsets[b] + s.
__global__ void __launch_bounds__(256) minsnap_pack_kernel(const double* __restrict__ coeffs, const double* __restrict__ times,
                                                           const int* __restrict__ n_wp, int B, int max_wp, const int* __restrict__ seg_offsets,
                                                           double* __restrict__ coeffs_out, double* __restrict__ times_out) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int P = max_wp - 1;
  const long long bs = g / 24;                            // (mission, spline slot); 24 threads copy one spline
  const int k = (int)(g - bs * 24);
  if (bs >= (long long)B * P) return;
  const int b = (int)(bs / P), s = (int)(bs - (long long)b * P);
  if (s >= n_wp[b] - 1) return;
  const size_t dst = (size_t)seg_offsets[b] + s;
  coeffs_out[dst * 24 + k] = coeffs[(size_t)bs * 24 + k];
  if (k == 0) times_out[dst] = times[bs];
}

// `n` sizes the grid (an upper bound of the list length); the kernel reads the length itself from `n_dev`
static int launch_solve_list(int bucket, const double* wp, const int* n_wp, const double* vel, const int* list, const int* n_dev, int n, int max_wp,
                             double factor, double* coeffs, double* times, int* status, cudaStream_t st) {
  const int grid = div_up(n, kCorrectThreads);
  switch (bucket) {
    case 0: minsnap_solve_list_kernel<4><<<grid, kCorrectThreads, 0, st>>>(wp, n_wp, vel, list, n_dev, max_wp, factor, coeffs, times, status); break;
    case 1: minsnap_solve_list_kernel<8><<<grid, kCorrectThreads, 0, st>>>(wp, n_wp, vel, list, n_dev, max_wp, factor, coeffs, times, status); break;
    case 2: minsnap_solve_list_kernel<16><<<grid, kCorrectThreads, 0, st>>>(wp, n_wp, vel, list, n_dev, max_wp, factor, coeffs, times, status); break;
    default: minsnap_solve_list_kernel<UAVB_MAX_SPLINES><<<grid, kCorrectThreads, 0, st>>>(wp, n_wp, vel, list, n_dev, max_wp, factor, coeffs, times, status); break;
  }
  UAVB_CUDA_OK(cudaGetLastError());
  return UAVB_OK;
}

// The loop behind uavb_minsnap_correct_f64 (arguments as there).  The first round is launched blind -- every bucket's kernels over a
// grid sized for B, each reading its list length from device memory -- so a batch in which nothing is hit costs ONE host
// synchronisation; later rounds are sized from the counters read back.  ctrl (optional): a zeroed device block of ints whose first
// 2 x 4 words serve as the counters and whose other words belong to the caller (who may place n_waypoints / status in it); every
// round's read-back copies the whole block to ctrl->host (pinned) in ONE transfer.  after_first_round (optional) enqueues the caller's follow-up work between round 1
// and its read-back, so that a plan in which nothing is hit is complete after that one synchronisation (plan_shared_tables);
// n_wp_known (optional, [B] host ints): the initial n_waypoints when the host has them, to launch only the buckets in use.
int correct_missions(double* waypoints, int* n_waypoints, const double* velocity, int B, int max_wp, double factor, double dt, const double* cuboids,
                     int n_obs, long long cuboid_stride, double* coeffs_out, double* times_out, int* status_out, int* rounds_out, cudaStream_t st,
                     const std::function<int()>* after_first_round, const int* n_wp_known, const CtrlBlock* ctrl) {
  int result = UAVB_OK;
  cudaError_t e = cudaSuccess;
  {
    DevPool pool(st);
    int* obs_idx = pool.alloc<int>(B);
    int* lists = pool.alloc<int>((size_t)2 * kBuckets * B);
    int* counts = ctrl ? ctrl->dev : pool.alloc<int>(2 * kBuckets);
    if (pool.err != cudaSuccess) return set_error(UAVB_ENOMEM, "minsnap_correct: %s", cudaGetErrorString(pool.err));
    WorkLists wl[2] = {{lists, counts}, {lists + (size_t)kBuckets * B, counts + kBuckets}};
    if (!ctrl) UAVB_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int) * 2 * kBuckets, st));      // a control block arrives zeroed
    correct_classify_kernel<<<div_up(B, 256), 256, 0, st>>>(n_waypoints, B, max_wp, obs_idx, status_out, wl[0]);
    e = cudaGetLastError();
    int host_counts[kBuckets] = {B, B, B, B};             // round 1: upper bounds, the kernels read the real lengths
    if (n_wp_known) {                                     // ... unless the caller knows the spline counts: only the buckets in use
      for (int k = 0; k < kBuckets; ++k) host_counts[k] = 0;
      for (int b = 0; b < B; ++b)
        if (n_wp_known[b] >= 2 && n_wp_known[b] <= max_wp) ++host_counts[bucket_of(n_wp_known[b] - 1)];
    }
    int cur = 0, rounds = 0;
    while (e == cudaSuccess && !result) {
      long long todo = 0;
      for (int k = 0; k < kBuckets; ++k) todo += host_counts[k];
      if (todo == 0) break;
      ++rounds;
      const WorkLists& in = wl[cur];
      const WorkLists& out = wl[cur ^ 1];
      if (rounds > 1) e = cudaMemsetAsync(out.count, 0, sizeof(int) * kBuckets, st);
      long long n_ctas = 0;
      for (int k = 0; k < kBuckets && e == cudaSuccess && !result; ++k) {
        if (host_counts[k] == 0) continue;
        n_ctas += host_counts[k];
        result = launch_solve_list(k, waypoints, n_waypoints, velocity, in.list + (size_t)k * B, in.count + k, host_counts[k], max_wp, factor,
                                   coeffs_out, times_out, status_out, st);
      }
      if (e == cudaSuccess && !result && n_obs > 0 && n_ctas > 0) {   // one sweep launch for every bucket (grid: the sum of the lists' bounds)
        correct_sweep_kernel<<<(unsigned)n_ctas, 32 * kSweepWarps, 0, st>>>(coeffs_out, times_out, waypoints, n_waypoints, obs_idx, in, B, max_wp, dt,
                                                                           cuboids, n_obs, cuboid_stride, status_out, out);
        e = cudaGetLastError();
      }
      // work the caller wants behind round 1 and in front of its read-back (valid if nothing was hit: rounds_out == 1)
      if (e == cudaSuccess && !result && rounds == 1 && after_first_round) result = (*after_first_round)();
      if (e == cudaSuccess && !result) {
        if (ctrl) {                                       // the whole control block (counters + the caller's words) in one copy
          e = cudaMemcpyAsync(ctrl->host, ctrl->dev, sizeof(int) * ctrl->ints, cudaMemcpyDeviceToHost, st);
          if (ctrl->no_wait) break;                       // the caller reads the block later (speculative plan: round 1 only)
          if (e == cudaSuccess) e = cudaStreamSynchronize(st);
          for (int k = 0; k < kBuckets; ++k) host_counts[k] = ctrl->host[(cur ^ 1) * kBuckets + k];
        } else {
          e = cudaMemcpyAsync(host_counts, out.count, sizeof(host_counts), cudaMemcpyDeviceToHost, st);
          if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
      }
      cur ^= 1;
    }
    if (rounds_out) *rounds_out = rounds;
  }
  if (e != cudaSuccess && !result) {
    cudaStreamSynchronize(st);
    cudaGetLastError();
    result = set_error(UAVB_ECUDA, "minsnap_correct: %s", cudaGetErrorString(e));
  }
  return result;
}

// seg_table / seg_yaw0 of a shared mission: 1 and the table's look-ahead yaw at the first segment of every table, 0 elsewhere.
__global__ void shared_seg_flags_kernel(const int* __restrict__ offs, const double* __restrict__ yaw0, int T, int n_seg, int* __restrict__ seg_table,
                                        double* __restrict__ seg_yaw0) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  int flag = 0;
  double y = 0.0;
  for (int k = 0; k < T; ++k)
    if (offs[k] == s && offs[k + 1] > offs[k]) { flag = 1; y = yaw0[k]; }
  seg_table[s] = flag;
  seg_yaw0[s] = y;
}

// One mission made of T consecutive MinimumSnap tables (take-off + course, main.py:80-83), each planned with the correction
// loop, packed into the segment arrays of a shared-mission rollout.  ONE host synchronisation when nothing is hit.
int plan_shared_tables(int T, const double* const* d_waypoints, const int* n_wp_in, const double* d_velocity, double factor, double dt,
                       const double* d_cuboids, int n_obs, int cap_seg, double* seg_coeffs, double* seg_times, int* seg_rows, int* seg_table,
                       double* seg_yaw0, int* n_seg_out, int* rows_out, int* status_out, int* rounds_out, cudaStream_t st, int* async_report) {
  constexpr int kMaxWp = UAVB_MAX_SPLINES + 1;
  int n_seg0 = 0;
  for (int k = 0; k < T; ++k) {
    UAVB_REQUIRE(d_waypoints[k] != nullptr && n_wp_in[k] >= 2 && n_wp_in[k] <= kMaxWp, "plan_shared: every table needs 2 .. UAVB_MAX_SPLINES + 1 waypoints");
    n_seg0 += n_wp_in[k] - 1;
  }
  UAVB_REQUIRE(n_seg0 <= cap_seg, "plan_shared: cap_seg is smaller than the mission");
  int result = UAVB_OK;
  {
    // ONE scratch allocation, ONE upload and ONE read-back per round.  Control block (ints): [0, 8) the loop's counters,
    // then n_waypoints [T], status [T], table rows [T], segment offsets [T + 1]; its host mirror is pinned.
    const int o_nwp = 2 * kBuckets, o_status = o_nwp + T, o_total = o_status + T, o_offs = o_total + T, n_ints = o_offs + T + 1;
    const size_t b_ints = ((size_t)n_ints * sizeof(int) + 15) / 16 * 16;
    const size_t b_fixed = sizeof(double) * (size_t)T * kMaxWp * 3, b_cf = sizeof(double) * (size_t)T * UAVB_MAX_SPLINES * 24;
    const size_t b_tf = sizeof(double) * (size_t)T * UAVB_MAX_SPLINES, b_yaw = sizeof(double) * (size_t)T;
    DevPool pool(st);
    char* blob = pool.alloc<char>(b_fixed + b_cf + b_tf + b_yaw + b_ints);
    int* host = static_cast<int*>(pinned_scratch(0, sizeof(int) * 2 * n_ints));   // [0, n_ints) read-back mirror, [n_ints, 2 n_ints) upload image
    if (pool.err != cudaSuccess || host == nullptr) return set_error(UAVB_ENOMEM, "plan_shared: scratch allocation failed");
    double* d_fixed = reinterpret_cast<double*>(blob);
    double* d_cf = reinterpret_cast<double*>(blob + b_fixed);
    double* d_tf = reinterpret_cast<double*>(blob + b_fixed + b_cf);
    double* d_yaw0 = reinterpret_cast<double*>(blob + b_fixed + b_cf + b_tf);
    int* d_ctrl = reinterpret_cast<int*>(blob + b_fixed + b_cf + b_tf + b_yaw);
    int* d_n_wp = d_ctrl + o_nwp; int* d_status = d_ctrl + o_status; int* d_total = d_ctrl + o_total; int* d_offs = d_ctrl + o_offs;
    // the upload image must outlive the asynchronous copy that reads it: in the speculative form it lives in the caller's block
    int* up = async_report ? async_report + UAVB_PLAN_REPORT_INTS / 2 : host + n_ints;
    for (int i = 0; i < n_ints; ++i) up[i] = 0;
    up[o_offs] = 0;
    for (int k = 0; k < T; ++k) { up[o_nwp + k] = n_wp_in[k]; up[o_offs + k + 1] = up[o_offs + k] + n_wp_in[k] - 1; }
    UAVB_CUDA_OK(cudaMemcpyAsync(d_ctrl, up, sizeof(int) * n_ints, cudaMemcpyHostToDevice, st));       // zeroed counters, n_waypoints, offsets
    for (int k = 0; k < T; ++k)
      UAVB_CUDA_OK(cudaMemcpyAsync(d_fixed + (size_t)k * kMaxWp * 3, d_waypoints[k], sizeof(double) * 3 * n_wp_in[k], cudaMemcpyDeviceToDevice, st));
    // pack + table geometry + segment flags for the segment offsets now in d_offs
    auto tail = [&](int n_seg) -> int {
      int rc = uavb_minsnap_pack_f64(d_cf, d_tf, d_n_wp, T, kMaxWp, d_offs, seg_coeffs, seg_times, st);
      if (!rc) rc = uavb_minsnap_table_meta_f64(seg_coeffs, seg_times, d_offs, T, dt, seg_rows, d_yaw0, d_total, st);
      if (rc) return rc;
      shared_seg_flags_kernel<<<div_up(n_seg, 128), 128, 0, st>>>(d_offs, d_yaw0, T, n_seg, seg_table, seg_yaw0);
      UAVB_CUDA_OK(cudaGetLastError());
      return UAVB_OK;
    };
    const std::function<int()> speculative = [&]() { return tail(n_seg0); };
    const CtrlBlock ctrl{d_ctrl, async_report ? async_report : host, n_ints, async_report != nullptr};
    int rounds = 0, n_seg = n_seg0;
    result = correct_missions(d_fixed, d_n_wp, d_velocity, T, kMaxWp, factor, dt, d_cuboids, n_obs, 0, d_cf, d_tf, d_status, &rounds, st, &speculative, n_wp_in,
                              &ctrl);
    if (async_report) {                                     // speculative: nothing was waited for; the caller checks the report later
      *n_seg_out = n_seg0;
      if (rounds_out) *rounds_out = 1;
      return result;
    }
    if (!result && rounds > 1) {                            // midpoints were inserted: lay the mission out again
      n_seg = 0;
      for (int k = 0; k < T; ++k) { up[o_offs + k] = n_seg; n_seg += host[o_nwp + k] - 1; }
      up[o_offs + T] = n_seg;
      if (n_seg > cap_seg) return set_error(UAVB_EINVAL, "plan_shared: the corrected mission has %d splines, cap_seg is %d", n_seg, cap_seg);
      UAVB_CUDA_OK(cudaMemcpyAsync(d_offs, up + o_offs, sizeof(int) * (T + 1), cudaMemcpyHostToDevice, st));
      result = tail(n_seg);
      if (!result) UAVB_CUDA_OK(cudaMemcpyAsync(host, d_ctrl, sizeof(int) * n_ints, cudaMemcpyDeviceToHost, st));
      if (!result) UAVB_CUDA_OK(cudaStreamSynchronize(st));
    }
    if (!result) {
      for (int k = 0; k < T; ++k) { status_out[k] = host[o_status + k]; rows_out[k] = host[o_total + k]; }
      *n_seg_out = n_seg;
      if (rounds_out) *rounds_out = rounds;
    }
  }
  return result;
}

}  // namespace uavb

using namespace uavb;

extern "C" int uavb_minsnap_correct_f64(double* waypoints, int* n_waypoints, const double* velocity, int B, int max_wp, double factor, double dt,
                                        const double* cuboids, int n_obs, long long cuboid_stride, double* coeffs_out, double* times_out,
                                        int* status_out, int* rounds_out, void* stream) {
  UAVB_REQUIRE(B >= 0 && max_wp >= 2 && max_wp <= UAVB_MAX_SPLINES + 1, "minsnap_correct: B >= 0 and 2 <= max_wp <= UAVB_MAX_SPLINES + 1 required");
  UAVB_REQUIRE(B == 0 || (waypoints && n_waypoints && velocity && coeffs_out && times_out && status_out), "minsnap_correct: NULL pointer");
  UAVB_REQUIRE(dt > 0.0, "minsnap_correct: dt > 0 required");
  UAVB_REQUIRE(n_obs >= 0 && (n_obs == 0 || cuboids != nullptr), "minsnap_correct: n_obs > 0 needs cuboids");
  UAVB_REQUIRE(cuboid_stride == 0 || cuboid_stride >= 6LL * n_obs, "minsnap_correct: cuboid_stride must be 0 or >= 6 n_obs");
  int rc = require_device();
  if (rc) return rc;
  if (rounds_out) *rounds_out = 0;
  if (B == 0) return UAVB_OK;
  return correct_missions(waypoints, n_waypoints, velocity, B, max_wp, factor, dt, cuboids, n_obs, cuboid_stride, coeffs_out, times_out, status_out,
                          rounds_out, static_cast<cudaStream_t>(stream), nullptr, nullptr, nullptr);
}

extern "C" int uavb_minsnap_pack_f64(const double* coeffs, const double* times, const int* n_waypoints, int B, int max_wp, const int* seg_offsets,
                                     double* coeffs_out, double* times_out, void* stream) {
  UAVB_REQUIRE(B >= 0 && max_wp >= 2, "minsnap_pack: B >= 0 and max_wp >= 2 required");
  UAVB_REQUIRE(B == 0 || (coeffs && times && n_waypoints && seg_offsets && coeffs_out && times_out), "minsnap_pack: NULL pointer");
  int rc = require_device();
  if (rc) return rc;
  if (B == 0) return UAVB_OK;
  const long long n = (long long)B * (max_wp - 1) * 24;
  minsnap_pack_kernel<<<div_up(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(coeffs, times, n_waypoints, B, max_wp, seg_offsets, coeffs_out,
                                                                                     times_out);
  UAVB_CUDA_OK(cudaGetLastError());
  return UAVB_OK;
}

extern "C" int uavb_plan_shared_f64(int n_tables, const double* const* table_waypoints, const int* table_n_waypoints, const double* table_velocity,
                                    double factor, double dt, const double* cuboids, int n_obs, int cap_seg, double* seg_coeffs, double* seg_times,
                                    int* seg_rows, int* seg_table, double* seg_yaw0, int* n_seg_out, int* rows_out, int* status_out, int* rounds_out,
                                    int* async_report, void* stream) {
  UAVB_REQUIRE(table_waypoints && table_n_waypoints && table_velocity && seg_coeffs && seg_times && seg_rows && seg_table && seg_yaw0 && n_seg_out,
               "plan_shared: NULL pointer");
  UAVB_REQUIRE(async_report != nullptr || (rows_out && status_out), "plan_shared: rows_out and status_out are required unless async_report is given");
  UAVB_REQUIRE(n_tables >= 1 && n_tables <= kMaxSharedTables, "plan_shared: 1 .. 8 tables");
  UAVB_REQUIRE(dt > 0.0 && n_obs >= 0 && (n_obs == 0 || cuboids != nullptr), "plan_shared: dt > 0; n_obs > 0 needs cuboids");
  int rc = require_device();
  if (rc) return rc;
  return plan_shared_tables(n_tables, table_waypoints, table_n_waypoints, table_velocity, factor, dt, cuboids, n_obs, cap_seg, seg_coeffs, seg_times,
                            seg_rows, seg_table, seg_yaw0, n_seg_out, rows_out, status_out, rounds_out, static_cast<cudaStream_t>(stream), async_report);
}

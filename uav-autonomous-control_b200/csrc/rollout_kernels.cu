// rollout_kernels.cu -- K2 host side: argument checks, launch policy (time slicing, scratch), the set-point table of a
// shared mission, and the C entry points.  The device code lives in rollout_impl.cuh and is instantiated in
// rollout_sliced.cu / rollout_log.cu / rollout_f64.cu.
#ifdef UAVB_DEV
#include <stdlib.h>
#endif

#include "rollout_impl.cuh"

namespace uavb {

// ---------------------------------------------------------------------------------------------
// Set-point table of a shared mission.  Pass 1 (one thread per row): polynomial values and the row's own heading
// (velocity direction where the horizontal speed reaches the threshold, NaN elsewhere).  Pass 2: every row takes the
// heading of the last valid row at or before it inside its MinimumSnap table, or the table's look-ahead yaw when there is
// none -- exactly the values cursor_target carries from row to row.
__global__ void __launch_bounds__(128) target_rows_kernel(const double* __restrict__ coeffs, const int* __restrict__ rows,
                                                          const int* __restrict__ table, int n_seg, double dt_outer,
                                                          TargetRow* __restrict__ out, float2* __restrict__ raw, int* __restrict__ tstart,
                                                          int n_rows) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_rows) return;
  int seg = 0, first = 0, tab_first = 0;                    // segment of row g, its first row, first row of its table
  int last_seg = 0, last_first = 0, last_tab = 0;            // last segment that owns rows (for rows past the end of the mission)
  for (; seg < n_seg; ++seg) {
    if (table[seg]) tab_first = first;
    if (rows[seg] > 0) { last_seg = seg; last_first = first; last_tab = tab_first; }
    if (g < first + rows[seg]) break;
    first += rows[seg];
  }
  int local = g - first;
  if (seg == n_seg) { seg = last_seg; first = last_first; tab_first = last_tab; local = rows[seg] - 1; }   // hold the last row (main.py:61)
  const double* cf = coeffs + (size_t)seg * 24;
  double p[3], v[3], a[3];
  eval_row([cf](int i) { return __ldg(cf + i); }, (double)local * dt_outer, p, v, a);
  TargetRow r;
  r.x = p[0]; r.y = p[1]; r.z = p[2];
  r.vx = (float)v[0]; r.vy = (float)v[1]; r.vz = (float)v[2];
  r.ax = (float)a[0]; r.ay = (float)a[1]; r.az = (float)a[2];
  r.yc = 0.f; r.ys = 0.f;
  out[g] = r;
  const bool valid = speed2_unfused(v[0], v[1]) >= kSpeed2Min;
  raw[g] = valid ? make_float2((float)v[0], (float)v[1]) : make_float2(nanf(""), nanf(""));
  tstart[g] = (tab_first << 8) | seg;                       // n_seg <= 2 * UAVB_MAX_SPLINES < 256
}

__global__ void __launch_bounds__(128) target_heading_kernel(const float2* __restrict__ raw, const int* __restrict__ tstart,
                                                             const double* __restrict__ yaw0, const int* __restrict__ rows, int n_seg,
                                                             TargetRow* __restrict__ out, int n_rows) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_rows) return;
  const int tab_first = tstart[g] >> 8;
  int r = g;
  float2 h = raw[r];
  while (h.x != h.x && r > tab_first) h = raw[--r];
  if (h.x != h.x) {                                         // no valid row yet in this table: its look-ahead yaw
    int seg = 0, first = 0;
    while (seg < n_seg && first != tab_first) first += rows[seg++];
    const double y0 = yaw0[seg];
    h = make_float2((float)cos(y0), (float)sin(y0));
  }
  out[g].yc = h.x; out[g].ys = h.y;
}

static int check_args(const uavb_rollout_args* a, bool f64) {
  UAVB_REQUIRE(a != nullptr, "rollout: args is NULL");
  UAVB_REQUIRE(a->B >= 0 && a->n_ticks >= 0, "rollout: B and n_ticks must be >= 0");
  UAVB_REQUIRE(a->inner_per_outer >= 1, "rollout: inner_per_outer must be >= 1");
  UAVB_REQUIRE(a->seg_coeffs && a->seg_rows && a->seg_table && a->seg_yaw0, "rollout: mission segment arrays are required");
  UAVB_REQUIRE((a->mission_seg_begin == nullptr) == (a->mission_seg_count == nullptr),
               "rollout: mission_seg_begin and mission_seg_count must be given together");
  UAVB_REQUIRE(a->mission_seg_begin != nullptr || a->n_seg_shared >= 1, "rollout: n_seg_shared must be >= 1 for a shared mission");
  UAVB_REQUIRE(a->resume || a->start != nullptr, "rollout: start is required when resume = 0");
  UAVB_REQUIRE(a->start_stride == 0 || a->start_stride == 3, "rollout: start_stride must be 0 or 3");
  UAVB_REQUIRE(a->goal == nullptr || a->goal_stride == 0 || a->goal_stride == 3, "rollout: goal_stride must be 0 or 3");
  UAVB_REQUIRE(!a->resume || a->carry != nullptr, "rollout: resume = 1 needs a carry block");
  UAVB_REQUIRE(!(f64 && (a->resume || a->carry)), "rollout_f64: the fp64 validation rollout has no carry/resume");
  UAVB_REQUIRE(a->log_stride >= 0, "rollout: log_stride must be >= 0");
  UAVB_REQUIRE(a->n_slices >= 0, "rollout: n_slices must be >= 0");
  UAVB_REQUIRE(a->log_stride == 0 || a->log_out != nullptr, "rollout: log_stride > 0 needs log_out");
  UAVB_REQUIRE(a->log_stride == 0 || a->B <= 300000000LL, "rollout: a state log supports at most 3e8 rollouts per launch");
  UAVB_REQUIRE(a->traj_out == nullptr || (!f64 && a->log_stride == 0), "rollout: traj_out is for fp32 rollouts without a state log");
  UAVB_REQUIRE(a->traj_out == nullptr || (a->traj_max_samples >= 1 && a->traj_interval >= 0.0), "rollout: traj_out needs traj_max_samples >= 1 and traj_interval >= 0");
  UAVB_REQUIRE(a->n_obs >= 0 && a->n_obs <= 1024, "rollout: n_obs out of range");
  UAVB_REQUIRE(a->n_obs == 0 || (a->aabbs != nullptr && a->n_obs_sets >= 1), "rollout: n_obs > 0 needs aabbs and n_obs_sets >= 1");
  UAVB_REQUIRE(a->dt_outer > 0.0 && a->veh.dt > 0.0 && a->veh.mass > 0.0, "rollout: dt_outer, veh.dt and veh.mass must be positive");
  UAVB_REQUIRE(a->shared_targets == nullptr || (a->mission_seg_begin == nullptr && a->n_target_rows >= 1),
               "rollout: shared_targets needs a shared mission and n_target_rows >= 1");
  return UAVB_OK;
}

// Stream-ordered scratch that frees itself when the launcher returns (the free is ordered after the kernel).
struct StreamScratch {
  cudaStream_t st;
  void* p = nullptr;
  explicit StreamScratch(cudaStream_t s) : st(s) {}
  ~StreamScratch() { if (p) cudaFreeAsync(p, st); }
  cudaError_t alloc(size_t bytes) {
    cudaMemPool_t pool = scratch_pool();
    return pool ? cudaMallocFromPoolAsync(&p, bytes, pool, st) : cudaMallocAsync(&p, bytes, st);
  }
};

template <class R> static int launch_rollout(const uavb_rollout_args* a, void* stream) {
  int rc = check_args(a, sizeof(R) == 8);
  if (rc) return rc;
  rc = require_device();
  if (rc) return rc;
  if (a->B == 0) return UAVB_OK;
  RolloutDev<R> p;
  p.a = *a;
  make_vehu<R>(p.u, a->veh, a->dt_outer);
  McValues mc;
  mc_from_vehicle(mc, a->veh);
  make_vehp<R>(p.vp, a->veh, mc);
  const bool log = a->log_stride > 0;
  // fp32 metrics-only rollouts with per-rollout missions fly one drone per thread (rollout_sliced_scalar_kernel), everything else in fp32 a pair
  const bool scalar32 = sizeof(R) == 4 && !log && a->traj_out == nullptr && a->mission_seg_begin != nullptr && a->pair_kernel_only == 0;
  const int threads = sizeof(R) == 8 ? kRolloutThreadsF64 : (log ? kRolloutThreadsLog : (scalar32 ? kScalarThreads : kRolloutThreads));
  // fp32: a thread flies a PAIR of drones (rollout_pair.cuh), so a CTA (= one work group of the slice scheduler) covers 2 x threads
  const int per_cta = (sizeof(R) == 8 || scalar32) ? threads : 2 * threads;
  const int grid = div_up(a->B, per_cta);
  size_t smem = (a->n_obs > 0 && a->aabb_set == nullptr) ? sizeof(float) * 6 * a->n_obs : 0;
  p.coeff_cache_offset = -1;
  p.vehp_cache = nullptr;
  p.vehp_stride = 0;
  const bool from_table = a->shared_targets != nullptr && a->mission_seg_begin == nullptr;
  if (!from_table && sizeof(R) == 4) {                            // fp32 kernels stage the current spline in shared memory
    smem = (smem + 7) / 8 * 8;
    p.coeff_cache_offset = (int)(smem / 8);
    smem += sizeof(double) * 24 * per_cta;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool mc_any = a->mc_mass || a->mc_inertia || a->mc_gains || a->mc_wind;
  if constexpr (sizeof(R) == 8) {                      // validation build
    launch_rollout_f64(log, mc_any, grid, smem, st, p);
  } else {
    int sms = 0;
    rc = sm_count_cached(&sms);
    if (rc) return rc;
    // fp32 launches ALWAYS run the time-sliced persistent kernel (16 warps per SM at 128 registers, CTA shapes in rollout_impl.cuh), with a single
    // slice when slicing has nothing to gain; a state log is written slice by slice into its place.  One compiled body for every batch size keeps
    // per-rollout results independent of how a job is sharded (ptxas fuses mul+add differently under different register
    // caps, so differently compiled variants are NOT bit-identical to each other).
    const int slots = sms * (log ? kRolloutCtasPerSmLog : (scalar32 ? kScalarCtasPerSm : kRolloutCtasPerSm));
    // ~32 items per resident CTA keep the tail near 3 % of the launch; slices are whole outer periods of >= 100 ticks
    constexpr int kMinChunkTicks = 100;
    long long want = grid > slots ? (32LL * slots + grid - 1) / grid : 1;
#ifdef UAVB_DEV
    if (const char* force = getenv("UAVB_ROLLOUT_CHUNKS")) {       // development builds only (build.py --variant dev -DUAVB_DEV): force the slice count
      const int f = atoi(force);
      if (f > 0) want = f < 4096 ? f : 4096;
    }
#endif
    if (a->n_slices > 0) want = a->n_slices < 4096 ? a->n_slices : 4096;         // uavb.h: results are independent of the slice count
    const int period = a->inner_per_outer;
    int chunk = (int)((a->n_ticks + want - 1) / want);
    chunk = ((chunk + period - 1) / period) * period;
    if (chunk < kMinChunkTicks) chunk = ((kMinChunkTicks + period - 1) / period) * period;
    const int n_chunks = a->n_ticks > 0 ? (a->n_ticks + chunk - 1) / chunk : 1;
    StreamScratch flags(st), carry(st), consts(st);
    const size_t n_flags = (size_t)grid + 1;
    cudaError_t e = flags.alloc(n_flags * sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(flags.p, 0, n_flags * sizeof(int), st);
    SliceSched sch;
    sch.final_carry = a->carry != nullptr;
    if (e == cudaSuccess && a->carry == nullptr && n_chunks > 1) {
      e = carry.alloc(sizeof(float) * UAVB_CARRY_WORDS * (size_t)a->B);
      p.a.carry = static_cast<float*>(carry.p);
    }
    if (e == cudaSuccess && mc_any && n_chunks > 1 && !scalar32) {      // the pairs' per-rollout constants, made by slice 0 (rollout_impl.cuh)
      const long long padded = 2LL * div_up(a->B, 2);
      e = consts.alloc(sizeof(float) * kVehpWords * (size_t)padded);
      p.vehp_cache = static_cast<float*>(consts.p);
      p.vehp_stride = padded;
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      return set_error(UAVB_ENOMEM, "rollout: scratch allocation failed: %s", cudaGetErrorString(e));
    }
    sch.counter = static_cast<int*>(flags.p);
    sch.done = sch.counter + 1;
    sch.n_groups = grid; sch.n_chunks = n_chunks; sch.chunk_ticks = chunk;
    sch.n_tab = 0;
    // Slices of DECREASING length (unless the caller forces a slice count): every slice costs a carry round trip
    // (~5 us per 1e5 rollouts), the tail of the launch is one slice of the last kind long -- so each slice takes a fifth of what is left
    // (whole outer periods), down to a length that is <= 0.5 % of a resident CTA's share of the launch and >= 200 ticks.  Measured on the
    // headline launch (1e5 x 10 760 ticks, 1 563 groups on 1 184 resident CTAs): 15 such slices 4.605 ms against 4.653 ms for the 25 equal
    // slices of the rule above (tools/pipeline_probe.py).  Results do not depend on the slicing (tests/test_rollout_gpu.py).
    if (a->n_slices <= 0 && n_chunks > 1) {
      const double share = (double)grid / (double)slots * (double)a->n_ticks;       // ticks a resident CTA flies in this launch
      int min_len = share / 200.0 < (double)a->n_ticks ? (int)(share / 200.0) : a->n_ticks;     // (never longer than the launch: no overflow below)
      int floor_len = 200, part = 5;
#ifdef UAVB_DEV
      if (const char* e = getenv("UAVB_SLICE_MIN")) floor_len = atoi(e) > 0 ? atoi(e) : floor_len;      // development builds only
      if (const char* e = getenv("UAVB_SLICE_PART")) part = atoi(e) > 1 ? atoi(e) : part;
#endif
      if (min_len < floor_len) min_len = floor_len;
      min_len = (int)(((long long)min_len + period - 1) / period * period);
      int t = 0, n = 0;
      sch.tab[0] = 0;
      while (t < a->n_ticks) {
        const int left = a->n_ticks - t;
        int len = left / part;
        if (len < min_len) len = min_len;
        const long long whole = ((long long)len + period - 1) / period * period;      // whole outer periods
        len = whole < left ? (int)whole : left;
        if (n == kSliceTab - 1) len = left;
        if (left - len < min_len / 2) len = left;                                    // no sliver at the end
        t += len;
        sch.tab[++n] = t;
      }
      sch.n_tab = n; sch.n_chunks = n;
    }
    const int pgrid = slots < grid ? slots : grid;
    // state log: through the TMA unit when the [samples x 13][B] log is a legal tensor (rows a multiple of 16 bytes, 16-byte aligned)
    LogTma maps;
    bool tma_log = false;
    if (log && a->log_tma >= 0) {
      const unsigned long long n_samples = (unsigned long long)(a->n_ticks / a->log_stride);
      if (a->B % 4 == 0 && n_samples > 0 && n_samples * 13ull < 2000000000ull) {
        tma_log = make_tensor_map_2d(&maps.box_k, a->log_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (unsigned long long)a->B, n_samples * 13ull, 4ull * a->B, 64,
                                     13 * kLogTmaSamples, CU_TENSOR_MAP_SWIZZLE_NONE) &&
                  make_tensor_map_2d(&maps.box_1, a->log_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (unsigned long long)a->B, n_samples * 13ull, 4ull * a->B, 64, 13,
                                     CU_TENSOR_MAP_SWIZZLE_NONE);
      }
      if (tma_log) {
        smem = (smem + 127) / 128 * 128;
        maps.stage_offset = (int)smem;
        smem += sizeof(float) * kLogTmaWarpFloats * (kRolloutThreadsLog / 32);
      }
    }
    if (log) launch_rollout_sliced_log(mc_any, from_table, pgrid, smem, st, *reinterpret_cast<RolloutDev<float>*>(&p), sch, tma_log ? &maps : nullptr);
    else if (scalar32) launch_rollout_sliced_scalar(mc_any, pgrid, smem, st, *reinterpret_cast<RolloutDev<float>*>(&p), sch);
    else if (a->traj_out) launch_rollout_sliced_traj(mc_any, from_table, pgrid, smem, st, *reinterpret_cast<RolloutDev<float>*>(&p), sch);
    else launch_rollout_sliced(mc_any, from_table, pgrid, smem, st, *reinterpret_cast<RolloutDev<float>*>(&p), sch);
  }
  UAVB_CUDA_OK(cudaGetLastError());
  return UAVB_OK;
}

}  // namespace uavb

extern "C" int uavb_rollout_targets_f64(const double* seg_coeffs, const int* seg_rows, const int* seg_table, const double* seg_yaw0, int n_seg,
                                        double dt_outer, void* targets_out, int n_rows, void* stream) {
  using namespace uavb;
  static_assert(sizeof(TargetRow) == UAVB_TARGET_ROW_BYTES, "TargetRow layout");
  UAVB_REQUIRE(seg_coeffs && seg_rows && seg_table && seg_yaw0 && targets_out, "rollout_targets: NULL pointer");
  UAVB_REQUIRE(n_seg >= 1 && n_seg <= 255 && n_rows >= 0 && dt_outer > 0.0, "rollout_targets: 1 <= n_seg <= 255, n_rows >= 0, dt_outer > 0 required");
  int rc = require_device();
  if (rc) return rc;
  if (n_rows == 0) return UAVB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamScratch tmp(st);
  if (tmp.alloc((size_t)n_rows * (sizeof(float2) + sizeof(int))) != cudaSuccess) {
    cudaGetLastError();
    return set_error(UAVB_ENOMEM, "rollout_targets: scratch allocation failed");
  }
  float2* raw = static_cast<float2*>(tmp.p);
  int* tstart = reinterpret_cast<int*>(raw + n_rows);
  TargetRow* out = static_cast<TargetRow*>(targets_out);
  target_rows_kernel<<<div_up(n_rows, 128), 128, 0, st>>>(seg_coeffs, seg_rows, seg_table, n_seg, dt_outer, out, raw, tstart, n_rows);
  UAVB_CUDA_OK(cudaGetLastError());
  target_heading_kernel<<<div_up(n_rows, 128), 128, 0, st>>>(raw, tstart, seg_yaw0, seg_rows, n_seg, out, n_rows);
  UAVB_CUDA_OK(cudaGetLastError());
  return UAVB_OK;
}

extern "C" int uavb_rollout_f32(const uavb_rollout_args* args, void* stream) { return uavb::launch_rollout<float>(args, stream); }
extern "C" int uavb_rollout_f64(const uavb_rollout_args* args, void* stream) { return uavb::launch_rollout<double>(args, stream); }
extern "C" void uavb_vehicle_defaults(uavb_vehicle* veh) {
  if (veh) uavb::vehicle_defaults(veh);
}

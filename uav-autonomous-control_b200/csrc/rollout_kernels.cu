// rollout_kernels.cu -- K2: persistent closed-loop rollout, one thread per drone (sm_100a).
//
// Each thread keeps its drone's 13-state, position low part, rotor speeds, controller integrator,
// commands, table cursor and metric accumulators in registers for the whole launch (thousands of
// 1 kHz ticks), evaluates the min-snap set-point in fp64 at the 100 Hz outer rate, runs the cascade,
// allocation, motor lag and rigid-body step in fp32 every tick and touches HBM only for
//   * its Monte-Carlo parameters (once), the 24 doubles of the current spline (once per outer
//     period, broadcast through L1 when the mission is shared),
//   * the optional decimated state log [sample][field][B] (one 4-byte store per field and thread:
//     a warp writes one full 128-byte line per field, no partial sectors),
//   * metrics / final state / carry (once).
// Nothing here is a dense contraction, so no tensor-core path exists; the bound is the FP32 issue
// rate (metrics-only) or HBM (full-rate log).  See DESIGN.md "K2".
#include <stdlib.h>

#include "rollout_core.cuh"
#include "uavb_common.cuh"
#include "veh_setup.cuh"

namespace uavb {

constexpr int kRolloutThreads = 64;
// Residency variants: K CTAs of 64 threads per SM, i.e. a register cap of 128 (K = 8), 96 (K = 10) or 80 (K = 12).
// Measured on B200 (tools/k_sweep.py, profiles/): K = 8 sustains the highest tick rate -- the spill-free 128-register
// body with 16 warps per SM beats 24 warps at 80 registers (139 vs 126 G ticks/s at 5e5 rollouts) -- so it is the
// default; the other variants remain selectable for experiments (UAVB_ROLLOUT_K).
constexpr int kRolloutCtasMin = 8, kRolloutCtasMax = 12;
// registers per thread for K resident CTAs (allocated per warp in units of 512, i.e. 16 per thread: 80 -> 12 CTAs, 96 -> 10, 128 -> 8)
constexpr int rollout_regs(int K) { return K >= 12 ? 80 : K >= 10 ? 96 : 128; }

template <class R> struct RolloutDev {
  uavb_rollout_args a;
  VehU<R> u;      // launch-uniform constants (constant bank)
  VehP<R> vp;     // per-rollout constants when no Monte-Carlo override is given (constant bank)
  int coeff_cache_offset;   // offset (in doubles) of the [24][64] coefficient staging area in dynamic shared memory, -1 = none
};

// Obstacle set: SHARED = one set for the launch, staged in the CTA's dynamic shared memory (LDS.64 broadcast reads);
// otherwise a per-rollout set in global memory.  A box is three (min, max) pairs; inclusive bounds exactly as
// is_collision_cuboid (minimum_snap.py:352-357).
template <bool SHARED> struct BoxesT {
  static constexpr bool kAny = true;
  const float* b;
  int n;
  __device__ __forceinline__ const float2* box(int i) const {
    if constexpr (SHARED) {
      extern __shared__ float2 s_box_pairs[];              // the same dynamic shared memory stage_shared_boxes fills
      return s_box_pairs + 3 * i;
    } else {
      return reinterpret_cast<const float2*>(b) + 3 * i;
    }
  }
  template <class R> __device__ __forceinline__ bool hit(R x, R y, R z) const {
    bool h = false;
    for (int i = 0; i < n; ++i) {
      const float2* q = box(i);
      const float2 bx = q[0], by = q[1], bz = q[2];
      h |= (bx.x <= x) & (x <= bx.y) & (by.x <= y) & (y <= by.y) & (bz.x <= z) & (z <= bz.y);
    }
    return h;
  }
  // true when some box comes within `reach` of the point in every axis (conservative: Chebyshev gap)
  template <class R> __device__ __forceinline__ bool within(R x, R y, R z, R reach) const {
    bool w = false;
    for (int i = 0; i < n; ++i) {
      const float2* q = box(i);
      const float2 bx = q[0], by = q[1], bz = q[2];
      const R gx = fmax((R)bx.x - x, x - (R)bx.y), gy = fmax((R)by.x - y, y - (R)by.y), gz = fmax((R)bz.x - z, z - (R)bz.y);
      w |= !(fmax(gx, fmax(gy, gz)) > reach);          // NaN positions keep watching
    }
    return w;
  }
};

// [sample][13][B] state log, one sample after every `stride` ticks.
template <class R> struct GlobalLog {
  static constexpr bool kNormEveryTick = true;
  R* out;
  long long B;
  int stride, left;
  __device__ __forceinline__ void tick(const Drone<R>& d) {
    if (--left) return;
    left = stride;
    // streaming stores: the log is written once and never re-read by the kernel, keep it out of the L2 working set
    R* o = out;
    __stcs(o + 0 * B, (R)(d.px + (double)d.dx)); __stcs(o + 1 * B, (R)(d.py + (double)d.dy)); __stcs(o + 2 * B, (R)(d.pz + (double)d.dz));
    __stcs(o + 3 * B, d.q0); __stcs(o + 4 * B, d.q1); __stcs(o + 5 * B, d.q2); __stcs(o + 6 * B, d.q3);
    __stcs(o + 7 * B, d.vx); __stcs(o + 8 * B, d.vy); __stcs(o + 9 * B, d.vz);
    __stcs(o + 10 * B, d.wx); __stcs(o + 11 * B, d.wy); __stcs(o + 12 * B, d.wz);
    out += 13 * B;
  }
};

__device__ __forceinline__ float i2f(int x) { return __int_as_float(x); }
__device__ __forceinline__ int f2i(float x) { return __float_as_int(x); }

// Resumable carry block, [UAVB_CARRY_WORDS][B] 32-bit words (fp32 rollout only).
struct Carry {
  float* p;
  long long B;
  __device__ __forceinline__ float& w(int k) const { return p[(long long)k * B]; }
  __device__ __forceinline__ void put64(int k, double x) const { w(k) = i2f(__double2loint(x)); w(k + 1) = i2f(__double2hiint(x)); }
  __device__ __forceinline__ double get64(int k) const { return __hiloint2double(f2i(w(k + 1)), f2i(w(k))); }
  __device__ void store(const Drone<float>& d, const Cursor<float>& c, const Accum<float>& a, int tick) const {
    put64(0, d.px); put64(2, d.py); put64(4, d.pz);
    w(6) = d.q0; w(7) = d.q1; w(8) = d.q2; w(9) = d.q3;
    w(10) = d.vx; w(11) = d.vy; w(12) = d.vz; w(13) = d.wx; w(14) = d.wy; w(15) = d.wz;
    w(16) = d.om0; w(17) = d.om1; w(18) = d.om2; w(19) = d.om3;
    w(20) = d.integral; w(21) = d.thrust_cmd; w(22) = d.pc; w(23) = d.qc; w(24) = d.rc;
    w(25) = d.zbx; w(26) = d.zby; w(27) = d.zbz; w(28) = c.hx;
    w(29) = i2f(c.seg); w(30) = i2f(c.row); w(31) = i2f(c.phase);
    put64(32, c.tx); put64(34, c.ty); put64(36, c.tz);
    w(38) = a.sum_e; w(39) = a.sum_e2; w(40) = a.max_e;
    w(41) = i2f(a.periods); w(42) = i2f(a.collided); w(43) = i2f(a.first_hit); w(44) = i2f(a.status);
    w(45) = i2f(tick);
    w(46) = d.dx; w(47) = d.dy; w(48) = d.dz; w(49) = c.hy;
  }
  __device__ void load(Drone<float>& d, Cursor<float>& c, Accum<float>& a, const VehU<float>& u, int* tick) const {
    d.px = get64(0); d.py = get64(2); d.pz = get64(4);
    d.q0 = w(6); d.q1 = w(7); d.q2 = w(8); d.q3 = w(9);
    d.vx = w(10); d.vy = w(11); d.vz = w(12); d.wx = w(13); d.wy = w(14); d.wz = w(15);
    d.om0 = w(16); d.om1 = w(17); d.om2 = w(18); d.om3 = w(19);
    d.integral = w(20); set_thrust_cmd<float>(d, u, w(21)); d.pc = w(22); d.qc = w(23); d.rc = w(24);
    d.zbx = w(25); d.zby = w(26); d.zbz = w(27); c.hx = w(28);
    c.seg = f2i(w(29)); c.row = f2i(w(30)); c.phase = f2i(w(31)); c.cached_seg = -1;
    c.tx = get64(32); c.ty = get64(34); c.tz = get64(36);
    a.sum_e = w(38); a.sum_e2 = w(39); a.max_e = w(40);
    a.periods = f2i(w(41)); a.collided = f2i(w(42)); a.first_hit = f2i(w(43)); a.status = f2i(w(44));
    *tick = f2i(w(45));
    d.dx = w(46); d.dy = w(47); d.dz = w(48); c.hy = w(49);
  }
};

// One drone, one slice of its mission: `n_ticks` ticks starting from the carry block (from_carry) or from the start
// pose; writes the carry block (to_carry) and / or the final outputs (finish).
// MC: some per-rollout override (mass / inertia / gains / wind) is present; otherwise every vehicle constant is a
// constant-bank operand.
template <class R, bool LOG, bool MC, bool TABLE>
__device__ __forceinline__ void drone_slice(const RolloutDev<R>& p, const float* s_boxes, long long i, int n_ticks, bool from_carry,
                                            bool to_carry, bool finish) {
  const uavb_rollout_args& a = p.a;
  const long long B = a.B;
  const bool shared_boxes = a.n_obs > 0 && a.aabb_set == nullptr;

  // per-rollout constants
  VehP<R> vloc;
  if constexpr (MC) {
    McValues mc;
    mc_from_vehicle(mc, a.veh);
    if (a.mc_mass) mc.mass = (double)a.mc_mass[i];
    if (a.mc_inertia) { mc.inertia[0] = (double)a.mc_inertia[i]; mc.inertia[1] = (double)a.mc_inertia[B + i]; mc.inertia[2] = (double)a.mc_inertia[2 * B + i]; }
    if (a.mc_gains) {
#pragma unroll
      for (int k = 0; k < UAVB_N_GAINS; ++k) mc.gains[k] = (double)a.mc_gains[k * B + i];
    }
    if (a.mc_wind) { mc.wind[0] = (double)a.mc_wind[i]; mc.wind[1] = (double)a.mc_wind[B + i]; mc.wind[2] = (double)a.mc_wind[2 * B + i]; }
    make_vehp<R>(vloc, a.veh, mc);
  }
  const VehP<R>& v = MC ? vloc : p.vp;
  const VehU<R>& u = p.u;

  MissionView m;
  m.coeffs = a.seg_coeffs; m.rows = a.seg_rows; m.table = a.seg_table; m.yaw0 = a.seg_yaw0;
  m.seg_begin = a.mission_seg_begin ? a.mission_seg_begin[i] : 0;
  m.seg_count = a.mission_seg_count ? a.mission_seg_count[i] : a.n_seg_shared;
  m.dt_outer = a.dt_outer;
  m.trows = a.mission_seg_begin ? nullptr : static_cast<const TargetRow*>(a.shared_targets);
  m.n_trows = a.n_target_rows;
  m.cache = nullptr; m.cache_stride = 0;
  if (p.coeff_cache_offset >= 0) {             // on-the-fly evaluation: this thread's column of the CTA's coefficient staging area
    extern __shared__ double s_dyn_f64[];
    m.cache = s_dyn_f64 + p.coeff_cache_offset + threadIdx.x;
    m.cache_stride = kRolloutThreads;
  }

  Drone<R> d;
  Cursor<R> c;
  Accum<R> acc;
  int tick0 = 0;
  bool resumed = false;
  if constexpr (sizeof(R) == 4) {
    if (from_carry) {
      Carry cb{a.carry + i, B};
      cb.load(d, c, acc, u, &tick0);
      resumed = true;
    }
  }
  if (!resumed) {
    const double* s = a.start + (size_t)a.start_stride * i;
    drone_init<R>(d, u, s[0], s[1], s[2]);
    cursor_init<R>(c);
    accum_init<R>(acc);
  }

  auto fly = [&](const auto& obst) {
    if constexpr (LOG) {
      GlobalLog<R> lg;
      lg.out = reinterpret_cast<R*>(a.log_out) + i;
      lg.B = B; lg.stride = a.log_stride; lg.left = a.log_stride;
      rollout_run<R, TABLE>(d, c, acc, u, v, m, tick0, n_ticks, a.inner_per_outer, a.thrust_frame_lag, obst, lg);
    } else {
      NoLog lg;
      rollout_run<R, TABLE>(d, c, acc, u, v, m, tick0, n_ticks, a.inner_per_outer, a.thrust_frame_lag, obst, lg);
    }
  };
  if (a.n_obs > 0) {
    if (shared_boxes) {
      fly(BoxesT<true>{nullptr, a.n_obs});
    } else {
      fly(BoxesT<false>{a.aabbs + (size_t)a.aabb_set[i] * a.n_obs * 6, a.n_obs});
    }
  } else {
    fly(NoObstacles{});
  }

  if constexpr (sizeof(R) == 4) {
    if (to_carry) {
      Carry cb{a.carry + i, B};
      cb.store(d, c, acc, tick0 + n_ticks);
    }
  }
  if (!finish) return;
  const double fx = d.px + (double)d.dx, fy = d.py + (double)d.dy, fz = d.pz + (double)d.dz;
  if (a.state_out) {
    R* o = reinterpret_cast<R*>(a.state_out) + i;
    o[0 * B] = (R)fx; o[1 * B] = (R)fy; o[2 * B] = (R)fz;
    if (!LOG && c.phase != 0) renormalise_q<R>(d);           // the reported state is unit even in the middle of an outer period
    o[3 * B] = d.q0; o[4 * B] = d.q1; o[5 * B] = d.q2; o[6 * B] = d.q3;
    o[7 * B] = d.vx; o[8 * B] = d.vy; o[9 * B] = d.vz;
    o[10 * B] = d.wx; o[11 * B] = d.wy; o[12 * B] = d.wz;
  }
  if (a.metrics_out) {
    R fd = R(0);
    if (a.goal) {
      const double* g = a.goal + (size_t)a.goal_stride * i;
      const R ex = (R)(g[0] - fx), ey = (R)(g[1] - fy), ez = (R)(g[2] - fz);
      fd = Math<R>::sqrt(ex * ex + ey * ey + ez * ez);
    }
    const R np = acc.periods > 0 ? R(1) / (R)acc.periods : R(0);
    R* o = reinterpret_cast<R*>(a.metrics_out) + (size_t)i * UAVB_N_METRICS;
    o[UAVB_M_FINAL_DIST] = fd;
    o[UAVB_M_COLLISION] = (R)acc.collided;
    o[UAVB_M_RMSE] = Math<R>::sqrt(acc.sum_e2 * np);
    o[UAVB_M_MEAN_ERR] = acc.sum_e * np;
    o[UAVB_M_MAX_ERR] = acc.max_e;
    o[UAVB_M_STATUS] = (R)acc.status;
    o[UAVB_M_FIRST_HIT] = (R)acc.first_hit;
    o[UAVB_M_PERIODS] = (R)acc.periods;
  }
}

__device__ __forceinline__ void stage_shared_boxes(const uavb_rollout_args& a, float* s_boxes) {
  if (a.n_obs > 0 && a.aabb_set == nullptr) {
    for (int j = threadIdx.x; j < a.n_obs * 6; j += blockDim.x) s_boxes[j] = a.aabbs[j];
    __syncthreads();
  }
}

// One-shot launch: thread i flies drone i for the whole launch.  K = residency variant (CTAs per SM).
template <class R, bool LOG, bool MC, int K>
__global__ void __maxnreg__(rollout_regs(K)) rollout_kernel(const __grid_constant__ RolloutDev<R> p) {
  extern __shared__ float s_boxes[];
  stage_shared_boxes(p.a, s_boxes);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.a.B) return;
  if (p.a.shared_targets != nullptr && p.a.mission_seg_begin == nullptr)
    drone_slice<R, LOG, MC, true>(p, s_boxes, i, p.a.n_ticks, p.a.resume != 0, p.a.carry != nullptr, true);
  else
    drone_slice<R, LOG, MC, false>(p, s_boxes, i, p.a.n_ticks, p.a.resume != 0, p.a.carry != nullptr, true);
}

// Time-sliced persistent launch (every metrics-only fp32 rollout).  When the batch needs between one and a few waves of
// CTAs, a one-shot launch ends with a long tail: every CTA lives for the whole mission, so the last partial wave costs a full
// mission time at a fraction of the machine.  Here the grid is exactly the resident capacity, the mission is cut into
// `n_chunks` slices of `chunk_ticks` ticks, and CTAs pull (chunk, group) items from an atomic counter in chunk-major
// order; between slices a drone's state rests in the carry block (208 B per drone and slice, ~0.2 B per tick).  Item
// (c, g) needs (c-1, g), which was handed out one full sweep of the groups earlier, so the wait on its completion flag
// practically never spins -- and cannot deadlock, because whoever holds the earlier item is running.  The tail shrinks
// from one mission to one slice.
struct SliceSched {
  int* counter;        // next item
  int* done;           // [n_groups] slices completed per group
  int n_groups, n_chunks, chunk_ticks;
  int final_carry;     // the caller asked for the carry block of the end state
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <bool MC, bool TABLE, int K>
__global__ void __maxnreg__(rollout_regs(K)) rollout_sliced_kernel(const __grid_constant__ RolloutDev<float> p, const SliceSched sch) {
  extern __shared__ float s_boxes[];
  __shared__ int s_item;
  stage_shared_boxes(p.a, s_boxes);
  const int n_items = sch.n_groups * sch.n_chunks;
  for (;;) {
    if (threadIdx.x == 0) {
      const int it = atomicAdd(sch.counter, 1);
      if (it < n_items) {
        const int c = it / sch.n_groups, g = it - c * sch.n_groups;
        while (ld_acquire_gpu(sch.done + g) < c) __nanosleep(200);
      }
      s_item = it;
    }
    __syncthreads();
    const int it = s_item;
    if (it >= n_items) return;
    const int c = it / sch.n_groups, g = it - c * sch.n_groups;
    const long long i = (long long)g * blockDim.x + threadIdx.x;
    if (i < p.a.B) {
      const bool last = c == sch.n_chunks - 1;
      const int ticks = last ? p.a.n_ticks - c * sch.chunk_ticks : sch.chunk_ticks;
      drone_slice<float, false, MC, TABLE>(p, s_boxes, i, ticks, c > 0 || p.a.resume != 0, !last || sch.final_carry != 0, last);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(sch.done + g, c + 1);
  }
}

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
  UAVB_REQUIRE(a->log_stride == 0 || a->log_out != nullptr, "rollout: log_stride > 0 needs log_out");
  UAVB_REQUIRE(a->n_obs >= 0 && a->n_obs <= 1024, "rollout: n_obs out of range");
  UAVB_REQUIRE(a->n_obs == 0 || (a->aabbs != nullptr && a->n_obs_sets >= 1), "rollout: n_obs > 0 needs aabbs and n_obs_sets >= 1");
  UAVB_REQUIRE(a->dt_outer > 0.0 && a->veh.dt > 0.0 && a->veh.mass > 0.0, "rollout: dt_outer, veh.dt and veh.mass must be positive");
  UAVB_REQUIRE(a->shared_targets == nullptr || (a->mission_seg_begin == nullptr && a->n_target_rows >= 1),
               "rollout: shared_targets needs a shared mission and n_target_rows >= 1");
  return UAVB_OK;
}

// SM count of the current device (one query per device and process).
static int sm_count_cached(int* sms) {
  static int cache[64] = {0};
  int dev = 0;
  UAVB_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || cache[dev] == 0) {
    int n = 0;
    UAVB_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < 64) cache[dev] = n;
    *sms = n;
    return UAVB_OK;
  }
  *sms = cache[dev];
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

template <bool MC, bool TABLE> static void launch_sliced(int k, int grid, size_t smem, cudaStream_t st, const RolloutDev<float>& p, const SliceSched& sch) {
  switch (k) {
    case 8: rollout_sliced_kernel<MC, TABLE, 8><<<grid, kRolloutThreads, smem, st>>>(p, sch); break;
    default: rollout_sliced_kernel<MC, TABLE, 12><<<grid, kRolloutThreads, smem, st>>>(p, sch); break;
  }
}

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
  const int grid = div_up(a->B, kRolloutThreads);
  size_t smem = (a->n_obs > 0 && a->aabb_set == nullptr) ? sizeof(float) * 6 * a->n_obs : 0;
  p.coeff_cache_offset = -1;
  const bool from_table = a->shared_targets != nullptr && a->mission_seg_begin == nullptr;
  if (!from_table && sizeof(R) == 4 && a->log_stride == 0) {     // metrics-only fp32 kernels stage the current spline in shared memory
    smem = (smem + 7) / 8 * 8;
    p.coeff_cache_offset = (int)(smem / 8);
    smem += sizeof(double) * 24 * kRolloutThreads;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool mc_any = a->mc_mass || a->mc_inertia || a->mc_gains || a->mc_wind;
  const bool log = a->log_stride > 0;
  if constexpr (sizeof(R) == 8) {                      // validation build: one residency variant
    if (log && mc_any) rollout_kernel<R, true, true, 8><<<grid, kRolloutThreads, smem, st>>>(p);
    else if (log) rollout_kernel<R, true, false, 8><<<grid, kRolloutThreads, smem, st>>>(p);
    else if (mc_any) rollout_kernel<R, false, true, 8><<<grid, kRolloutThreads, smem, st>>>(p);
    else rollout_kernel<R, false, false, 8><<<grid, kRolloutThreads, smem, st>>>(p);
  } else if (log) {
    if (mc_any) rollout_kernel<R, true, true, 8><<<grid, kRolloutThreads, smem, st>>>(p);
    else rollout_kernel<R, true, false, 8><<<grid, kRolloutThreads, smem, st>>>(p);
  } else {
    int sms = 0;
    rc = sm_count_cached(&sms);
    if (rc) return rc;
    // Metrics-only fp32 launches ALWAYS run the time-sliced persistent kernel at K = 8 (128 registers, no spills: the
    // highest sustained tick rate), with a single slice when slicing has nothing to gain.  One compiled body for every
    // batch size keeps per-rollout results independent of how a job is sharded (ptxas fuses mul+add differently under
    // different register caps, so the residency variants are NOT bit-identical to each other).
    int k = 8;
    if (const char* force = getenv("UAVB_ROLLOUT_K")) {                  // development override for residency experiments
      const int f = atoi(force);
      if (f >= kRolloutCtasMin && f <= kRolloutCtasMax) k = f;
    }
    if (k != 8) k = 12;                                                  // compiled residencies: 8 (production) and 12 (experiments)
    const int slots = sms * k;
    // ~32 items per resident CTA keep the tail near 3 % of the launch; slices are whole outer periods of >= 100 ticks
    constexpr int kMinChunkTicks = 100;
    long long want = grid > slots ? (32LL * slots + grid - 1) / grid : 1;
    if (const char* force = getenv("UAVB_ROLLOUT_CHUNKS")) want = atoi(force) > 0 ? atoi(force) : want;   // development override
    const int period = a->inner_per_outer;
    int chunk = (int)((a->n_ticks + want - 1) / want);
    chunk = ((chunk + period - 1) / period) * period;
    if (chunk < kMinChunkTicks) chunk = ((kMinChunkTicks + period - 1) / period) * period;
    const int n_chunks = a->n_ticks > 0 ? (a->n_ticks + chunk - 1) / chunk : 1;
    StreamScratch flags(st), carry(st);
    const size_t n_flags = (size_t)grid + 1;
    cudaError_t e = flags.alloc(n_flags * sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(flags.p, 0, n_flags * sizeof(int), st);
    SliceSched sch;
    sch.final_carry = a->carry != nullptr;
    if (e == cudaSuccess && a->carry == nullptr && n_chunks > 1) {
      e = carry.alloc(sizeof(float) * UAVB_CARRY_WORDS * (size_t)a->B);
      p.a.carry = static_cast<float*>(carry.p);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      return set_error(UAVB_ENOMEM, "rollout: scratch allocation failed: %s", cudaGetErrorString(e));
    }
    sch.counter = static_cast<int*>(flags.p);
    sch.done = sch.counter + 1;
    sch.n_groups = grid; sch.n_chunks = n_chunks; sch.chunk_ticks = chunk;
    const int pgrid = slots < grid ? slots : grid;
    const RolloutDev<float>& pf = *reinterpret_cast<RolloutDev<float>*>(&p);
    if (from_table) {
      if (mc_any) launch_sliced<true, true>(k, pgrid, smem, st, pf, sch);
      else launch_sliced<false, true>(k, pgrid, smem, st, pf, sch);
    } else {
      if (mc_any) launch_sliced<true, false>(k, pgrid, smem, st, pf, sch);
      else launch_sliced<false, false>(k, pgrid, smem, st, pf, sch);
    }
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

// rollout_core.cuh -- the persistent per-drone mission loop of K2: table cursor, outer/inner
// schedule, tracking metrics, collision flag.  Used by rollout_kernels.cu (one thread per drone).
//
// Schedule reproduced (paths relative to /root/reference):
//   TrajectoryController.step (uav_ac/main.py:37-45): outer loop iff inner_step % frequency == 0,
//   using table row `trajectory_index`, then index = min(index+1, N-1) (:61); body-rate loop and
//   set_propeller_speed every tick; then MujocoSimulation.step (mujoco_sim.py:144-151).
//   Tracking error |p - target[:3]| after each outer period (tests/integration/
//   test_mujoco_trajectory_tracking.py:27-31); final distance to goal (main.py:115).
//   Collision flag: inclusive point-in-AABB of minimum_snap.py:327-357 on the body origin after
//   every tick, sticky (BASELINE.json north_star; replaces MuJoCo contacts, mujoco_sim.py:220-230).
#pragma once

#include "flight_core.cuh"

namespace uavb {

// Packed mission segments, reference coefficient layout (MinimumSnap.coeffs rows 8*i+j, 3 axes).
struct MissionView {
  const double* coeffs;   // [n_seg][24]
  const int* rows;        // [n_seg] rows of each segment = len(np.arange(0, T_i, dt)) (minimum_snap.py:104)
  const int* table;       // [n_seg] 1 = first segment of a MinimumSnap table
  const double* yaw0;     // [n_seg] look-ahead yaw of the table starting here (minimum_snap.py:134-135)
  int seg_begin, seg_count;
  double dt_outer;
};

template <class R> struct Cursor {
  int seg, row;           // table row the NEXT outer update will use (main.py:48 trajectory_index)
  int phase;              // inner_step % frequency (main.py:25,39), kept incrementally
  R yaw_hold;             // yaw of the last valid row (minimum_snap.py:126-136 hold-last-valid)
  double tx, ty, tz;      // position set-point of the row used by the current outer period
};

template <class R> struct Accum {
  R sum_e, sum_e2, max_e;
  int periods;
  int collided, first_hit;
  int status;
};

struct NoLog {
  template <class R> UAVB_HD void tick(const Drone<R>&) {}
};

struct NoObstacles {
  template <class R> UAVB_HD bool hit(R, R, R) const { return false; }
};

// Advance the cursor one row with the end clamp of main.py:61.
UAVB_HD void cursor_advance(int* seg, int* row, const MissionView& m) {
  if (*row + 1 < m.rows[m.seg_begin + *seg]) { ++*row; return; }
  int s = *seg + 1;
  while (s < m.seg_count && m.rows[m.seg_begin + s] <= 0) ++s;   // zero-length segments own no rows
  if (s < m.seg_count) { *seg = s; *row = 0; }
}

// Fetch the set-point of the cursor row: polynomial values plus the yaw rule of _calculate_yaws.
template <class R> UAVB_HD void cursor_target(Cursor<R>& c, const MissionView& m, Target* t) {
  const int sg = m.seg_begin + c.seg;
  const double* cf = m.coeffs + (size_t)sg * 24;
#if defined(__CUDA_ARCH__)
  auto ld = [cf](int i) { return __ldg(cf + i); };
#else
  auto ld = [cf](int i) { return cf[i]; };
#endif
  eval_row(ld, (double)c.row * m.dt_outer, t);
  if (c.row == 0 && m.table[sg]) c.yaw_hold = (R)m.yaw0[sg];
  if (sqrt(t->vx * t->vx + t->vy * t->vy) >= 1e-3) c.yaw_hold = Math<R>::atan2((R)t->vy, (R)t->vx);
  t->yaw = (double)c.yaw_hold;
}

// n_ticks ticks of the closed loop for one drone.  `tick0` is the global index of the first tick
// (first_hit and the log phase are relative to the start of the mission, not of this launch).
// The tick loop is split at outer-period boundaries so the 1 kHz body stays branch-light.
template <class R, class OBST, class LOG>
UAVB_HD void rollout_run(Drone<R>& d, Cursor<R>& c, Accum<R>& a, const Veh<R>& v, const MissionView& m, int tick0,
                         int n_ticks, int freq, int lag, const OBST& obst, LOG& logger) {
  typedef Math<R> M;
  int k = 0;
  while (k < n_ticks) {
    if (c.phase == 0) {
      Target t;
      cursor_target<R>(c, m, &t);
      outer_update<R>(d, v, t);
      c.tx = t.x; c.ty = t.y; c.tz = t.z;
      cursor_advance(&c.seg, &c.row, m);
    }
    const int n = (freq - c.phase < n_ticks - k) ? (freq - c.phase) : (n_ticks - k);
    for (int j = 0; j < n; ++j) {
      inner_tick<R>(d, v, lag);
      if (!a.collided && obst.hit(d.px, d.py, d.pz)) { a.collided = 1; a.first_hit = tick0 + k + j; }
      logger.tick(d);
    }
    k += n;
    c.phase += n;
    if (c.phase == freq) {
      c.phase = 0;
      const R ex = pos_err<R>(c.tx, d.px, d.plx), ey = pos_err<R>(c.ty, d.py, d.ply), ez = pos_err<R>(c.tz, d.pz, d.plz);
      const R e2 = ex * ex + ey * ey + ez * ez;
      const R e = M::sqrt(e2);
      a.sum_e += e; a.sum_e2 += e2; a.max_e = M::fmax(a.max_e, e);
      ++a.periods;
      if (!M::finite(e2)) a.status |= 1;
      else if (d.px * d.px + d.py * d.py + d.pz * d.pz > R(1e8)) a.status |= 2;
    }
  }
}

template <class R> UAVB_HD void drone_init(Drone<R>& d, double sx, double sy, double sz) {
  d.px = (R)sx; d.py = (R)sy; d.pz = (R)sz;
  d.plx = (R)(sx - (double)d.px); d.ply = (R)(sy - (double)d.py); d.plz = (R)(sz - (double)d.pz);
  d.q0 = R(1); d.q1 = d.q2 = d.q3 = R(0);            // quad.py:78-80
  d.vx = d.vy = d.vz = R(0);
  d.wx = d.wy = d.wz = R(0);
  d.om0 = d.om1 = d.om2 = d.om3 = R(0);              // quad.py:85
  d.integral = R(0);                                 // controller.py:20
  d.thrust_cmd = R(0); d.pc = d.qc = d.rc = R(0);    // main.py:26-27
  body_z<R>(d, &d.zbx, &d.zby, &d.zbz);              // mj_forward in MujocoSimulation.__init__ (mujoco_sim.py:81)
}

template <class R> UAVB_HD void cursor_init(Cursor<R>& c) {
  c.seg = 0; c.row = 0; c.phase = 0; c.yaw_hold = R(0); c.tx = c.ty = c.tz = 0.0;
}

template <class R> UAVB_HD void accum_init(Accum<R>& a) {
  a.sum_e = a.sum_e2 = a.max_e = R(0);
  a.periods = 0; a.collided = 0; a.first_hit = -1; a.status = 0;
}

}  // namespace uavb

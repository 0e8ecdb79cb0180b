// veh_setup.cuh -- turn the ABI vehicle description (+ optional per-rollout Monte-Carlo overrides)
// into the constant sets of flight_core.cuh: VehU (uniform across a launch, passed in the kernel
// parameter block) and VehP (per rollout).  Derived constants are formed in fp64 and rounded once.
// Reference: Quad.__init__ (uav_ac/quadrotor/quad.py:36-73), motor lag response
// 1 - exp(-dt/tau) (quad.py:102), lab_course.xml:3,8-13,100,116-119.
#pragma once

#include "flight_core.cuh"
#include "../../include/uavb.h"

namespace uavb {

// Values that may be overridden per rollout (BASELINE configs[2..3]); wind is a force in N.
struct McValues {
  double mass;
  double inertia[3];
  double gains[UAVB_N_GAINS];
  double wind[3];
};

UAVB_HD void vehicle_defaults(uavb_vehicle* v) {
  v->g = 9.81; v->dt = 0.001;                       // lab_course.xml:3
  v->mass = 0.5; v->inertia[0] = 0.0023; v->inertia[1] = 0.0023; v->inertia[2] = 0.0046;   // :100
  v->arm = 0.120208; v->kf = 1.0; v->kappa = 0.016;  // :116-119, :9-10
  v->min_thrust = 0.1; v->max_thrust = 4.5;         // :11
  v->tau_rise = 0.0125; v->tau_fall = 0.025;        // :12
  v->max_ascent = 3.0; v->max_descent = 2.0; v->max_speed_xy = 3.0; v->max_horiz_accel = 12.0; v->max_tilt = 0.7;  // :13
  // quad.py:54-73: second_order_gains(tau, zeta) = (1/tau^2, 2 zeta/tau)
  v->gains[0] = 1.0 / (0.25 * 0.25); v->gains[1] = 2.0 * 0.875 / 0.25;     // kp_xy kd_xy
  v->gains[2] = 1.0 / (0.2 * 0.2);   v->gains[3] = 2.0 * 0.8 / 0.2;        // kp_z kd_z
  v->gains[4] = 0.1;                                                       // ki_z
  v->gains[5] = 1.0 / 0.07; v->gains[6] = 1.0 / 0.07; v->gains[7] = 1.0 / 0.25;   // kp_roll kp_pitch kp_yaw
  v->gains[8] = 1.0 / 0.008; v->gains[9] = 1.0 / 0.008; v->gains[10] = 1.0 / 0.09; // kp_p kp_q kp_r
  v->integral_limit = 10.0;                         // controller.py:10
}

UAVB_HD void mc_from_vehicle(McValues& o, const uavb_vehicle& u) {
  o.mass = u.mass;
  for (int i = 0; i < 3; ++i) { o.inertia[i] = u.inertia[i]; o.wind[i] = 0.0; }
  for (int i = 0; i < UAVB_N_GAINS; ++i) o.gains[i] = u.gains[i];
}

// Launch-uniform constants; built on the host (exp() for the motor-lag responses of quad.py:102).
template <class R> inline void make_vehu(VehU<R>& v, const uavb_vehicle& u, double dt_outer) {
  v.dt = (R)u.dt; v.half_dt = (R)(0.5 * u.dt); v.half_dt_sq = v.half_dt * v.half_dt; v.dt_outer = (R)dt_outer; v.g = (R)u.g;
  v.half_dt_cu3 = (R)(0.125 * u.dt * u.dt * u.dt / 3.0); v.small_rot_wn2 = (R)(5e-4 / (0.25 * u.dt * u.dt));
  v.kf = (R)u.kf; v.inv_kf = (R)(1.0 / u.kf); v.arm_kf = (R)(u.arm * u.kf); v.kappa_kf = (R)(u.kappa * u.kf);
  v.inv_arm4 = (R)(0.25 / u.arm); v.inv_kappa4 = (R)(0.25 / u.kappa);
  v.fmin = (R)u.min_thrust; v.fmax = (R)u.max_thrust; v.fmin4 = (R)(4.0 * u.min_thrust); v.fmax4 = (R)(4.0 * u.max_thrust);
  v.w2min = (R)(u.min_thrust / u.kf); v.w2max = (R)(u.max_thrust / u.kf); v.quarter_inv_kf = (R)(0.25 / u.kf);
  const double ar = 1.0 - exp(-u.dt / u.tau_rise), af = 1.0 - exp(-u.dt / u.tau_fall);
  v.a_rise = (R)ar; v.a_fall = (R)af; v.a_mean = (R)(0.5 * (ar + af)); v.a_hdiff = (R)(0.5 * (ar - af));
  v.max_ascent = (R)u.max_ascent; v.max_descent = (R)u.max_descent; v.max_speed_xy = (R)u.max_speed_xy;
  v.max_acc_xy = (R)u.max_horiz_accel; v.max_tilt = (R)u.max_tilt; v.integral_limit = (R)u.integral_limit;
}

// Per-rollout constants from the (possibly perturbed) mass, inertia, gains and wind.
template <class R> UAVB_HD void make_vehp(VehP<R>& v, const uavb_vehicle& u, const McValues& o) {
  const double m = o.mass, Ix = o.inertia[0], Iy = o.inertia[1], Iz = o.inertia[2];
  v.kf_dt_over_m = (R)(u.kf * u.dt / m); v.kf_dt_over_m2 = (R)(2.0 * u.kf * u.dt / m);
  v.dIx = (R)(Iz - Iy); v.dIy = (R)(Ix - Iz); v.dIz = (R)(Iy - Ix);
  v.Ikp_p = (R)(Ix * o.gains[8]); v.Ikp_q = (R)(Iy * o.gains[9]); v.Ikp_r = (R)(Iz * o.gains[10]);
  v.dt_invIx = (R)(u.dt / Ix); v.dt_invIy = (R)(u.dt / Iy); v.dt_invIz = (R)(u.dt / Iz);
  const double ia = 1.0 / (4.0 * u.arm * u.kf), ik = 1.0 / (4.0 * u.kappa * u.kf);   // rotor units, see inner_tick
  v.Gx = (R)((Iz - Iy) * ia); v.Gy = (R)((Ix - Iz) * ia); v.Gz = (R)(-(Iy - Ix) * ik);
  v.Jp = (R)(Ix * o.gains[8] * ia); v.Jq = (R)(Iy * o.gains[9] * ia); v.Jr = (R)(-Iz * o.gains[10] * ik);
  v.Wx = (R)(u.dt * u.arm * u.kf / Ix); v.Wy = (R)(u.dt * u.arm * u.kf / Iy); v.Wz = (R)(u.dt * u.kappa * u.kf / Iz);
  v.Kx = (R)(-u.dt * (Iz - Iy) / Ix); v.Ky = (R)(-u.dt * (Ix - Iz) / Iy); v.Kz = (R)(-u.dt * (Iy - Ix) / Iz);
  v.dvx = (R)(u.dt * o.wind[0] / m); v.dvy = (R)(u.dt * o.wind[1] / m); v.dvz = (R)(u.dt * (o.wind[2] / m + u.g));
  v.mass = (R)m;
  v.kp_xy = (R)o.gains[0]; v.kd_xy = (R)o.gains[1]; v.kp_z = (R)o.gains[2]; v.kd_z = (R)o.gains[3]; v.ki_z = (R)o.gains[4];
  v.kp_roll = (R)o.gains[5]; v.kp_pitch = (R)o.gains[6]; v.kp_yaw = (R)o.gains[7];
  const double wn = sqrt(o.wind[0] * o.wind[0] + o.wind[1] * o.wind[1] + o.wind[2] * o.wind[2]);
  v.acc_max = (R)((4.0 * u.max_thrust + wn) / m + u.g);
}

}  // namespace uavb

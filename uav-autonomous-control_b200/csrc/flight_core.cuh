// flight_core.cuh -- per-drone closed-loop arithmetic shared by the rollout kernel (K2) and the
// stage kernels.  One drone per thread; everything here is straight-line register code.
//
// Reference semantics (paths relative to /root/reference):
//   outer loop   uav_ac/main.py:47-61  -> controller.py:26-56 (altitude), :58-97 (lateral),
//                :132-154 (roll/pitch), :156-168 (yaw), quad.py:129-155 (R), :189-213 (Euler angles)
//   inner loop   uav_ac/main.py:42-45  -> controller.py:115-130 (body rates),
//                quad.py:105-122 (allocation), :88-103 (motor lag)
//   physics      uav_ac/simulation/mujoco_sim.py:232-251 (rotor wrench) + MuJoCo free-joint
//                semi-implicit Euler step (restated)
//
// The type parameter R is the state/arithmetic type: float for the production rollout, double for
// the validation build.  In float mode three things keep the closed loop within 1e-4 m of an fp64
// run over 16k ticks (DESIGN.md "fp32 budget"):
//   * position is carried in fp64 and advanced once per outer period; inside the period the 1 kHz
//     ticks accumulate the displacement since the last fold in an fp32 register triple (|d| < 0.1 m,
//     so its rounding is ~1e-9 m per tick instead of 1e-6 m at |p| ~ 20 m);
//   * set-points come from fp64 Horner evaluation and position errors are formed in fp64;
//   * the quaternion is advanced by adding q*(dq-1), so only the final add rounds at 1 ulp of q.
//
// Instruction diet of the 1 kHz body (DESIGN.md "K2 optimisation log"): per-launch constants live in
// the kernel parameter block (constant-bank operands, no registers), products of constants are
// folded on the host (kf dt/m, dt/I, arm kf ...), the gyroscopic term uses the diagonal-inertia
// identity w x (I w) = ((Iz-Iy) wy wz, (Ix-Iz) wz wx, (Iy-Ix) wx wy), the allocation takes a
// division-free path whenever no rotor limit binds, sqrt/rcp/rsqrt are single MUFU instructions, and the
// persistent rollout's tick (inner_tick) works in "rotor units" -- moments / (4 arm kf), forces / kf -- so the
// mixer, the limits and the rotor speed commands need no scaling; the SI forms of the same stages serve the
// stage kernels (one controller / vehicle method per launch).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define UAVB_HD __host__ __device__ __forceinline__
#else
#define UAVB_HD inline
#endif

namespace uavb {

constexpr double kTwoPi = 6.283185307179586476925286766559;
constexpr double kPi = 3.141592653589793238462643383279;

template <class R> struct Math;
template <> struct Math<float> {
  static UAVB_HD float sqrt(float x) { return sqrtf(x); }
  // single-MUFU forms (max relative error 2^-23 .. 2^-22, PTX ISA "sqrt.approx / rcp.approx / rsqrt.approx")
  static UAVB_HD float sqrt_fast(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return sqrtf(x);
#endif
  }
  static UAVB_HD float rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / x;
#endif
  }
  static UAVB_HD float rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / sqrtf(x);
#endif
  }
  static UAVB_HD float div(float a, float b) { return a * rcp_fast(b); }
  static UAVB_HD float fma(float a, float b, float c) { return fmaf(a, b, c); }
  // atan2 without a division: a = min(|x|,|y|) / max(|x|,|y|) through the reciprocal above, atan(a) = a P(a^2) with an
  // equi-ripple degree-7 P on [0, 1] (max error 3.7e-8 rad before rounding; 3.3e-7 rad in fp32 over all quadrants, the same
  // as atan2f, whose error is also set by the rounding of results near pi), then the octant / quadrant reflections.
  // atan2(0, 0) = 0 like atan2f(+0, +0).
  static UAVB_HD float atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = (mx > 0.f) ? mn * rcp_fast(mx) : 0.f;
    const float t = a * a;
    float p = -0.0040545563519447094f;
    p = fmaf(p, t, 0.02186292376737154f);
    p = fmaf(p, t, -0.055912287992173626f);
    p = fmaf(p, t, 0.09642195584271772f);
    p = fmaf(p, t, -0.13908629508211973f);
    p = fmaf(p, t, 0.19946565845760894f);
    p = fmaf(p, t, -0.33329860832632324f);
    p = fmaf(p, t, 0.9999993356075512f);
    float r = a * p;
    r = (ay > ax) ? 1.57079632679489662f - r : r;
    r = (x < 0.f) ? 3.14159265358979324f - r : r;
    return copysignf(r, y);
  }
  static UAVB_HD void sincos(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    sincosf(x, s, c);
#else
    *s = sinf(x); *c = cosf(x);
#endif
  }
  static UAVB_HD void sincos_fast(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    *s = __sinf(x); *c = __cosf(x);
#else
    *s = sinf(x); *c = cosf(x);
#endif
  }
  static UAVB_HD float floor(float x) { return floorf(x); }
  static UAVB_HD float abs(float x) { return fabsf(x); }
  static UAVB_HD float fmin(float a, float b) { return fminf(a, b); }
  static UAVB_HD float fmax(float a, float b) { return fmaxf(a, b); }
  static UAVB_HD bool finite(float x) { return fabsf(x) <= 3.0e38f; }
};
template <> struct Math<double> {
  static UAVB_HD double sqrt(double x) { return ::sqrt(x); }
  static UAVB_HD double sqrt_fast(double x) { return ::sqrt(x); }
  static UAVB_HD double rcp_fast(double x) { return 1.0 / x; }
  static UAVB_HD double rsqrt(double x) { return 1.0 / ::sqrt(x); }
  static UAVB_HD double div(double a, double b) { return a / b; }
  static UAVB_HD double fma(double a, double b, double c) { return ::fma(a, b, c); }
  static UAVB_HD double atan2(double y, double x) { return ::atan2(y, x); }
  static UAVB_HD void sincos(double x, double* s, double* c) { *s = ::sin(x); *c = ::cos(x); }
  static UAVB_HD void sincos_fast(double x, double* s, double* c) { *s = ::sin(x); *c = ::cos(x); }
  static UAVB_HD double floor(double x) { return ::floor(x); }
  static UAVB_HD double abs(double x) { return ::fabs(x); }
  static UAVB_HD double fmin(double a, double b) { return ::fmin(a, b); }
  static UAVB_HD double fmax(double a, double b) { return ::fmax(a, b); }
  static UAVB_HD bool finite(double x) { return ::fabs(x) <= 1.0e300; }
};

template <class R> UAVB_HD R clampr(R x, R lo, R hi) { return Math<R>::fmin(Math<R>::fmax(x, lo), hi); }

// Constants that are uniform across a launch.  The rollout kernel receives this struct by value in
// its parameter block, so every field is a constant-bank operand.
template <class R> struct VehU {
  R dt, half_dt, half_dt_sq, dt_outer, g;
  R half_dt_cu3, small_rot_wn2;                    // (dt/2)^3 / 3 and the |w|^2 below which the pair tick's two-term tan series holds (flight_pair.cuh)
  R kf, inv_kf, arm_kf, kappa_kf;                  // arm*kf, kappa*kf: torque per unit of summed w^2
  R inv_arm4, inv_kappa4;                          // 1/(4 arm), 1/(4 kappa): mixer division by 4 folded in (quad.py:112)
  R fmin, fmax, fmin4, fmax4, a_rise, a_fall;      // a_* = 1 - exp(-dt/tau) (quad.py:102)
  R a_mean, a_hdiff;                               // (a_rise + a_fall)/2, (a_rise - a_fall)/2
  R w2min, w2max, quarter_inv_kf;                  // rotor units (see "rotor units" below): fmin/kf, fmax/kf, 1/(4 kf)
  R max_ascent, max_descent, max_speed_xy, max_acc_xy, max_tilt, integral_limit;
};

// Constants that differ between rollouts under Monte-Carlo perturbation (uniform otherwise).
template <class R> struct VehP {
  // 1 kHz body
  R kf_dt_over_m, kf_dt_over_m2;                   // kf dt / m: velocity gained per unit of summed w^2; twice that
  R dIx, dIy, dIz;                                 // Iz-Iy, Ix-Iz, Iy-Ix (gyroscopic term, diagonal inertia)
  R Ikp_p, Ikp_q, Ikp_r;                           // I * kp of the body-rate loop (controller.py:128)
  R dt_invIx, dt_invIy, dt_invIz;                  // dt / I
  // the same three groups in rotor units (the persistent rollout's 1 kHz body; the SI ones serve the stage kernels)
  R Gx, Gy, Gz;                                    // dIx/(4 arm kf), dIy/(4 arm kf), -dIz/(4 kappa kf)
  R Jp, Jq, Jr;                                    // Ikp_p/(4 arm kf), Ikp_q/(4 arm kf), -Ikp_r/(4 kappa kf)
  R Wx, Wy, Wz;                                    // dt arm kf/Ix, dt arm kf/Iy, dt kappa kf/Iz
  R Kx, Ky, Kz;                                    // -dt dIx/Ix, -dt dIy/Iy, -dt dIz/Iz: rate gained per unit of wy wz, wz wx, wx wy
  R dvx, dvy, dvz;                                 // dt * (wind/m + g e_z): velocity gained per tick without thrust
  // 100 Hz outer loop
  R mass, kp_xy, kd_xy, kp_z, kd_z, ki_z, kp_roll, kp_pitch, kp_yaw;
  R acc_max;                                       // bound on |acceleration| (4 fmax/m + g + |wind|/m): obstacle culling
};

// Persistent per-drone state (registers across the whole rollout).
template <class R> struct Drone {
  double px, py, pz;   // position at the last fold (start of the current outer period)
  R dx, dy, dz;        // displacement accumulated since the last fold
  R q0, q1, q2, q3;    // attitude, scalar first, FRD->NED
  R vx, vy, vz;        // world velocity
  R wx, wy, wz;        // body rates p q r
  R om0, om1, om2, om3;  // rotor speeds (quad.py:85)
  R integral;          // altitude integrator (controller.py:20)
  R thrust_cmd;        // main.py:26
  R coll;              // clip(thrust_cmd, 4 fmin, 4 fmax) / (4 kf) (quad.py:107,113) in rotor units, refreshed with thrust_cmd
  R pc, qc, rc;        // pqr_cmd (main.py:27)
  R cp, cq, cr;        // the same in rotor units (Jp pc, Jq qc, Jr rc): what the persistent rollout carries between ticks
  R zbx, zby, zbz;     // thrust direction MuJoCo last computed (stale body z axis) as the half axis (a, b, c) of
                       // half_axis(): z = (2a, 2b, 1 - 2c)
};

// Set-point of one table row (minimum_snap.py:122-123 columns 0..9): fp64 position (errors are formed in fp64), the rest
// already rounded to the arithmetic type.  The yaw column (:9) travels as a heading direction (yc, ys) = k (cos yaw,
// sin yaw), k > 0 arbitrary.
template <class R> struct Target {
  double x, y, z;
  R vx, vy, vz, ax, ay, az;
  R yc, ys;
};

// The same set-point as it is stored in a precomputed table (shared missions): 56 bytes per row.
struct TargetRow {
  double x, y, z;
  float vx, vy, vz, ax, ay, az, yc, ys;
};

template <class R> UAVB_HD void set_thrust_cmd(Drone<R>& d, const VehU<R>& u, R c) {
  d.thrust_cmd = c;
  d.coll = u.quarter_inv_kf * clampr<R>(c, u.fmin4, u.fmax4);
}

// Fold the displacement into the fp64 position (done at outer-period boundaries).
template <class R> UAVB_HD void fold_position(Drone<R>& d) {
  d.px += (double)d.dx; d.py += (double)d.dy; d.pz += (double)d.dz;
  d.dx = d.dy = d.dz = R(0);
}

// Half of the third column of R(q) (quad.py:153): z = (2a, 2b, 1 - 2c).
template <class R> UAVB_HD void half_axis(const Drone<R>& d, R* a, R* b, R* c) {
  typedef Math<R> M;
  *a = M::fma(d.q1, d.q3, d.q0 * d.q2);
  *b = M::fma(d.q2, d.q3, -(d.q0 * d.q1));
  *c = M::fma(d.q1, d.q1, d.q2 * d.q2);
}

// Third column of R(q) for a normalised quaternion (quad.py:153): body z axis in the world frame.
template <class R> UAVB_HD void body_z(const Drone<R>& d, R* zx, R* zy, R* zz) {
  typedef Math<R> M;
  const R a = M::fma(d.q1, d.q3, d.q0 * d.q2);
  const R b = M::fma(d.q2, d.q3, -(d.q0 * d.q1));
  const R c = M::fma(d.q1, d.q1, d.q2 * d.q2);
  *zx = a + a;
  *zy = b + b;
  *zz = M::fma(R(-2), c, R(1));
}

// ---------------------------------------------------------------------------------------------
// Outer-loop stages.  Each is one method of CascadedController; outer_update chains them exactly like
// TrajectoryController._update_outer_loop (main.py:47-61), and the stage kernels expose them one by one.

// Entries of R(q) the outer loop reads (quad.py:153, unit quaternion).
template <class R> struct RotE {
  R R00, R01, R02, R10, R11, R12, R22;
};
template <class R> UAVB_HD RotE<R> rot_entries(R q0, R q1, R q2, R q3) {
  RotE<R> r;
  r.R00 = R(1) - R(2) * (q2 * q2 + q3 * q3); r.R01 = R(2) * (q1 * q2 - q0 * q3); r.R02 = R(2) * (q1 * q3 + q0 * q2);
  r.R10 = R(2) * (q1 * q2 + q0 * q3); r.R11 = R(1) - R(2) * (q1 * q1 + q3 * q3); r.R12 = R(2) * (q2 * q3 - q0 * q1);
  r.R22 = R(1) - R(2) * (q1 * q1 + q2 * q2);
  return r;
}

// CascadedController.altitude (controller.py:26-56): collective thrust command; updates the integrator.
// ez = z_des - z (formed in fp64 by the caller and rounded once).
template <class R>
UAVB_HD R altitude_cmd(R& integral, const VehU<R>& u, const VehP<R>& v, R ez, R vz, R vz_des, R az_des, R inv_R22) {
  const R climb = clampr<R>(vz_des, -u.max_ascent, u.max_descent);
  const R ezd = climb - vz;
  integral = clampr<R>(integral + ez * u.dt_outer, -u.integral_limit, u.integral_limit);
  R acc_z = v.kp_z * ez + v.ki_z * integral + v.kd_z * ezd + az_des - u.g;
  acc_z = acc_z * inv_R22;
  return clampr<R>(-v.mass * acc_z, u.fmin4, u.fmax4);
}

// CascadedController.lateral (controller.py:58-97): commanded R02, R12.  ex, ey = p_des - p.
template <class R>
UAVB_HD void lateral_cmd(const VehU<R>& u, const VehP<R>& v, R ex, R ey, R vx, R vy, R vxd, R vyd, R axd, R ayd, R c, R* bx, R* by) {
  typedef Math<R> M;
  const R vm2 = vxd * vxd + vyd * vyd;
  if (vm2 > u.max_speed_xy * u.max_speed_xy) {
    const R s = u.max_speed_xy * M::rsqrt(vm2);
    vxd *= s; vyd *= s;
  }
  R ax = v.kp_xy * ex + v.kd_xy * (vxd - vx) + axd;
  R ay = v.kp_xy * ey + v.kd_xy * (vyd - vy) + ayd;
  const R am2 = ax * ax + ay * ay;
  if (am2 > u.max_acc_xy * u.max_acc_xy) {
    const R s = u.max_acc_xy * M::rsqrt(am2);
    ax *= s; ay *= s;
  }
  const R inv_accz = -v.mass * M::rcp_fast(c);    // 1 / (-c/m)
  *bx = clampr<R>(ax * inv_accz, -u.max_tilt, u.max_tilt);
  *by = clampr<R>(ay * inv_accz, -u.max_tilt, u.max_tilt);
}

// CascadedController.roll_pitch_controller (controller.py:132-154).
template <class R> UAVB_HD void roll_pitch_cmd(const VehP<R>& v, R bx, R by, const RotE<R>& r, R inv_R22, R* p_c, R* q_c) {
  const R bdx = v.kp_roll * (bx - r.R02);
  const R bdy = v.kp_pitch * (by - r.R12);
  *p_c = (r.R10 * bdx - r.R00 * bdy) * inv_R22;
  *q_c = (r.R11 * bdx - r.R01 * bdy) * inv_R22;
}

// CascadedController.yaw_controller (controller.py:156-168) with the Euler angles of quad.py:189-213 taken without the
// trig round trip: phi = atan2(a, b) -> sin = a/h, cos = b/h; theta = asin(s) -> cos = sqrt(1 - s^2).
// The yaw error wrap_to_pi(wrap_to_2pi(psi_des) - psi) (controller.py:164-165) is ONE atan2 of the rotated heading:
// with (cp, sp) ~ (cos psi, sin psi) from quad.py:208-213, atan2(ys cp - yc sp, yc cp + ys sp) is the angle from psi
// to psi_des in (-pi, pi] (the reference's floored modulo gives [-pi, pi): they differ only at exactly +-pi).
template <class R> UAVB_HD R yaw_rate_cmd(const VehP<R>& v, R q0, R q1, R q2, R q3, R yc, R ys, R q_c) {
  typedef Math<R> M;
  const R R22 = R(1) - R(2) * (q1 * q1 + q2 * q2);
  const R sa = R(2) * (q0 * q1 + q2 * q3);
  const R ih = M::rsqrt(sa * sa + R22 * R22);
  const R sin_phi = sa * ih, cos_phi = R22 * ih;
  const R sin_th = clampr<R>(R(2) * (q0 * q2 - q3 * q1), R(-1), R(1));
  const R cos_th = M::sqrt_fast(R(1) - sin_th * sin_th);
  const R sp = R(2) * (q0 * q3 + q1 * q2), cp = R(1) - R(2) * (q2 * q2 + q3 * q3);
  const R e_yaw = M::atan2(ys * cp - yc * sp, yc * cp + ys * sp);
  return (v.kp_yaw * e_yaw * cos_th - q_c * sin_phi) * M::rcp_fast(cos_phi);
}

// The same for a UNIT quaternion (what the persistent rollout's outer loop always sees): with sa = sin(phi) cos(theta) and
// R22 = cos(phi) cos(theta), cos(theta)^2 = sa^2 + R22^2, so
//   (kp e cos(theta) - q_c sin(phi)) / cos(phi) = (kp e (sa^2 + R22^2) - q_c sa) / R22
// and neither Euler angle nor a square root is needed; inv_R22 is the reciprocal the altitude loop already formed.
template <class R> UAVB_HD R yaw_rate_cmd_unit(const VehP<R>& v, R q0, R q1, R q2, R q3, R yc, R ys, R q_c, R R22, R inv_R22) {
  typedef Math<R> M;
  const R sa = R(2) * (q0 * q1 + q2 * q3);
  const R sp = R(2) * (q0 * q3 + q1 * q2), cp = R(1) - R(2) * (q2 * q2 + q3 * q3);
  const R e_yaw = M::atan2(ys * cp - yc * sp, yc * cp + ys * sp);
  return (v.kp_yaw * e_yaw * (sa * sa + R22 * R22) - q_c * sa) * inv_R22;
}

// TrajectoryController._update_outer_loop (main.py:47-61) on the fresh state; (ex, ey, ez) = set-point - position, formed
// in fp64 by the caller and rounded once.
template <class R> UAVB_HD void outer_update(Drone<R>& d, const VehU<R>& u, const VehP<R>& v, const Target<R>& t, R ex, R ey, R ez) {
  const RotE<R> r = rot_entries<R>(d.q0, d.q1, d.q2, d.q3);
  const R inv_R22 = Math<R>::rcp_fast(r.R22);
  const R c = altitude_cmd<R>(d.integral, u, v, ez, d.vz, t.vz, t.az, inv_R22);
  set_thrust_cmd<R>(d, u, c);
  R bx, by;
  lateral_cmd<R>(u, v, ex, ey, d.vx, d.vy, t.vx, t.vy, t.ax, t.ay, c, &bx, &by);
  R p_c, q_c;
  roll_pitch_cmd<R>(v, bx, by, r, inv_R22, &p_c, &q_c);
  d.pc = p_c; d.qc = q_c;
  d.rc = yaw_rate_cmd_unit<R>(v, d.q0, d.q1, d.q2, d.q3, t.yc, t.ys, q_c, r.R22, inv_R22);
  d.cp = v.Jp * d.pc; d.cq = v.Jq * d.qc; d.cr = v.Jr * d.rc;
}

// ---------------------------------------------------------------------------------------------
// Inner-loop stages (main.py:42-44).

// CascadedController.body_rate_controller (controller.py:115-130): M = I kp (cmd - w) + w x (I w).  With a diagonal
// inertia w x (I w) = (dIx wy wz, dIy wz wx, dIz wx wy); it is returned because the physics step reuses it.
template <class R> UAVB_HD void body_rate_moment(const Drone<R>& d, const VehP<R>& v, R* g, R* Mo) {
  typedef Math<R> M;
  g[0] = v.dIx * (d.wy * d.wz); g[1] = v.dIy * (d.wz * d.wx); g[2] = v.dIz * (d.wx * d.wy);
  Mo[0] = M::fma(v.Ikp_p, d.pc - d.wx, g[0]);
  Mo[1] = M::fma(v.Ikp_q, d.qc - d.wy, g[1]);
  Mo[2] = M::fma(v.Ikp_r, d.rc - d.wz, g[2]);
}

// Mixer rows (+,+,+) (-,+,-) (-,-,+) (+,-,-) on the scaled moments (p_bar, q_bar, r_bar)/4 plus the rotor limits of
// quad.py:114-121.  Unit-agnostic: `coll`, `lo`, `hi` and the result share one unit (N in the stage form, rotor units in the
// persistent rollout).
template <class R> UAVB_HD void mix_and_limit(R pb, R qb, R rb, R coll, R lo, R hi, R* f) {
  typedef Math<R> M;
  // unclipped outputs first (coll +- r_bar, then +- the roll/pitch sums: 8 additions): when all four lie inside the limits
  // no ratio of quad.py:116-119 is below 1, the scale is 1 and the final clip is the identity
  const R s1 = pb + qb, s2 = pb - qb, t1 = coll + rb, t2 = coll - rb;
  const R f0 = t1 + s1, f1 = t2 - s2, f2 = t1 - s1, f3 = t2 + s2;
  const R f_hi = M::fmax(M::fmax(f0, f1), M::fmax(f2, f3)), f_lo = M::fmin(M::fmin(f0, f1), M::fmin(f2, f3));
  if (f_hi <= hi && f_lo >= lo) {
    f[0] = f0; f[1] = f1; f[2] = f2; f[3] = f3;
  } else {
    const R m0 = s1 + rb, m1 = -(s2 + rb), m2 = rb - s1, m3 = s2 - rb;
    const R room_hi = hi - coll, room_lo = lo - coll;      // >= 0 and <= 0: coll is a clipped collective share
    // ratios of quad.py:116-119: room_hi / m for m > 0, room_lo / m for m < 0, 1 for m = 0, and the scale is their minimum
    // clipped to [0, 1].  With room_hi >= 0 >= room_lo the minimum over the positive moments is the ratio of the LARGEST one
    // and the minimum over the negative moments that of the most negative one (room * (1/m) is monotonic in m on either
    // side), so two reciprocals give the same bits as four.  Straight-line selects: a few lanes of a warp saturating must
    // not cost the warp a chain of branches.  fmax / fmin drop NaN moments, which therefore keep ratio 1.
    const R m_hi = M::fmax(M::fmax(m0, m1), M::fmax(m2, m3)), m_lo = M::fmin(M::fmin(m0, m1), M::fmin(m2, m3));
    const R l_hi = (m_hi > R(0)) ? M::div(room_hi, m_hi) : R(1);
    const R l_lo = (m_lo < R(0)) ? M::div(room_lo, m_lo) : R(1);
    const R s = clampr<R>(M::fmin(l_hi, l_lo), R(0), R(1));
    f[0] = clampr<R>(M::fma(s, m0, coll), lo, hi);
    f[1] = clampr<R>(M::fma(s, m1, coll), lo, hi);
    f[2] = clampr<R>(M::fma(s, m2, coll), lo, hi);
    f[3] = clampr<R>(M::fma(s, m3, coll), lo, hi);
  }
}

// Quad._allocate_rotor_forces (quad.py:105-122) in newtons; `coll` = clip(thrust_cmd, 4 fmin, 4 fmax) / 4.
template <class R> UAVB_HD void allocate_forces(const VehU<R>& u, R coll, const R* Mo, R* f) {
  mix_and_limit<R>(Mo[0] * u.inv_arm4, Mo[1] * u.inv_arm4, -Mo[2] * u.inv_kappa4, coll, u.fmin, u.fmax, f);
}

// Asymmetric first-order lag of the rotor speeds toward their commands (quad.py:98-103): w += a (c - w) with a = a_rise
// when c > w and a_fall otherwise, i.e. a (c - w) = a_mean (c - w) + a_hdiff |c - w| with a_mean = (a_rise + a_fall)/2 and
// a_hdiff = (a_rise - a_fall)/2 -- two fused multiply-adds and no select.
template <class R> UAVB_HD void lag_toward(Drone<R>& d, const VehU<R>& u, R c0, R c1, R c2, R c3) {
  typedef Math<R> M;
  const R e0 = c0 - d.om0, e1 = c1 - d.om1, e2 = c2 - d.om2, e3 = c3 - d.om3;
  d.om0 = M::fma(u.a_hdiff, M::abs(e0), M::fma(u.a_mean, e0, d.om0));
  d.om1 = M::fma(u.a_hdiff, M::abs(e1), M::fma(u.a_mean, e1, d.om1));
  d.om2 = M::fma(u.a_hdiff, M::abs(e2), M::fma(u.a_mean, e2, d.om2));
  d.om3 = M::fma(u.a_hdiff, M::abs(e3), M::fma(u.a_mean, e3, d.om3));
}

// Quad.set_propeller_speed after the allocation (quad.py:95-103): w_cmd = sqrt(f / kf), then the lag.
template <class R> UAVB_HD void motor_lag(Drone<R>& d, const VehU<R>& u, const R* f, R* cmd_out) {
  typedef Math<R> M;
  const R c0 = M::sqrt_fast(f[0] * u.inv_kf), c1 = M::sqrt_fast(f[1] * u.inv_kf), c2 = M::sqrt_fast(f[2] * u.inv_kf), c3 = M::sqrt_fast(f[3] * u.inv_kf);
  lag_toward<R>(d, u, c0, c1, c2, c3);
  if (cmd_out) { cmd_out[0] = c0; cmd_out[1] = c1; cmd_out[2] = c2; cmd_out[3] = c3; }
}

// 1.5 - 0.5 |q|^2 = 1/|q| + O((|q|^2 - 1)^2): renormalisation of a quaternion whose norm is 1 up to accumulated rounding.
template <class R> UAVB_HD void renormalise_q(Drone<R>& d) {
  typedef Math<R> M;
  const R nn = M::fma(d.q0, d.q0, M::fma(d.q1, d.q1, M::fma(d.q2, d.q2, d.q3 * d.q3)));
  const R rn = M::fma(R(-0.5), nn, R(1.5));
  d.q0 *= rn; d.q1 *= rn; d.q2 *= rn; d.q3 *= rn;
}

// Sums of squared rotor speeds behind the wrench (mujoco_sim.py:235-247, rotor layout lab_course.xml:116-119):
//   tot = s0+s1+s2+s3,  tx = (s0+s3)-(s1+s2),  ty = (s0+s1)-(s2+s3),  tzn = -((s1+s3)-(s0+s2)),  s_i = w_i^2
// from one butterfly (8 additions).
template <class R> UAVB_HD void rotor_sums(const Drone<R>& d, R* tot, R* tx, R* ty, R* tzn) {
  const R s0 = d.om0 * d.om0, s1 = d.om1 * d.om1, s2 = d.om2 * d.om2, s3 = d.om3 * d.om3;
  const R a = s0 + s1, b = s2 + s3, c = s0 - s1, e = s2 - s3;
  *tot = a + b; *ty = a - b; *tx = c - e; *tzn = c + e;
}

// NORM: renormalise the quaternion in this step (mujoco_sim.py:36-42 does so every tick).  The persistent rollout passes
// false and renormalises once per outer period instead: q * dq of a unit q and the unit dq below stays unit up to
// ~6e-8 per tick, the drift over 10 ticks (< 1e-6 in |q|^2) scales the thrust axis and R by the same factor, far inside
// the fp32 noise of the step, and the outer loop always sees a freshly normalised q.
// Semi-implicit Euler with the velocity increments of the tick (ivx, ivy, ivz) and the new body rates (nwx, nwy, nwz).
// The increments are formed by the caller and added once: near hover thrust and gravity cancel inside them.
template <class R, bool NORM = true>
UAVB_HD void integrate(Drone<R>& d, const VehU<R>& u, R ivx, R ivy, R ivz, R nwx, R nwy, R nwz) {
  typedef Math<R> M;
  // velocities first (semi-implicit Euler)
  d.vx += ivx; d.vy += ivy; d.vz += ivz;
  d.wx = nwx; d.wy = nwy; d.wz = nwz;
  // positions with the new velocity
  d.dx = M::fma(u.dt, d.vx, d.dx);
  d.dy = M::fma(u.dt, d.vy, d.dy);
  d.dz = M::fma(u.dt, d.vz, d.dz);
  // q <- q * [cos(a/2), sin(a/2) w/|w|], a = dt |w|   (mju_quatIntegrate), written as q += q*(dq-1)
  const R wn2 = M::fma(d.wx, d.wx, M::fma(d.wy, d.wy, d.wz * d.wz));
  const R h = u.half_dt;
  const R x2 = u.half_dt_sq * wn2;                                // (a/2)^2
  R sf, cm1;
  if (x2 < R(1e-3)) {                                        // |a/2| < 0.0316
    if constexpr (sizeof(R) == 4) {
      // fp32: the next series terms (x2^2/120, x2^3/720 < 1e-8 relative) are below half an ulp
      sf = h * M::fma(x2, R(-1.0 / 6), R(1));
      cm1 = x2 * M::fma(x2, R(1.0 / 24), R(-0.5));
    } else {                                                 // series exact to < 1e-13 relative
      sf = h * M::fma(x2, M::fma(x2, M::fma(x2, R(-1.0 / 5040), R(1.0 / 120)), R(-1.0 / 6)), R(1));
      cm1 = x2 * M::fma(x2, M::fma(x2, M::fma(x2, R(1.0 / 40320), R(-1.0 / 720)), R(1.0 / 24)), R(-0.5));
    }
  } else {
    // |w| > 63 rad/s: a tumbling vehicle.  Single-MUFU forms (fp32: rsqrt + sin/cos approximations, abs. error ~1e-6) keep
    // this cold branch free of subroutine calls, so the hot loop around it keeps its constants in (uniform) registers.
    const R iw = M::rsqrt(wn2);
    R sn, cs;
    M::sincos_fast(h * (wn2 * iw), &sn, &cs);
    sf = sn * iw;
    cm1 = cs - R(1);
  }
  const R bx = sf * d.wx, by = sf * d.wy, bz = sf * d.wz;
  const R q0 = d.q0, q1 = d.q1, q2 = d.q2, q3 = d.q3;
  const R n0 = q0 + M::fma(-q3, bz, M::fma(-q2, by, M::fma(-q1, bx, q0 * cm1)));
  const R n1 = q1 + M::fma(-q3, by, M::fma(q2, bz, M::fma(q0, bx, q1 * cm1)));
  const R n2 = q2 + M::fma(q3, bx, M::fma(-q1, bz, M::fma(q0, by, q2 * cm1)));
  const R n3 = q3 + M::fma(-q2, bx, M::fma(q1, by, M::fma(q0, bz, q3 * cm1)));
  if constexpr (NORM) {
    // re-normalise (mujoco_sim.py:36-42).  The input quaternion is unit and dq is unit up to the series remainder, so
    // |n|^2 = 1 + e with |e| at rounding level and 1/sqrt(1+e) = 1.5 - 0.5 |n|^2 + O(e^2) is exact to working precision.
    const R nn = M::fma(n0, n0, M::fma(n1, n1, M::fma(n2, n2, n3 * n3)));
    const R rn = M::fma(R(-0.5), nn, R(1.5));
    d.q0 = n0 * rn; d.q1 = n1 * rn; d.q2 = n2 * rn; d.q3 = n3 * rn;
  } else {
    d.q0 = n0; d.q1 = n1; d.q2 = n2; d.q3 = n3;
  }
}

// Inner loop, part 2 in SI units (stage kernels): rotor wrench with the given thrust axis + free-body step
// (mujoco_sim.py:232-255 + MuJoCo Euler; oracle/freebody.py states the same equations in fp64).
template <class R, bool NORM = true>
UAVB_HD void physics_step(Drone<R>& d, const VehU<R>& u, const VehP<R>& v, R zx, R zy, R zz, R gx, R gy, R gz) {
  typedef Math<R> M;
  R tot, tx, ty, tzn;
  rotor_sums<R>(d, &tot, &tx, &ty, &tzn);
  const R dvt = -tot * v.kf_dt_over_m;                         // dt * specific thrust along -z body
  integrate<R, NORM>(d, u, M::fma(zx, dvt, v.dvx), M::fma(zy, dvt, v.dvy), M::fma(zz, dvt, v.dvz),
                     M::fma(v.dt_invIx, M::fma(u.arm_kf, tx, -gx), d.wx),
                     M::fma(v.dt_invIy, M::fma(u.arm_kf, ty, -gy), d.wy),
                     M::fma(v.dt_invIz, M::fma(-u.kappa_kf, tzn, -gz), d.wz));
}

// One full inner tick in the reference order (SURVEY 8(a) "exact tick order"): body-rate loop (controller.py:115-130),
// allocation (quad.py:105-122), motor lag (quad.py:95-103), wrench with the stale (lag=1) or fresh (lag=0) thrust axis
// (mujoco_sim.py:232-255), integration.
//
// Rotor units.  Everything between the body-rate error and the rotor speeds is linear, so the tick works with moments
// divided by (4 arm kf) -- (4 kappa kf) and a flipped sign for yaw -- and forces divided by kf, i.e. directly with squared
// rotor speeds: the mixer inputs need no scaling, the speed commands are sqrt() of the limited mixer outputs, and the
// gyroscopic products re-enter the rate update with their own constants K.  Same equations as body_rate_moment ->
// allocate_forces -> motor_lag -> physics_step with the constant factors folded into VehP (Gx.., Jp.., Wx..) once per
// rollout.
template <class R, bool NORM = true> UAVB_HD void inner_tick(Drone<R>& d, const VehU<R>& u, const VehP<R>& v, int thrust_frame_lag) {
  typedef Math<R> M;
  // gyroscopic products; M' = J (cmd - w) + G (w x w) is formed as (J cmd + G prod) - J w with J cmd carried from the
  // outer loop, and the products re-enter the rate update through K
  const R yz = d.wy * d.wz, zx_ = d.wz * d.wx, xy = d.wx * d.wy;
  R w2[4];
  mix_and_limit<R>(M::fma(-v.Jp, d.wx, M::fma(v.Gx, yz, d.cp)), M::fma(-v.Jq, d.wy, M::fma(v.Gy, zx_, d.cq)),
                   M::fma(-v.Jr, d.wz, M::fma(v.Gz, xy, d.cr)), d.coll, u.w2min, u.w2max, w2);
  lag_toward<R>(d, u, M::sqrt_fast(w2[0]), M::sqrt_fast(w2[1]), M::sqrt_fast(w2[2]), M::sqrt_fast(w2[3]));
  // thrust axis of X_k (what mj_step's forward pass will compute) as the half axis (a, b, c): z = (2a, 2b, 1 - 2c), so that
  // dv = z dvt + dv0 = (a dvt2 + dvx, b dvt2 + dvy, (dvt2/2 + dvz) - c dvt2) with dvt2 = 2 dvt and no doubling of the axis
  R ha, hb, hc;
  half_axis<R>(d, &ha, &hb, &hc);
  const R ua = thrust_frame_lag ? d.zbx : ha, ub = thrust_frame_lag ? d.zby : hb, uc = thrust_frame_lag ? d.zbz : hc;
  d.zbx = ha; d.zby = hb; d.zbz = hc;
  R tot, tx, ty, tzn;
  rotor_sums<R>(d, &tot, &tx, &ty, &tzn);
  const R dvt2 = -tot * v.kf_dt_over_m2;                       // 2 dt * specific thrust along -z body
  integrate<R, NORM>(d, u, M::fma(ua, dvt2, v.dvx), M::fma(ub, dvt2, v.dvy), M::fma(-uc, dvt2, M::fma(R(0.5), dvt2, v.dvz)),
                     M::fma(v.Wx, tx, M::fma(v.Kx, yz, d.wx)),
                     M::fma(v.Wy, ty, M::fma(v.Ky, zx_, d.wy)),
                     M::fma(-v.Wz, tzn, M::fma(v.Kz, xy, d.wz)));
}

// ---------------------------------------------------------------------------------------------
// Table row on the fly (minimum_snap.py:104-110): t = j*dt in fp64, Horner in fp64.
// c points at the 24 coefficients of the segment in the reference layout [power][axis].
// Position, velocity and acceleration come from one pass of the nested Horner recurrence
//   d2 <- d2 t + d1 ; d1 <- d1 t + p ; p <- p t + c_k      (k = 6 .. 0;  p' = d1, p'' = 2 d2)
// i.e. 3 FMAs per coefficient and no multiplications by the derivative factors k, k(k-1).
template <class LOAD> UAVB_HD void eval_row(LOAD ld, double t, double* p, double* v, double* a) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int ax = 0; ax < 3; ++ax) {
    double pp = ld(21 + ax), d1 = 0.0, d2 = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 6; k >= 0; --k) {
      d2 = ::fma(d2, t, d1);
      d1 = ::fma(d1, t, pp);
      pp = ::fma(pp, t, ld(3 * k + ax));
    }
    p[ax] = pp; v[ax] = d1; a[ax] = d2 + d2;
  }
}

}  // namespace uavb

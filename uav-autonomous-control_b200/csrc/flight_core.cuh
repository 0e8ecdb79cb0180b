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
//   * position is carried as an unevaluated sum hi+lo and advanced with a compensated add;
//   * set-points come from fp64 Horner evaluation and position errors are formed in fp64;
//   * the quaternion is advanced by adding q*(dq-1), so only the final add rounds at 1 ulp of q.
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
  static UAVB_HD float rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
  }
  static UAVB_HD float div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);
#else
    return a / b;
#endif
  }
  static UAVB_HD float atan2(float y, float x) { return atan2f(y, x); }
  static UAVB_HD void sincos(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    sincosf(x, s, c);
#else
    *s = sinf(x); *c = cosf(x);
#endif
  }
  static UAVB_HD float rint(float x) { return rintf(x); }
  static UAVB_HD float floor(float x) { return floorf(x); }
  static UAVB_HD float abs(float x) { return fabsf(x); }
  static UAVB_HD float fmin(float a, float b) { return fminf(a, b); }
  static UAVB_HD float fmax(float a, float b) { return fmaxf(a, b); }
  static UAVB_HD bool finite(float x) { return fabsf(x) <= 3.0e38f; }
};
template <> struct Math<double> {
  static UAVB_HD double sqrt(double x) { return ::sqrt(x); }
  static UAVB_HD double rsqrt(double x) { return 1.0 / ::sqrt(x); }
  static UAVB_HD double div(double a, double b) { return a / b; }
  static UAVB_HD double atan2(double y, double x) { return ::atan2(y, x); }
  static UAVB_HD void sincos(double x, double* s, double* c) { *s = ::sin(x); *c = ::cos(x); }
  static UAVB_HD double rint(double x) { return ::rint(x); }
  static UAVB_HD double floor(double x) { return ::floor(x); }
  static UAVB_HD double abs(double x) { return ::fabs(x); }
  static UAVB_HD double fmin(double a, double b) { return ::fmin(a, b); }
  static UAVB_HD double fmax(double a, double b) { return ::fmax(a, b); }
  static UAVB_HD bool finite(double x) { return ::fabs(x) <= 1.0e300; }
};

template <class R> UAVB_HD R clampr(R x, R lo, R hi) { return Math<R>::fmin(Math<R>::fmax(x, lo), hi); }

// Per-drone constants.  Fields above the marker may differ between rollouts (Monte-Carlo); the rest
// is uniform across a launch.
template <class R> struct Veh {
  // per rollout
  R mass, inv_mass;
  R Ix, Iy, Iz, inv_Ix, inv_Iy, inv_Iz;
  R Ikp_p, Ikp_q, Ikp_r;            // I * kp of the body-rate loop (controller.py:128)
  R kp_xy, kd_xy, kp_z, kd_z, ki_z, kp_roll, kp_pitch, kp_yaw;
  R wax, way, waz;                  // wind force / mass (extension; zero for reference runs)
  // uniform
  R g, dt, dt_outer;
  R arm, inv_arm4, kf, inv_kf, kappa, inv_kappa4;   // inv_*4 = 1/(4 x): mixer division by 4 folded in (quad.py:112)
  R fmin, fmax, a_rise, a_fall;
  R max_ascent, max_descent, max_speed_xy, max_acc_xy, max_tilt, integral_limit;
};

// Persistent per-drone state (registers across the whole rollout).
template <class R> struct Drone {
  R px, py, pz;        // position (hi part)
  R plx, ply, plz;     // position low-order part (float mode; always 0 in double mode)
  R q0, q1, q2, q3;    // attitude, scalar first, FRD->NED
  R vx, vy, vz;        // world velocity
  R wx, wy, wz;        // body rates p q r
  R om0, om1, om2, om3;  // rotor speeds (quad.py:85)
  R integral;          // altitude integrator (controller.py:20)
  R thrust_cmd;        // main.py:26
  R pc, qc, rc;        // pqr_cmd (main.py:27)
  R zbx, zby, zbz;     // thrust direction MuJoCo last computed (stale body z axis)
};

// Set-point of one table row (minimum_snap.py:122-123 columns 0..9), fp64 from the Horner evaluation.
struct Target {
  double x, y, z, vx, vy, vz, ax, ay, az;
  double yaw;
};

// Third column of R(q) for a normalised quaternion (quad.py:153): body z axis in the world frame.
template <class R> UAVB_HD void body_z(const Drone<R>& d, R* zx, R* zy, R* zz) {
  *zx = R(2) * (d.q1 * d.q3 + d.q0 * d.q2);
  *zy = R(2) * (d.q2 * d.q3 - d.q0 * d.q1);
  *zz = R(1) - R(2) * (d.q1 * d.q1 + d.q2 * d.q2);
}

// Position error p_des - p formed in fp64 from the hi+lo pair, then rounded once.
template <class R> UAVB_HD R pos_err(double des, R hi, R lo) { return (R)((des - (double)hi) - (double)lo); }

// yaw error: wrap_to_pi(wrap_to_2pi(psi_des) - psi) (controller.py:164-165, :170-178).  Python's %
// is a floored modulo, reproduced with floor().
template <class R> UAVB_HD R yaw_error(R psi_des, R psi) {
  const R two_pi = (R)kTwoPi, pi = (R)kPi;
  R a = psi_des - two_pi * Math<R>::floor(psi_des / two_pi);        // [0, 2pi)
  R e = a - psi + pi;
  e = e - two_pi * Math<R>::floor(e / two_pi);
  return e - pi;
}

// ---------------------------------------------------------------------------------------------
// Outer loop: TrajectoryController._update_outer_loop (main.py:47-61) on the fresh state.
template <class R> UAVB_HD void outer_update(Drone<R>& d, const Veh<R>& v, const Target& t) {
  typedef Math<R> M;
  const R q0 = d.q0, q1 = d.q1, q2 = d.q2, q3 = d.q3;
  // quad.py:153 for a unit quaternion (the state is re-normalised every tick)
  const R R00 = R(1) - R(2) * (q2 * q2 + q3 * q3), R01 = R(2) * (q1 * q2 - q0 * q3), R02 = R(2) * (q1 * q3 + q0 * q2);
  const R R10 = R(2) * (q1 * q2 + q0 * q3), R11 = R(1) - R(2) * (q1 * q1 + q3 * q3), R12 = R(2) * (q2 * q3 - q0 * q1);
  const R R22 = R(1) - R(2) * (q1 * q1 + q2 * q2);
  const R inv_R22 = R(1) / R22;

  // altitude (controller.py:26-56)
  const R climb = clampr<R>((R)t.vz, -v.max_ascent, v.max_descent);
  const R ez = pos_err<R>(t.z, d.pz, d.plz);
  const R ezd = climb - d.vz;
  d.integral = clampr<R>(d.integral + ez * v.dt_outer, -v.integral_limit, v.integral_limit);
  R acc_z = v.kp_z * ez + v.ki_z * d.integral + v.kd_z * ezd + (R)t.az - v.g;
  acc_z = acc_z * inv_R22;
  const R c = clampr<R>(-v.mass * acc_z, R(4) * v.fmin, R(4) * v.fmax);
  d.thrust_cmd = c;

  // lateral (controller.py:58-97)
  R vxd = (R)t.vx, vyd = (R)t.vy;
  const R vmag = M::sqrt(vxd * vxd + vyd * vyd);
  if (vmag > v.max_speed_xy) {
    const R s = v.max_speed_xy / vmag;
    vxd *= s; vyd *= s;
  }
  R ax = v.kp_xy * pos_err<R>(t.x, d.px, d.plx) + v.kd_xy * (vxd - d.vx) + (R)t.ax;
  R ay = v.kp_xy * pos_err<R>(t.y, d.py, d.ply) + v.kd_xy * (vyd - d.vy) + (R)t.ay;
  const R amag = M::sqrt(ax * ax + ay * ay);
  if (amag > v.max_acc_xy) {
    const R s = v.max_acc_xy / amag;
    ax *= s; ay *= s;
  }
  const R inv_accz = -v.mass / c;                 // 1 / (-c/m)
  const R bx = clampr<R>(ax * inv_accz, -v.max_tilt, v.max_tilt);
  const R by = clampr<R>(ay * inv_accz, -v.max_tilt, v.max_tilt);

  // roll / pitch rates (controller.py:132-154)
  const R bdx = v.kp_roll * (bx - R02);
  const R bdy = v.kp_pitch * (by - R12);
  const R p_c = (R10 * bdx - R00 * bdy) * inv_R22;
  const R q_c = (R11 * bdx - R01 * bdy) * inv_R22;

  // yaw rate (controller.py:156-168); Euler angles of quad.py:189-213 without the trig round trip:
  // phi = atan2(a, b) -> sin = a/h, cos = b/h; theta = asin(s) -> cos = sqrt(1-s^2)
  const R sa = R(2) * (q0 * q1 + q2 * q3);
  const R ih = M::rsqrt(sa * sa + R22 * R22);
  const R sin_phi = sa * ih, cos_phi = R22 * ih;
  const R sin_th = clampr<R>(R(2) * (q0 * q2 - q3 * q1), R(-1), R(1));
  const R cos_th = M::sqrt(R(1) - sin_th * sin_th);
  const R psi = M::atan2(R(2) * (q0 * q3 + q1 * q2), R(1) - R(2) * (q2 * q2 + q3 * q3));
  const R e_yaw = yaw_error<R>((R)t.yaw, psi);
  const R r_c = (v.kp_yaw * e_yaw * cos_th - q_c * sin_phi) / cos_phi;

  d.pc = p_c; d.qc = q_c; d.rc = r_c;
}

// ---------------------------------------------------------------------------------------------
// Inner loop, part 1: body-rate controller + allocation + motor lag (main.py:42-44).
// Returns the gyroscopic term w x (I w) of the CURRENT state, which the physics step reuses.
template <class R>
UAVB_HD void inner_control(Drone<R>& d, const Veh<R>& v, R* gx, R* gy, R* gz, R* moment_out, R* forces_out) {
  typedef Math<R> M;
  // controller.py:115-130
  const R Iwx = v.Ix * d.wx, Iwy = v.Iy * d.wy, Iwz = v.Iz * d.wz;
  *gx = d.wy * Iwz - d.wz * Iwy;
  *gy = d.wz * Iwx - d.wx * Iwz;
  *gz = d.wx * Iwy - d.wy * Iwx;
  const R Mx = v.Ikp_p * (d.pc - d.wx) + *gx;
  const R My = v.Ikp_q * (d.qc - d.wy) + *gy;
  const R Mz = v.Ikp_r * (d.rc - d.wz) + *gz;
  // quad.py:105-122
  const R c_bar = clampr<R>(d.thrust_cmd, R(4) * v.fmin, R(4) * v.fmax);
  const R coll = R(0.25) * c_bar;
  const R pb = Mx * v.inv_arm4, qb = My * v.inv_arm4, rb = -Mz * v.inv_kappa4;
  const R m0 = pb + qb + rb, m1 = qb - pb - rb, m2 = rb - pb - qb, m3 = pb - qb - rb;
  const R room_hi = v.fmax - coll, room_lo = v.fmin - coll;
  R s = R(1);
  {
    const R l0 = (m0 > R(0)) ? M::div(room_hi, m0) : ((m0 < R(0)) ? M::div(room_lo, m0) : R(1));
    const R l1 = (m1 > R(0)) ? M::div(room_hi, m1) : ((m1 < R(0)) ? M::div(room_lo, m1) : R(1));
    const R l2 = (m2 > R(0)) ? M::div(room_hi, m2) : ((m2 < R(0)) ? M::div(room_lo, m2) : R(1));
    const R l3 = (m3 > R(0)) ? M::div(room_hi, m3) : ((m3 < R(0)) ? M::div(room_lo, m3) : R(1));
    s = clampr<R>(M::fmin(M::fmin(l0, l1), M::fmin(l2, l3)), R(0), R(1));
  }
  const R f0 = clampr<R>(coll + s * m0, v.fmin, v.fmax);
  const R f1 = clampr<R>(coll + s * m1, v.fmin, v.fmax);
  const R f2 = clampr<R>(coll + s * m2, v.fmin, v.fmax);
  const R f3 = clampr<R>(coll + s * m3, v.fmin, v.fmax);
  // quad.py:88-103
  const R c0 = M::sqrt(f0 * v.inv_kf), c1 = M::sqrt(f1 * v.inv_kf), c2 = M::sqrt(f2 * v.inv_kf), c3 = M::sqrt(f3 * v.inv_kf);
  d.om0 += ((c0 > d.om0) ? v.a_rise : v.a_fall) * (c0 - d.om0);
  d.om1 += ((c1 > d.om1) ? v.a_rise : v.a_fall) * (c1 - d.om1);
  d.om2 += ((c2 > d.om2) ? v.a_rise : v.a_fall) * (c2 - d.om2);
  d.om3 += ((c3 > d.om3) ? v.a_rise : v.a_fall) * (c3 - d.om3);
  if (moment_out) { moment_out[0] = Mx; moment_out[1] = My; moment_out[2] = Mz; }
  if (forces_out) { forces_out[0] = f0; forces_out[1] = f1; forces_out[2] = f2; forces_out[3] = f3; }
}

// compensated p += inc for the hi+lo pair (float mode); plain add in double mode
UAVB_HD void pos_add(float& hi, float& lo, float inc) {
  const float y = inc + lo;
  const float t = hi + y;
  lo = y - (t - hi);
  hi = t;
}
UAVB_HD void pos_add(double& hi, double& lo, double inc) { hi += inc; (void)lo; }

// Inner loop, part 2: rotor wrench with the given thrust axis + free-body semi-implicit Euler step
// (mujoco_sim.py:232-255 + MuJoCo Euler; oracle/freebody.py states the same equations in fp64).
template <class R>
UAVB_HD void physics_step(Drone<R>& d, const Veh<R>& v, R zx, R zy, R zz, R gx, R gy, R gz) {
  typedef Math<R> M;
  const R F0 = v.kf * d.om0 * d.om0, F1 = v.kf * d.om1 * d.om1, F2 = v.kf * d.om2 * d.om2, F3 = v.kf * d.om3 * d.om3;
  const R a_t = -(F0 + F1 + F2 + F3) * v.inv_mass;          // specific thrust along -z body
  const R tx = v.arm * ((F0 + F3) - (F1 + F2));
  const R ty = v.arm * ((F0 + F1) - (F2 + F3));
  const R tz = v.kappa * ((F1 + F3) - (F0 + F2));
  // velocities first
  d.vx += v.dt * (zx * a_t + v.wax);
  d.vy += v.dt * (zy * a_t + v.way);
  d.vz += v.dt * (zz * a_t + v.waz + v.g);
  d.wx += v.dt * ((tx - gx) * v.inv_Ix);
  d.wy += v.dt * ((ty - gy) * v.inv_Iy);
  d.wz += v.dt * ((tz - gz) * v.inv_Iz);
  // positions with the new velocity
  pos_add(d.px, d.plx, v.dt * d.vx);
  pos_add(d.py, d.ply, v.dt * d.vy);
  pos_add(d.pz, d.plz, v.dt * d.vz);
  // q <- q * [cos(a/2), sin(a/2) w/|w|], a = dt |w|   (mju_quatIntegrate), written as q += q*(dq-1)
  const R wn2 = d.wx * d.wx + d.wy * d.wy + d.wz * d.wz;
  const R h = R(0.5) * v.dt;
  const R x2 = h * h * wn2;                                  // (a/2)^2
  R sf, cm1;
  if (x2 < R(1e-3)) {                                        // |a/2| < 0.0316: series exact to < 1e-13 relative
    sf = h * (R(1) - x2 * (R(1.0 / 6) - x2 * (R(1.0 / 120) - x2 * R(1.0 / 5040))));
    cm1 = -x2 * (R(0.5) - x2 * (R(1.0 / 24) - x2 * (R(1.0 / 720) - x2 * R(1.0 / 40320))));
  } else {
    const R wn = M::sqrt(wn2);
    R sn, cs;
    M::sincos(h * wn, &sn, &cs);
    sf = sn / wn;
    cm1 = cs - R(1);
  }
  const R bx = sf * d.wx, by = sf * d.wy, bz = sf * d.wz;
  const R q0 = d.q0, q1 = d.q1, q2 = d.q2, q3 = d.q3;
  R n0 = q0 + (q0 * cm1 - q1 * bx - q2 * by - q3 * bz);
  R n1 = q1 + (q1 * cm1 + q0 * bx + q2 * bz - q3 * by);
  R n2 = q2 + (q2 * cm1 + q0 * by - q1 * bz + q3 * bx);
  R n3 = q3 + (q3 * cm1 + q0 * bz + q1 * by - q2 * bx);
  const R rn = M::rsqrt(n0 * n0 + n1 * n1 + n2 * n2 + n3 * n3);
  d.q0 = n0 * rn; d.q1 = n1 * rn; d.q2 = n2 * rn; d.q3 = n3 * rn;
}

// One full inner tick in the reference order (SURVEY 8(a) "exact tick order"): body-rate loop,
// allocation, motor lag, wrench with the stale (lag=1) or fresh (lag=0) thrust axis, integration.
template <class R> UAVB_HD void inner_tick(Drone<R>& d, const Veh<R>& v, int thrust_frame_lag) {
  R gx, gy, gz;
  inner_control<R>(d, v, &gx, &gy, &gz, nullptr, nullptr);
  R zx, zy, zz;
  body_z<R>(d, &zx, &zy, &zz);                               // axis of X_k: what mj_step's forward pass will compute
  const R ux = thrust_frame_lag ? d.zbx : zx, uy = thrust_frame_lag ? d.zby : zy, uz = thrust_frame_lag ? d.zbz : zz;
  d.zbx = zx; d.zby = zy; d.zbz = zz;
  physics_step<R>(d, v, ux, uy, uz, gx, gy, gz);
}

// ---------------------------------------------------------------------------------------------
// Table row on the fly (minimum_snap.py:104-110): t = j*dt in fp64, Horner in fp64.
// c points at the 24 coefficients of the segment in the reference layout [power][axis].
template <class LOAD> UAVB_HD void eval_row(LOAD ld, double t, Target* out) {
  double p[3], v[3], a[3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int ax = 0; ax < 3; ++ax) {
    const double c7 = ld(21 + ax), c6 = ld(18 + ax), c5 = ld(15 + ax), c4 = ld(12 + ax);
    const double c3 = ld(9 + ax), c2 = ld(6 + ax), c1 = ld(3 + ax), c0 = ld(ax);
    p[ax] = ((((((c7 * t + c6) * t + c5) * t + c4) * t + c3) * t + c2) * t + c1) * t + c0;
    v[ax] = (((((7.0 * c7 * t + 6.0 * c6) * t + 5.0 * c5) * t + 4.0 * c4) * t + 3.0 * c3) * t + 2.0 * c2) * t + c1;
    a[ax] = ((((42.0 * c7 * t + 30.0 * c6) * t + 20.0 * c5) * t + 12.0 * c4) * t + 6.0 * c3) * t + 2.0 * c2;
  }
  out->x = p[0]; out->y = p[1]; out->z = p[2];
  out->vx = v[0]; out->vy = v[1]; out->vz = v[2];
  out->ax = a[0]; out->ay = a[1]; out->az = a[2];
}

}  // namespace uavb

// flight_pair.cuh -- the 1 kHz tick of the fp32 production rollout for TWO drones per thread, every state variable a float2
// (lane x = first drone, lane y = second drone), written with the packed fp32x2 instructions of sm_100
// (FFMA2 / FADD2 / FMUL2: fma.rn.f32x2, add.rn.f32x2, mul.rn.f32x2).
//
// Why two drones per thread (measured, profiles/r02_ffma2_probe.md): a packed instruction occupies the FMA pipe for two
// cycles but only ONE issue slot, and it has no lane swizzle -- both lanes must be the same operation on aligned register
// pairs.  The tick of one drone is full of horizontal products (cross products, the quaternion product, the mixer and the
// rotor-sum butterflies), so packing inside one drone costs as many moves as it saves; across two drones every operation
// of the tick is vertical.  The scalar tick was issue-bound (148 issue slots per drone-tick, FMA pipe 61 % busy); the pair
// tick is FMA-pipe-bound.
//
// Same equations, same order, same reference lines as inner_tick<float> in flight_core.cuh (controller.py:115-130,
// quad.py:88-122, mujoco_sim.py:232-255 + the restated free-body step) in the same "rotor units".  Every operation is an
// explicit IEEE round-to-nearest intrinsic (no compiler contraction), so a drone's trajectory does not depend on which
// lane it sits in, on its partner, or on the template instantiation that flies it.
//
// Device-only: the 100 Hz outer loop stays the scalar code of flight_core.cuh, run once per lane.
#pragma once

#include "flight_core.cuh"

namespace uavb {

typedef float2 V2;

#define UAVB_DEV __device__ __forceinline__

UAVB_DEV V2 as2(V2 x) { return x; }
UAVB_DEV V2 as2(float x) { return make_float2(x, x); }            // becomes a 32-bit operand broadcast to both lanes (R.F32 / UR.F32)
UAVB_DEV V2 neg2(V2 x) { return make_float2(-x.x, -x.y); }        // folds into the operand's negate modifier
UAVB_DEV float neg2(float x) { return -x; }
UAVB_DEV V2 abs2(V2 x) { return make_float2(fabsf(x.x), fabsf(x.y)); }   // folds into the |abs| modifier
template <class A, class B, class C> UAVB_DEV V2 fma2(A a, B b, C c) { return __ffma2_rn(as2(a), as2(b), as2(c)); }
template <class A, class B> UAVB_DEV V2 add2(A a, B b) { return __fadd2_rn(as2(a), as2(b)); }
template <class A, class B> UAVB_DEV V2 sub2(A a, B b) { return __fadd2_rn(as2(a), neg2(as2(b))); }
template <class A, class B> UAVB_DEV V2 mul2(A a, B b) { return __fmul2_rn(as2(a), as2(b)); }
UAVB_DEV V2 sqrt2(V2 x) { return make_float2(Math<float>::sqrt_fast(x.x), Math<float>::sqrt_fast(x.y)); }
template <int L> UAVB_DEV float& lane(V2& v) { return L ? v.y : v.x; }
template <int L> UAVB_DEV float lane(const V2& v) { return L ? v.y : v.x; }

// Per-rollout constants of the 1 kHz body for a pair with Monte-Carlo overrides (lane = drone).  Without overrides the
// tick reads the scalar VehP<float> of the kernel parameter block instead and every constant is a broadcast operand.
struct VehP2 {
  V2 kf_dt_over_m2, Gx, Gy, Gz, Jp, Jq, Jr, Wx, Wy, Wz, Kx, Ky, Kz, dvx, dvy, dvz;
};
UAVB_DEV void zip_vehp(VehP2& o, const VehP<float>& a, const VehP<float>& b) {
  o.kf_dt_over_m2 = make_float2(a.kf_dt_over_m2, b.kf_dt_over_m2);
  o.Gx = make_float2(a.Gx, b.Gx); o.Gy = make_float2(a.Gy, b.Gy); o.Gz = make_float2(a.Gz, b.Gz);
  o.Jp = make_float2(a.Jp, b.Jp); o.Jq = make_float2(a.Jq, b.Jq); o.Jr = make_float2(a.Jr, b.Jr);
  o.Wx = make_float2(a.Wx, b.Wx); o.Wy = make_float2(a.Wy, b.Wy); o.Wz = make_float2(a.Wz, b.Wz);
  o.Kx = make_float2(a.Kx, b.Kx); o.Ky = make_float2(a.Ky, b.Ky); o.Kz = make_float2(a.Kz, b.Kz);
  o.dvx = make_float2(a.dvx, b.dvx); o.dvy = make_float2(a.dvy, b.dvy); o.dvz = make_float2(a.dvz, b.dvz);
}

// Persistent state of the pair (registers across a whole slice).  Fields as in Drone<float>; the fp64 fold position and
// the 100 Hz scalars are kept per lane.
struct Drone2 {
  double px[2], py[2], pz[2];
  V2 dx, dy, dz;
  V2 q0, q1, q2, q3;
  V2 vx, vy, vz;
  V2 wx, wy, wz;
  V2 om0, om1, om2, om3;
  V2 coll, cp, cq, cr;
  V2 zbx, zby, zbz;
  float integral[2], thrust_cmd[2];
};

template <int L> UAVB_DEV void get_lane(const Drone2& p, Drone<float>& d) {
  d.px = p.px[L]; d.py = p.py[L]; d.pz = p.pz[L];
  d.dx = lane<L>(p.dx); d.dy = lane<L>(p.dy); d.dz = lane<L>(p.dz);
  d.q0 = lane<L>(p.q0); d.q1 = lane<L>(p.q1); d.q2 = lane<L>(p.q2); d.q3 = lane<L>(p.q3);
  d.vx = lane<L>(p.vx); d.vy = lane<L>(p.vy); d.vz = lane<L>(p.vz);
  d.wx = lane<L>(p.wx); d.wy = lane<L>(p.wy); d.wz = lane<L>(p.wz);
  d.om0 = lane<L>(p.om0); d.om1 = lane<L>(p.om1); d.om2 = lane<L>(p.om2); d.om3 = lane<L>(p.om3);
  d.integral = p.integral[L]; d.thrust_cmd = p.thrust_cmd[L]; d.coll = lane<L>(p.coll);
  d.pc = d.qc = d.rc = 0.f;                           // SI body-rate commands: not carried by the persistent rollout
  d.cp = lane<L>(p.cp); d.cq = lane<L>(p.cq); d.cr = lane<L>(p.cr);
  d.zbx = lane<L>(p.zbx); d.zby = lane<L>(p.zby); d.zbz = lane<L>(p.zbz);
}
template <int L> UAVB_DEV void put_lane(Drone2& p, const Drone<float>& d) {
  p.px[L] = d.px; p.py[L] = d.py; p.pz[L] = d.pz;
  lane<L>(p.dx) = d.dx; lane<L>(p.dy) = d.dy; lane<L>(p.dz) = d.dz;
  lane<L>(p.q0) = d.q0; lane<L>(p.q1) = d.q1; lane<L>(p.q2) = d.q2; lane<L>(p.q3) = d.q3;
  lane<L>(p.vx) = d.vx; lane<L>(p.vy) = d.vy; lane<L>(p.vz) = d.vz;
  lane<L>(p.wx) = d.wx; lane<L>(p.wy) = d.wy; lane<L>(p.wz) = d.wz;
  lane<L>(p.om0) = d.om0; lane<L>(p.om1) = d.om1; lane<L>(p.om2) = d.om2; lane<L>(p.om3) = d.om3;
  p.integral[L] = d.integral; p.thrust_cmd[L] = d.thrust_cmd; lane<L>(p.coll) = d.coll;
  lane<L>(p.cp) = d.cp; lane<L>(p.cq) = d.cq; lane<L>(p.cr) = d.cr;
  lane<L>(p.zbx) = d.zbx; lane<L>(p.zby) = d.zby; lane<L>(p.zbz) = d.zbz;
}
// What the outer loop writes (outer_update): integrator, thrust command and the body-rate commands in rotor units.
template <int L> UAVB_DEV void put_lane_commands(Drone2& p, const Drone<float>& d) {
  p.integral[L] = d.integral; p.thrust_cmd[L] = d.thrust_cmd; lane<L>(p.coll) = d.coll;
  lane<L>(p.cp) = d.cp; lane<L>(p.cq) = d.cq; lane<L>(p.cr) = d.cr;
}

// Mixer + rotor limits (quad.py:105-122) in rotor units for the pair.  The unclipped outputs come from the packed
// butterfly; a lane whose outputs leave [lo, hi] re-runs the scalar mix_and_limit<float> (same adds, same bits for the
// unclipped part, then the ratio scaling and the clip), the other lane keeps its packed values.
UAVB_DEV void mix_and_limit_pair(V2 pb, V2 qb, V2 rb, V2 coll, float lo, float hi, V2& f0, V2& f1, V2& f2, V2& f3) {
  const V2 s1 = add2(pb, qb), s2 = sub2(pb, qb), t1 = add2(coll, rb), t2 = sub2(coll, rb);
  f0 = add2(t1, s1); f1 = sub2(t2, s2); f2 = sub2(t1, s1); f3 = add2(t2, s2);
  const float hi8 = fmaxf(fmaxf(fmaxf(f0.x, f0.y), fmaxf(f1.x, f1.y)), fmaxf(fmaxf(f2.x, f2.y), fmaxf(f3.x, f3.y)));
  const float lo8 = fminf(fminf(fminf(f0.x, f0.y), fminf(f1.x, f1.y)), fminf(fminf(f2.x, f2.y), fminf(f3.x, f3.y)));
  if (!(hi8 <= hi && lo8 >= lo)) {
    float f[4];
    mix_and_limit<float>(pb.x, qb.x, rb.x, coll.x, lo, hi, f);       // identity on a lane that is inside its limits
    f0.x = f[0]; f1.x = f[1]; f2.x = f[2]; f3.x = f[3];
    mix_and_limit<float>(pb.y, qb.y, rb.y, coll.y, lo, hi, f);
    f0.y = f[0]; f1.y = f[1]; f2.y = f[2]; f3.y = f[3];
  }
}

// Asymmetric motor lag (quad.py:98-103): w += a_mean e + a_hdiff |e|, e = cmd - w.
UAVB_DEV void lag_toward_pair(Drone2& d, const VehU<float>& u, V2 c0, V2 c1, V2 c2, V2 c3) {
  const V2 e0 = sub2(c0, d.om0), e1 = sub2(c1, d.om1), e2 = sub2(c2, d.om2), e3 = sub2(c3, d.om3);
  d.om0 = fma2(u.a_hdiff, abs2(e0), fma2(u.a_mean, e0, d.om0));
  d.om1 = fma2(u.a_hdiff, abs2(e1), fma2(u.a_mean, e1, d.om1));
  d.om2 = fma2(u.a_hdiff, abs2(e2), fma2(u.a_mean, e2, d.om2));
  d.om3 = fma2(u.a_hdiff, abs2(e3), fma2(u.a_mean, e3, d.om3));
}

UAVB_DEV void renormalise_q_pair(Drone2& d) {
  const V2 nn = fma2(d.q0, d.q0, fma2(d.q1, d.q1, fma2(d.q2, d.q2, mul2(d.q3, d.q3))));
  const V2 rn = fma2(-0.5f, nn, 1.5f);
  d.q0 = mul2(d.q0, rn); d.q1 = mul2(d.q1, rn); d.q2 = mul2(d.q2, rn); d.q3 = mul2(d.q3, rn);
}

UAVB_DEV void half_axis_pair(const Drone2& d, V2& a, V2& b, V2& c) {
  a = fma2(d.q1, d.q3, mul2(d.q0, d.q2));
  b = fma2(d.q2, d.q3, neg2(mul2(d.q0, d.q1)));
  c = fma2(d.q1, d.q1, mul2(d.q2, d.q2));
}

// Large-rotation branch of the quaternion map for one lane (|w| > 63 rad/s, a tumbling vehicle): single-MUFU forms.
UAVB_DEV void half_angle_large(float wn2, float h, float* sf, float* cm1) {
  const float iw = Math<float>::rsqrt(wn2);
  float sn, cs;
  Math<float>::sincos_fast(__fmul_rn(h, __fmul_rn(wn2, iw)), &sn, &cs);
  *sf = __fmul_rn(sn, iw);
  *cm1 = __fadd_rn(cs, -1.f);
}

// One inner tick of the pair in the reference order; see inner_tick<float> for the derivation of the rotor-unit forms.
// VP: VehP2 (per-lane constants) or VehP<float> (launch-uniform constants, broadcast).  LAG: thrust along the body axis
// of the previous tick (stale data.xmat) or of this one.
template <bool NORM, bool LAG, class VP> UAVB_DEV void inner_tick_pair(Drone2& d, const VehU<float>& u, const VP& v) {
  const V2 yz = mul2(d.wy, d.wz), zx = mul2(d.wz, d.wx), xy = mul2(d.wx, d.wy);
  V2 w0, w1, w2, w3;
  mix_and_limit_pair(fma2(neg2(v.Jp), d.wx, fma2(v.Gx, yz, d.cp)), fma2(neg2(v.Jq), d.wy, fma2(v.Gy, zx, d.cq)),
                     fma2(neg2(v.Jr), d.wz, fma2(v.Gz, xy, d.cr)), d.coll, u.w2min, u.w2max, w0, w1, w2, w3);
  lag_toward_pair(d, u, sqrt2(w0), sqrt2(w1), sqrt2(w2), sqrt2(w3));
  V2 ha, hb, hc;
  half_axis_pair(d, ha, hb, hc);
  const V2 ua = LAG ? d.zbx : ha, ub = LAG ? d.zby : hb, uc = LAG ? d.zbz : hc;
  d.zbx = ha; d.zby = hb; d.zbz = hc;
  // rotor sums (mujoco_sim.py:235-247): a = s0+s1, c = s0-s1, b = s2+s3, e = s2-s3 with s_i = w_i^2
  const V2 s0 = mul2(d.om0, d.om0), s2 = mul2(d.om2, d.om2);
  const V2 a = fma2(d.om1, d.om1, s0), c = fma2(neg2(d.om1), d.om1, s0);
  const V2 b = fma2(d.om3, d.om3, s2), e = fma2(neg2(d.om3), d.om3, s2);
  const V2 tot = add2(a, b), ty = sub2(a, b), tx = sub2(c, e), tzn = add2(c, e);
  const V2 dvt2 = mul2(neg2(tot), v.kf_dt_over_m2);
  // semi-implicit Euler: velocity increments formed first (thrust and gravity cancel inside them near hover), new rates
  d.vx = add2(d.vx, fma2(ua, dvt2, v.dvx));
  d.vy = add2(d.vy, fma2(ub, dvt2, v.dvy));
  d.vz = add2(d.vz, fma2(neg2(uc), dvt2, fma2(0.5f, dvt2, v.dvz)));
  d.wx = fma2(v.Wx, tx, fma2(v.Kx, yz, d.wx));
  d.wy = fma2(v.Wy, ty, fma2(v.Ky, zx, d.wy));
  d.wz = fma2(neg2(v.Wz), tzn, fma2(v.Kz, xy, d.wz));
  d.dx = fma2(u.dt, d.vx, d.dx);
  d.dy = fma2(u.dt, d.vy, d.dy);
  d.dz = fma2(u.dt, d.vz, d.dz);
  // q <- q * [cos(a/2), sin(a/2) w/|w|], a = dt |w|, written as q += q*(dq-1) (mju_quatIntegrate)
  const V2 wn2 = fma2(d.wx, d.wx, fma2(d.wy, d.wy, mul2(d.wz, d.wz)));
  const V2 x2 = mul2(u.half_dt_sq, wn2);
  V2 sf = mul2(u.half_dt, fma2(x2, -1.0f / 6, 1.f));
  V2 cm1 = mul2(x2, fma2(x2, 1.0f / 24, -0.5f));
  if (!(fmaxf(x2.x, x2.y) < 1e-3f)) {
    if (!(x2.x < 1e-3f)) half_angle_large(wn2.x, u.half_dt, &sf.x, &cm1.x);
    if (!(x2.y < 1e-3f)) half_angle_large(wn2.y, u.half_dt, &sf.y, &cm1.y);
  }
  const V2 bx = mul2(sf, d.wx), by = mul2(sf, d.wy), bz = mul2(sf, d.wz);
  const V2 q0 = d.q0, q1 = d.q1, q2 = d.q2, q3 = d.q3;
  const V2 n0 = add2(q0, fma2(neg2(q3), bz, fma2(neg2(q2), by, fma2(neg2(q1), bx, mul2(q0, cm1)))));
  const V2 n1 = add2(q1, fma2(neg2(q3), by, fma2(q2, bz, fma2(q0, bx, mul2(q1, cm1)))));
  const V2 n2 = add2(q2, fma2(q3, bx, fma2(neg2(q1), bz, fma2(q0, by, mul2(q2, cm1)))));
  const V2 n3 = add2(q3, fma2(neg2(q2), bx, fma2(q1, by, fma2(q0, bz, mul2(q3, cm1)))));
  if constexpr (NORM) {
    const V2 nn = fma2(n0, n0, fma2(n1, n1, fma2(n2, n2, mul2(n3, n3))));
    const V2 rn = fma2(-0.5f, nn, 1.5f);
    d.q0 = mul2(n0, rn); d.q1 = mul2(n1, rn); d.q2 = mul2(n2, rn); d.q3 = mul2(n3, rn);
  } else {
    d.q0 = n0; d.q1 = n1; d.q2 = n2; d.q3 = n3;
  }
}

}  // namespace uavb

// rollout_pair.cuh -- the fp32 production rollout, two drones per thread: thread j of the launch flies drones 2j and
// 2j+1 in the two lanes of float2 registers (flight_pair.cuh: the 1 kHz tick in packed fp32x2 instructions); the 100 Hz
// outer loop, the table cursor, the metrics and the collision flag are the scalar code of rollout_core.cuh run once per
// lane.  Both drones of a pair tick in lock-step (same tick count, same outer/inner schedule, main.py:37-45), so the
// schedule below is rollout_run's with every per-drone step done twice.
//
// A pair's two rollouts never mix: each lane runs exactly the operation sequence a lone drone would, so per-rollout
// results do not depend on the partner, on the batch size or on how a job is sharded (tests/test_rollout_gpu.py).
#pragma once

#include "flight_pair.cuh"
#include "rollout_core.cuh"

namespace uavb {

// [sample][13][B] state log of a pair: drones 2j, 2j+1 are neighbours in every field row, so a field leaves as ONE 8-byte
// streaming store per thread when the rows are 8-byte aligned (B even) -- a warp writes 256 contiguous bytes per field.
// `out` points at element 2j of field 0 of the next sample.
struct PairLog {
  static constexpr bool kNormEveryTick = true;
  float* out;
  unsigned B;              // field stride in elements
  int stride, left;
  bool vec, second;        // 8-byte stores possible; the pair's second drone exists
  UAVB_DEV void put(float* o, V2 x) const {
    if (vec) { __stcs(reinterpret_cast<float2*>(o), x); return; }
    __stcs(o, x.x);
    if (second) __stcs(o + 1, x.y);
  }
  UAVB_DEV void tick(const Drone2& d) {
    if (--left) return;
    left = stride;
    float* o = out;
    const unsigned b = B;
    put(o, make_float2((float)(d.px[0] + (double)d.dx.x), (float)(d.px[1] + (double)d.dx.y)));
    put(o + (size_t)b, make_float2((float)(d.py[0] + (double)d.dy.x), (float)(d.py[1] + (double)d.dy.y)));
    put(o + (size_t)(2u * b), make_float2((float)(d.pz[0] + (double)d.dz.x), (float)(d.pz[1] + (double)d.dz.y)));
    put(o + (size_t)(3u * b), d.q0); put(o + (size_t)(4u * b), d.q1); put(o + (size_t)(5u * b), d.q2); put(o + (size_t)(6u * b), d.q3);
    put(o + (size_t)(7u * b), d.vx); put(o + (size_t)(8u * b), d.vy); put(o + (size_t)(9u * b), d.vz);
    put(o + (size_t)(10u * b), d.wx); put(o + (size_t)(11u * b), d.wy); put(o + (size_t)(12u * b), d.wz);
    out = o + (size_t)(13u * b);
  }
};

struct NoPairLog {
  static constexpr bool kNormEveryTick = false;
  UAVB_DEV void tick(const Drone2&) {}
};

// Outer update of lane L (TrajectoryController._update_outer_loop, main.py:47-61) on the fresh state of that drone.
template <int L, bool TABLE> UAVB_DEV void pair_outer(Drone2& d, Cursor<float>& c, const VehU<float>& u, const VehP<float>& v,
                                                     const MissionView& m, const Target<float>* shared_t) {
  Target<float> t;
  if constexpr (TABLE) {
    t = *shared_t;                                           // shared mission: one row for both lanes
  } else {
    cursor_target<float>(c, m, &t);
    cursor_advance(&c.seg, &c.row, m);
  }
  c.ex = (float)(t.x - d.px[L]); c.ey = (float)(t.y - d.py[L]); c.ez = (float)(t.z - d.pz[L]);   // the position is folded here
  Drone<float> s;
  get_lane<L>(d, s);
  outer_update<float>(s, u, v, t, c.ex, c.ey, c.ez);
  put_lane_commands<L>(d, s);
}

// End of an outer period for lane L: tracking error |set-point - p| (test_mujoco_trajectory_tracking.py:27-31), fold.
template <int L> UAVB_DEV void pair_period_end(Drone2& d, Cursor<float>& c, Accum<float>& a) {
  const float ex = c.ex - lane<L>(d.dx), ey = c.ey - lane<L>(d.dy), ez = c.ez - lane<L>(d.dz);
  d.px[L] += (double)lane<L>(d.dx); d.py[L] += (double)lane<L>(d.dy); d.pz[L] += (double)lane<L>(d.dz);
  const float e2 = ex * ex + ey * ey + ez * ez;
  const float e = Math<float>::sqrt_fast(e2);
  a.sum_e += e; a.sum_e2 += e2; a.max_e = fmaxf(a.max_e, e);
  ++a.periods;
}

template <int L, class OBST> UAVB_DEV bool pair_watch(const Drone2& d, const Accum<float>& a, const VehU<float>& u, const VehP<float>& v,
                                                     const OBST& obst, int n, float& clear) {
  if (a.collided) return false;
  const float T = (float)n * u.dt;
  const float vx = lane<L>(d.vx), vy = lane<L>(d.vy), vz = lane<L>(d.vz);
  const float speed = Math<float>::sqrt_fast(vx * vx + vy * vy + vz * vz);
  const float reach = 1.01f * (speed + v.acc_max * T) * T + 1e-4f;
  if (!(clear > reach))
    clear = obst.gap((float)(d.px[L] + (double)lane<L>(d.dx)), (float)(d.py[L] + (double)lane<L>(d.dy)), (float)(d.pz[L] + (double)lane<L>(d.dz)));
  const bool watch = !(clear > reach);                       // NaN positions keep measuring and watching
  clear -= reach;
  return watch;
}

template <int L, class OBST> UAVB_DEV void pair_hit(const Drone2& d, Accum<float>& a, const OBST& obst, int tick) {
  if (!a.collided && obst.hit((float)(d.px[L] + (double)lane<L>(d.dx)), (float)(d.py[L] + (double)lane<L>(d.dy)), (float)(d.pz[L] + (double)lane<L>(d.dz)))) {
    a.collided = 1; a.first_hit = tick;
  }
}

// n_ticks ticks of the closed loop for the pair; the schedule, the obstacle culling and the metrics are rollout_run's
// (rollout_core.cuh), see there.  c[0].phase is the phase of both lanes.  VP2: VehP2 or VehP<float> (see inner_tick_pair).
template <bool TABLE, class VP2, class OBST, class LOG>
UAVB_DEV void rollout_run_pair(Drone2& d, Cursor<float> (&c)[2], Accum<float> (&a)[2], const VehU<float>& u, const VehP<float>& va,
                               const VehP<float>& vb, const VP2& v2, const MissionView& ma, const MissionView& mb, int tick0,
                               int n_ticks, int freq, int lag, const OBST& oa, const OBST& ob, LOG& logger) {
  int k = 0;
  float clear_a = 0.f, clear_b = 0.f;                        // not part of the carry: every launch / slice measures first
  while (k < n_ticks) {
    if (c[0].phase == 0) {
      Target<float> t;
      if constexpr (TABLE) {
        table_target<float>(ma.trows, c[0].row, &t);
        if (c[0].row + 1 < ma.n_trows) ++c[0].row;           // index clamp of main.py:61
        c[1].row = c[0].row;
      }
      pair_outer<0, TABLE>(d, c[0], u, va, ma, &t);
      pair_outer<1, TABLE>(d, c[1], u, vb, mb, &t);
    }
    const int n = (freq - c[0].phase < n_ticks - k) ? (freq - c[0].phase) : (n_ticks - k);
    bool watch = false;
    if (OBST::kAny) {
      const bool wa = pair_watch<0>(d, a[0], u, va, oa, n, clear_a);
      const bool wb = pair_watch<1>(d, a[1], u, vb, ob, n, clear_b);
      watch = __any_sync(__activemask(), wa || wb);          // one decision per warp; watching is always correct
    }
    auto stretch = [&](auto watch_c, auto lag_c) {
      constexpr bool kWatch = decltype(watch_c)::value, kLag = decltype(lag_c)::value;
#pragma unroll 2
      for (int j = 0; j < n; ++j) {
        inner_tick_pair<LOG::kNormEveryTick, kLag>(d, u, v2);
        if constexpr (kWatch) {
          pair_hit<0>(d, a[0], oa, tick0 + k + j);
          pair_hit<1>(d, a[1], ob, tick0 + k + j);
        }
        logger.tick(d);
      }
    };
    if (watch) {
      if (lag) stretch(BoolC<true>{}, BoolC<true>{}); else stretch(BoolC<true>{}, BoolC<false>{});
    } else {
      if (lag) stretch(BoolC<false>{}, BoolC<true>{}); else stretch(BoolC<false>{}, BoolC<false>{});
    }
    k += n;
    c[0].phase += n;
    if (c[0].phase == freq) {
      c[0].phase = 0;
      if (!LOG::kNormEveryTick) renormalise_q_pair(d);       // once per outer period (not per launch: chunked runs stay bit-identical)
      pair_period_end<0>(d, c[0], a[0]);
      pair_period_end<1>(d, c[1], a[1]);
      d.dx = d.dy = d.dz = make_float2(0.f, 0.f);
    }
    c[1].phase = c[0].phase;
  }
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    if (!Math<float>::finite(a[l].sum_e2)) a[l].status |= 1;
    if (a[l].max_e > 1e4f) a[l].status |= 2;                 // more than 1e4 m from its set-point
  }
}

}  // namespace uavb

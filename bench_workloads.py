"""Workload builders shared by bench.py and tools/run_configs.py: the BASELINE.json configs at full size on the product path
(uav_ac_b200.kernels over the C ABI), each returning a closure to time plus what is needed to check its result.

configs[1]  10^6 random 5-waypoint solves (K1)
configs[2]  10^5 lab_course rollouts per GPU, Monte-Carlo gains x U(0.8,1.2), mass / inertia x U(0.9,1.1)   (the headline workload)
configs[3]  10^6 rollouts with per-rollout missions, wind and 64 obstacle sets x 6 AABBs
configs[4]  10^7 x 60 s x 1 kHz sharded 8 ways: one GPU's share is 1.25e6 rollouts x 60 000 ticks in chunked launches
K3          sampled (N, 11) tables of 10^5 config-1 missions;  RRT*: 16 384 lab-volume missions

Nothing here imports oracle/ except parity_block(), which is the checker leg of the bench (outside every timed region).
"""
from __future__ import annotations

import os
import statistics

import numpy as np

FREQUENCY = 10
MC_LO, MC_HI = [0.8] * 11 + [0.9] * 4, [1.2] * 11 + [1.1] * 4        # 11 gains, mass, 3 inertia


def event_ms(fn, reps=3, warm=1, before=None):
    """Mean CUDA-event time of fn() over `reps` calls after `warm` untimed ones (events on the current torch stream)."""
    import torch
    out, ts = None, []
    for i in range(warm + reps):
        if before is not None:
            before(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return statistics.mean(ts), out


def mc_vehicle_arrays(kernels, nat, dev, B, seed=20261017, index_base=0):
    """[15, B] fp32: 11 gains, mass, 3 inertia of B Monte-Carlo vehicles, keyed by the global rollout index."""
    import torch
    veh = nat.default_vehicle()
    base = torch.tensor(list(veh.gains) + [veh.mass] + list(veh.inertia), dtype=torch.float32, device=dev)[:, None]
    return (kernels.mc_uniform(seed, B, MC_LO, MC_HI, index_base=index_base, device=dev) * base).contiguous()


def lab_course(kernels, dev, velocity=3.0, table_rows=None):
    """Shared lab_course mission (take-off + course tables) planned on the device; returns (plan, kwargs of kernels.rollout)."""
    import torch
    from uav_ac_b200.simulation.scene import LAB_COURSE_GOAL, LAB_COURSE_OBSTACLES, LAB_COURSE_START, LAB_COURSE_WAYPOINTS
    f64 = dict(dtype=torch.float64, device=dev)
    wp = torch.tensor(LAB_COURSE_WAYPOINTS, **f64)
    vel = torch.tensor([velocity], **f64)
    plan = kernels.plan_missions([(wp[None, :2].contiguous(), vel), (wp[None, 1:].contiguous(), vel)], 0.01, shared=True, table_rows=table_rows,
                                 obstacles=torch.tensor(LAB_COURSE_OBSTACLES, **f64))
    kw = dict(start=torch.tensor(LAB_COURSE_START, **f64), goal=torch.tensor(LAB_COURSE_GOAL, **f64),
              obstacles=torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=dev), want_state=False)
    return plan, kw


def config3(kernels, dev, B=1_000_000, seed=31, index_base=0, boxes_seed=8):
    """BASELINE configs[3]: per-rollout 5-waypoint missions (take-off from the ground below the first waypoint), constant wind
    +-0.08 N per axis, 64 obstacle sets of 6 random AABBs.  Returns (fly, n_ticks, info)."""
    import torch
    wp, vel = kernels.mc_missions(seed, B, 4, index_base=index_base, device=dev)
    ground = wp[:, 0].clone()
    ground[:, 2] = -0.021
    plan = kernels.plan_missions([(torch.stack((ground, wp[:, 0]), dim=1).contiguous(), vel), (wp, vel)], 0.01)
    wind = kernels.mc_uniform(seed + 1, B, [-0.08] * 3, [0.08] * 3, index_base=index_base, device=dev)
    rng = np.random.default_rng(boxes_seed)
    ctr, half = rng.uniform([2, 2, -5], [22, 12, -1], (64, 6, 3)), rng.uniform(0.3, 1.2, (64, 6, 3))
    boxes = np.stack((ctr[..., 0] - half[..., 0], ctr[..., 0] + half[..., 0], ctr[..., 1] - half[..., 1], ctr[..., 1] + half[..., 1],
                      ctr[..., 2] - half[..., 2], ctr[..., 2] + half[..., 2]), axis=-1).astype(np.float32)
    boxes_t = torch.tensor(boxes, device=dev)
    idx = torch.arange(index_base, index_base + B, device=dev, dtype=torch.int64)
    sets = ((idx * 2654435761) % 64).to(torch.int32)
    n_ticks = FREQUENCY * int(plan.total_rows.max().item())
    start, goal = ground.contiguous(), wp[:, -1].contiguous()

    def fly(out=None):
        return kernels.rollout(plan, B, n_ticks, start=start, goal=goal, mc_wind=wind, obstacles=boxes_t, obstacle_set=sets, want_state=False, out=out)
    return fly, n_ticks, dict(plan=plan, wind=wind, boxes=boxes_t, sets=sets, start=start, goal=goal)


def config4_share(kernels, nat, dev, B=1_250_000, ticks=60_000, chunk=20_000, index_base=0):
    """One GPU's share of BASELINE configs[4]: B lab_course rollouts with Monte-Carlo vehicles, 60 s at 1 kHz, metrics only, flown as
    ticks / chunk launches through the resumable carry block (the shared mission ends after ~10.8 s; the drones then hold the last
    set-point, main.py:61).  Returns (fly, ticks)."""
    import torch
    plan, kw = lab_course(kernels, dev)
    mc = mc_vehicle_arrays(kernels, nat, dev, B, seed=5, index_base=index_base)
    carry = torch.empty((nat.CARRY_WORDS, B), dtype=torch.float32, device=dev)
    result = kernels.RolloutResult(torch.empty((B, nat.N_METRICS), dtype=torch.float32, device=dev), None, None, None)

    def fly():
        for k in range(ticks // chunk):
            kernels.rollout(plan, B, chunk, mc_gains=mc[:11], mc_mass=mc[11], mc_inertia=mc[12:15], carry=carry, resume=k > 0, out=result, **kw)
        return result
    return fly, ticks


def sample_table_rate(kernels, dev, peaks, B=100_000):
    """K3: the `get_trajectory()` tables of B config-1 missions in one pass (88 B per row written)."""
    import torch
    wp, vel = kernels.mc_missions(99, B, 4, device=dev)
    c, t, _ = kernels.minsnap_solve(wp, vel)
    offs = torch.arange(B + 1, dtype=torch.int32, device=dev) * 4
    rows, yaw0, total = kernels.table_meta(c, t.reshape(-1), offs, 0.01)
    roff = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    roff[1:] = torch.cumsum(total, 0)
    n_rows = int(roff[-1])
    ms, _ = event_ms(lambda: kernels.minsnap_sample(c, t.reshape(-1), offs, rows, roff, 0.01), reps=4, warm=3)
    gbs = n_rows * 88.0 / (ms * 1e-3) / 1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    return {"metric": "sampled table rows/s", "value": n_rows / (ms * 1e-3), "unit": "rows/s", "missions": B, "rows": n_rows, "kernel_ms": ms,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                         "kernel": "sample_table_kernel", "note": "88 B per (N, 11) row written; write-only traffic against the read+write copy peak"}}


def correction_rate(kernels, nat, dev, B=100_000, n_obs=4):
    """The device-side obstacle-correction loop (uavb_minsnap_correct_f64) on B config-1 missions x n_obs shared boxes placed in the
    mission volume: plan, sweep the sampled points, insert midpoints, re-plan only the missions that were hit, until clean.
    Wall-clock of the whole call (it synchronises once per round)."""
    import time
    import torch
    wp, vel = kernels.mc_missions(123, B, 4, device=dev)
    rng = np.random.default_rng(5)
    ctr, half = rng.uniform([4, 3, -4.5], [20, 11, -1.5], (n_obs, 3)), rng.uniform(0.15, 0.4, (n_obs, 3))
    boxes = torch.tensor(np.stack((ctr[:, 0] - half[:, 0], ctr[:, 0] + half[:, 0], ctr[:, 1] - half[:, 1], ctr[:, 1] + half[:, 1],
                                   ctr[:, 2] - half[:, 2], ctr[:, 2] + half[:, 2]), axis=-1), dtype=torch.float64, device=dev)
    cap = 17

    def run():
        w, n = kernels.fixed_pitch(wp, cap, dev)
        return kernels.minsnap_correct(w, n, vel, 0.01, boxes) + (n,)
    run()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(3):
        t0 = time.perf_counter(); c, t, status, rounds, n = run(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return {"metric": "collision-free plans/s", "value": B / best, "unit": "missions/s", "missions": B, "obstacles": n_obs, "ms": best * 1e3,
            "plan_rounds": rounds, "grown_fraction": float((n > 5).float().mean()), "too_many_fraction": float((status == nat.SOLVE_TOO_MANY).float().mean()),
            "waypoint_capacity": cap, "call": "uavb_minsnap_correct_f64 (K1 over work lists bucketed by spline count + sampled-point sweep + midpoint "
            "insertion on the device; one host synchronisation per round)"}


def rrt_rate(dev, B=16_384):
    """RRT*: B lab-volume missions, 1 500 iterations cap, through RRTStar.run() (host path buffers included)."""
    import time
    import torch
    from uav_ac_b200.planning.rrt import RRTStar
    from uav_ac_b200.simulation.scene import LAB_COURSE_OBSTACLES as OBS, PLANNING_BOUNDS as LIM
    rng = np.random.default_rng(0)
    s = np.round(rng.uniform(LIM[0] + [0.5, 0.5, 0.3], [3.0, 13.5, -0.5], (B, 3)), 2)
    g = np.round(rng.uniform([21.0, 0.5, -5.5], LIM[1] - [0.5, 0.5, 0.5], (B, 3)), 2)
    r = RRTStar(LIM, s, g, 1.5, 1500, OBS, seed=1)
    r.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); r.run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    it = r.stats[:, 0].astype(float)
    return {"metric": "RRT* missions/s", "value": B / dt, "unit": "missions/s", "missions": B, "ms": dt * 1e3, "found_fraction": float((r.status == 0).mean()),
            "tree_iterations_per_s": float(it.sum() / dt), "call": "RRTStar.run() incl. device->host copy of the used path prefix"}


def parity_block(kernels, nat, dev, n_rollouts=2048, n_missions=256):
    """Checker leg of the bench (outside every timed region): the CUDA path against the oracle on seeded inputs.
    * K1: n_missions config-1 missions against the reference's own MinimumSnap when oracle/_ref is built (else the NumPy oracle):
      norm-wise error vs method="solve" (gate 1e-9) and the fraction of missions on which the default lstsq branch agrees to 1e-6.
    * K2: n_rollouts Monte-Carlo lab_course rollouts (whole mission, 4 AABBs) against oracle/oracle_c.c: max |dpos|, max attitude angle,
      collision flags bit-exact."""
    import torch
    from oracle import c_port, flight_np, minsnap_np, ref_arm
    from uav_ac_b200.simulation.scene import LAB_COURSE_GOAL, LAB_COURSE_OBSTACLES, LAB_COURSE_START, LAB_COURSE_WAYPOINTS
    out = {}
    wp, vel = kernels.mc_missions(7, n_missions, 4, device=dev)
    c, t, st = kernels.minsnap_solve(wp, vel)
    cg, wpn, veln = c.cpu().numpy(), wp.cpu().numpy(), vel.cpu().numpy()
    use_ref = ref_arm.available()
    err_solve, ok_lstsq = 0.0, 0
    for i in range(n_missions):
        if use_ref:
            cs, _ = ref_arm.solve_lstsq(wpn[i], veln[i], "solve")
            cl, _ = ref_arm.solve_lstsq(wpn[i], veln[i], "lstsq")
        else:
            cs, _ = minsnap_np.solve_coeffs(wpn[i], veln[i], "solve")
            cl, _ = minsnap_np.solve_coeffs(wpn[i], veln[i], "lstsq")
        scale = np.abs(cs).max()
        err_solve = max(err_solve, float(np.abs(cg[i] - cs).max() / scale))
        ok_lstsq += bool(np.abs(cg[i] - cl).max() / scale < 1e-6)
    out["k1"] = {"missions": n_missions, "against": "reference MinimumSnap (oracle/_ref)" if use_ref else "oracle/minsnap_np.py",
                 "coeff_err_solve": err_solve, "gate": 1e-9, "lstsq_pass_fraction_1e-6": ok_lstsq / n_missions, "failed": int((st != 0).sum())}
    # K2
    B = n_rollouts
    rng = np.random.default_rng(11)
    gs, ms_, is_ = rng.uniform(0.8, 1.2, (B, 11)), rng.uniform(0.9, 1.1, B), rng.uniform(0.9, 1.1, (B, 3))
    veh = nat.default_vehicle()
    gains = torch.tensor((np.array(list(veh.gains))[None] * gs).T.astype(np.float32), device=dev).contiguous()
    mass = torch.tensor((veh.mass * ms_).astype(np.float32), device=dev)
    inertia = torch.tensor((np.array(list(veh.inertia))[None] * is_).T.astype(np.float32), device=dev).contiguous()
    plan, kw = lab_course(kernels, dev)
    n_ticks = FREQUENCY * int(plan.total_rows.item())
    kw["want_state"] = True
    res = kernels.rollout(plan, B, n_ticks, mc_gains=gains, mc_mass=mass, mc_inertia=inertia, **kw)
    torch.cuda.synchronize()
    tab = minsnap_np.mission_table(LAB_COURSE_WAYPOINTS, None, 3.0, 0.01, method="solve")
    # the vehicles exactly as the kernel saw them: fp32-rounded Monte-Carlo values
    g32, m32, i32 = gains.cpu().numpy().T.astype(np.float64), mass.cpu().numpy().astype(np.float64), inertia.cpu().numpy().T.astype(np.float64)
    vehicles = [flight_np.Vehicle().with_values(g32[i], m32[i], i32[i]) for i in range(B)]
    m_ref, X_ref = c_port.closed_loop_batch(vehicles, tab, np.asarray(LAB_COURSE_START, float), obstacles=LAB_COURSE_OBSTACLES, goal=LAB_COURSE_GOAL,
                                            threads=os.cpu_count() or 1)
    X = res.state.double().cpu().numpy().T                      # [B, 13]
    dpos = float(np.abs(X[:, :3] - X_ref[:, :3]).max())
    qa, qb = X[:, 3:7] / np.linalg.norm(X[:, 3:7], axis=1, keepdims=True), X_ref[:, 3:7] / np.linalg.norm(X_ref[:, 3:7], axis=1, keepdims=True)
    # angle of the relative rotation from the VECTOR part of conj(qb) * qa (well conditioned near zero, unlike acos of the dot product)
    vec = qb[:, :1] * qa[:, 1:] - qa[:, :1] * qb[:, 1:] - np.cross(qb[:, 1:], qa[:, 1:])
    dang = float((2 * np.arcsin(np.linalg.norm(vec, axis=1).clip(0, 1))).max())
    m = res.metrics.cpu().numpy()
    out["k2"] = {"rollouts": B, "ticks": n_ticks, "against": "oracle/oracle_c.c (fp64)", "dpos_max_m": dpos, "dang_max_rad": dang, "tol": 1e-4,
                 "flags_exact": bool(np.array_equal(m[:, 1] != 0, m_ref[:, 1] != 0)), "drmse_max": float(np.abs(m[:, 2] - m_ref[:, 2]).max())}
    return out

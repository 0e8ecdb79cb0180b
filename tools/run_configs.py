#!/usr/bin/env python
"""Run BASELINE.json configs[1..4] at FULL size on one GPU and report throughput plus size-independent checks.

    python tools/run_configs.py [--out profiles/r01_configs.json] [--c5-share 8]

configs[1]  10^6 random 5-waypoint solves (K1)                       -> solves/s, constraint residuals
configs[2]  10^5 lab_course rollouts, Monte-Carlo gains/mass/inertia  -> steps/s, mission report
configs[3]  10^6 rollouts, random waypoint sets + wind + AABB sets    -> steps/s, collision fraction, determinism under re-sharding
configs[4]  10^7 x 60 s x 1 kHz sharded 8 ways: ONE GPU's share (1.25e6 rollouts x 60 000 ticks, metrics only, chunked launches)
Timing: CUDA events on the launching stream, one warm-up launch.  Not the driver's bench (bench.py is); this script
documents that the full-size workloads fit and run, and what they deliver.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=2):
    import torch
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return best, out


def main():
    import numpy as np
    import torch
    from uav_ac_b200 import _native as nat, kernels, sharding
    from uav_ac_b200.simulation.scene import LAB_COURSE_GOAL, LAB_COURSE_OBSTACLES, LAB_COURSE_START, LAB_COURSE_WAYPOINTS
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--c5-share", type=int, default=8, help="number of GPUs configs[4] is sharded over (this process runs one share)")
    ap.add_argument("--skip-c5", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    report = {"gpu": torch.cuda.get_device_name(0)}
    f64 = dict(dtype=torch.float64, device=dev)

    # ---------------------------------------------------------------- configs[1]
    B, S = 1_000_000, 4
    wp, vel = kernels.mc_missions(2026, B, S)
    ms, (c, t, st) = timed(lambda: kernels.minsnap_solve(wp, vel))
    c4 = c.reshape(B, S, 8, 3)
    pw = torch.stack([t ** j for j in range(8)], dim=-1)
    end = (c4 * pw[..., None]).sum(dim=2)
    scale = c4.abs().amax(dim=(1, 2, 3)).clamp_min(1.0)
    report["configs[1]"] = {"missions": B, "splines": S, "kernel_ms": ms, "solves_per_s": B / (ms * 1e-3), "GBps_algorithmic": 928.0 * B / (ms * 1e-3) / 1e9,
                            "failed": int((st != 0).sum()), "max_endpoint_residual_rel": float(((end - wp[:, 1:]).abs().amax(dim=(1, 2)) / scale).max())}
    del c, t, c4, pw, end

    # ---------------------------------------------------------------- configs[2]
    B = 100_000
    veh = nat.default_vehicle()
    base = torch.tensor(list(veh.gains) + [veh.mass] + list(veh.inertia), dtype=torch.float32, device=dev)[:, None]
    mc = (kernels.mc_uniform(20261017, B, [0.8] * 11 + [0.9] * 4, [1.2] * 11 + [1.1] * 4) * base).contiguous()
    wpl = torch.tensor(LAB_COURSE_WAYPOINTS, **f64)
    v3 = torch.tensor([3.0], **f64)
    plan = kernels.plan_missions([(wpl[None, :2].contiguous(), v3), (wpl[None, 1:].contiguous(), v3)], 0.01, shared=True)
    n_ticks = 10 * int(plan.total_rows.item())
    obs = torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=dev)
    kw = dict(start=torch.tensor(LAB_COURSE_START, **f64), goal=torch.tensor(LAB_COURSE_GOAL, **f64), obstacles=obs, want_state=False)
    ms, res = timed(lambda: kernels.rollout(plan, B, n_ticks, mc_gains=mc[:11], mc_mass=mc[11], mc_inertia=mc[12:15], **kw))
    report["configs[2]"] = {"rollouts": B, "ticks": n_ticks, "kernel_ms": ms, "steps_per_s": B * n_ticks / (ms * 1e-3), **sharding.summarize(res.metrics)}

    # ---------------------------------------------------------------- configs[3]
    B = 1_000_000
    t0 = time.perf_counter()
    wp, vel = kernels.mc_missions(31, B, 4)
    ground = wp[:, 0].clone()
    ground[:, 2] = -0.021
    tk = torch.stack((ground, wp[:, 0]), dim=1).contiguous()
    plan4 = kernels.plan_missions([(tk, vel), (wp, vel)], 0.01)
    wind = kernels.mc_uniform(32, B, [-0.08] * 3, [0.08] * 3)
    rng = np.random.default_rng(8)
    ctr, half = rng.uniform([2, 2, -5], [22, 12, -1], (64, 6, 3)), rng.uniform(0.3, 1.2, (64, 6, 3))
    boxes = np.stack((ctr[..., 0] - half[..., 0], ctr[..., 0] + half[..., 0], ctr[..., 1] - half[..., 1], ctr[..., 1] + half[..., 1],
                      ctr[..., 2] - half[..., 2], ctr[..., 2] + half[..., 2]), axis=-1).astype(np.float32)
    boxes_t = torch.tensor(boxes, device=dev)
    sets = (torch.arange(B, device=dev, dtype=torch.int32) * 2654435761 % 64).to(torch.int32).abs() % 64
    torch.cuda.synchronize()
    plan_s = time.perf_counter() - t0
    n4 = 10 * int(plan4.total_rows.max().item())
    goal4 = wp[:, -1].contiguous()
    ms, res4 = timed(lambda: kernels.rollout(plan4, B, n4, start=ground.contiguous(), goal=goal4, mc_wind=wind, obstacles=boxes_t, obstacle_set=sets,
                                             want_state=False), reps=1)
    m4 = res4.metrics
    # determinism under re-sharding: rollouts [300000, 300512) flown alone must reproduce their rows bit for bit
    lo, n = 300_000, 512
    wps, vels = kernels.mc_missions(31, n, 4, index_base=lo)
    g2 = wps[:, 0].clone(); g2[:, 2] = -0.021
    plan_s2 = kernels.plan_missions([(torch.stack((g2, wps[:, 0]), dim=1).contiguous(), vels), (wps, vels)], 0.01)
    sub = kernels.rollout(plan_s2, n, n4, start=g2.contiguous(), goal=wps[:, -1].contiguous(), mc_wind=kernels.mc_uniform(32, n, [-0.08] * 3, [0.08] * 3, index_base=lo),
                          obstacles=boxes_t, obstacle_set=sets[lo:lo + n].contiguous(), want_state=False)
    report["configs[3]"] = {"rollouts": B, "ticks": n4, "kernel_ms": ms, "steps_per_s": B * n4 / (ms * 1e-3), "plan_and_inputs_s": plan_s,
                            "collision_fraction": float((m4[:, 1] > 0).float().mean()), "nonfinite_fraction": float((m4[:, 5] != 0).float().mean()),
                            "median_final_dist": float(m4[:, 0].median()), "resharded_rows_bit_identical": bool(torch.equal(sub.metrics, m4[lo:lo + n]))}
    del plan4, res4, m4, wp, tk, wind

    # ---------------------------------------------------------------- configs[4] (one GPU's share)
    if not args.skip_c5:
        B = 10_000_000 // args.c5_share
        ticks, chunk = 60_000, 20_000
        mc = (kernels.mc_uniform(5, B, [0.8] * 11 + [0.9] * 4, [1.2] * 11 + [1.1] * 4) * base).contiguous()
        carry = torch.empty((nat.CARRY_WORDS, B), dtype=torch.float32, device=dev)

        def fly():
            out = None
            for k in range(ticks // chunk):                        # chunked launches through the resumable carry block
                out = kernels.rollout(plan, B, chunk, mc_gains=mc[:11], mc_mass=mc[11], mc_inertia=mc[12:15], carry=carry, resume=k > 0, **kw)
            return out
        ms, res5 = timed(fly, reps=1)
        m5 = res5.metrics
        report["configs[4]"] = {"share_of": args.c5_share, "rollouts": B, "ticks": ticks, "chunks": ticks // chunk, "kernel_ms": ms,
                                "steps_per_s": B * ticks / (ms * 1e-3), "periods_ok": bool((m5[:, 7] == ticks // 10).all()),
                                "hover_final_dist_max": float(m5[:, 0].max()), **sharding.summarize(m5)}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()

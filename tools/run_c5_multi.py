#!/usr/bin/env python
"""BASELINE configs[4] as a whole: 10^7 rollouts x 60 s x 1 kHz (6e11 ticks), metrics only, sharded over the ranks of one node.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/run_c5_multi.py

Every rank flies its contiguous share of the global rollout index range (Monte-Carlo parameters keyed by the global index, so the
job's per-rollout results do not depend on the number of ranks), in three chunked launches through the resumable carry block; the
only collective is the final all-gather of the [B, 8] metrics.  Timing: CUDA events per rank around the rank's launches, maximum
over the ranks; the all-gather is timed separately.  Rank 0 prints one JSON object.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from uav_ac_b200 import _native as nat, kernels, sharding
    from uav_ac_b200.simulation.scene import LAB_COURSE_GOAL, LAB_COURSE_OBSTACLES, LAB_COURSE_START, LAB_COURSE_WAYPOINTS
    rank, local, world = sharding.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    total = int(os.environ.get("UAVB_C5_ROLLOUTS", 10_000_000))
    ticks, chunk = 60_000, 20_000
    lo, hi = sharding.shard_range(total, rank, world)
    B = hi - lo
    f64 = dict(dtype=torch.float64, device=dev)
    veh = nat.default_vehicle()
    base = torch.tensor(list(veh.gains) + [veh.mass] + list(veh.inertia), dtype=torch.float32, device=dev)[:, None]
    mc = (kernels.mc_uniform(5, B, [0.8] * 11 + [0.9] * 4, [1.2] * 11 + [1.1] * 4, index_base=lo) * base).contiguous()
    wpl = torch.tensor(LAB_COURSE_WAYPOINTS, **f64)
    v3 = torch.tensor([3.0], **f64)
    plan = kernels.plan_missions([(wpl[None, :2].contiguous(), v3), (wpl[None, 1:].contiguous(), v3)], 0.01, shared=True)
    obs = torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=dev)
    kw = dict(start=torch.tensor(LAB_COURSE_START, **f64), goal=torch.tensor(LAB_COURSE_GOAL, **f64), obstacles=obs, want_state=False)
    carry = torch.empty((nat.CARRY_WORDS, B), dtype=torch.float32, device=dev)

    def fly():
        out = None
        for k in range(ticks // chunk):
            out = kernels.rollout(plan, B, chunk, mc_gains=mc[:11], mc_mass=mc[11], mc_inertia=mc[12:15], carry=carry, resume=k > 0, **kw)
        return out

    fly()                                                      # warm-up (allocations, first-use set-up)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    res = fly()
    b.record()
    metrics = sharding.gather_metrics(res.metrics, total) if world > 1 else res.metrics
    c.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b), b.elapsed_time(c)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        fly_ms, gather_ms = float(t[0]), float(t[1])
        rep = {"workload": "BASELINE configs[4]", "rollouts": total, "ticks": ticks, "n_gpus": world, "rollouts_per_gpu": B, "chunks": ticks // chunk,
               "fly_ms_max_over_ranks": fly_ms, "gather_ms": gather_ms, "steps_per_s": total * float(ticks) / ((fly_ms + gather_ms) * 1e-3),
               "steps_per_s_per_gpu": total * float(ticks) / ((fly_ms + gather_ms) * 1e-3) / world,
               "periods_ok": bool((metrics[:, 7] == ticks // 10).all()), "hover_final_dist_max": float(metrics[:, 0].max()),
               **sharding.summarize(metrics)}
        print(json.dumps(rep, indent=1))
        out = os.path.join(ROOT, "gpurun_out", f"c5_n{world}.json")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out, "w") as f:
            json.dump(rep, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

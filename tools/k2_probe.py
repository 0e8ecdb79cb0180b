"""Development probe: the headline K2 workload (BASELINE configs[2]: shared lab_course mission, Monte-Carlo gains / mass / inertia,
4 AABBs, metrics only) flown a few times with CUDA-event timing -- the command the ncu captures of K2 wrap.

    python tools/k2_probe.py [B] [reps] [per_rollout_missions=0|1]
"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uav_ac_b200 import _native as nat, kernels
from uav_ac_b200.simulation.scene import LAB_COURSE_GOAL, LAB_COURSE_OBSTACLES, LAB_COURSE_START, LAB_COURSE_WAYPOINTS
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
f64 = dict(dtype=torch.float64, device=dev)
veh = nat.default_vehicle()
base = torch.tensor(list(veh.gains) + [veh.mass] + list(veh.inertia), dtype=torch.float32, device=dev)[:, None]
wpl = torch.tensor(LAB_COURSE_WAYPOINTS, **f64); v3 = torch.tensor([3.0], **f64)
plan = kernels.plan_missions([(wpl[None, :2].contiguous(), v3), (wpl[None, 1:].contiguous(), v3)], 0.01, shared=True)
n_ticks = 10 * int(plan.total_rows.item())
obs = torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=dev)
mc = (kernels.mc_uniform(1, B, [0.8] * 11 + [0.9] * 4, [1.2] * 11 + [1.1] * 4) * base).contiguous()
kw = dict(start=torch.tensor(LAB_COURSE_START, **f64), goal=torch.tensor(LAB_COURSE_GOAL, **f64), obstacles=obs, want_state=False,
          mc_gains=mc[:11], mc_mass=mc[11], mc_inertia=mc[12:15])
ts = []
for i in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); res = kernels.rollout(plan, B, n_ticks, **kw); b.record(); torch.cuda.synchronize()
    if i: ts.append(a.elapsed_time(b))
t = statistics.mean(ts) if ts else float("nan")
print(f"B={B} ticks={n_ticks}: {t:.3f} ms -> {B * n_ticks / t / 1e6:.1f} G ticks/s; reached {float((res.metrics[:, 0] < 0.5).float().mean()):.3f}", flush=True)

"""Development probe: the time-sliced rollout at several batch sizes and slice counts (UAVB_ROLLOUT_CHUNKS; 1 = a single slice,
i.e. the behaviour of a one-shot launch)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uav_ac_b200 import _native as nat, kernels
from uav_ac_b200.simulation.scene import LAB_COURSE_GOAL, LAB_COURSE_OBSTACLES, LAB_COURSE_START, LAB_COURSE_WAYPOINTS
dev = torch.device("cuda", 0)
f64 = dict(dtype=torch.float64, device=dev)
veh = nat.default_vehicle()
base = torch.tensor(list(veh.gains) + [veh.mass] + list(veh.inertia), dtype=torch.float32, device=dev)[:, None]
wpl = torch.tensor(LAB_COURSE_WAYPOINTS, **f64); v3 = torch.tensor([3.0], **f64)
plan = kernels.plan_missions([(wpl[None, :2].contiguous(), v3), (wpl[None, 1:].contiguous(), v3)], 0.01, shared=True)
n_ticks = 10 * int(plan.total_rows.item())
obs = torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=dev)
def run(B, **env):
    for k in ("UAVB_ROLLOUT_CHUNKS",):
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    mc = (kernels.mc_uniform(1, B, [0.8] * 11 + [0.9] * 4, [1.2] * 11 + [1.1] * 4) * base).contiguous()
    kw = dict(start=torch.tensor(LAB_COURSE_START, **f64), goal=torch.tensor(LAB_COURSE_GOAL, **f64), obstacles=obs, want_state=False,
              mc_gains=mc[:11], mc_mass=mc[11], mc_inertia=mc[12:15])
    ts = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); kernels.rollout(plan, B, n_ticks, **kw); b.record(); torch.cuda.synchronize()
        if i: ts.append(a.elapsed_time(b))
    t = statistics.mean(ts)
    print(f"B={B} {env}: {t:.3f} ms -> {B * n_ticks / t / 1e6:.1f} G ticks/s", flush=True)
for B in (37888, 75776, 100000, 151552):
    run(B, UAVB_ROLLOUT_CHUNKS=1)
for ch in (2, 3, 4, 6, 9, 12, 15, 25, 50):
    run(100000, UAVB_ROLLOUT_CHUNKS=ch)
run(100000)
run(500000, UAVB_ROLLOUT_CHUNKS=1)
run(500000, UAVB_ROLLOUT_CHUNKS=8)
run(500000)

"""Development probe: full-rate state log (52 B/tick) throughput of the rollout kernel at several batch sizes."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uav_ac_b200 import _native as nat, kernels
from uav_ac_b200.simulation.scene import LAB_COURSE_GOAL, LAB_COURSE_OBSTACLES, LAB_COURSE_START, LAB_COURSE_WAYPOINTS
dev = torch.device("cuda", 0)
f64 = dict(dtype=torch.float64, device=dev)
wpl = torch.tensor(LAB_COURSE_WAYPOINTS, **f64); v3 = torch.tensor([3.0], **f64)
plan = kernels.plan_missions([(wpl[None, :2].contiguous(), v3), (wpl[None, 1:].contiguous(), v3)], 0.01, shared=True)
obs = torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=dev)
for B, ticks, stride in [(16384, 4000, 1), (75776, 800, 1), (151552, 400, 1), (303104, 200, 1), (75776, 8000, 10), (75776, 10760, 50)]:
    log = torch.empty((ticks // stride, 13, B), dtype=torch.float32, device=dev)
    res = kernels.RolloutResult(None, None, log, None)
    ts = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        kernels.rollout(plan, B, ticks, start=torch.tensor(LAB_COURSE_START, **f64), goal=torch.tensor(LAB_COURSE_GOAL, **f64), obstacles=obs,
                        want_state=False, want_metrics=False, log_stride=stride, out=res, log_tma=int(os.environ.get('LOG_TMA', '0')))
        b.record(); torch.cuda.synchronize()
        if i: ts.append(a.elapsed_time(b))
    t = statistics.mean(ts)
    print(f"B={B} ticks={ticks} stride={stride}: {t:.3f} ms, log {52.0 * B * (ticks // stride) / t / 1e6:.0f} GB/s, {B * ticks / t / 1e6:.1f} G ticks/s", flush=True)
    del log, res

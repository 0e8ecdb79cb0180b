"""Development probe: how the event-timed duration of the full-rate log launch depends on what surrounds it."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uav_ac_b200 import kernels
from uav_ac_b200.simulation.scene import LAB_COURSE_GOAL, LAB_COURSE_OBSTACLES, LAB_COURSE_START, LAB_COURSE_WAYPOINTS
dev = torch.device("cuda", 0)
f64 = dict(dtype=torch.float64, device=dev)
wpl = torch.tensor(LAB_COURSE_WAYPOINTS, **f64); v3 = torch.tensor([3.0], **f64)
plan = kernels.plan_missions([(wpl[None, :2].contiguous(), v3), (wpl[None, 1:].contiguous(), v3)], 0.01, shared=True)
obs = torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=dev)
B, ticks = 151552, 400
log = torch.empty((ticks, 13, B), dtype=torch.float32, device=dev)
res = kernels.RolloutResult(None, None, log, None)
start, goal = torch.tensor(LAB_COURSE_START, **f64), torch.tensor(LAB_COURSE_GOAL, **f64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
spacer = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
def call():
    kernels.rollout(plan, B, ticks, start=start, goal=goal, obstacles=obs, want_state=False, want_metrics=False, log_stride=1, out=res)
def timed(pre, n=1):
    ts = []
    for i in range(5):
        pre(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): call()
        b.record(); torch.cuda.synchronize()
        if i: ts.append(a.elapsed_time(b) / n)
    return statistics.mean(ts)
gb = 52.0 * B * ticks / 1e6
for name, pre, n in (("bare", lambda i: None, 1), ("flush 256 MB", lambda i: flush.fill_(i), 1), ("8 x flush", lambda i: [flush.fill_(i + k) for k in range(8)], 1),
                     ("1 GB fill spacer", lambda i: spacer.fill_(i), 1), ("1 GB read spacer", lambda i: spacer.sum(), 1), ("4 back to back", lambda i: None, 4)):
    t = timed(pre, n)
    print(f"{name:18s}: {t:.4f} ms -> {gb / t:.0f} GB/s", flush=True)

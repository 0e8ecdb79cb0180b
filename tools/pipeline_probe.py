"""Development probe: where the 0.2 ms between the K2 launch (4.77 ms) and the bench step (4.97 ms) goes, and what removes it.

  A. K2 alone against the number of EQUAL time slices (rollout(..., n_slices=n); n = 0 is the library's policy: slices of decreasing
     length; results do not depend on the slicing)
  B. the speculative planner alone (device time of one plan_missions call with the queue primed)
  C. K steps of plan + K2: on one stream (bench.py of round 2 so far) / the NEXT step's plan on a side stream while K2 flies

    python tools/pipeline_probe.py [B] [K]
"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uav_ac_b200 import _native as nat, kernels
from uav_ac_b200.simulation.scene import LAB_COURSE_GOAL, LAB_COURSE_OBSTACLES, LAB_COURSE_START, LAB_COURSE_WAYPOINTS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda", 0)
f64 = dict(dtype=torch.float64, device=dev)
veh = nat.default_vehicle()
base = torch.tensor(list(veh.gains) + [veh.mass] + list(veh.inertia), dtype=torch.float32, device=dev)[:, None]
wpl = torch.tensor(LAB_COURSE_WAYPOINTS, **f64); v3 = torch.tensor([3.0], **f64)
obs = torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=dev)
obs64 = obs.double()
tables = [(wpl[None, :2].contiguous(), v3), (wpl[None, 1:].contiguous(), v3)]
plan0 = kernels.plan_missions(tables, 0.01, shared=True, obstacles=obs64)
rows = int(plan0.total_rows.item())
n_ticks = 10 * rows
mc = (kernels.mc_uniform(1, B, [0.8] * 11 + [0.9] * 4, [1.2] * 11 + [1.1] * 4) * base).contiguous()
kw = dict(start=torch.tensor(LAB_COURSE_START, **f64), goal=torch.tensor(LAB_COURSE_GOAL, **f64), obstacles=obs, want_state=False,
          mc_gains=mc[:11], mc_mass=mc[11], mc_inertia=mc[12:15])
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = [kernels.RolloutResult(torch.empty((B, nat.N_METRICS), dtype=torch.float32, device=dev), None, None, None) for _ in range(2)]


def space(i):
    for k in range(8):
        flush.fill_((i + k) & 0xFF)


def ev():
    return torch.cuda.Event(enable_timing=True)


# ---- A
for ns in [0] + [int(x) for x in os.environ.get("SLICES", "13,19,25,31,37,50,75,100").split(",")]:
    ts = []
    for i in range(5):
        space(i)
        a, b = ev(), ev()
        a.record(); kernels.rollout(plan0, B, n_ticks, n_slices=ns, out=out[0], **kw); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
    print(f"A  n_slices={ns:4d}: K2 {statistics.mean(ts):.4f} ms (min {min(ts):.4f})", flush=True)

if os.environ.get("ONLY_A"): sys.exit(0)

# ---- B
ts = []
plans = []
for i in range(6):
    space(i)
    a, b = ev(), ev()
    a.record(); plans.append(kernels.plan_missions(tables, 0.01, shared=True, table_rows=rows, obstacles=obs64)); b.record(); torch.cuda.synchronize()
    if i >= 2: ts.append(a.elapsed_time(b))
for p in plans: p.verify()
print(f"B  speculative plan alone: {statistics.mean(ts) * 1e3:.1f} us device time (min {min(ts) * 1e3:.1f})", flush=True)


# ---- C
def run_serial(K):
    plans = []
    evs = [(ev(), ev()) for _ in range(K)]
    space(0)
    for s in range(K):
        flush.fill_(s & 0xFF)
        evs[s][0].record()
        p = kernels.plan_missions(tables, 0.01, shared=True, table_rows=rows, obstacles=obs64)
        plans.append(p)
        kernels.rollout(p, B, n_ticks, out=out[s % 2], **kw)
        evs[s][1].record()
    torch.cuda.synchronize()
    for p in plans: p.verify()
    return sum(a.elapsed_time(b) for a, b in evs) / K


side = torch.cuda.Stream(dev)


def run_piped(K):
    main = torch.cuda.current_stream(dev)
    plans, ready = [], []
    evs = [(ev(), ev()) for _ in range(K)]

    def issue(after_event):
        side.wait_event(after_event)
        with torch.cuda.stream(side):
            p = kernels.plan_missions(tables, 0.01, shared=True, table_rows=rows, obstacles=obs64)
            e = torch.cuda.Event(); e.record(side)
        plans.append(p); ready.append(e)

    space(0)
    for s in range(K):
        flush.fill_(s & 0xFF)
        evs[s][0].record()
        if s == 0:
            issue(evs[0][0])                                  # the first plan of the region starts inside the region
        main.wait_event(ready[s])
        kernels.rollout(plans[s], B, n_ticks, out=out[s % 2], **kw)
        if s + 1 < K:
            issue(evs[s][0])                                  # behind this step's K2 launch in host order; the side stream does NOT wait for K2
        evs[s][1].record()
    torch.cuda.synchronize()
    for p in plans: p.verify()
    return sum(a.elapsed_time(b) for a, b in evs) / K


for name, fn in (("serial", run_serial), ("piped ", run_piped), ("serial", run_serial), ("piped ", run_piped)):
    fn(2)
    t = fn(K)
    print(f"C  {name}: {t:.4f} ms per step over {K} steps -> {B * n_ticks / t / 1e6:.1f} G steps/s", flush=True)
m0 = out[0].metrics.clone()
run_serial(2); ms = out[1].metrics.clone()
run_piped(2); mp = out[1].metrics.clone()
print("C  metrics identical serial/piped:", bool(torch.equal(ms, mp)), flush=True)

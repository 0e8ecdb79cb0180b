"""Development probe: BASELINE configs[3]-style rollouts (per-rollout missions, wind, per-rollout obstacle sets) for timing / ncu."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uav_ac_b200 import kernels
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
wp, vel = kernels.mc_missions(31, B, 4)
ground = wp[:, 0].clone(); ground[:, 2] = -0.021
plan = kernels.plan_missions([(torch.stack((ground, wp[:, 0]), dim=1).contiguous(), vel), (wp, vel)], 0.01)
wind = kernels.mc_uniform(32, B, [-0.08] * 3, [0.08] * 3)
rng = np.random.default_rng(8)
ctr, half = rng.uniform([2, 2, -5], [22, 12, -1], (64, 6, 3)), rng.uniform(0.3, 1.2, (64, 6, 3))
boxes = torch.tensor(np.stack((ctr[..., 0] - half[..., 0], ctr[..., 0] + half[..., 0], ctr[..., 1] - half[..., 1], ctr[..., 1] + half[..., 1],
                               ctr[..., 2] - half[..., 2], ctr[..., 2] + half[..., 2]), axis=-1).astype(np.float32), device=dev)
sets = (torch.arange(B, device=dev) % 64).to(torch.int32)
n = 10 * int(plan.total_rows.max().item())
ts = []
for i in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); kernels.rollout(plan, B, n, start=ground.contiguous(), goal=wp[:, -1].contiguous(), mc_wind=wind, obstacles=boxes, obstacle_set=sets, want_state=False); b.record()
    torch.cuda.synchronize()
    if i: ts.append(a.elapsed_time(b))
t = statistics.mean(ts)
print(f"configs[3]-style B={B} ticks={n}: {t:.2f} ms -> {B * n / t / 1e6:.1f} G ticks/s")

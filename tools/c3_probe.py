"""Development probe: BASELINE configs[3] (per-rollout missions, wind, 64 obstacle sets x 6 AABBs) at a chosen batch size.
    python tools/c3_probe.py [B] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_workloads as wl
from uav_ac_b200 import kernels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
fly, n_ticks, info = wl.config3(kernels, dev, B=B)
res = kernels.RolloutResult(torch.empty((B, 8), dtype=torch.float32, device=dev), None, None, None)
ms, _ = wl.event_ms(lambda: fly(res), reps=reps, warm=2)
m = res.metrics
print(f"B={B} ticks={n_ticks}: {ms:.2f} ms -> {B * n_ticks / ms / 1e6:.1f} G ticks/s; collisions {float((m[:, 1] > 0).float().mean()):.4f} failed {float((m[:, 5] != 0).float().mean()):.4f}")

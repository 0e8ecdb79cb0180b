"""Development probe: launch K1 on 10^6 config-2 missions a few times (for ncu captures and timing)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uav_ac_b200 import kernels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wp, vel = kernels.mc_missions(99, B, S)
c = torch.empty((B, 8 * S, 3), dtype=torch.float64, device=wp.device)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=wp.device)
ts = []
for i in range(6):
    flush.fill_(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); kernels.minsnap_solve(wp, vel); b.record(); torch.cuda.synchronize()
    if i > 1: ts.append(a.elapsed_time(b))
t = statistics.mean(ts)
print(f"K1 B={B} S={S}: {t:.4f} ms -> {B / t / 1e6:.3f} G solves/s, {(24 * S * 8 + (S + 1) * 24 + 8 + S * 8) * B / t / 1e6:.1f} GB/s algorithmic")

"""Development probe: table geometry + sampled (N, 11) tables (K3) for a batch of config-2 missions."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uav_ac_b200 import kernels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
wp, vel = kernels.mc_missions(99, B, 4)
c, t, _ = kernels.minsnap_solve(wp, vel)
offs = torch.arange(B + 1, dtype=torch.int32, device=wp.device) * 4
def ev(fn, n=8):
    ts = []
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        if i: ts.append(a.elapsed_time(b))
    return statistics.median(ts[2:]), out          # the first two calls allocate the (double-buffered) output
ms_meta, (rows, yaw0, total) = ev(lambda: kernels.table_meta(c, t.reshape(-1), offs, 0.01))
roff = torch.zeros(B + 1, dtype=torch.int32, device=wp.device); roff[1:] = torch.cumsum(total, 0)
n_rows = int(roff[-1])
ms_s, tab = ev(lambda: kernels.minsnap_sample(c, t.reshape(-1), offs, rows, roff, 0.01))
print(f"B={B}: rows {n_rows} ({n_rows * 88 / 1e9:.2f} GB)  table_meta {ms_meta:.3f} ms   sample {ms_s:.3f} ms -> {n_rows * 88 / ms_s / 1e6:.0f} GB/s written, {n_rows / ms_s / 1e6:.2f} G rows/s")

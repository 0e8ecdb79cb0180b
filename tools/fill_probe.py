"""Development probe: write-only HBM bandwidth (torch fill of a 4 GiB tensor) next to the read+write copy figure of MEASURED_PEAKS.json."""
import torch
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
for name, fn, nbytes in (("fill", lambda: x.fill_(1.0), x.numel() * 4), ("copy", lambda: y.copy_(x), 2 * x.numel() * 4)):
    ts = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{name}: {nbytes / min(ts) / 1e6:.0f} GB/s")

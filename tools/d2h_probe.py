"""Pinned-host copy bandwidth of this box (context for bench.py's e2e number)."""
import torch
n = 8294400
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 50
    print(f"{name}: {ms * 1e3:.1f} us per 8.3 MB frame = {n / ms * 1e-6:.1f} GB/s")

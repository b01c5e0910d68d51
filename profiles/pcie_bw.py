"""Pinned host <-> device copy bandwidth of the box (context for the e2e number of bench.py)."""
import torch

x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for name, src, dst in (("H2D", x, d), ("D2H", d, x)):
    for size in (8 << 20, 64 << 20, 256 << 20):
        dst[:size].copy_(src[:size], non_blocking=True)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dst[:size].copy_(src[:size], non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name} {size >> 20:4d} MiB: {5 * size / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")

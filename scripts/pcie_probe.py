"""H2D bandwidth of a 4 MB pinned buffer (what the e2e step moves per step at cfg2), one copy and 4 chunks."""
import time, torch
x = torch.empty(1_000_000, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
s = torch.cuda.Stream()
for chunks in (1, 4, 8):
    with torch.cuda.stream(s):
        for _ in range(5):
            d.copy_(x, non_blocking=True)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record(s)
        for _ in range(n):
            for c in range(chunks):
                a, b = c * x.numel() // chunks, (c + 1) * x.numel() // chunks
                d[a:b].copy_(x[a:b], non_blocking=True)
        e1.record(s)
        s.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        print(f"H2D 4 MB pinned in {chunks} chunk(s): {us:.1f} us -> {4.0 / us * 1e3:.1f} GB/s")

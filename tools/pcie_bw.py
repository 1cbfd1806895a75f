#!/usr/bin/env python
"""Developer tool: raw pinned D2H / H2D bandwidth of the box (context for the e2e number)."""
import time
import torch
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    print("%s pinned 1 GiB x5: %.1f GB/s" % (name, 5 * n / (time.perf_counter() - t0) / 1e9))
# chunked D2H on a side stream like CountEngine.draw_to_host
s = torch.cuda.Stream()
t0 = time.perf_counter()
with torch.cuda.stream(s):
    for i in range(0, n, 1 << 28):
        h[i:i + (1 << 28)].copy_(d[i:i + (1 << 28)], non_blocking=True)
s.synchronize()
print("D2H 4 x 256 MiB chunks on a side stream: %.1f GB/s" % (n / (time.perf_counter() - t0) / 1e9))

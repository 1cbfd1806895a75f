#!/usr/bin/env python
"""Time the streaming epilogues (transform, CSR fill) on a bench-like count matrix."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from prosstt_b200 import formats, stats as pstats

dev = torch.device("cuda:0")
n, G = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200000, 20000)
g = torch.Generator(device=dev).manual_seed(1)
X = (torch.rand((n, G), device=dev, generator=g) ** 6 * 30).to(torch.int32)      # ~45 % zeros, mean ~4
s = torch.exp(torch.randn(n, device=dev, generator=g) * 0.7)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = torch.empty((n, G), dtype=torch.float32, device=dev)
for mode in ("normalize", "normalize_log1p", "log1p"):
    ms = timed(lambda: pstats.transform_counts(X, s, mode, out=out))
    print("transform %-16s n=%d G=%d: %.2f ms  %.0f GB/s (read+write)" % (mode, n, G, ms, 8.0 * n * G / ms / 1e6))
st = pstats.count_stats(X)
nnz = int((G - st["cell_zeros"].long()).sum().item())
ms = timed(lambda: formats.to_csr(X, stats=st))
print("to_csr (given stats)      nnz=%.3e (%.0f%% zeros): %.2f ms  %.0f GB/s (read 4 B/count + write 8 B/nnz)"
      % (nnz, 100 - 100.0 * nnz / (n * G), ms, (4.0 * n * G + 8.0 * nnz) / ms / 1e6))

#!/usr/bin/env python
"""Developer tool: achieved HBM bandwidth of pst_rel_means (rel = W.H, exp*scale, per-gene max)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prosstt_b200 import _native as nat
dev = torch.device("cuda", 0)
for P, K, G in ((750, 10, 20000), (51000, 10, 30000), (20000, 70, 20000)):
    W = torch.randn(P, K, dtype=torch.float64, device=dev) * 0.3
    H = torch.rand(K, G, dtype=torch.float64, device=dev) * 0.2
    gs = torch.rand(G, dtype=torch.float64, device=dev) + 0.5
    rel = torch.empty(P, G, dtype=torch.float64, device=dev)
    m32 = torch.empty(P, G, dtype=torch.float32, device=dev)
    cmax = torch.full((G,), float("-inf"), dtype=torch.float64, device=dev)
    st = nat.stream_ptr(dev)
    for name, args, nbytes in (
        ("rel f64 + colmax", (nat.ptr(rel), None, None, nat.ptr(cmax)), 8 * P * G),
        ("means f32 only", (None, None, nat.ptr(m32), None), 4 * P * G),
        ("rel f64 + means f32 + colmax", (nat.ptr(rel), None, nat.ptr(m32), nat.ptr(cmax)), 12 * P * G)):
        for _ in range(2):
            nat.call("pst_rel_means", nat.ptr(W), nat.ptr(H), nat.ptr(gs), 0, P, K, G, *args, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            nat.call("pst_rel_means", nat.ptr(W), nat.ptr(H), nat.ptr(gs), 0, P, K, G, *args, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("P=%6d K=%2d G=%5d  %-30s %8.3f ms  %7.1f GB/s written" % (P, K, G, name, ms, nbytes / ms / 1e6))
    assert torch.allclose(rel, W @ H, rtol=1e-12, atol=1e-12)

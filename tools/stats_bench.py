#!/usr/bin/env python
"""Developer tool: achieved HBM read bandwidth of pst_count_stats."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prosstt_b200.stats import count_stats
for n, G in ((200000, 20000), (1000000, 20000)):
    X = torch.randint(0, 50, (n, G), dtype=torch.int32, device="cuda")
    count_stats(X); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        count_stats(X)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("n=%d G=%d: %.2f ms  %.0f GB/s read" % (n, G, ms, 4.0 * n * G / ms / 1e6))
    del X

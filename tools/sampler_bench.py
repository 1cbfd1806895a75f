#!/usr/bin/env python
"""Developer tool: time pst_draw_counts for each sampler on the bench tree (GPU box only)
and sanity-check the totals.  Not part of the product or the driver contract."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from prosstt_b200.session import DensitySession  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=200000)
ap.add_argument("--genes", type=int, default=20000)
ap.add_argument("--samplers", default="gamma_poisson,hybrid")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--scale-mean", type=float, default=0.0, help="mean of log library size (depth regime)")
a = ap.parse_args()
args = dict(bench.WORKLOADS["c4"], G=a.genes)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
t0 = time.time()
tree = bench.build_tree_gpu(args, dev)
alpha, beta = bench.gene_hyper(a.genes)
print("tree built in %.1fs" % (time.time() - t0), flush=True)
for name in a.samplers.split(","):
    s = DensitySession(tree, alpha, beta, a.cells, device=dev, sampler=name, scale_mean=a.scale_mean)
    s.step(1)
    torch.cuda.synchronize()
    ms = []
    for i in range(a.reps):
        s.step(2 + i)
        torch.cuda.synchronize()
        ms.append(s.last_draw_ms())
    s.engine.check()
    X = s.X
    # expected totals from the model: mu = means[row]*scaling
    mu_tot = (s.engine.means.double()[s.rows.long()].sum(dim=1) * s.s64).sum().item()
    tot = X.sum(dtype=torch.int64).item()
    zero = (X == 0).float().mean().item()
    # a position-weighted checksum: equal for two libraries only if they draw the same matrix
    weights = (torch.arange(a.genes, device=dev, dtype=torch.int64) % 1009) + 1
    check = int((X.to(torch.int64) * weights).sum().item()) % (1 << 61)
    print("%-14s draw %.2f ms  -> %.3e counts/s  (%.1f GB/s)  total/expected %.5f  zeros %.4f  max %d  checksum %x"
          % (name, min(ms), a.cells * a.genes / (min(ms) / 1e3), 4 * a.cells * a.genes / (min(ms) / 1e3) / 1e9,
             tot / mu_tot, zero, int(X.max().item()), check), flush=True)
    del s, X
    torch.cuda.empty_cache()

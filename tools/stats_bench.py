#!/usr/bin/env python
"""Cost of the per-gene summaries: the draw alone, the draw with the summaries fused in, and the draw followed
by the second pass (pst_count_stats) - bench tree, 200 000 cells x 20 000 genes unless told otherwise."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from prosstt_b200.session import DensitySession  # noqa: E402
from prosstt_b200.stats import count_stats, new_gene_stats  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=200000)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
w = dict(bench.WORKLOADS["c4"])
tree = bench.build_tree_gpu(w, dev)
alpha, beta = bench.gene_hyper(w["G"])
s = DensitySession(tree, alpha, beta, a.cells, device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


plain = timed(lambda: s.step(5))
fused = new_gene_stats(w["G"], dev)
with_fused = timed(lambda: s.step(5, gene_stats=fused))
second = timed(lambda: (s.step(5), count_stats(s.X)))
check = new_gene_stats(w["G"], dev)
s.step(5, gene_stats=check)
ref = count_stats(s.X)
same = all(torch.equal(check[k], ref[k]) for k in check)
print("draw %.2f ms | draw with fused per-gene summaries %.2f ms (+%.1f %%) | draw + pst_count_stats %.2f ms (+%.1f %%) | "
      "fused == second pass: %s" % (plain, with_fused, 100 * (with_fused / plain - 1), second, 100 * (second / plain - 1), same))

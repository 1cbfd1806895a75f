#!/usr/bin/env python
"""Developer tool: where does the host time of one sample_density(host_out=...) call go?"""
import argparse, cProfile, pstats, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from prosstt_b200 import simulation as sim
a = argparse.Namespace(genes=20000)
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
tree = bench.build_tree_gpu(dict(bench.WORKLOADS["c4"]), dev)
alpha, beta = bench.gene_hyper(a.genes)
n = 131072
bufs = (torch.empty((n, a.genes), dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int64).pin_memory(),
        torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory())
call = lambda s: sim.sample_density(tree, n, alpha=alpha, beta=beta, seed=s, device=dev, dtype=np.int32, host_out=bufs)
call(1); call(2)
t0 = time.perf_counter(); call(3); print("one call: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile(); pr.enable(); call(4); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

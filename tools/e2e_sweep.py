#!/usr/bin/env python
"""Developer tool: where the time of the staged device->host path goes.  Times simulation.sample_density into a
pinned int32 host matrix (uint8 transport) for a few staging-buffer sizes and host-thread counts, and the
stages of one call on their own (sampling + narrowing without the copy, copy alone, expansion alone)."""
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import bench
    from prosstt_b200 import simulation as sim
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    args = dict(bench.WORKLOADS["c4"])
    tree = bench.build_tree_gpu(args, dev)
    alpha, beta = bench.gene_hyper(args["G"])
    n, G = 131072, args["G"]
    hX = torch.empty((n, G), dtype=torch.int32).pin_memory()
    hpt = torch.empty(n, dtype=torch.int64).pin_memory()
    hco = torch.empty(n, dtype=torch.int32).pin_memory()
    hsc = torch.empty(n, dtype=torch.float64).pin_memory()
    for threads in [int(t) for t in os.environ.get("SWEEP_THREADS", "0").split(",")]:
        def call(seed):
            sim.sample_density(tree, n, alpha=alpha, beta=beta, device=dev, seed=seed, dtype=np.int32,
                               host_out=(hX, hpt, hco, hsc), host_transport=os.environ.get("SWEEP_TRANSPORT", "u8"),
                               host_threads=threads)
        call(1)
        t0 = time.perf_counter()
        for i in range(5):
            call(2 + i)
        dt = (time.perf_counter() - t0) / 5
        print("stage %s MB transport %s threads %d: %.1f ms per call -> %.3e counts/s"
              % (os.environ.get("PST_STAGE_MB", "256"), os.environ.get("SWEEP_TRANSPORT", "u8"), threads, dt * 1e3, n * G / dt),
              flush=True)
    sys.exit(0)

for mb in ("64", "128", "256", "512"):
    env = dict(os.environ, PST_STAGE_MB=mb, SWEEP_THREADS="0" if mb != "256" else "0,8,12")
    subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env)

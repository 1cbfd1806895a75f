#!/usr/bin/env python
"""Bit-exactness across GPU partitions on real multi-GPU hardware (run under torchrun).

Every rank samples its shard of one sample_density / sample_whole_tree / sample_pseudotime_series
call on its own GPU; rank 0 additionally samples the whole thing alone and compares the gathered
shards with it bit for bit (NCCL all_gather of the slabs = the optional epilogue of DESIGN.md §5)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from prosstt_b200 import simulation as sim  # noqa: E402
from prosstt_b200.sharding import allreduce_gene_stats, gather_counts  # noqa: E402
from prosstt_b200.stats import count_stats  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
w = dict(bp=3, chain=0, T=40, G=4000, K=8)
tree = bench.build_tree_gpu(w, dev)
alpha, beta = bench.gene_hyper(w["G"])
ok = True
for name, call in (
    ("sample_density", lambda **kw: sim.sample_density(tree, 8 * 1000 + 3, alpha=alpha, beta=beta, seed=5, device=dev, out="torch", **kw)),
    ("sample_whole_tree", lambda **kw: sim.sample_whole_tree(tree, 5, alpha=alpha, beta=beta, seed=6, device=dev, out="torch", **kw)),
    ("sample_pseudotime_series", lambda **kw: sim.sample_pseudotime_series(tree, [3000, 2001, 1000], [0, 40, 100], [3.0, 5.0, 8.0], alpha=alpha, beta=beta, seed=7, device=dev, out="torch", **kw)),
):
    X, pt, codes, sc = call(shard=(rank, world))
    total = torch.tensor([X.shape[0]], device=dev)
    dist.all_reduce(total)
    gathered = gather_counts(X, int(total.item()))               # NCCL over NVLink
    # the cheap epilogue: per-gene summaries of the local slab, summed over the ranks (NCCL all_reduce)
    st, n_tot = allreduce_gene_stats(count_stats(X), X.shape[0])
    if rank == 0:
        whole = call()[0]
        same = torch.equal(gathered, whole)
        ws = count_stats(whole)
        stats_same = n_tot == whole.shape[0] and all(torch.equal(st[k], ws[k]) for k in ("gene_sum", "gene_sumsq", "gene_zeros"))
        ok &= same and stats_same
        print("%-26s world=%d cells=%d genes=%d  gathered shards == single-GPU result: %s  all-reduced gene stats == "
              "single-GPU stats: %s  (sum %d)"
              % (name, world, whole.shape[0], whole.shape[1], same, stats_same, int(whole.sum(dtype=torch.int64))),
              flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

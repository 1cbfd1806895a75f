#!/usr/bin/env python
"""The README quick start as a runnable script (needs a B200 and the built library)."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prosstt_b200 import formats, stats, tree, simulation as sim, sim_utils as sut  # noqa: E402
from prosstt_b200.stats import count_stats, gene_mean_var  # noqa: E402

np.random.seed(42)
t = tree.Tree(topology=[[0, 1], [0, 2]], time={0: 50, 1: 50, 2: 50}, num_branches=3,
              branch_points=1, modules=10, G=10000)
uMs, Ws, H = sim.simulate_lineage(t, a=0.05)
scale = sut.simulate_base_gene_exp(t, uMs)
t.add_genes({b: np.exp(uMs[b]) * scale for b in t.branches})
X, pseudotime, branches, scalings = sim.sample_density(t, 10000, alpha=0.2, beta=2.0)
print("X", X.shape, X.dtype, "mean count %.3f" % X.mean(), "zeros %.3f" % (X == 0).mean())

# keep everything on the GPU and summarise it there
Xd, pt, codes, s = sim.sample_density(t, 10000, alpha=0.2, beta=2.0, out="torch")
st = count_stats(Xd)
mean, var = gene_mean_var(st, Xd.shape[0])
print("library size median %d, per-gene mean of means %.3f" % (int(st["cell_total"].median()), float(mean.mean())))

# normalise / log-transform on the device, export as CSR
logX = stats.normalize(Xd, s, log=True)
path = formats.save_sparse_npz(os.path.join(tempfile.mkdtemp(), "counts.npz"), Xd)
import scipy.sparse  # noqa: E402
back = scipy.sparse.load_npz(path)
assert back.shape == tuple(Xd.shape) and back.nnz == int((Xd != 0).sum())
print("logX", tuple(logX.shape), logX.dtype, "| csr nnz", back.nnz, "->", path)

# per-gene summaries accumulated inside the draw (no second pass over the matrix)
from prosstt_b200.session import DensitySession  # noqa: E402
from prosstt_b200.stats import new_gene_stats  # noqa: E402
sess = DensitySession(t, 0.2, 2.0, 10000)
gs = new_gene_stats(t.G, sess.dev)
sess.step(seed=7, gene_stats=gs)
sess.engine.check()
assert all(bool((gs[k] == count_stats(sess.X)[k]).all()) for k in gs)
print("fused per-gene summaries: total counts", int(gs["gene_sum"].sum()), "zeros", int(gs["gene_zeros"].sum()))

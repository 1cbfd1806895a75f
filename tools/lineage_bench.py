#!/usr/bin/env python
"""simulate_lineage at the C2 and C5 shapes: wall time of the level-batched device loop
beside the reference's algorithm on the host (the oracle port: Python walks + G scipy pearsonr calls per
sibling pair and attempt).  The oracle runs the C2 shape in full; at the C5 shape it is timed on one branch
of walks and one sibling pair and scaled (51 branches, 25 pairs, one attempt each: a lower bound)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import prosstt_oracle as orc  # noqa: E402
from prosstt_b200 import _native as nat, simulation as sim, tree as ptree  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
out = {}
for name in ("c2", "c5"):
    w = dict(bench.WORKLOADS[name])
    top, time_ = bench.topology(w)
    t = ptree.Tree(topology=top, time=time_, num_branches=len(time_), branch_points=w["bp"], modules=w["K"], G=w["G"])
    for rep in range(3):
        np.random.seed(bench.SEEDS["lineage"])
        torch.cuda.synchronize()
        l0 = nat.launch_count()
        t0 = time.perf_counter()
        state, H = sim.simulate_lineage(t, a=0.05, seed=bench.SEEDS["lineage"] + rep, device=dev, _return_state=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        launches = nat.launch_count() - l0
        del state
        torch.cuda.empty_cache()
    rec = {"branches": len(time_), "T": w["T"], "K": w["K"], "G": w["G"], "device_s": dt, "launches": launches}
    # host side
    ot = orc.OTree(top, time_, G=w["G"], modules=w["K"])
    rng = np.random.RandomState(bench.SEEDS["lineage"])
    if name == "c2":
        t0 = time.perf_counter()
        orc.simulate_lineage(ot, rng, a=0.05)
        rec["oracle_s"] = time.perf_counter() - t0
        rec["oracle_how"] = "full run of the oracle's simulate_lineage"
    else:
        t0 = time.perf_counter()
        wA = np.stack([orc.diffusion(w["T"], rng) for _ in range(w["K"])]).T
        walk_s = time.perf_counter() - t0
        wB = np.stack([orc.diffusion(w["T"], rng) for _ in range(w["K"])]).T
        Hh = rng.standard_gamma(0.05, size=w["K"] * w["G"]).reshape(w["K"], w["G"])
        t0 = time.perf_counter()
        orc.pearson_anticorrelated(np.dot(wA, Hh), np.dot(wB, Hh))
        pair_s = time.perf_counter() - t0
        nb = len(time_)
        rec["oracle_s"] = nb * walk_s + (nb - 1) // 2 * pair_s
        rec["oracle_how"] = ("scaled: %d branches x %.2f s of walks + %d sibling pairs x %.1f s of per-gene pearsonr, one "
                             "attempt each (lower bound)" % (nb, walk_s, (nb - 1) // 2, pair_s))
    # the reference itself calls scipy.stats.pearsonr once per gene and sibling pair (sim_utils.py:145-168): timed
    # here on 500 genes and scaled to G genes x sibling pairs (one attempt each) + one Python-level rvs call per
    # walk step (simulation.py:114-121, ~30 us each: SURVEY.md section 6)
    import scipy.stats
    T = w["T"]
    xa, xb = rng.normal(size=(T, 500)), rng.normal(size=(T, 500))
    t0 = time.perf_counter()
    for g in range(500):
        scipy.stats.pearsonr(xa[:, g], xb[:, g])
    per_gene = (time.perf_counter() - t0) / 500
    nb = len(time_)
    t0 = time.perf_counter()
    for _ in range(2000):
        scipy.stats.norm.rvs(loc=0, scale=0.1)
    per_step = (time.perf_counter() - t0) / 2000
    rec["reference_style_s"] = (nb - 1) // 2 * w["G"] * per_gene + nb * w["K"] * T * per_step
    rec["reference_style_how"] = ("%d sibling pairs x %d genes x %.0f us per scipy.stats.pearsonr call + %d walk steps x %.0f us "
                                  "per scipy rvs call, one attempt per branch (lower bound; the reference package is not "
                                  "on the GPU box)" % ((nb - 1) // 2, w["G"], per_gene * 1e6, nb * w["K"] * T, per_step * 1e6))
    rec["speedup_vs_oracle_port"] = rec["oracle_s"] / rec["device_s"]
    rec["speedup_vs_reference_style"] = rec["reference_style_s"] / rec["device_s"]
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "lineage_bench.json"), "w"))

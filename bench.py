#!/usr/bin/env python
"""Headline benchmark: NB counts sampled per second by sample_density on the BASELINE.json
config-4 shape (15-branch tree, T=50, K=10, G=20 000, 1M cells per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One process per GPU (torchrun for N>1; NCCL only for the barrier / max-over-ranks timing:
the path shards by cells with no data-path collective, so scaling is "weak": every rank
samples its own --cells slab with globally numbered cells).  A "step" is one complete
sample_density pass: Philox uniforms -> density index map -> library sizes -> NB draw of
cells x genes counts into HBM.  Prints ONE JSON line on rank 0.

--impl reference times the reference's CPU algorithm (the oracle port: NumPy gather +
get_pr_umi + legacy RandomState.negative_binomial, i.e. exactly what scipy's nbinom.rvs
executes at prosstt/simulation.py:647-648) on all host cores, on a bounded cell sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "NB counts sampled/sec (cells x genes / s), sample_density"
UNIT = "counts/s"
SEEDS = dict(tree=42, lineage=43, sampling=44)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=1000000, help="cells per GPU per step")
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--branch-points", type=int, default=7, help="7 bifurcations = 15 branches")
    ap.add_argument("--steps-per-branch", type=int, default=50)
    ap.add_argument("--programs", type=int, default=10)
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4: the headline config; c5: 50-branch tree, 1000 steps/branch, G=30000 (means table 6 GB)")
    ap.add_argument("--sampler", default=None)
    ap.add_argument("--e2e-cells", type=int, default=131072, help="cells per e2e step (host buffers)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-cells", type=int, default=4096, help="cells of the cpu_baseline sample")
    ap.add_argument("--ref-cells-per-core", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def apply_workload(a):
    if a.workload == "c5":        # BASELINE config 5 shape (many_branches_cells), cells per step as given
        a.branch_points, a.steps_per_branch, a.genes = 25, 1000, 30000      # 51 branches


def workload_name(a):
    return ("%s sample_density: %d-branch random binary tree x %d steps, K=%d, G=%d, %d cells/GPU"
            % (a.workload.upper(), 2 * a.branch_points + 1, a.steps_per_branch, a.programs, a.genes, a.cells))


def gene_hyper(G):
    """alpha, beta as in examples/generate_simN.py:94-95."""
    rng = np.random.RandomState(SEEDS["sampling"])
    alpha = np.exp(rng.normal(np.log(0.2), np.log(1.5), size=G))
    beta = np.exp(rng.normal(np.log(1.0), np.log(1.5), size=G)) + 1
    return alpha, beta


def topology(a):
    from prosstt_b200 import tree as ptree
    np.random.seed(SEEDS["tree"])
    top = [[int(p), int(c)] for p, c in ptree.Tree.gen_random_topology(a.branch_points)]
    time_ = {b: a.steps_per_branch for b in range(2 * a.branch_points + 1)}
    return top, time_


# ------------------------------------------------------------------------------ ours
def build_tree_gpu(a, dev):
    from prosstt_b200 import simulation as sim, sim_utils as sut, tree as ptree
    top, time_ = topology(a)
    t = ptree.Tree(topology=top, time=time_, num_branches=len(time_), branch_points=a.branch_points,
                   modules=a.programs, G=a.genes)
    np.random.seed(SEEDS["lineage"])
    if t.G * sum(time_.values()) > 5e7:
        # big tables stay in HBM (no host round trip of the P x G arrays)
        sim.default_gene_expression_on_device(t, seed=SEEDS["lineage"], device=dev)
        return t
    rel, W, H = sim.simulate_lineage(t, a=0.05, seed=SEEDS["lineage"], device=dev)
    scale = sut.simulate_base_gene_exp(t, rel)
    t.add_genes({b: np.exp(rel[b]) * scale for b in t.branches})
    return t


class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, read+write copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(cells, genes):
    """dram bytes per launch of the draw kernel from the committed ncu capture, scaled to
    this launch size (profiles/traffic.json: {"bytes_per_count": x, "source": ...})."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        rec = json.load(fh)
    return float(rec["bytes_per_count"]) * cells * genes


def cpu_baseline(a, sample_cells):
    """The oracle port of sample_density on ONE host core (the reference is single-threaded)."""
    from oracle import prosstt_oracle as orc
    ot, alpha, beta = build_tree_cpu(a)
    rng = np.random.RandomState(SEEDS["sampling"])
    orc.sample_density(ot, 64, alpha, beta, rng)           # warm-up
    t0 = time.perf_counter()
    X, _, _, _ = orc.sample_density(ot, sample_cells, alpha, beta, rng)
    dt = time.perf_counter() - t0
    return {"value": X.size / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d cells x %d genes of the same tree in %.1f s (oracle sample_density: "
                      "NumPy gather + get_pr_umi + legacy RandomState.negative_binomial)"
                      % (sample_cells, a.genes, dt)}


def run_ours(a):
    import torch
    import torch.distributed as dist
    from prosstt_b200 import _native as nat
    from prosstt_b200.session import DensitySession
    from prosstt_b200 import simulation as sim

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (ours) needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == a.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % a.gpus

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = a.sampler or sim.DEFAULT_SAMPLER
    t_setup = time.perf_counter()
    tree = build_tree_gpu(a, dev)
    torch.cuda.synchronize(dev)
    setup_s = time.perf_counter() - t_setup     # Tree + simulate_lineage + base expression + means (untimed set-up)
    alpha, beta = gene_hyper(a.genes)
    sess = DensitySession(tree, alpha, beta, a.cells, first=rank * a.cells, device=dev, sampler=sampler)
    seed = SEEDS["sampling"]

    for i in range(a.warmup):
        sess.step(seed + i)
    barrier()
    sess.engine.check()
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.25)
    launches0 = nat.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    draw_ms = []
    barrier()
    w0 = time.time()
    ev0.record()
    draw_events = []
    for i in range(a.steps):
        sess.t_draw = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        sess.step(seed + 100 + i)
        draw_events.append(sess.t_draw)
    ev1.record()
    barrier()
    w1 = time.time()
    launches = nat.launch_count() - launches0
    clock_rec = clocks.stop(w0, w1)
    sess.engine.check()
    ms = ev0.elapsed_time(ev1)
    draw_ms = [e0.elapsed_time(e1) for e0, e1 in draw_events]
    if world > 1:
        tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    counts_per_step = float(a.cells) * a.genes * world
    value = counts_per_step * a.steps / (ms / 1e3)

    # ---- end to end through the public API path with HOST buffers -----------------------
    e2e = None
    if not a.no_e2e:
        ecells = min(a.e2e_cells, a.cells)
        hpt = torch.empty(ecells, dtype=torch.int64).pin_memory()
        hco = torch.empty(ecells, dtype=torch.int32).pin_memory()
        hsc = torch.empty(ecells, dtype=torch.float64).pin_memory()
        h2d = 16 * sess.tables.P + 8 * a.genes    # per call: cdf f64 + pos_pt/pos_branch i32 [P], alpha/beta-1 f32 [G]

        def run_e2e(x_dtype):
            hX = torch.empty((ecells, a.genes), dtype=x_dtype).pin_memory()
            overflow = {}

            def api_call(seed_):
                # the reference-facing call: sample_density(tree, no_cells, alpha, beta) with host outputs
                return sim.sample_density(tree, ecells * world, alpha=alpha, beta=beta, seed=seed_, device=dev,
                                          shard=(rank, world), dtype=np.int32, sampler=sampler,
                                          host_out=(hX, hpt, hco, hsc, overflow))
            api_call(seed)                                        # warm-up
            barrier()
            t0 = time.perf_counter()
            for i in range(a.e2e_steps):
                api_call(seed + 200 + i)
            barrier()
            dt = time.perf_counter() - t0
            if world > 1:
                tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dt = float(tmax.item())
            listed = len(overflow.get("index", ()))
            d2h = hX.numel() * hX.element_size() + ecells * (8 + 4 + 8) + listed * 12
            return float(ecells) * a.genes * world * a.e2e_steps / dt, d2h, listed

        v32, d2h, _ = run_e2e(torch.int32)
        v16, d2h16, listed = run_e2e(torch.uint16)
        v8, d2h8, listed8 = run_e2e(torch.uint8)
        e2e = {"value": v32, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "cells_per_step_per_gpu": ecells, "steps": a.e2e_steps,
               "note": "simulation.sample_density(tree, N, alpha, beta, host_out=pinned buffers): tree tables, "
                       "cdf and gene parameters uploaded per call; int32 counts + pseudotime + branch + "
                       "scalings land in host memory (chunked, copy overlapped with sampling)",
               "narrow_u16": {"value": v16, "unit": UNIT, "d2h_bytes_per_step": int(d2h16),
                              "overflow_entries_last_step": int(listed),
                              "note": "same call with a uint16 host matrix: min(count, 65535) plus an exact "
                                      "(index, value) list of the saturated elements; lossless, half the PCIe bytes"},
               "narrow_u8": {"value": v8, "unit": UNIT, "d2h_bytes_per_step": int(d2h8),
                             "overflow_entries_last_step": int(listed8),
                             "note": "uint8 host matrix, min(count, 255) plus the exact list of the counts >= 255; "
                                     "lossless, a quarter of the PCIe bytes"}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    kernel_ms = float(np.mean(draw_ms))
    algo_bytes = 4.0 * a.cells * a.genes
    achieved = algo_bytes / (kernel_ms / 1e3) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "sampler": sampler, "output": "int32 (N,G) resident in HBM",
                   "seeds": SEEDS, "cells_per_gpu": a.cells, "genes": a.genes,
                   "tree_lineage_means_setup_s": round(setup_s, 3),
                   "l2": "every step writes a %.1f GB count slab >> 126 MB L2 (self-flushing); fp32 means "
                         "table %.0f MB, cells visited grouped by tree row"
                         % (algo_bytes / 1e9, tree.G * sess.tables.P * 4 / 1e6)},
        "clocks": clock_rec, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": profiled_traffic(a.cells, a.genes),
                     "kernel": "draw_counts (%s)" % sampler, "kernel_ms": kernel_ms,
                     "kernel_share_of_step": kernel_ms * a.steps / ms,
                     "algorithmic_bytes_per_count": 4, "peak_source": peak_src},
    }
    if not a.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(a, a.cpu_cells)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ reference arm
_REF = {}


def build_tree_cpu(a):
    """Same tree shape on the CPU with the oracle's simulate_lineage (legacy stream)."""
    from oracle import prosstt_oracle as orc
    top, time_ = topology(a)
    ot = orc.OTree(top, time_, G=a.genes, modules=a.programs)
    rng = np.random.RandomState(SEEDS["lineage"])
    rel, W, H = orc.simulate_lineage(ot, rng, a=0.05)
    cap = orc.max_rel_exp(ot, rel)
    scale, _ = orc.base_gene_exp_from_normals(cap, rng.normal(0.8, 1.0, size=20 * a.genes))
    ot.means = orc.absolute_means(rel, scale)
    alpha, beta = gene_hyper(a.genes)
    return ot, alpha, beta


def _ref_worker(job):
    from oracle import prosstt_oracle as orc
    seed, cells = job
    ot, alpha, beta = _REF["state"]
    X, _, _, _ = orc.sample_density(ot, cells, alpha, beta, np.random.RandomState(seed))
    return int(X.size), int(X.sum() & 0xFFFF)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    _REF["state"] = build_tree_cpu(a)
    ctx = mp.get_context("fork")
    per = a.ref_cells_per_core
    with ctx.Pool(cores) as pool:
        def step(i):
            jobs = [(1000003 * (i + 1) + w, per) for w in range(cores)]
            return sum(n for n, _ in pool.map(_ref_worker, jobs, chunksize=1))
        for i in range(a.warmup):
            step(i)
        t0 = time.perf_counter()
        total = 0
        for i in range(a.steps):
            total += step(a.warmup + i)
        dt = time.perf_counter() - t0
    value = total / dt
    sample = ("%d host processes x %d cells x %d genes per step (disjoint cell shards, own seeds); "
              "oracle port of draw_counts = NumPy gather + get_pr_umi + legacy "
              "RandomState.negative_binomial" % (cores, per, a.genes))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
           "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": {"workload": workload_name(a), "cells_per_step": per * cores},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


if __name__ == "__main__":
    args = parse_args()
    apply_workload(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

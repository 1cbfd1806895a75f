#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

The reference (soedinglab/prosstt 1.2.0) ships no tests or golden vectors for
the simulation hot path, so the pins are produced here
from the live reference under fixed `np.random.seed`s.  Two shims are needed to
import it with this image's libraries:
  * `newick` is not installed      -> tests/golden/_shim/newick.py
  * NumPy >= 2 removed `np.Inf`     -> patched below (tree_utils.py:224-225)

Every random stage is recorded together with the raw draws it consumed (by
replaying the legacy MT19937 state), so that the deterministic kernels can be
fed the reference's own draws (north star, correctness part 1 and 2).

Outputs (all small, committed):
  maps.json        integer maps of a set of trees
  lineage_*.npz    walk draws, raw walks, carried programs, H, rel means, M
  sampling_*.npz   index maps of the three samplers with the uniforms/normals used
  counts_*.npz     a complete reference sample_density / whole_tree run under a seed
  nbparams.npz     get_pr_umi input/output vectors
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, "/root/reference")
np.Inf = np.inf  # reference uses np.Inf (tree_utils.py:224-225)

warnings.filterwarnings("ignore")

from prosstt import tree as rtree  # noqa: E402
from prosstt import simulation as rsim  # noqa: E402
from prosstt import sim_utils as rsut  # noqa: E402
from prosstt import count_model as rcm  # noqa: E402

assert rtree.__file__.startswith("/root/reference"), rtree.__file__


def jsonable(x):
    if isinstance(x, dict):
        return [[jsonable(k), jsonable(v)] for k, v in x.items()]
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, np.ndarray):
        return [jsonable(v) for v in x.tolist()]
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    if isinstance(x, (np.str_,)):
        return str(x)
    return x


# --------------------------------------------------------------------------
# 1. integer maps
# --------------------------------------------------------------------------
def tree_cases():
    cases = []
    cases.append(dict(name="doc_default", topology=[[0, 1], [0, 2]],
                      time={0: 40, 1: 40, 2: 40}, branch_points=1))
    cases.append(dict(name="doc_assign5", topology=[[0, 1], [0, 2], [2, 3], [2, 4]],
                      time={0: 10, 1: 20, 2: 8, 3: 25, 4: 15}, branch_points=2))
    cases.append(dict(name="probe_abc", topology=[["A", "B"], ["A", "C"]],
                      time={"A": 3, "B": 4, "C": 2}, branch_points=1))
    cases.append(dict(name="linear3", topology=[[0, 1], [1, 2]],
                      time={0: 7, 1: 5, 2: 9}, branch_points=0))
    cases.append(dict(name="multi5", topology=[["r", "a"], ["r", "b"], ["r", "c"], ["r", "d"], ["r", "e"]],
                      time={"r": 6, "a": 3, "b": 9, "c": 9, "d": 2, "e": 5}, branch_points=1))
    cases.append(dict(name="fork_chains", topology=[[0, 1], [0, 2], [1, 3], [2, 4], [4, 5]],
                      time={0: 5, 1: 4, 2: 6, 3: 8, 4: 2, 5: 3}, branch_points=1))
    cases.append(dict(name="equal_T50_bp2", topology=None, seed=11, bp=2, T=50))
    for s, bp in [(1, 1), (2, 2), (3, 3), (4, 5), (5, 7), (6, 12)]:
        cases.append(dict(name="random_s%d_bp%d" % (s, bp), topology=None, seed=s, bp=bp, T=None))
    return cases


def build_ref_tree(case, G=7, modules=3):
    if case["topology"] is None:
        np.random.seed(case["seed"])
        top = rtree.Tree.gen_random_topology(case["bp"])
        nb = 2 * case["bp"] + 1
        if case["T"] is None:
            lens = np.random.randint(2, 60, size=nb)
        else:
            lens = [case["T"]] * nb
        time = {int(b): int(l) for b, l in zip(range(nb), lens)}
        top = [[int(a), int(b)] for a, b in top]
        bp = case["bp"]
    else:
        top, time, bp = case["topology"], case["time"], case["branch_points"]
    t = rtree.Tree(topology=top, time=time, num_branches=len(time),
                   branch_points=bp, modules=modules, G=G)
    return t, top, time, bp


def gen_maps():
    out = []
    for case in tree_cases():
        t, top, time, bp = build_ref_tree(case)
        tz = t.populate_timezone()
        bt = t.branch_times()
        assign = rsut.assign_branches(bt, tz)
        pt, br = rsim.cover_whole_tree(t)
        rec = dict(
            name=case["name"], topology=top, time=jsonable(time), branch_points=bp,
            branches=jsonable(t.branches), root=jsonable(t.root),
            branch_times=jsonable(dict(bt)), timezone=jsonable(tz),
            assignments=jsonable(dict(assign)),
            cover_pt=jsonable(pt), cover_br=jsonable(br),
            max_time=int(t.get_max_time()),
            bfs=jsonable(rsut.breadth_first_branches(t)),
            paths=jsonable(t.paths(t.root)),
            parallel=jsonable({k: v for k, v in t.get_parallel_branches().items()}),
            density_sum=float(sum(np.sum(v) for v in t.density.values())),
        )
        out.append(rec)
    with open(os.path.join(HERE, "maps.json"), "w") as fh:
        json.dump(out, fh, indent=0, separators=(",", ":"))
    print("maps.json:", len(out), "trees")


# --------------------------------------------------------------------------
# 2. lineage: walks + carry + W.H + exp*scale, with the raw draws recorded
# --------------------------------------------------------------------------
class DiffusionTap:
    """Wraps reference `simulation.diffusion` (simulation.py:89-124): saves the MT
    state, runs the reference, then replays the state to recover the raw draws in
    the documented order U(0,1.5), N(0,.2), U(0,1), (T-1) x N(0, 2/T)."""

    def __init__(self):
        self.orig = rsim.diffusion
        self.calls = []

    def __call__(self, steps):
        state = np.random.get_state()
        walk = self.orig(steps)
        after = np.random.get_state()
        np.random.set_state(state)
        u0 = np.random.uniform(0, 1.5)
        v0 = np.random.normal(0, 0.2)
        eta = np.random.uniform()
        eps = np.array([np.random.normal(0, 2 / steps) for _ in range(steps - 1)])
        # the replay must land on the same state and rebuild the same walk
        assert all(np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b
                   for a, b in zip(np.random.get_state(), after))
        chk = np.zeros(steps)
        vel = np.zeros(steps)
        chk[0] = np.log(u0)
        vel[0] = v0
        for t in range(steps - 1):
            chk[t + 1] = chk[t] + vel[t]
            vel[t + 1] = eta * vel[t] + eps[t]
        assert np.array_equal(chk, walk)
        self.calls.append(dict(steps=steps, u0=u0, v0=v0, eta=eta, eps=eps, walk=walk.copy()))
        return walk


def gen_lineage(name, seed, case, G, K, a, cutoff):
    np.random.seed(seed)
    t, top, time, bp = build_ref_tree(case, G=G, modules=K)
    tap = DiffusionTap()
    rsim.diffusion = tap
    accepted = {}
    orig_adjust = rsut.adjust_to_parent

    def adjust_tap(programs, current, topology):
        # called right after sim_expr_branch (simulation.py:265-266 / 275-276):
        # the last K diffusion calls are this attempt's raw walks
        accepted[current] = (len(tap.calls) - K, programs[current].copy())
        return orig_adjust(programs, current, topology)

    rsut.adjust_to_parent = adjust_tap
    try:
        rel, Ws, H = rsim.simulate_lineage(t, a=a, rel_exp_cutoff=cutoff, intra_branch_tol=0)
    finally:
        rsim.diffusion = tap.orig
        rsut.adjust_to_parent = orig_adjust
    bfs = list(rel.index)
    gene_scale = rsut.simulate_base_gene_exp(t, rel)
    maxes = np.max(rsut.max_relat_exp(t, rel), axis=1)
    M = {b: np.exp(rel[b]) * gene_scale for b in t.branches}
    t.add_genes(M)
    rec = dict(seed=seed, G=G, K=K, a=a, cutoff=cutoff,
               topology=np.array(top), branches=np.array(t.branches),
               times=np.array([time[b] for b in t.branches]),
               bfs=np.array(bfs), H=H, gene_scale=gene_scale, max_rel_exp=maxes,
               attempts=len(tap.calls) // K)
    for b in t.branches:
        first, raw = accepted[b]
        calls = tap.calls[first:first + K]
        T = time[b]
        rec["u0_%s" % b] = np.array([c["u0"] for c in calls])
        rec["v0_%s" % b] = np.array([c["v0"] for c in calls])
        rec["eta_%s" % b] = np.array([c["eta"] for c in calls])
        rec["eps_%s" % b] = np.stack([c["eps"] for c in calls]) if T > 1 else np.zeros((K, 0))
        rec["raw_%s" % b] = raw                # (T, K) before the parent carry
        rec["W_%s" % b] = Ws[b]                # (T, K) after the parent carry
        rec["rel_%s" % b] = rel[b]             # (T, G)
        rec["M_%s" % b] = M[b]                 # (T, G)
        assert np.array_equal(raw, np.stack([c["walk"] for c in calls]).T)
    np.savez_compressed(os.path.join(HERE, "lineage_%s.npz" % name), **rec)
    print("lineage_%s.npz: attempts=%d" % (name, rec["attempts"]))
    return t, rel, H, gene_scale


# --------------------------------------------------------------------------
# 3. samplers' index maps with the draws they consumed
# --------------------------------------------------------------------------
def gen_sampling(name, t, seed):
    rec = {}
    G = t.G
    # non-uniform density (density_sampling.ipynb cell 8 style: assigned directly)
    rng = np.random.RandomState(1000 + seed)
    dens = {b: rng.uniform(0.2, 1.0, size=t.time[b]) for b in t.branches}
    tot = sum(np.sum(v) for v in dens.values())
    dens = {b: v / tot for b, v in dens.items()}
    t.set_density(dens)
    rec["density"] = np.concatenate([dens[b] for b in t.branches])
    alpha = np.exp(rng.normal(np.log(0.2), np.log(1.5), size=G))
    beta = np.exp(rng.normal(np.log(1.0), np.log(1.5), size=G)) + 1
    rec["alpha"], rec["beta"] = alpha, beta

    # --- sample_density (simulation.py:416-471) ---
    N = 400
    np.random.seed(seed)
    st = np.random.get_state()
    X, pt, br, sc = rsim.sample_density(t, N, alpha=alpha, beta=beta)
    np.random.set_state(st)
    u = np.random.random_sample(N)            # inside random.choice (simulation.py:464)
    z = np.random.normal(0.0, 0.7, size=N)    # calc_scalings (sim_utils.py:495)
    assert np.array_equal(np.exp(z), sc)
    rec.update(dens_seed=seed, dens_N=N, dens_u=u, dens_z=z, dens_pt=pt,
               dens_br=np.array(br), dens_scalings=sc, dens_X=X)

    # --- sample_pseudotime_series (simulation.py:319-413; sim_utils.py:342-403) ---
    mt = t.get_max_time()
    points = [0, mt // 3, (2 * mt) // 3, mt - 1]
    cells = [60, 50, 70, 40]
    std = [3.0, 5.0, 2.5, 6.0]
    np.random.seed(seed + 1)
    st = np.random.get_state()
    X, pt, br, sc = rsim.sample_pseudotime_series(t, cells, points, std, alpha=alpha, beta=beta)
    np.random.set_state(st)
    zt = np.concatenate([np.random.normal(p, s, size=n) for p, n, s in zip(points, cells, std)])
    up = np.random.random_sample(len(pt))     # one uniform per cell in pick_branch (sim_utils.py:399)
    z = np.random.normal(0.0, 0.7, size=len(pt))
    assert np.array_equal(np.exp(z), sc)
    rec.update(ser_seed=seed + 1, ser_points=np.array(points), ser_cells=np.array(cells),
               ser_std=np.array(std), ser_zt=zt, ser_upick=up, ser_z=z, ser_pt=np.array(pt),
               ser_br=np.array(br), ser_scalings=sc, ser_X=X)

    # --- sample_whole_tree (simulation.py:474-548) ---
    np.random.seed(seed + 2)
    X, pt, br, sc = rsim.sample_whole_tree(t, 2, alpha=alpha, beta=beta)
    rec.update(wt_seed=seed + 2, wt_n=2, wt_pt=np.array(pt), wt_br=np.array(br),
               wt_scalings=sc, wt_X=X)
    np.savez_compressed(os.path.join(HERE, "sampling_%s.npz" % name), **rec)
    print("sampling_%s.npz" % name, "X mean", rec["dens_X"].mean())


def gen_nbparams():
    rng = np.random.RandomState(7)
    G = 64
    a = np.exp(rng.normal(np.log(0.2), np.log(1.5), size=G))
    b = np.exp(rng.normal(np.log(2.0), np.log(1.5), size=G)) + 1
    m = np.exp(rng.normal(0.5, 3.0, size=G))
    m[:4] = [1e-10, 1e-6, 3e4, 1.6e5]
    p, r = rcm.get_pr_umi(a, b, m)
    # Poisson limit used by linear.ipynb cell 23: alpha=0, beta=1+10e-9
    a2 = np.zeros(G)
    b2 = np.full(G, 1 + 10e-9)
    p2, r2 = rcm.get_pr_umi(a2, b2, m)
    np.random.seed(5)
    ga, gb = rcm.generate_negbin_params(type("T", (), {"G": G})(), mean_alpha=0.2, mean_beta=2)
    np.random.seed(5)
    za = np.random.normal(np.log(0.2), np.log(1.5), size=G)
    zb = np.random.normal(np.log(2), np.log(1.5), size=G)
    assert np.array_equal(np.exp(za), ga) and np.array_equal(np.exp(zb) + 1, gb)
    np.savez_compressed(os.path.join(HERE, "nbparams.npz"), a=a, b=b, m=m, p=p, r=r,
                        a2=a2, b2=b2, p2=p2, r2=r2, gen_za=za, gen_zb=zb, gen_alpha=ga, gen_beta=gb)
    print("nbparams.npz")


def gen_written_files():
    """Files written by the reference's own writers (tree_utils.py:59-173): the byte-level pin of the
    on-disk formats."""
    import tempfile
    from prosstt import tree_utils as rtu
    sys.path.insert(0, os.path.dirname(HERE))
    from conftest import writer_inputs
    X, uMs, H, labs, brns, scal, gscale, alpha, beta = writer_inputs()
    t = rtree.Tree(topology=[["A", "B"]], time={"A": 3, "B": 2}, num_branches=2, branch_points=0, modules=2, G=5)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        rtu.save_matrices("job", tmp, X, uMs, H)
        rtu.save_cell_params("job", tmp, labs, brns, scal)
        rtu.save_gene_params("job", tmp, gscale, alpha, beta)
        rtu.save_params("job", tmp, t, 7)
        for name in sorted(os.listdir(tmp)):
            with open(os.path.join(tmp, name)) as fh:
                out[name] = fh.read()
    with open(os.path.join(HERE, "written_files.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("written_files.json", sorted(out))


if __name__ == "__main__":
    gen_maps()
    gen_written_files()
    cases = {c["name"]: c for c in tree_cases()}
    small = dict(name="abc", topology=[["A", "B"], ["A", "C"]],
                 time={"A": 9, "B": 12, "C": 7}, branch_points=1)
    t, rel, H, gs = gen_lineage("abc", seed=92, case=small, G=48, K=4, a=0.05, cutoff=8)
    gen_sampling("abc", t, seed=21)
    t, rel, H, gs = gen_lineage("bp2", seed=42, case=cases["doc_assign5"], G=64, K=6, a=0.05, cutoff=8)
    gen_sampling("bp2", t, seed=31)
    t, rel, H, gs = gen_lineage("fork", seed=2018, case=cases["fork_chains"], G=40, K=10, a=0.05, cutoff=8)
    gen_sampling("fork", t, seed=41)
    gen_nbparams()

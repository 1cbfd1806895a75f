"""Simulation utilities: tree walking, parent carry, relative means, base expression,
library sizes, branch picking.

Mirror of the hot-path part of prosstt/sim_utils.py.  Numeric work runs on the GPU
through the C ABI; tree bookkeeping stays on the host.  Out of scope (SURVEY.md section 2,
row 4): print_progress, learn_data_summary, commited_branches.
"""
import collections
import numbers
from collections import defaultdict

import numpy as np
import torch

from prosstt_b200 import _native as nat
from prosstt_b200.device import TreeTables, tree_tables


# ----------------------------------------------------------------------------- tree walking
def breadth_first_branches(tree):
    """Branches ordered by depth below the root; ties keep tree.branches order, branches
    the topology never reaches come first (sim_utils.py:545-567)."""
    depth = {b: -1 for b in tree.branches}
    depth[tree.root] = 0
    kids = defaultdict(list)
    for parent, child in tree.topology:
        kids[parent].append(child)
    frontier, seen = collections.deque([tree.root]), set()
    while frontier:
        node = frontier.popleft()
        if node in seen:
            continue
        seen.add(node)
        for child in kids.get(node, ()):
            depth[child] = depth[node] + 1
            frontier.append(child)
    ranked = sorted(depth.items(), key=lambda kv: kv[1])
    return np.array([b for b, _ in ranked])


def bfs_finder(graph, start):
    """Rows of the (n,2) connection list `graph` in breadth-first order from `start`
    (sim_utils.py:570-608)."""
    graph = np.asarray(graph)
    out, seen, todo = [], set(), collections.deque([start])
    while todo:
        key = todo.popleft()
        if key in seen:
            continue
        seen.add(key)
        rows = graph[graph[:, 0] == key]
        out.extend(rows.tolist())
        todo.extend(rows[:, 1].tolist())
    return np.array(out, dtype=graph.dtype).reshape(-1, 2)


def bifurc_adjust(child, parent):
    """Shift `child` so that its first row equals the last row of `parent`
    (sim_utils.py:129-142)."""
    return child - (child[0] - parent[-1])


def adjust_to_parent(relative_means, current, topology):
    """Carry the parent's end point into branch `current` (sim_utils.py:611-640); the
    parent is the first topology row whose child is `current`."""
    topology = np.asarray(topology)
    hits = np.nonzero(topology[:, 1] == current)[0]
    if hits.size == 0:
        return relative_means[current]
    parent = topology[hits[0], 0]
    parent = parent.item() if hasattr(parent, "item") else parent
    return bifurc_adjust(relative_means[current], relative_means[parent])


def find_parallel(tree, programs, branch):
    """Siblings of `branch` (same parent) that already have programs, itself included
    (sim_utils.py:643-667)."""
    for siblings in tree.get_parallel_branches().values():
        if branch in siblings:
            return np.intersect1d(siblings, list(programs.keys()))
    return [branch, None]


def flat_order(n):
    """Index triples (flat, i, j) of the strict upper triangle of an n x n matrix
    (sim_utils.py:171-187)."""
    ii, jj = np.triu_indices(n, k=1)
    return np.stack([np.arange(len(ii)), ii, jj], axis=1).astype(int)


def test_correlation(W, k, cutoff):
    """The reference's intra-branch check is a no-op: its loop `range(k-1, 0)` is empty
    for k >= 1 (sim_utils.py:90-94).  Kept for API parity."""
    return False


test_correlation.__test__ = False  # not a pytest test


# ----------------------------------------------------------------------------- GPU numerics
def _pearson_negative_count(rel_a, rel_b, dev):
    n = min(rel_a.shape[0], rel_b.shape[0])
    G = rel_a.shape[1]
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    nat.call("pst_pearson_anticorr", nat.ptr(rel_a), nat.ptr(rel_b), n, G, nat.ptr(count),
             nat.stream_ptr(dev))
    return int(count.item())


def pearson_between_programs(genes, prog1, prog2, device=None):
    """Per-gene Pearson r over the common rows of two (T, G) arrays
    (sim_utils.py:145-168), on the GPU."""
    dev = nat.device(device)
    n = min(prog1.shape[0], prog2.shape[0])
    a = nat.to_dev(np.asarray(prog1)[:n, :genes], torch.float64, dev)
    b = nat.to_dev(np.asarray(prog2)[:n, :genes], torch.float64, dev)
    a = a - a.mean(dim=0)
    b = b - b.mean(dim=0)
    r = (a * b).sum(dim=0) / torch.sqrt((a * a).sum(dim=0) * (b * b).sum(dim=0))
    return r.cpu().numpy()


def diverging_parallel(branches, programs, genes, tol=0.5, device=None):
    """For every pair of parallel branches: does the fraction of genes with negative
    Pearson r exceed tol?  (sim_utils.py:216-252).  `programs[b]` are (T, G) arrays
    (host or device)."""
    dev = nat.device(device)
    branches = [b for b in branches if b is not None]
    if len(branches) == 1:
        return [True]
    dev_arr = {}
    for b in branches:
        v = programs[b]
        dev_arr[b] = v if isinstance(v, torch.Tensor) else nat.to_dev(v, torch.float64, dev)
    pairs = flat_order(len(branches))
    diverging = np.zeros(len(pairs), dtype=bool)
    for index, i, j in pairs:
        neg = _pearson_negative_count(dev_arr[branches[i]], dev_arr[branches[j]], dev)
        diverging[index] = neg / (genes * 1.0) > tol
    return diverging


def calc_relat_means(tree, programs, coefficients, device=None):
    """rel_means[b] = programs[b] . coefficients for every branch (sim_utils.py:190-213),
    one pst_rel_means launch over the packed tree."""
    dev = nat.device(device)
    tables = tree_tables(tree, dev)
    H = nat.to_dev(coefficients, torch.float64, dev)
    K, G = H.shape
    W = torch.cat([nat.to_dev(np.asarray(programs[b]).reshape(-1, K), torch.float64, dev)
                   for b in tables.names])
    rel = torch.empty((tables.P, G), dtype=torch.float64, device=dev)
    nat.call("pst_rel_means", nat.ptr(W), nat.ptr(H), None, 0, tables.P, K, G, nat.ptr(rel),
             None, None, None, nat.stream_ptr(dev))
    host = rel.cpu().numpy()
    return {b: host[int(tables.row_base[i]):int(tables.row_base[i] + tables.T[i])]
            for i, b in enumerate(tables.names)}


def max_relat_exp(tree, relative_means):
    """(G, B) maximum of exp(relative mean) per gene and branch (sim_utils.py:406-426)."""
    return np.stack([np.max(np.exp(np.asarray(relative_means[b])), axis=0)
                     for b in tree.branches], axis=1)


def base_gene_exp_on_device(cap, seed, abs_max=5000, gene_mean=0.8, gene_std=1, max_tries=100000):
    """Device form of the redraw loop (sim_utils.py:463-469): `cap` is the (G,) fp64 device vector
    of max exp(relative mean); attempt a of gene g reads the Philox normal at element a*G + g of
    the stream keyed by `seed`.  Returns the (G,) fp64 device vector of base expressions."""
    G = cap.numel()
    dev = cap.device
    base = torch.empty(G, dtype=torch.float64, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    nat.call("pst_base_gene_exp", nat.split_seed(seed), nat.TAG_BASE_Z, cap, G, float(abs_max),
             float(gene_mean), float(gene_std), int(max_tries), base, None, flags, nat.stream_ptr(dev))
    if int(flags.item()) != 0:
        raise RuntimeError("a gene cannot satisfy abs_max=%g within %d draws" % (abs_max, max_tries))
    return base


def simulate_base_gene_exp(tree, relative_means, abs_max=5000, gene_mean=0.8, gene_std=1, seed=None,
                           device=None):
    """Per-gene base expression exp(N(gene_mean, gene_std)), redrawn while
    base * max relative expression > abs_max (sim_utils.py:429-470).  With seed=None this is the
    O(G) host loop on the global legacy numpy stream in the reference's draw order (gene by gene,
    redraws in place), so np.random.seed(s) reproduces its values; with a seed the draws are
    counter-based and run on the device (base_gene_exp_on_device)."""
    cap = np.max(max_relat_exp(tree, relative_means), axis=1)
    if seed is not None:
        dev = nat.device(device)
        return base_gene_exp_on_device(nat.to_dev(cap, torch.float64, dev), seed, abs_max, gene_mean,
                                       gene_std).cpu().numpy()
    return _base_gene_exp_legacy(cap, abs_max, gene_mean, gene_std)


def _base_gene_exp_legacy(cap, abs_max, gene_mean, gene_std):
    base = np.zeros(len(cap))
    for gene in range(len(cap)):
        value = np.exp(np.random.normal(gene_mean, gene_std))
        tries = 0
        while value * cap[gene] > abs_max:
            tries += 1
            if tries > 100000:
                raise RuntimeError("gene %d cannot satisfy abs_max=%g" % (gene, abs_max))
            value = np.exp(np.random.normal(gene_mean, gene_std))
        base[gene] = value
    return base


def calc_scalings(cells, scale=True, scale_mean=0, scale_v=0.7, seed=None, first=0, device=None,
                  return_device=False):
    """Library-size factors exp(N(scale_mean, scale_v)) per cell, or ones
    (sim_utils.py:473-498).  Philox stream keyed by the global cell index `first + i`."""
    dev = nat.device(device)
    st = nat.stream_ptr(dev)
    s64 = torch.empty(cells, dtype=torch.float64, device=dev)
    s32 = torch.empty(cells, dtype=torch.float32, device=dev)
    if scale:
        z = torch.empty(cells, dtype=torch.float64, device=dev)
        nat.call("pst_normal_f64", nat.split_seed(seed), nat.TAG_SCALING_Z, int(first), cells,
                 float(scale_mean), float(scale_v), None, None, nat.ptr(z), st)
        nat.call("pst_scalings", nat.ptr(z), cells, nat.ptr(s64), nat.ptr(s32), st)
    else:
        nat.call("pst_scalings", None, cells, nat.ptr(s64), nat.ptr(s32), st)
    if return_device:
        return s64, s32
    return s64.cpu().numpy()


# ----------------------------------------------------------------------------- cells -> branches
def assign_branches(branch_times, timezone):
    """zone index -> branches alive during the whole zone, in branch_times order
    (sim_utils.py:274-315)."""
    live = defaultdict(list)
    for i, zone in enumerate(timezone):
        for name, span in branch_times.items():
            if belongs_to(zone, span):
                live[i].append(name)
    return live


def belongs_to(timezone, branch):
    """sim_utils.py:318-339."""
    return timezone[0] >= branch[0] and timezone[1] <= branch[1]


def pick_branches(tree, pseudotime, seed=None, first=0, device=None, uniforms=None):
    """One branch per pseudotime value, chosen among the branches alive at that time
    with probability proportional to their density (sim_utils.py:342-403, incl. the
    zone-relative density index).  Names longer than the first branch name
    are NOT truncated (reference bug).  `uniforms` (one per cell) replaces
    the Philox stream - used to replay the reference's own draws."""
    codes, _, flags = _pick_branch_codes(tree, pseudotime, seed, first, device, uniforms)
    check_flags(flags)
    tables = tree_tables(tree, nat.device(device))
    return tables.branch_names(codes.cpu().numpy())


def _pick_branch_codes(tree, pseudotime, seed, first, device, uniforms=None, tables=None):
    dev = nat.device(device)
    tables = tables or tree_tables(tree, dev)
    st = nat.stream_ptr(dev)
    pt = pseudotime if isinstance(pseudotime, torch.Tensor) else \
        nat.to_dev(np.asarray(pseudotime), torch.int64, dev)
    n = int(pt.numel())
    if uniforms is None:
        u = torch.empty(n, dtype=torch.float64, device=dev)
        nat.call("pst_uniform_f64", nat.split_seed(seed), nat.TAG_PICK_U, int(first), n, nat.ptr(u), st)
    else:
        u = nat.to_dev(uniforms, torch.float64, dev)
    dens = nat.to_dev(tables.density_packed(tree), torch.float64, dev)
    codes = torch.empty(n, dtype=torch.int32, device=dev)
    rows = torch.empty(n, dtype=torch.int32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    nat.call("pst_pick_branch", nat.ptr(pt), nat.ptr(u), n, len(tables.zone_lo),
             nat.ptr(tables.d("zone_lo")), nat.ptr(tables.d("zone_hi")),
             nat.ptr(tables.d("cand_off")), nat.ptr(tables.d("cand_branch")), tables.max_cand,
             nat.ptr(tables.d("branch_start")), nat.ptr(tables.d("row_base")),
             nat.ptr(tables.d("T")), nat.ptr(dens), nat.ptr(codes), nat.ptr(rows),
             nat.ptr(flags), st)
    # the status word is NOT read here (that would stall the stream in front of the count draw): the caller
    # checks it with check_flags once the rest of its work has been queued
    return codes, rows, flags


def check_flags(flags):
    """Read a device status word (one device->host sync) and raise what it reports."""
    from prosstt_b200.device import raise_flags
    word = int(flags.item())
    if word:
        raise_flags(word)


def pick_branch(tree, pseudotime, timezones=None, assignments=None, seed=None, device=None):
    """Single-cell form of pick_branches (sim_utils.py:367-403)."""
    return pick_branches(tree, [pseudotime], seed=seed, device=device)[0]


def process_timeseries_input(series_points, cells, point_std):
    """Broadcast the inputs of sample_pseudotime_series (sim_utils.py:501-542): an int
    `cells` is split as int(cells / n_points) per point; a scalar `point_std` is divided
    by the number of points (sic)."""
    n = len(series_points)
    if isinstance(cells, collections.abc.Iterable):
        cells = np.array(cells, dtype=int)
    elif isinstance(cells, numbers.Number):
        cells = np.array([cells / n] * n, dtype=int)
    if isinstance(point_std, collections.abc.Iterable):
        point_std = np.array(point_std, dtype=float)
    elif isinstance(point_std, numbers.Number):
        point_std = np.array([point_std / n] * n, dtype=float)
    if not isinstance(series_points, np.ndarray):
        series_points = np.array(series_points, dtype=int)
    return series_points, cells, point_std


# ----------------------------------------------------------------------------- groups (Beta variant)
def random_partition(k, iterable):
    """Each value goes to one of k groups uniformly at random (sim_utils.py:52-73)."""
    groups = [[] for _ in range(k)]
    for value in iterable:
        groups[np.random.randint(k)].append(value)
    return groups


def create_groups(no_programs, no_genes):
    """Two independent random partitions of the genes over the programs, concatenated
    per program (sim_utils.py:97-126)."""
    first = random_partition(no_programs, np.random.permutation(no_genes))
    second = random_partition(no_programs, np.random.permutation(no_genes))
    return [a + b for a, b in zip(first, second)]

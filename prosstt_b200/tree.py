"""Lineage tree: topology, branch lengths, cell density and mean expression.

Host-side mirror of the reference's `prosstt.tree.Tree` (prosstt/tree.py:19-446): same
constructor, attributes, methods, defaults and exceptions, so scripts written for the
reference run unchanged.  The integer maps (branch_times, populate_timezone, paths, ...)
feed the device tables in `prosstt_b200.device` and are bit-exact with the reference
(tests/test_host_maps.py against tests/golden/maps.json).
"""
from collections import defaultdict

from collections.abc import Mapping

import numpy as np
import pandas as pd


class Tree(object):
    """A lineage tree (reference: prosstt/tree.py:19-80).

    Attributes: topology (list of [parent, child]), time (pd.Series branch -> length),
    num_branches, branch_points, modules (K), G, means (dict branch -> (T_b, G) array),
    branches (list), root, density (dict branch -> (T_b,) array).
    """

    def_time = 40
    def_genes = 500

    def __init__(self, topology=None, time=None, num_branches=3, branch_points=1,
                 modules=None, G=def_genes, density=None, root=None):
        if topology is None:
            topology = [["A", "B"], ["A", "C"]]
        if time is None:
            time = {"A": self.def_time, "B": self.def_time, "C": self.def_time}
        self.topology = topology
        self.time = pd.Series(time, name="time")
        self.num_branches = num_branches
        self.branch_points = branch_points
        self.G = G
        self.means = None
        self.branches = list(time.keys())
        # tree.py:67-68: one draw from the global legacy stream when K is not given
        self.modules = (5 * branch_points + np.random.randint(1, 20)) if modules is None else modules
        self.root = self.branches[0] if root is None else root
        self.density = self.default_density() if density is None else density
        self._device_cache = {}

    # ------------------------------------------------------------------ constructors
    @staticmethod
    def gen_random_topology(branch_points, branch_names=None):
        """Random binary topology with `branch_points` bifurcations (tree.py:82-113):
        each step picks one current leaf with np.random.choice and hangs the next two
        unused labels under it; rows are [parent, child], parents first."""
        total = 2 * branch_points + 1
        names = np.arange(total) if branch_names is None else branch_names
        leaves = [0]
        rows = []
        for first in range(1, total, 2):
            parent = np.random.choice(leaves)
            rows.append([names[parent], names[first]])
            rows.append([names[parent], names[first + 1]])
            leaves.remove(parent)
            leaves.extend((first, first + 1))
        return rows

    @classmethod
    def from_newick(cls, newick_tree, modules=None, genes=def_genes, density=None):
        """Tree from a Newick string, e.g. "(A:50,B:50)C:50;" (tree.py:115-126)."""
        from prosstt_b200 import tree_utils as tu
        parsed = tu.newick_loads(newick_tree)
        top, time, branches, br_points, root = tu.parse_newick(parsed, cls.def_time)
        return cls(top, time, branches, br_points, modules, genes, density, root)

    @classmethod
    def from_random_topology(cls, branch_points, time, modules, genes):
        """tree.py:128-136."""
        topology = cls.gen_random_topology(branch_points, branch_names=list(time.keys()))
        num_branches = len(np.unique(topology))
        return cls(topology, time, num_branches, branch_points, modules, genes)

    # ------------------------------------------------------------------ density / means
    def default_density(self):
        """Uniform density 1/P at every tree position (tree.py:138-151)."""
        total = 0
        for length in self.time.values:
            total += length
        return {b: np.array([1. / total] * int(self.time[b])) for b in self.time.keys()}

    def add_genes(self, *args):
        """add_genes(dict of (T_b,G) means) or add_genes(relative_means, base_expr)
        (tree.py:154-163)."""
        if len(args) == 1 and isinstance(args[0], Mapping):      # dict, or the device-resident DeviceMeans
            self._add_genes_from_average(args[0])
        if len(args) == 2 and isinstance(args[1], np.ndarray):
            self._add_genes_from_relative(args[0], args[1])

    def _add_genes_from_relative(self, relative_means, base_gene_expr):
        """means[b] = exp(relative_means[b]) * base_gene_expr (tree.py:166-183)."""
        self._add_genes_from_average(
            {b: np.exp(relative_means[b]) * base_gene_expr for b in self.branches})

    def _add_genes_from_average(self, average_expression):
        """Shape checks of tree.py:186-213, then store."""
        if len(average_expression) != self.num_branches:
            raise ValueError("The number of arrays in average_expression must be equal to "
                             "the number of branches in the topology")
        for branch in average_expression:
            shape = np.shape(average_expression[branch])
            want = (self.time[branch], self.G)
            if shape != want:
                raise ValueError("Branch %s was expected to have a shape %s and instead is %s"
                                 % (str(branch), str(want), str(shape)))
        self.means = average_expression
        self.invalidate_device_cache()

    def set_density(self, density):
        """tree.py:216-238."""
        self._check_per_branch(density, "density")
        self.density = density

    def set_velocity(self, velocity):
        """Density as the reverse of a velocity profile (tree.py:241-264)."""
        from prosstt_b200 import tree_utils as tu
        self._check_per_branch(velocity, "velocity")
        self.density = tu._density_from_velocity(tu.sanitize_velocity(velocity))

    def _check_per_branch(self, per_branch, what):
        if len(per_branch) != len(self.branches):
            raise ValueError("The number of arrays in %s must be equal to the number of "
                             "branches in the topology" % what)
        for b in per_branch:
            if len(per_branch[b]) != self.time[b]:
                raise ValueError("Branch %s was expected to have a length %s and instead is %s"
                                 % (str(b), str(self.time[b]), str(np.shape(per_branch[b]))))

    def invalidate_device_cache(self):
        """Drop device copies of `means` (call after editing tree.means in place)."""
        self._device_cache = {}

    # ------------------------------------------------------------------ integer maps
    def as_dictionary(self):
        """parent -> [children] in topology order (tree.py:287-300)."""
        kids = defaultdict(list)
        for parent, child in self.topology:
            kids[parent].append(child)
        return kids

    def paths(self, start):
        """All paths from `start` to the leaves below it, depth first in topology order
        (tree.py:302-330)."""
        kids = self.as_dictionary()
        done, stack = [], [[start]]
        while stack:
            path = stack.pop()
            below = kids.get(path[-1], [])
            if not below:
                done.append(path)
            else:
                stack.extend(path + [c] for c in reversed(below))
        return done

    def branch_times(self):
        """Absolute [start, end] pseudotime of every branch, root first then children in
        topology order (tree.py:376-399).  Rows must list parents before children."""
        span = defaultdict(list)
        span[self.root] = [0, self.time[self.root] - 1]
        for parent, child in self.topology:
            if parent not in span:
                raise ValueError("topology row [%s, %s] comes before the row that defines %s"
                                 % (str(parent), str(child), str(parent)))
            stop = span[parent][1]
            span[child] = [stop + 1, stop + self.time[child]]
        return span

    def get_max_time(self):
        """Length of the longest root-to-leaf path (tree.py:267-285)."""
        return int(max(stop for _, stop in self.branch_times().values())) + 1

    def populate_timezone(self):
        """Maximal pseudotime intervals over which the set of live branches is constant
        (tree.py:332-374).  Every branch end is a cut point, so the zones are the gaps
        between consecutive distinct branch ends."""
        cuts = sorted({stop + 1 for _, stop in self.branch_times().values()})
        zones, lo = [], 0
        for cut in cuts:
            zones.append([lo, cut - 1])
            lo = cut
        return zones

    def morph_stack(self, stack):
        """Branch lengths along a path -> [start, end) pairs (tree.py:402-423)."""
        begin = 0
        for i, length in enumerate(stack):
            stack[i] = [begin, begin + length]
            begin += length
        return stack

    def get_parallel_branches(self):
        """parent -> array of its children (tree.py:425-434)."""
        top = np.array(self.topology)
        return {p: top[top[:, 0] == p, 1] for p in np.unique(top[:, 0])}

    def default_gene_expression(self):
        """simulate_lineage(a=0.05) + base expression + add_genes (tree.py:436-446).  The (P, G)
        tables stay in HBM; tree.means is a dict-like that copies a branch to the host on first
        access (simulation.DeviceMeans)."""
        from prosstt_b200 import simulation as sim
        sim.default_gene_expression_on_device(self, a=0.05)

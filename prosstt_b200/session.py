"""Resident sampling session: sample_density with every buffer preallocated.

`simulation.sample_density` allocates its outputs per call (reference behaviour: fresh
arrays).  For repeated or very large runs a DensitySession keeps the replicated state
(means table, per-gene parameters, tree tables, cdf) and the output slab resident in HBM
and re-samples in place; `step_to_host` streams the slab to a (pinned) host buffer in
cell chunks, overlapping the device->host copy with sampling.  Same kernels, same
counter-based streams: results equal simulation.sample_density(..., seed=seed,
shard=...) bit for bit (tests/test_gpu_parity.py).
"""
import numpy as np
import torch

from prosstt_b200 import _native as nat
from prosstt_b200.device import CountEngine, TreeTables, choice_cdf


class DensitySession(object):
    def __init__(self, tree, alpha, beta, cells, first=0, device=None, sampler=None,
                 scale=True, scale_mean=0.0, scale_v=0.7, resident_output=True):
        if sampler is None:                       # same default as simulation.sample_density
            from prosstt_b200.simulation import DEFAULT_SAMPLER as sampler
        self.dev = nat.device(device)
        self.tree = tree
        self.n, self.first = int(cells), int(first)
        self.capacity = self.n
        self.scale, self.scale_mean, self.scale_v = scale, float(scale_mean), float(scale_v)
        self.tables = TreeTables(tree, self.dev)
        self.engine = CountEngine(tree, self.tables, alpha, beta, self.dev, sampler=sampler)
        self.alpha_host, self.beta_host = alpha, beta
        self.G = self.engine.G
        dev, n = self.dev, self.n
        self.cdf = nat.to_dev(choice_cdf(self.tables.density_packed(tree)), torch.float64, dev)
        self.u = torch.empty(n, dtype=torch.float64, device=dev)
        self.z = torch.empty(n, dtype=torch.float64, device=dev)
        self.rows = torch.empty(n, dtype=torch.int32, device=dev)
        self.pt = torch.empty(n, dtype=torch.int64, device=dev)
        self.codes = torch.empty(n, dtype=torch.int32, device=dev)
        self.s64 = torch.empty(n, dtype=torch.float64, device=dev)
        self.s32 = torch.empty(n, dtype=torch.float32, device=dev)
        self.X = torch.empty((n, self.G), dtype=torch.int32, device=dev) if resident_output else None
        self.t_draw = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))

    def set_range(self, first, cells=None):
        """Re-target the session at the global cell range [first, first + cells), cells <= the size it
        was built with: a long run is sampled chunk after chunk through one set of buffers."""
        n = self.capacity if cells is None else int(cells)
        if not 0 <= n <= self.capacity:
            raise ValueError("a session built for %d cells cannot sample %d" % (self.capacity, n))
        self.n, self.first = n, int(first)

    # -- per-call host inputs of sample_density(tree, N, alpha, beta): density + gene params
    def upload_inputs(self):
        """Host -> device copy of what one API call uploads (cdf, alpha, beta-1).  Returns
        the number of bytes copied."""
        from prosstt_b200.device import gene_params
        cdf = choice_cdf(self.tables.density_packed(self.tree))
        self.cdf.copy_(torch.from_numpy(cdf), non_blocking=True)
        a, b = gene_params(self.alpha_host, self.beta_host, self.G, self.dev)
        self.engine.alpha.copy_(a)
        self.engine.beta_m1.copy_(b)
        return cdf.nbytes + 2 * 4 * self.G

    def index_and_scalings(self, seed):
        st = nat.stream_ptr(self.dev)
        tb, n = self.tables, self.n
        nat.call("pst_uniform_f64", nat.derive_seed(seed, 0), nat.TAG_DENSITY_U, self.first, n,
                 nat.ptr(self.u), st)
        nat.call("pst_density_index", nat.ptr(self.cdf), tb.P, nat.ptr(self.u), n,
                 nat.ptr(tb.d("pos_pt")), nat.ptr(tb.d("pos_branch")), nat.ptr(self.rows),
                 nat.ptr(self.pt), nat.ptr(self.codes), st)
        if self.scale:
            nat.call("pst_normal_f64", nat.derive_seed(seed, 1), nat.TAG_SCALING_Z, self.first, n,
                     self.scale_mean, self.scale_v, None, None, nat.ptr(self.z), st)
            nat.call("pst_scalings", nat.ptr(self.z), n, nat.ptr(self.s64), nat.ptr(self.s32), st)
        else:
            nat.call("pst_scalings", None, n, nat.ptr(self.s64), nat.ptr(self.s32), st)

    def step(self, seed, gene_stats=None):
        """One sample_density pass into the resident slab.  Stream-ordered, no sync.  gene_stats
        (stats.new_gene_stats): per-gene sum / sum of squares / zeros accumulated inside the draw."""
        self.index_and_scalings(seed)
        n = self.n
        self.t_draw[0].record()
        self.engine.draw(self.rows[:n], self.s32[:n], nat.derive_seed(seed, 2), self.first, out=self.X[:n],
                         gene_stats=gene_stats)
        self.t_draw[1].record()
        return self.X[:n]

    def last_draw_ms(self):
        """Device time of the last draw kernel (call after a synchronize)."""
        return self.t_draw[0].elapsed_time(self.t_draw[1])

    def step_to_host(self, seed, host_X, host_pt=None, host_codes=None, host_scal=None,
                     chunk_cells=None):
        """One pass streamed into host buffers (CPU tensors, ideally pinned).  Returns the
        bytes copied device->host.  Synchronises before returning.  host_X may be int32, uint16 or
        uint8 (narrow transfer formats; the saturated elements are then in `self.engine.overflow`)."""
        self.index_and_scalings(seed)
        n = self.n
        self.engine.draw_to_host(self.rows[:n], self.s32[:n], nat.derive_seed(seed, 2), self.first, host_X,
                                 chunk_cells=chunk_cells)
        nbytes = host_X.numel() * host_X.element_size()
        for src, dst in ((self.pt, host_pt), (self.codes, host_codes), (self.s64, host_scal)):
            if dst is not None:
                dst.copy_(src[:n], non_blocking=True)
                nbytes += dst.numel() * dst.element_size()
        torch.cuda.current_stream(self.dev).synchronize()
        self.engine.check()
        return nbytes

    def results(self):
        """(X, pseudotime, branches, scalings) of the last step as numpy (reference types)."""
        self.engine.check()
        n = self.n
        return (self.X[:n].cpu().numpy(), self.pt[:n].cpu().numpy(),
                self.tables.branch_names(self.codes[:n].cpu().numpy()), self.s64[:n].cpu().numpy())

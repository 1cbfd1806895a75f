"""Device-side view of a Tree: packed tables in HBM and the count engine.

Layout (include/prosstt_b200.h): branches in `tree.branches` order, branch b owns packed
rows [row_base[b], row_base[b]+T_b); P = sum T_b.  Everything here is small and
replicated on every GPU (a few MB; the means table is P x G fp32); only cells shard.
"""
import os

import numpy as np
import torch

from prosstt_b200 import _native as nat


def tree_tables(tree, dev):
    """TreeTables of `tree` on `dev`, cached on the tree: the integer maps depend only on the topology,
    the branch lengths, the branch order and the root, which form the key - an edited tree gets new tables."""
    try:
        key = (str(dev), tuple((p, c) for p, c in tree.topology), tuple((b, int(tree.time[b])) for b in tree.branches),
               tree.root)
        hash(key)
    except TypeError:
        return TreeTables(tree, dev)
    cache = tree.__dict__.setdefault("_tables_cache", {})
    hit = cache.get(key)
    if hit is None:
        cache.clear()
        hit = cache[key] = TreeTables(tree, dev)
    return hit


class TreeTables(object):
    """Integer maps of a tree flattened for the kernels (host numpy + device copies)."""

    def __init__(self, tree, dev):
        self.dev = dev
        self.names = list(tree.branches)
        self.index = {b: i for i, b in enumerate(self.names)}
        if len(self.index) != len(self.names):
            raise ValueError("branch names must be unique")
        B = len(self.names)
        self.B = B
        self.T = np.array([int(tree.time[b]) for b in self.names], dtype=np.int32)
        if np.any(self.T < 1):
            raise ValueError("every branch needs at least one pseudotime step")
        self.row_base = np.zeros(B, dtype=np.int32)
        self.row_base[1:] = np.cumsum(self.T)[:-1]
        self.P = int(self.T.sum())
        bt = tree.branch_times()
        missing = [b for b in self.names if b not in bt or len(bt[b]) != 2]
        if missing:
            raise ValueError("branches %s are not connected to the root by the topology" % missing)
        self.branch_start = np.array([bt[b][0] for b in self.names], dtype=np.int32)
        # sample_density's concatenation (simulation.py:454-461)
        self.pos_branch = np.repeat(np.arange(B, dtype=np.int32), self.T)
        self.pos_pt = (np.arange(self.P, dtype=np.int32) - np.repeat(self.row_base, self.T)
                       + np.repeat(self.branch_start, self.T)).astype(np.int32)
        # timezones and their live branches in branch_times order (sim_utils.py:274-339)
        zones = tree.populate_timezone()
        order = [self.index[b] for b in bt.keys() if b in self.index]
        self.zone_lo = np.array([z[0] for z in zones], dtype=np.int32)
        self.zone_hi = np.array([z[1] for z in zones], dtype=np.int32)
        cand, off = [], [0]
        for lo, hi in zones:
            live = [b for b in order
                    if lo >= self.branch_start[b] and hi <= self.branch_start[b] + self.T[b] - 1]
            cand.extend(live)
            off.append(len(cand))
        self.cand_branch = np.array(cand, dtype=np.int32)
        self.cand_off = np.array(off, dtype=np.int32)
        self.max_cand = int(np.max(np.diff(self.cand_off))) if len(zones) else 0
        # cover_whole_tree (simulation.py:520-548): zone-major, branch, time
        cpt, cbr = [], []
        for z, (lo, hi) in enumerate(zones):
            for b in self.cand_branch[self.cand_off[z]:self.cand_off[z + 1]]:
                cpt.append(np.arange(lo, hi + 1, dtype=np.int32))
                cbr.append(np.full(hi + 1 - lo, b, dtype=np.int32))
        self.cover_pt = np.concatenate(cpt) if cpt else np.zeros(0, np.int32)
        self.cover_branch = np.concatenate(cbr) if cbr else np.zeros(0, np.int32)
        self.cover_row = (self.row_base[self.cover_branch] + self.cover_pt
                          - self.branch_start[self.cover_branch]).astype(np.int32)
        self.max_time = int((self.branch_start + self.T).max())
        self._dev = {}

    def d(self, name, dtype=torch.int32):
        """Device copy of one of the host tables (cached)."""
        if name not in self._dev:
            self._dev[name] = nat.to_dev(getattr(self, name), dtype, self.dev)
        return self._dev[name]

    def density_packed(self, tree):
        dens = [np.asarray(tree.density[b], dtype=np.float64) for b in self.names]
        for b, v in zip(self.names, dens):
            if v.shape != (int(tree.time[b]),):
                raise ValueError("density of branch %s has shape %s, expected (%d,)"
                                 % (str(b), str(v.shape), int(tree.time[b])))
        return np.concatenate(dens)

    def branch_codes(self, branches):
        """Branch names (any iterable) -> int32 indices into tree.branches."""
        arr = np.asarray(branches)
        if arr.dtype.kind in "iu" and all(isinstance(b, (int, np.integer)) for b in self.names):
            lut_keys = np.array(self.names, dtype=np.int64)
            sorter = np.argsort(lut_keys)
            pos = np.searchsorted(lut_keys, arr, sorter=sorter)
            pos = np.clip(pos, 0, len(lut_keys) - 1)
            codes = sorter[pos]
            if np.any(lut_keys[codes] != arr):
                raise KeyError("unknown branch name in `branches`")
            return codes.astype(np.int32)
        try:
            return np.array([self.index[b.item() if hasattr(b, "item") else b] for b in arr],
                            dtype=np.int32)
        except KeyError as err:
            raise KeyError("unknown branch name in `branches`: %s" % err)

    def branch_names(self, codes):
        """int32 indices -> array of branch names with the dtype numpy gives the names
        (reference: possible_branches[sample], simulation.py:461-467)."""
        return np.array(self.names)[np.asarray(codes, dtype=np.int64)]


def choice_cdf(p):
    """The cdf numpy's legacy RandomState.choice builds (with its validity checks)."""
    p = np.asarray(p, dtype=np.float64)
    if p.ndim != 1 or p.size == 0:
        raise ValueError("'p' must be 1-dimensional and non-empty")
    if np.any(np.isnan(p)) or np.any(p < 0):
        raise ValueError("probabilities are not non-negative")
    if abs(float(np.sum(p)) - 1.0) > np.sqrt(np.finfo(np.float64).eps):
        raise ValueError("probabilities do not sum to 1")
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return cdf


def _content_key(m):
    """Cheap fingerprint of one branch's means for the device-table cache: buffer address, shape, the
    first and last pseudotime rows and ~4096 values at a fixed stride.  An in-place edit of the whole
    array, of a gene column (seen in the first row) or of a pseudotime row (>= G/stride samples fall in
    it) - ordinary NumPy usage in scripts written for the reference - changes it, unlike id(); after
    editing single elements call tree.invalidate_device_cache().  ~0.05 ms per branch."""
    if isinstance(m, torch.Tensor):
        return ("t", m.data_ptr(), tuple(m.shape), m._version)
    a = np.asarray(m)
    flat = a.reshape(-1)
    step = max(1, flat.size // 4096) | 1
    edge = (a[0].tobytes(), a[-1].tobytes()) if a.ndim >= 2 and a.shape[0] else ()
    return (a.__array_interface__["data"][0], a.shape, flat[::step].tobytes()) + edge


def means_table(tree, tables, dev):
    """(P, G) fp32 means table in HBM, built once per tree.means and cached on the tree."""
    if tree.means is None:
        raise ValueError("the tree has no mean expression yet: call add_genes() or "
                         "default_gene_expression() first")
    if hasattr(tree.means, "table32"):                 # simulation.DeviceMeans: already in HBM
        if tuple(tree.means.table32.shape) != (tables.P, int(tree.G)):
            raise ValueError("the device-resident means table has shape %s, the tree needs %s"
                             % (tuple(tree.means.table32.shape), (tables.P, int(tree.G))))
        if tree.means.table32.device == dev:
            return tree.means.table32
        return tree.means.table32.to(dev)
    key = ("means32", str(dev)) + tuple(_content_key(tree.means[b]) for b in tables.names)
    cache = tree.__dict__.setdefault("_device_cache", {})
    hit = cache.get(key)
    if hit is not None:
        return hit
    G = int(tree.G)
    out = torch.empty((tables.P, G), dtype=torch.float32, device=dev)
    st = nat.stream_ptr(dev)
    for i, b in enumerate(tables.names):
        m = tree.means[b]
        if isinstance(m, torch.Tensor):
            m64 = m.to(device=dev, dtype=torch.float64).contiguous()
        else:
            m64 = nat.to_dev(m, torch.float64, dev)
        if tuple(m64.shape) != (int(tables.T[i]), G):
            raise ValueError("means of branch %s have shape %s, expected %s"
                             % (str(b), tuple(m64.shape), (int(tables.T[i]), G)))
        dst = out[int(tables.row_base[i]):int(tables.row_base[i]) + int(tables.T[i])]
        nat.call("pst_f64_to_f32", nat.ptr(m64), m64.numel(), 1e-30, dst.data_ptr(), st)
    cache.clear()
    cache[key] = out
    return out


def gene_params(alpha, beta, G, dev):
    """Per-gene alpha and beta-1 as fp32 device vectors (beta-1 formed in fp64)."""
    a = np.broadcast_to(np.asarray(alpha, dtype=np.float64), (G,)) if np.ndim(alpha) == 0 \
        else np.asarray(alpha, dtype=np.float64)
    b = np.broadcast_to(np.asarray(beta, dtype=np.float64), (G,)) if np.ndim(beta) == 0 \
        else np.asarray(beta, dtype=np.float64)
    if a.shape != (G,) or b.shape != (G,):
        raise ValueError("alpha and beta must be scalars or have one value per gene (G=%d)" % G)
    return nat.to_dev(a, torch.float32, dev), nat.to_dev(b - 1.0, torch.float32, dev)


def raise_flags(word):
    """Translate the device status word into the reference's exceptions."""
    if word & nat.FLAG_ROW:
        raise IndexError("a cell's pseudotime lies outside its branch")
    if word & nat.FLAG_NOZONE:
        raise IndexError("a pseudotime value lies outside the lineage tree")
    if word & nat.FLAG_DOMAIN:
        raise ValueError("Domain error in arguments. The mean expression must be positive "
                         "and alpha*mean + beta must exceed 1 for every cell and gene.")
    if word & nat.FLAG_SCRATCH:
        raise RuntimeError("the tail list of pst_draw_counts overflowed (scratch too small)")
    if word & nat.FLAG_CLAMPED:
        raise OverflowError("a sampled count exceeded the int32 range")


_NP_TO_TORCH = {np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
                np.dtype(np.uint16): torch.uint16, np.dtype(np.uint8): torch.uint8}
_STAGE_BYTES = int(os.environ.get("PST_STAGE_MB", "256")) << 20   # one pinned staging buffer (two per device)
_STAGING = {}


_AUTO_STATE = {"narrow_overflowed": False}     # an automatically chosen uint8 transport overflowed its list once


def _shared_host_transport(threads):
    """Transport of a host matrix when nothing was asked for.  "direct" / "i32": the counts cross PCIe as
    int32, 4 B per count (55-57 GB/s = 1.4e10 counts/s for one B200).  "u8": they cross as uint8 + exact
    overflow list (a quarter of the bytes) and host threads expand them with non-temporal stores
    (pst_host_widen_stream: 5.7e9 counts/s per thread, 3.4e10 with 16).  Measured: one rank with 16 threads
    2.4e10 against 1.3e10 counts/s; eight ranks with 4 threads each on a 32-vCPU box, where the host MEMORY
    system is the limit, 3.0e10 against 2.3e10 (6 B of host traffic per count instead of the 10 B of an
    expansion with ordinary stores, which lost to "direct" there).  So "u8" wherever the CPU has the streaming
    stores and this rank has at least 3 threads, or 10 threads without them.  A fixed rule, not a timing trial:
    every rank of a job must pick the same transport (a trial run by eight ranks at once measured their
    contention and split them).  PST_HOST_TRANSPORT=direct|i32|u16|u8 overrides."""
    forced = os.environ.get("PST_HOST_TRANSPORT")
    if forced:
        return forced
    need = 3 if nat.load().pst_host_stream_stores() else 10
    return "u8" if threads >= need else "direct"


def _host_threads():
    """Host threads of this rank for the expansion: its share of the cores when several ranks share a host."""
    try:
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except ValueError:
        local_world = 1
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, cores // local_world)


def _pinned_staging(dev, nbytes):
    """Two pinned host buffers of at least nbytes for the staged device->host path, allocated once per
    device and kept (page-locking costs ~0.2 s per GB)."""
    key = str(dev)
    have = _STAGING.get(key)
    if have is None or have[0].numel() < nbytes:
        size = max(int(nbytes), 1)
        have = [torch.empty(size, dtype=torch.uint8).pin_memory() for _ in range(2)]
        _STAGING[key] = have
    return have


class CountEngine(object):
    """draw_counts on one GPU: replicated small state + a cell range to sample."""

    timers = None     # set to a list to collect (start, end) CUDA events around every draw (bench.py)

    def __init__(self, tree, tables, alpha, beta, dev, sampler="gamma_poisson"):
        self.dev = dev
        self.G = int(tree.G)
        self.P = tables.P
        self.means = means_table(tree, tables, dev)
        self.alpha, self.beta_m1 = gene_params(alpha, beta, self.G, dev)
        self.sampler = nat.SAMPLERS[sampler]
        self.flags = torch.zeros(4, dtype=torch.int32, device=dev)   # status word (+3 reserved)
        self._scratch = None
        self.overflow = None                                    # filled by draw_to_host (uint16 format)

    def draw(self, rows, scaling32, seed, cell0, out=None, gene_stats=None):
        """Sample X for the cells described by rows/scaling32 (device tensors); global
        index of the first cell is cell0.  Returns an (n, G) int32 device tensor.
        gene_stats: dict with int64 (G,) device tensors "gene_sum", "gene_sumsq", "gene_zeros"
        (stats.new_gene_stats) that the draw adds this call's per-gene summaries to - fused into
        the kernel, no second pass over the matrix; equals stats.count_stats(X) bit for bit."""
        n = int(rows.numel())
        if out is None:
            out = torch.empty((n, self.G), dtype=torch.int32, device=self.dev)
        st = nat.stream_ptr(self.dev)
        timers = CountEngine.timers
        if timers is not None:
            t0 = torch.cuda.Event(enable_timing=True)
            t0.record(torch.cuda.current_stream(self.dev))
        # scratch of the call: visiting order of the cells (grouped by tree row) and the list of the counts
        # that the fix-up kernel finishes with a 64-bit uniform (top 2^-14 of the uniforms)
        words = int(nat.load().pst_draw_scratch_words(n, self.G, self.P))
        if self._scratch is None or self._scratch.numel() < words:
            self._scratch = torch.empty(words, dtype=torch.int32, device=self.dev)
        nat.call("pst_draw_counts", nat.ptr(self.means), self.P, self.G, nat.ptr(rows),
                 nat.ptr(scaling32), nat.ptr(self.alpha), nat.ptr(self.beta_m1),
                 seed, int(cell0), n, out.data_ptr(), out.stride(0) if n else self.G,
                 nat.ptr(self.flags), self.sampler, self._scratch, words,
                 *((gene_stats["gene_sum"], gene_stats["gene_sumsq"], gene_stats["gene_zeros"])
                   if gene_stats is not None else (None, None, None)), st)
        if timers is not None:
            t1 = torch.cuda.Event(enable_timing=True)
            t1.record(torch.cuda.current_stream(self.dev))
            timers.append((t0, t1))
        return out

    def draw_to_host(self, rows, scaling32, seed, cell0, host_out, chunk_cells=None, overflow_cap=None,
                     transport=None, threads=0, fresh=False):
        """Sample in cell chunks and stream them into `host_out`, an (n, G) CPU tensor or C-contiguous
        NumPy array: sampling of chunk i+1 overlaps the transfer (and host-side expansion) of chunk i.

        host_out int32, pinned, transport None/"direct": the counts as sampled are copied straight into
          it (4 B per count over PCIe).
        host_out uint16 / uint8 (pinned): the narrow formats themselves (2 / 1 B per count):
          min(count, SAT) with SAT = 65535 / 255, and every element that reads SAT is listed exactly in
          `self.overflow` = (flat index into host_out, int32 value) NumPy arrays sorted by index
          (formats.widen rebuilds int32).  At default depth about 2 in 10^4 counts reach 255 and none
          reaches 65535.
        host_out int32 or int64, pageable or pinned, transport "u8" / "u16" / "i32": the chunk crosses
          PCIe in the transport format into pinned staging buffers (cached per device) and host threads
          (pst_host_widen) expand it into host_out, the overflow list is applied at the end: the caller
          receives exact int32 / int64 counts.  This is the path of the reference-shaped call, which
          returns a fresh (pageable) int64 array.  Default transport: the wide one ("direct" for a pinned
          int32 tensor, "i32" otherwise) or "u8", whichever is faster on this host (_shared_host_transport);
          an automatically chosen "u8" whose overflow list does not fit (deep regimes) is sampled again
          through the wide transport - the counts are the same either way.
          The expansion uses streaming stores (pst_host_widen_stream) unless `fresh` says that host_out was
          just allocated and never touched (its pages are zeroed into the cache by the first-touch faults)."""
        n = int(rows.numel())
        is_np = isinstance(host_out, np.ndarray)
        if is_np:
            if not host_out.flags.c_contiguous or not host_out.flags.writeable:
                raise ValueError("host_out must be a writeable C-contiguous array")
            hdt = _NP_TO_TORCH.get(host_out.dtype)
            pinned = False
        else:
            hdt = host_out.dtype
            pinned = host_out.is_pinned() and host_out.is_contiguous()
        if hdt not in (torch.int32, torch.int64, torch.uint16, torch.uint8) or tuple(host_out.shape) != (n, self.G):
            raise ValueError("host_out must be an int32, int64, uint16 or uint8 CPU matrix of shape (%d, %d)"
                             % (n, self.G))
        narrow_dst = hdt in (torch.uint16, torch.uint8)
        if not threads:
            threads = _host_threads()
        auto = transport is None
        wide = None
        if auto and not narrow_dst:
            # nothing asked for: the copy engine writes a pinned int32 matrix itself ("direct"), every other
            # target is filled by host threads from int32 staging ("i32") - or, where this rank's share of
            # the host cores expands uint8 well above the PCIe rate, from uint8 staging ("u8")
            wide = "direct" if (not is_np and hdt == torch.int32 and pinned) else "i32"
            transport = wide
            if not _AUTO_STATE["narrow_overflowed"] and _shared_host_transport(threads) == "u8":
                transport = "u8"
        if transport in (None, "direct") and not is_np and hdt != torch.int64 and \
                (pinned or narrow_dst or transport == "direct"):
            return self._draw_to_host_direct(rows, scaling32, seed, cell0, host_out, chunk_cells, overflow_cap)
        if narrow_dst:
            raise ValueError("a uint16 / uint8 host matrix must be a CPU tensor and takes transport 'direct'")
        try:
            return self._draw_to_host_staged(rows, scaling32, seed, cell0, host_out, hdt, chunk_cells, overflow_cap,
                                             transport or "i32", threads, fresh)
        except OverflowError:
            if not (auto and transport == "u8"):
                raise
        # A deep regime (more than 1 % of the counts above 254): the uint8 transport nobody asked for does not
        # pay here.  The draw is a pure function of (seed, cell, gene): sample again through the wide transport,
        # and keep to it for the rest of the process.
        _AUTO_STATE["narrow_overflowed"] = True
        if wide == "direct":
            return self._draw_to_host_direct(rows, scaling32, seed, cell0, host_out, chunk_cells, None)
        return self._draw_to_host_staged(rows, scaling32, seed, cell0, host_out, hdt, chunk_cells, None,
                                         "i32", threads, False)

    def stream_chunks(self, rows, scaling32, seed, cell0, consume, transport="i32", chunk_cells=None,
                      overflow_cap=None):
        """Sample in cell chunks, move each chunk over PCIe in the transport format ("i32", "u16", "u8")
        into one of two pinned staging buffers and hand it to `consume(staged, lo, hi)` on a worker
        thread (`staged`: pinned uint8 tensor holding rows [lo, hi) in the transport format) while the
        next chunk is sampled and copied.  The matrix is never resident anywhere: this is the streamed
        path for outputs larger than HBM (config 5: 600 GB).  Returns the overflow list of the narrow
        transports as CPU tensors (flat index, exact value), unsorted, or None."""
        import concurrent.futures
        n, G = int(rows.numel()), self.G
        tdt = {"u8": torch.uint8, "u16": torch.uint16, "i32": torch.int32}.get(transport)
        if tdt is None:
            raise ValueError("transport must be 'i32', 'u16' or 'u8'")
        width = torch.empty(0, dtype=tdt).element_size()
        narrow = width < 4
        ramp_up = chunk_cells is None
        if chunk_cells is None:
            chunk_cells = max(1, min(n, _STAGE_BYTES // max(1, width * G)))
        if narrow and overflow_cap is None:
            overflow_cap = (1 << 20) if width == 2 else max(1 << 20, n * G // 100)
        dev = self.dev
        stage_dev = [torch.empty((chunk_cells, G), dtype=tdt, device=dev) for _ in range(2)]
        stage_host = _pinned_staging(dev, chunk_cells * G * width)
        if narrow:
            work = torch.empty((chunk_cells, G), dtype=torch.int32, device=dev)
            ovf_index = torch.empty(max(1, overflow_cap), dtype=torch.int64, device=dev)
            ovf_value = torch.empty(max(1, overflow_cap), dtype=torch.int32, device=dev)
            ovf_count = torch.zeros(1, dtype=torch.int64, device=dev)
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        pool = concurrent.futures.ThreadPoolExecutor(max_workers=1)
        pending = [None, None]                      # consumer still reading stage_host[k]

        def landed(k, lo, hi, event):
            event.synchronize()                     # the chunk is in stage_host[k]
            consume(stage_host[k][:(hi - lo) * G * width], lo, hi)

        # the first chunks are smaller so that the copy and the consumer start almost immediately
        bounds, lo = [], 0
        ramp = [max(1, chunk_cells // 8), max(1, chunk_cells // 4), max(1, chunk_cells // 2)] if ramp_up else []
        while lo < n:
            size = ramp.pop(0) if ramp else chunk_cells
            bounds.append((lo, min(n, lo + size)))
            lo = bounds[-1][1]
        try:
            for i, (lo, hi) in enumerate(bounds):
                k = i & 1
                if pending[k] is not None:
                    pending[k].result()             # stage_host[k] (and so stage_dev[k]) is free again
                buf = stage_dev[k][:hi - lo]
                if narrow:
                    self.draw(rows[lo:hi], scaling32[lo:hi], seed, cell0 + lo, out=work[:hi - lo])
                    nat.call("pst_narrow_counts", work.data_ptr(), hi - lo, G, G, buf.data_ptr(), G, 8 * width, lo,
                             ovf_index, ovf_value, int(overflow_cap), ovf_count, nat.stream_ptr(dev))
                else:
                    self.draw(rows[lo:hi], scaling32[lo:hi], seed, cell0 + lo, out=buf)
                done = torch.cuda.Event()
                done.record(main)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done)
                    stage_host[k][:(hi - lo) * G * width].copy_(buf.view(torch.uint8).view(-1), non_blocking=True)
                    copied = torch.cuda.Event()
                    copied.record(copy_stream)
                pending[k] = pool.submit(landed, k, lo, hi, copied)
            for f in pending:
                if f is not None:
                    f.result()
        finally:
            pool.shutdown(wait=True)
        if not narrow:
            return None
        count = int(ovf_count.item())
        if count > overflow_cap:
            raise OverflowError("%d counts reach the saturation value of the %s transport but the overflow "
                                "list holds %d: use a wider transport or a larger overflow_cap"
                                % (count, transport, overflow_cap))
        return ovf_index[:count].cpu(), ovf_value[:count].cpu()

    def _draw_to_host_staged(self, rows, scaling32, seed, cell0, host_out, hdt, chunk_cells, overflow_cap,
                             transport, threads, fresh=False):
        n, G = int(rows.numel()), self.G
        width = {"u8": 1, "u16": 2, "i32": 4}.get(transport)
        if width is None:
            raise ValueError("transport must be 'direct', 'i32', 'u16' or 'u8'")
        dst_bits = 32 if hdt == torch.int32 else 64
        dst_ptr = host_out.ctypes.data if isinstance(host_out, np.ndarray) else host_out.data_ptr()
        lib = nat.load()
        widen = lib.pst_host_widen if fresh else lib.pst_host_widen_stream
        self.overflow = None

        def expand(staged, lo, hi):
            rc = widen(staged.data_ptr(), 8 * width, dst_ptr + lo * G * (dst_bits // 8), dst_bits,
                                    (hi - lo) * G, int(threads))
            if rc != 0:
                raise nat.NativeError("pst_host_widen failed")

        listed = self.stream_chunks(rows, scaling32, seed, cell0, expand, transport, chunk_cells, overflow_cap)
        if listed is not None:
            index, value = listed
            lib.pst_host_apply_overflow(dst_ptr, dst_bits, 0, n * G, index.data_ptr(), value.data_ptr(),
                                        int(index.numel()))
        return host_out

    def _draw_to_host_direct(self, rows, scaling32, seed, cell0, host_out, chunk_cells=None, overflow_cap=None):
        """Chunks copied straight into a pinned CPU tensor of the transfer dtype (see draw_to_host)."""
        n = int(rows.numel())
        narrow = host_out.dtype in (torch.uint16, torch.uint8)
        width = host_out.element_size()
        if narrow and overflow_cap is None:
            # uint8 lists every count >= 255 (2e-4 of them at default depth; allow 50x that)
            overflow_cap = (1 << 20) if width == 2 else max(1 << 20, n * self.G // 100)
        if chunk_cells is None:
            # ~1 GiB copies keep the copy engine at its large-transfer rate; the first chunks are
            # smaller so that the device->host stream starts almost immediately
            chunk_cells = max(1, min(n, (1 << 30) // max(1, width * self.G)))
            ramp = [max(1, chunk_cells // 8), max(1, chunk_cells // 4), max(1, chunk_cells // 2)]
        else:
            ramp = []
        bounds, lo = [], 0
        while lo < n:
            size = ramp.pop(0) if ramp else chunk_cells
            bounds.append((lo, min(n, lo + size)))
            lo = bounds[-1][1]
        # staging buffers (what the copy engine reads) are double-buffered; in the narrow format the
        # int32 chunk is consumed by the narrowing kernel on the sampling stream, so one is enough
        stage = [torch.empty((chunk_cells, self.G), dtype=host_out.dtype, device=self.dev) for _ in range(2)]
        self.overflow = None
        if narrow:
            work = torch.empty((chunk_cells, self.G), dtype=torch.int32, device=self.dev)
            ovf_index = torch.empty(max(1, overflow_cap), dtype=torch.int64, device=self.dev)
            ovf_value = torch.empty(max(1, overflow_cap), dtype=torch.int32, device=self.dev)
            ovf_count = torch.zeros(1, dtype=torch.int64, device=self.dev)
        copy_stream = torch.cuda.Stream(device=self.dev)
        main = torch.cuda.current_stream(self.dev)
        free = [torch.cuda.Event(), torch.cuda.Event()]
        for i, (lo, hi) in enumerate(bounds):
            buf = stage[i & 1][:hi - lo]
            if i >= 2:
                main.wait_event(free[i & 1])
            if narrow:
                self.draw(rows[lo:hi], scaling32[lo:hi], seed, cell0 + lo, out=work[:hi - lo])
                nat.call("pst_narrow_counts", work.data_ptr(), hi - lo, self.G, self.G, buf.data_ptr(), self.G,
                         8 * width, lo, ovf_index, ovf_value, int(overflow_cap), ovf_count, nat.stream_ptr(self.dev))
            else:
                self.draw(rows[lo:hi], scaling32[lo:hi], seed, cell0 + lo, out=buf)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                host_out[lo:hi].copy_(buf, non_blocking=True)
                free[i & 1].record(copy_stream)
        copy_stream.synchronize()
        if narrow:
            count = int(ovf_count.item())
            if count > overflow_cap:
                raise OverflowError("%d counts reach the saturation value of %s but the overflow list holds %d: "
                                    "use a wider host buffer or a larger overflow_cap"
                                    % (count, str(host_out.dtype).replace("torch.", ""), overflow_cap))
            index, order = torch.sort(ovf_index[:count])            # appended in no particular order
            self.overflow = (index.cpu().numpy(), ovf_value[:count][order].cpu().numpy())
        return host_out

    def check(self):
        """One device->host read of the status word; raises like the reference would."""
        word = int(self.flags[0].item())
        if word:
            self.flags[0].zero_()
            raise_flags(word)

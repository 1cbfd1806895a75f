"""Compact exchange formats for the count matrix.

The reference writes `<job>_simulation.txt`, a dense TSV (prosstt/tree_utils.py:113-146); that
writer is kept byte-compatible in `tree_utils.save_matrices` for small runs.  At the sizes this
package samples (10^10 counts and up) the text dump would take hours, so two binary forms are
added, both readable without this package:

* CSR, built on the device (`to_csr`, pst_csr_fill) and stored in SciPy's `save_npz` layout
  (`save_sparse_npz`): `scipy.sparse.load_npz(path)` / AnnData read it directly.  About 44 % of
  the counts are zero at default parameters and far more at low depth.
* dense int32 `.npy` shards written chunk by chunk from the pinned host buffers the samplers
  stream into (`NpyShardWriter`): `np.load(path, mmap_mode="r")` reads them.
"""
import os

import numpy as np
import torch

from prosstt_b200 import _native as nat
from prosstt_b200 import stats as pstats


def to_csr(X, stats=None):
    """Device CSR of a device count matrix.  X: (n, G) int32 CUDA tensor.  Returns
    (indptr int64 (n+1,), indices int32 (nnz,), data int32 (nnz,)) CUDA tensors, columns
    ascending inside each row.  `stats` may pass a count_stats(X) result to skip that pass."""
    X = pstats._check_counts(X, "to_csr")
    n, G = X.shape
    dev = X.device
    if stats is None:
        stats = pstats.count_stats(X)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    if n:
        torch.cumsum(G - stats["cell_zeros"].to(torch.int64), dim=0, out=indptr[1:])
    nnz = int(indptr[-1].item())
    indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    data = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    nat.call("pst_csr_fill", X.data_ptr(), n, G, X.stride(0) if n else G, indptr, indices, data, flags,
             nat.stream_ptr(dev))
    if int(flags.item()) != 0:
        raise RuntimeError("row pointer does not match the matrix (was X modified after count_stats?)")
    return indptr, indices[:nnz], data[:nnz]


def save_sparse_npz(path, X, compressed=False):
    """Write X (device int32 matrix, or a (indptr, indices, data, shape) tuple from to_csr) in the
    layout of scipy.sparse.save_npz for a csr_matrix."""
    if isinstance(X, torch.Tensor):
        shape = tuple(X.shape)
        indptr, indices, data = to_csr(X)
    else:
        indptr, indices, data, shape = X
    arrays = dict(indices=indices.cpu().numpy(), indptr=indptr.cpu().numpy(), format=np.array("csr".encode("ascii")),
                  shape=np.asarray(shape, dtype=np.int64), data=data.cpu().numpy())
    if not str(path).endswith(".npz"):
        path = str(path) + ".npz"
    with open(path, "wb") as fh:
        (np.savez_compressed if compressed else np.savez)(fh, **arrays)
    return path


def load_sparse_npz(path):
    """(indptr, indices, data, shape) NumPy arrays of a file written by save_sparse_npz (or by
    scipy.sparse.save_npz for a csr matrix)."""
    with np.load(path) as z:
        fmt = z["format"].item()
        fmt = fmt.decode("ascii") if isinstance(fmt, bytes) else str(fmt)
        if fmt != "csr":
            raise ValueError("expected a csr matrix, found %r" % fmt)
        return z["indptr"], z["indices"], z["data"], tuple(int(v) for v in z["shape"])


def csr_to_dense(indptr, indices, data, shape):
    """Dense NumPy matrix of a CSR triple (small matrices / tests)."""
    out = np.zeros(shape, dtype=np.asarray(data).dtype)
    rows = np.repeat(np.arange(shape[0]), np.diff(np.asarray(indptr)))
    out[rows, np.asarray(indices)] = np.asarray(data)
    return out


def widen(Xn, overflow):
    """Exact int32 matrix from a narrow transfer format: Xn (n, G) uint8 or uint16 with saturated
    elements reading 255 / 65535, overflow = (flat index, value) arrays or the dict filled by
    `sample_density(host_out=(..., dict))`."""
    if isinstance(overflow, dict):
        overflow = (overflow["index"], overflow["value"])
    Xn = np.asarray(Xn)
    if Xn.dtype not in (np.dtype(np.uint8), np.dtype(np.uint16)):
        raise TypeError("expected a uint8 or uint16 matrix")
    out = Xn.astype(np.int32)
    index, value = np.asarray(overflow[0]), np.asarray(overflow[1])
    saturated = np.flatnonzero(Xn.reshape(-1) == np.iinfo(Xn.dtype).max)
    if not np.array_equal(saturated, index):
        raise ValueError("overflow list does not match the saturated elements of the matrix")
    out.reshape(-1)[index] = value
    return out


class NpyShardWriter(object):
    """Append-only writer of one dense (n_cells, G) `.npy` file, filled in row blocks.

    The header is written up front for the final shape, so the file is a valid `.npy` once every
    row has been appended; blocks come straight from the pinned host buffers of
    `sample_density(..., host_out=...)` / `DensitySession.step_to_host`.  One writer per rank
    gives one shard per GPU (`<stem>.rank<k>.npy`)."""

    def __init__(self, path, n_cells, G, dtype=np.int32):
        self.path = str(path)
        self.shape = (int(n_cells), int(G))
        self.dtype = np.dtype(dtype)
        self.rows = 0
        self._fh = open(self.path, "wb")
        np.lib.format.write_array_header_2_0(
            self._fh, {"descr": np.lib.format.dtype_to_descr(self.dtype), "fortran_order": False,
                       "shape": self.shape})

    def append(self, block):
        block = np.ascontiguousarray(block, dtype=self.dtype)
        if block.ndim != 2 or block.shape[1] != self.shape[1]:
            raise ValueError("block must have shape (rows, %d)" % self.shape[1])
        if self.rows + block.shape[0] > self.shape[0]:
            raise ValueError("more rows than announced (%d)" % self.shape[0])
        self._fh.write(memoryview(block).cast("B"))
        self.rows += block.shape[0]

    def close(self):
        if self._fh is None:
            return
        self._fh.close()
        self._fh = None
        if self.rows != self.shape[0]:
            raise ValueError("%s: %d of %d rows written" % (self.path, self.rows, self.shape[0]))

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            self.close()
        elif self._fh is not None:
            self._fh.close()
            self._fh = None
        return False


def shard_path(stem, rank, world):
    """`<stem>.npy` for one rank, `<stem>.rank<k>of<w>.npy` otherwise."""
    return "%s.npy" % stem if world == 1 else "%s.rank%dof%d.npy" % (stem, rank, world)

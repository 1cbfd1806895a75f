"""Summary statistics of a count matrix in one streaming pass on the GPU (pst_count_stats).

Library sizes, zero fractions and per-gene mean/variance are what every PROSSTT notebook computes
right after sampling (`X.sum(axis=1)`, `(X == 0).mean()`, `X.var(axis=0)`); on an 80 GB matrix a
NumPy pass is minutes, this kernel is one HBM read (≈13 ms per 80 GB)."""
import torch

from prosstt_b200 import _native as nat


def count_stats(X):
    """X: (n, G) int32 CUDA tensor (row stride >= G allowed).  Returns a dict of device tensors:
    cell_total (n,) int64, cell_zeros (n,) int32, gene_sum / gene_sumsq / gene_zeros (G,) int64."""
    if not (isinstance(X, torch.Tensor) and X.is_cuda and X.dtype == torch.int32 and X.dim() == 2):
        raise TypeError("count_stats expects a 2-D int32 CUDA tensor")
    if X.stride(1) != 1:
        X = X.contiguous()
    n, G = X.shape
    dev = X.device
    out = {
        "cell_total": torch.zeros(n, dtype=torch.int64, device=dev),
        "cell_zeros": torch.zeros(n, dtype=torch.int32, device=dev),
        "gene_sum": torch.zeros(G, dtype=torch.int64, device=dev),
        "gene_sumsq": torch.zeros(G, dtype=torch.int64, device=dev),
        "gene_zeros": torch.zeros(G, dtype=torch.int64, device=dev),
    }
    nat.call("pst_count_stats", X.data_ptr(), n, G, X.stride(0) if n else G, out["cell_total"], out["cell_zeros"],
             out["gene_sum"], out["gene_sumsq"], out["gene_zeros"], nat.stream_ptr(dev))
    return out


def gene_mean_var(stats, n_cells):
    """Per-gene sample mean and (population) variance from count_stats output, as float64 tensors."""
    mean = stats["gene_sum"].double() / n_cells
    var = stats["gene_sumsq"].double() / n_cells - mean * mean
    return mean, var

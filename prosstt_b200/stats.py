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


def new_gene_stats(G, device):
    """Zeroed per-gene accumulators for the fused summaries of the draw (CountEngine.draw(gene_stats=...),
    DensitySession.step(gene_stats=...)): they are added to, so one set can span many chunks or steps."""
    return {k: torch.zeros(G, dtype=torch.int64, device=device) for k in ("gene_sum", "gene_sumsq", "gene_zeros")}


def gene_mean_var(stats, n_cells):
    """Per-gene sample mean and (population) variance from count_stats output, as float64 tensors."""
    mean = stats["gene_sum"].double() / n_cells
    var = stats["gene_sumsq"].double() / n_cells - mean * mean
    return mean, var


_MODES = {"normalize": 0, "normalize_log1p": 1, "log1p": 2}


def _check_counts(X, who):
    if not (isinstance(X, torch.Tensor) and X.is_cuda and X.dtype == torch.int32 and X.dim() == 2):
        raise TypeError("%s expects a 2-D int32 CUDA tensor" % who)
    return X if X.stride(1) == 1 else X.contiguous()


def transform_counts(X, scalings=None, mode="normalize", out=None):
    """fp32 (n, G) transform of a device count matrix in one pass (pst_transform_counts):
    mode "normalize" = X / scalings[:, None] (compare_axolotl.ipynb cell 14), "normalize_log1p" =
    log(X / scalings[:, None] + 1), "log1p" = log(X + 1) (minimal_example.ipynb cell 6).
    scalings: (n,) tensor or array (any float dtype); out: optional fp32 CUDA tensor to fill."""
    X = _check_counts(X, "transform_counts")
    if mode not in _MODES:
        raise ValueError("mode must be one of %s" % sorted(_MODES))
    n, G = X.shape
    dev = X.device
    s32 = None
    if mode != "log1p":
        if scalings is None:
            raise ValueError("scalings are required for mode %r" % mode)
        s32 = torch.as_tensor(scalings).to(device=dev, dtype=torch.float32).contiguous()
        if s32.shape != (n,):
            raise ValueError("scalings must have shape (%d,)" % n)
    if out is None:
        out = torch.empty((n, G), dtype=torch.float32, device=dev)
    elif not (out.is_cuda and out.dtype == torch.float32 and out.shape == (n, G) and out.stride(1) == 1):
        raise ValueError("out must be an fp32 CUDA tensor of shape (%d, %d) with unit column stride" % (n, G))
    nat.call("pst_transform_counts", X.data_ptr(), n, G, X.stride(0) if n else G, s32, _MODES[mode],
             out.data_ptr(), out.stride(0) if n else G, nat.stream_ptr(dev))
    return out


def normalize(X, scalings, log=False):
    """X / scalings per cell, optionally followed by log(. + 1)."""
    return transform_counts(X, scalings, "normalize_log1p" if log else "normalize")


def log1p(X):
    """log(X + 1) as fp32."""
    return transform_counts(X, None, "log1p")

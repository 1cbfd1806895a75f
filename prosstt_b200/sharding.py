"""Cells partitioned over the GPUs of one box.

The path has no data-path collective: every rank holds the small replicated state and samples
its own contiguous slice of the cells with `shard=(rank, world)`; all random streams are keyed by
the global cell index, so the union of the slices is bit-identical to the single-GPU result.
`gather_counts` is the optional epilogue: an all_gather of the slabs (NCCL over NVLink/NVSwitch
when the tensors live on GPUs, gloo on CPU tensors).  `allreduce_gene_stats` is the cheap one: the
per-gene sums of `stats.count_stats` (a few hundred KB) summed over the ranks, so every rank holds the
whole-job gene mean / variance / zero fraction without moving a count."""
import os

import torch
import torch.distributed as dist


def rank_and_world():
    """(rank, world) from torch.distributed if initialised, else from the torchrun environment."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n, rank, world):
    """Contiguous cell range [lo, hi) of `rank`; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("need 0 <= rank < world")
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_counts(local, n_total, group=None):
    """all_gather of per-rank slabs with (possibly) different numbers of rows -> (n_total, G) tensor
    on every rank, rows in global cell order.  `local` is this rank's (n_local, G) tensor."""
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    mine = local
    if local.shape[0] != pad:
        mine = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        mine[:local.shape[0]] = local
    slabs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(slabs, mine.contiguous(), group=group)
    return torch.cat([s[:hi - lo] for s, (lo, hi) in zip(slabs, sizes)])


def allreduce_gene_stats(stats, n_local, group=None):
    """Sum the per-gene fields of a `stats.count_stats` result (gene_sum, gene_sumsq, gene_zeros)
    and the number of cells over all ranks, in place.  Returns (stats, n_total).  The per-cell
    fields stay rank-local (they belong to this rank's cells)."""
    n = torch.tensor([int(n_local)], dtype=torch.int64, device=stats["gene_sum"].device)
    for t in (stats["gene_sum"], stats["gene_sumsq"], stats["gene_zeros"], n):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return stats, int(n.item())

"""Host-side sharding logic under torch.distributed (gloo, world_size 2, CPU only): the
per-rank cell ranges tile the global range, and rank-local results gathered in rank order equal
the single-process result.  The device kernels are keyed by the global cell index (GPU test
test_counts_bit_exact_across_partitions); here the same keying is exercised with the oracle's
Philox on the CPU so the N>1 control flow is covered without a GPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from prosstt_b200 import _native as nat
from prosstt_b200.simulation import _shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _global_stream(seed, first, n):
    """What pst_uniform_f64 computes for elements [first, first+n): Philox keyed by the global index."""
    from oracle import prosstt_oracle as orc
    idx = first + np.arange(n, dtype=np.uint64)
    ctr = np.stack([idx & 0xFFFFFFFF, idx >> 32, np.full(n, nat.TAG_DENSITY_U), np.zeros(n)], axis=1).astype(np.uint32)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32), (n, 2))
    w = orc.philox4x32_10(ctr, key).astype(np.uint64)
    return ((w[:, 0] >> 5) * 67108864.0 + (w[:, 1] >> 6)) / 9007199254740992.0


def _worker(rank, world, port, n_cells, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = _shard_range(n_cells, (rank, world))
    local = torch.from_numpy(_global_stream(nat.derive_seed(seed, 0), lo, hi - lo))
    from prosstt_b200.sharding import gather_counts, rank_and_world, shard_range
    assert rank_and_world() == (rank, world) and shard_range(n_cells, rank, world) == (lo, hi)
    full = gather_counts(local.reshape(-1, 1), n_cells).reshape(-1)      # the optional gather epilogue
    # the cheap epilogue: per-gene summaries summed over ranks (fields as stats.count_stats returns them)
    from prosstt_b200.sharding import allreduce_gene_stats
    Xl = (local.reshape(-1, 1) * torch.arange(1, 4)).to(torch.int64)      # (n_local, 3) pseudo counts
    st = {"gene_sum": Xl.sum(0), "gene_sumsq": (Xl * Xl).sum(0), "gene_zeros": (Xl == 0).sum(0)}
    st, n_tot = allreduce_gene_stats(st, hi - lo)
    Xf = (full.reshape(-1, 1) * torch.arange(1, 4)).to(torch.int64)
    assert n_tot == n_cells and torch.equal(st["gene_sum"], Xf.sum(0))
    assert torch.equal(st["gene_sumsq"], (Xf * Xf).sum(0)) and torch.equal(st["gene_zeros"], (Xf == 0).sum(0))
    tmax = torch.tensor([float(rank + 1)])
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)          # the bench's max-over-ranks timing
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), tmax.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_tile_the_cells():
    for n in (0, 1, 7, 1000, 1000003):
        for world in (1, 2, 3, 8):
            spans = [_shard_range(n, (r, world)) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    assert _shard_range(10, None) == (0, 10)
    import pytest
    with pytest.raises(ValueError):
        _shard_range(10, (2, 2))


def test_two_ranks_gather_equals_single_process(tmp_path):
    n_cells, seed, world = 1001, 77, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_cells, seed, str(tmp_path)), nprocs=world, join=True)
    gathered = np.load(tmp_path / "gathered.npy")
    single = _global_stream(nat.derive_seed(seed, 0), 0, n_cells)
    assert np.array_equal(gathered, single)
    assert np.load(tmp_path / "tmax.npy")[0] == world


def test_seed_derivation_is_stable():
    # stage keys must never change silently: they define the random streams
    assert nat.derive_seed(44, 0) == nat.derive_seed(44, 0)
    assert len({nat.derive_seed(44, s) for s in range(8)}) == 8
    assert nat.split_seed(2 ** 64 + 5) == 5
    np.random.seed(3)
    a = nat.split_seed(None)
    np.random.seed(3)
    assert a == nat.split_seed(None)

#!/usr/bin/env python
"""Host side of the device->host path on this box: pinned D2H rate (1 and 2 copy streams), and the rate of
pst_host_widen / pst_host_widen_stream (uint8 -> int32, int32 -> int64, uint8 -> int64, int32 copy) into pinned
and into fresh pageable memory for a few thread counts.  Run on every rank concurrently under torchrun to see the aggregate."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prosstt_b200 import _native as nat  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


lib = nat.load()
cores = len(os.sched_getaffinity(0))
nbytes = 1 << 30
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
h = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
out = ["rank %d/%d cores %d" % (rank, world, cores)]
for streams in (1, 2):
    ss = [torch.cuda.Stream(device=dev) for _ in range(streams)]
    barrier()
    t0 = time.perf_counter()
    for rep in range(6):
        for k, s in enumerate(ss):
            with torch.cuda.stream(s):
                part = nbytes // streams
                h[rep & 1][k * part:(k + 1) * part].copy_(d[k * part:(k + 1) * part], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    out.append("D2H pinned, %d stream(s): %.1f GB/s" % (streams, 6 * nbytes / dt / 1e9))
src = h[0]
n = nbytes
for name, sb, db, cnt in (("u8->i32", 8, 32, n // 4), ("i32->i64", 32, 64, n // 8), ("u8->i64", 8, 64, n // 8),
                          ("i32->i32", 32, 32, n // 4)):
    dst_pinned = h[1]
    for threads in sorted(set([1, 4, 8, max(1, cores // max(1, world)), cores])):
        line = "%s threads %2d:" % (name, threads)
        for label, fn in (("plain", lib.pst_host_widen), ("stream", lib.pst_host_widen_stream)):
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                fn(src.data_ptr(), sb, dst_pinned.data_ptr(), db, cnt, threads)
            dt = (time.perf_counter() - t0) / 3
            rates = []
            for advise in (False, True):
                fresh = np.empty(cnt * db // 8, dtype=np.uint8)
                if advise:
                    lib.pst_host_prepare(fresh.ctypes.data, fresh.nbytes)
                t0 = time.perf_counter()
                fn(src.data_ptr(), sb, fresh.ctypes.data, db, cnt, threads)
                rates.append(cnt / (time.perf_counter() - t0) / 1e9)
                del fresh
            line += "  %s: pinned %.2f Gcounts/s (%.1f GB/s written), fresh %.2f, fresh+hugepage advice %.2f;" % (
                label, cnt / dt / 1e9, cnt * db / 8 / dt / 1e9, rates[0], rates[1])
        out.append(line)
for r in range(world):
    barrier()
    if r == rank:
        print("\n".join(out), flush=True)
if world > 1:
    dist.destroy_process_group()

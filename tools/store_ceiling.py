#!/usr/bin/env python
"""Write-only HBM ceiling: time pst_store_fill (pure 128-bit stores) over a buffer far larger
than L2, beside a torch copy (read+write, what MEASURED_PEAKS.json quotes) and cudaMemset.  Writes
profiles/store_ceiling.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prosstt_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
n = 5 * (1 << 30)                       # 20 GiB of int32
buf = torch.empty(n, dtype=torch.int32, device=dev)
st = nat.stream_ptr(dev)


def best(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return min(out), sorted(out)[len(out) // 2]


fill_ms, fill_med = best(lambda: nat.call("pst_store_fill", buf, n, 7, st))
assert int(buf[-1].item()) == 7 and int(buf[12345].item()) == 7
memset_ms, _ = best(lambda: buf.zero_())
half = n // 2
copy_ms, _ = best(lambda: buf[:half].copy_(buf[half:]))
rec = {"store_gbs": 4 * n / fill_ms / 1e6, "store_gbs_median": 4 * n / fill_med / 1e6,
       "memset_gbs": 4 * n / memset_ms / 1e6, "copy_gbs_read_plus_write": 2 * 4 * half / copy_ms / 1e6,
       "bytes": 4 * n, "how": "pst_store_fill (128-bit stores, grid 148 x 8 CTAs of 256 threads) over 20 GiB, best of 10 "
                             "with CUDA events; torch zero_() and a 10 GiB -> 10 GiB copy_ beside it",
       "gpu": torch.cuda.get_device_name(0)}
print(json.dumps(rec))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "store_ceiling.json"), "w"))

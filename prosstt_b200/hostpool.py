"""Recycled host memory for the result matrices of the reference-shaped calls.

`sample_density(tree, N, alpha, beta)` returns a fresh int64 NumPy matrix like the reference does
(simulation.py:651): 8 B per count of host memory that nobody has touched yet, so the threads that
expand the counts into it take one page fault (and one page of zeroing) per 4 KiB - measured 5.5e9
counts/s with 16 threads, against 1.3-1.8e10 into memory that is already mapped
(profiles/r02_host_bw_1gpu.txt).  A script that samples more than once (every notebook of the
reference does) drops the previous result before or while it asks for the next one, so the buffer of
a result that has been garbage-collected is kept here and handed out again: same array semantics
(writeable, C-contiguous, views keep it alive), no page faults from the second call on.

Only matrices of at least 4 MiB are pooled; at most PST_HOST_POOL_GB (default: a quarter of the
physical memory, divided among the ranks that share the host) of free buffers is retained; `release()` returns them to the OS; PST_HOST_POOL_GB=0
disables the pool (every result is then `np.empty`)."""
import collections
import mmap
import os
import threading
import weakref

import numpy as np

_LOCK = threading.Lock()
_FREE = []                    # [nbytes, mmap] of results that were garbage-collected, pages mapped
_RETURNED = collections.deque()   # what finalizers hand back; moved into _FREE under the lock by the next caller
_MIN_BYTES = 4 << 20
_HUGE = 2 << 20


def _cap_bytes():
    env = os.environ.get("PST_HOST_POOL_GB")
    if env is not None:
        try:
            return max(0, int(float(env) * (1 << 30)))
        except ValueError:
            return 0
    try:
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))      # ranks of one job share the host
    except ValueError:
        ranks = 1
    try:
        return os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") // (4 * ranks)
    except (ValueError, OSError, AttributeError):
        return 0


def _give_back(buf, size):
    # runs wherever the last reference dies - possibly inside the garbage collector while this module holds
    # its lock: only an atomic append here
    _RETURNED.append([size, buf])


def _collect():
    """Move returned buffers into the free list and enforce the cap (call with _LOCK held)."""
    while True:
        try:
            _FREE.append(_RETURNED.popleft())
        except IndexError:
            break
    cap = _cap_bytes()
    total = sum(b[0] for b in _FREE)
    while _FREE and total > cap:                     # oldest first
        total -= _FREE.pop(0)[0]


def release():
    """Return every retained buffer to the OS."""
    with _LOCK:
        _collect()
        del _FREE[:]


def retained_bytes():
    with _LOCK:
        _collect()
        return sum(b[0] for b in _FREE)


def result_array(shape, dtype):
    """(array, fresh): a writeable C-contiguous array of `shape`/`dtype` and whether its memory has never
    been touched (then the first writes fault its pages in)."""
    dtype = np.dtype(dtype)
    count = 1
    for d in shape:
        count *= int(d)
    nbytes = count * dtype.itemsize
    if nbytes < _MIN_BYTES or _cap_bytes() == 0:
        return np.empty(shape, dtype=dtype), True
    need = (nbytes + _HUGE - 1) // _HUGE * _HUGE
    buf = None
    with _LOCK:
        _collect()
        best = None
        for i, (size, _) in enumerate(_FREE):        # smallest retained buffer that fits without wasting half
            if need <= size <= 2 * need and (best is None or size < _FREE[best][0]):
                best = i
        if best is not None:
            need, buf = _FREE.pop(best)
    fresh = buf is None
    if fresh:
        # private like any malloc'ed array: a forked child gets its own copy-on-write view
        buf = mmap.mmap(-1, need, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        if hasattr(buf, "madvise") and hasattr(mmap, "MADV_HUGEPAGE"):
            try:
                buf.madvise(mmap.MADV_HUGEPAGE)
            except OSError:
                pass
    # `owner` is the object every view of the result keeps alive (NumPy does not collapse a base chain
    # past a non-ndarray buffer); when the last of them is gone the mapping comes back to the pool
    owner = np.frombuffer(buf, dtype=np.uint8)
    weakref.finalize(owner, _give_back, buf, need).atexit = False
    return owner[:nbytes].view(dtype).reshape(shape), fresh

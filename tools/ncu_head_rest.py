#!/usr/bin/env python
"""Split the executed SASS of the draw kernel (ncu --set full --import-source on, --page source --csv) into
the always-executed body of the per-cell loop and everything else, by opcode and by execution frequency."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {n: i for i, n in enumerate(rows[hi])}
data = rows[hi + 1:]


def num(r, n):
    try:
        return float(r[col[n]])
    except (ValueError, IndexError):
        return 0.0


def opcode(r):
    s = r[col["Source"]].strip().split()
    if not s:
        return "?"
    return (s[1] if s[0].startswith("@") else s[0]).split(".")[0]


tot = sum(num(r, "Instructions Executed") for r in data)
thr = sum(num(r, "Thread Instructions Executed") for r in data)
c = collections.Counter(round(num(r, "Instructions Executed")) for r in data)
H = max((k for k in c if c[k] > 60), key=lambda k: k * c[k])
print("warp-instructions %.4e  thread-instructions %.4e  head iterations %d  per iteration %.1f" % (tot, thr, H, tot / H))
head, rest = collections.Counter(), collections.Counter()
for r in data:
    e = num(r, "Instructions Executed")
    (head if round(e) == H else rest)[opcode(r)] += e
print("always-executed head body: %.1f" % (sum(head.values()) / H))
print("  " + " ".join("%s %.0f" % (k, v / H) for k, v in head.most_common()))
print("everything else: %.1f" % (sum(rest.values()) / H))
print("  " + " ".join("%s %.1f" % (k, v / H) for k, v in rest.most_common(26)))
print("by execution frequency (per head iteration): frequency -> warp-instructions")
cls = collections.Counter()
for r in data:
    e = num(r, "Instructions Executed")
    if round(e) != H:
        cls[round(e / H, 2)] += e / H
print("  " + "  ".join("%.2f:%.1f" % kv for kv in sorted(cls.items(), key=lambda kv: -kv[1])[:24]))

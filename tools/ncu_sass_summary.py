#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: totals per stall reason and a per-region
listing (regions split at barriers / backward branches) with instruction and sample shares."""
import csv
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = rows[hdr_i + 1:]
def f(r, n):
    try:
        return float(r[col[n]])
    except (ValueError, IndexError):
        return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in data)
tot_samp = sum(f(r, "# Samples") for r in data)
print("instructions executed (warp-level): %.3e   samples: %d   sass lines: %d" % (tot_inst, tot_samp, len(data)))
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = sorted(((sum(f(r, n) for r in data), n) for n in stalls), reverse=True)
print("stall reasons (all samples): " + ", ".join("%s %.1f%%" % (n[6:], 100 * v / max(tot_samp, 1)) for v, n in agg[:8]))
thr = sum(f(r, "Thread Instructions Executed") for r in data)
print("avg active threads / warp-inst: %.1f" % (thr / max(tot_inst, 1)))
# regions
width = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("\nregion listing (every %d sass lines): idx  inst%%  samp%%  avg_thr  first-instruction" % width)
for s in range(0, len(data), width):
    chunk = data[s:s + width]
    ci = sum(f(r, "Instructions Executed") for r in chunk)
    cs = sum(f(r, "# Samples") for r in chunk)
    ct = sum(f(r, "Thread Instructions Executed") for r in chunk)
    ops = {}
    for r in chunk:
        op = r[col["Source"]].strip().split()[0] if r[col["Source"]].strip() else "?"
        if op.startswith("@"):
            op = r[col["Source"]].strip().split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + f(r, "Instructions Executed")
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
    print("%5d  %5.1f  %5.1f  %5.1f   %s" % (s, 100 * ci / tot_inst, 100 * cs / max(tot_samp, 1), ct / max(ci, 1),
                                            " ".join("%s:%.0f%%" % (k, 100 * v / max(ci, 1)) for k, v in top)))

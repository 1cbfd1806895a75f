#!/usr/bin/env python
"""Rebuild the tracked summaries under profiles/ from the captures a GPU run left in gpurun_out/:
  prof.ncu-rep (ncu --set full of the draw kernel), launches csv (ncu launch list of bench.py), bench json."""
import collections
import csv
import json
import subprocess
import sys

rep, launches, bench_json = sys.argv[1:4]
COUNTS = 4e9          # tools/sampler_bench.py --cells 200000 x 20000 genes
TAG = "r02"
WANT = """dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed gpu__time_duration.sum
l1tex__t_sector_hit_rate.pct launch__block_size launch__grid_size launch__registers_per_thread launch__shared_mem_per_block_static
lts__t_sector_hit_rate.pct sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__thread_inst_executed_per_inst_executed.ratio""".split()

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
head = ("# ncu --set full, draw_counts_kernel<10,true,false,false> (hybrid sampler), 200000 cells x 20000 genes (4e9 counts), "
        "final round-2 kernel (tools/sampler_bench.py --cells 200000 --samplers hybrid)")
with open("profiles/%s_draw_counts_ncu_raw.txt" % TAG, "w") as fh:
    fh.write(head + "\n" + "".join("%-95s %-12s %s\n" % (w, d[w][0], d[w][1]) for w in WANT if w in d))
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
rd = float(d["dram__bytes_read.sum"][1]) * scale[d["dram__bytes_read.sum"][0]]
wr = float(d["dram__bytes_write.sum"][1]) * scale[d["dram__bytes_write.sum"][0]]
json.dump({"bytes_per_count": round((rd + wr) / COUNTS, 3),
           "source": "ncu --set full on draw_counts_kernel (hybrid), 200000 cells x 20000 genes (4e9 counts), final round-2 "
                     "kernel: dram__bytes_read.sum %.3f GB + dram__bytes_write.sum %.3f GB "
                     "(profiles/%s_draw_counts_ncu_raw.txt)" % (rd / 1e9, wr / 1e9, TAG),
           "algorithmic_bytes_per_count": 4}, open("profiles/traffic.json", "w"))
for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
          "smsp__thread_inst_executed_per_inst_executed.ratio"):
    print(k, d[k])
print("thread-instructions per count: %.1f" % (float(d["smsp__inst_executed.sum"][1]) *
                                                float(d["smsp__thread_inst_executed_per_inst_executed.ratio"][1]) / COUNTS))

# executed SASS by opcode
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
H = max((k for k in c if c[k] > 100), key=lambda k: k * c[k])
headc, rest = collections.Counter(), collections.Counter()
for r in data:
    e = num(r, "Instructions Executed")
    (headc if round(e) == H else rest)[opcode(r)] += e
with open("profiles/%s_draw_counts_sass_mix.txt" % TAG, "w") as fh:
    fh.write("# draw_counts_kernel (hybrid), 200000 cells x 20000 genes (4e9 counts): executed SASS by opcode\n")
    fh.write("# source: ncu --set full --import-source on (tools/sampler_bench.py --cells 200000 --samplers hybrid), final round-2 kernel\n")
    fh.write("warp-instructions executed: %.4e  (%.2f per count, %.1f thread-instructions per count)\n" % (tot, tot / COUNTS, thr / COUNTS))
    fh.write("head iterations (one warp x one cell x 128 genes): %d ; warp-instructions per iteration: %.1f\n" % (H, tot / H))
    hs, rs = sum(headc.values()), sum(rest.values())
    fh.write("\nalways-executed head body: %.1f warp-instructions per iteration (%.1f%% of all)\n" % (hs / H, 100 * hs / tot))
    fh.write("".join("  %-10s %6.1f\n" % (k, v / H) for k, v in headc.most_common()))
    fh.write("\neverything else (enqueue blocks, queue drains, mixture, chunk set-up): %.1f warp-instructions per iteration "
             "(%.1f%% of all)\n" % (rs / H, 100 * rs / tot))
    fh.write("".join("  %-10s %6.1f\n" % (k, v / H) for k, v in rest.most_common(24)))
print("head body %.1f, rest %.1f warp-instructions per iteration" % (hs / H, rs / H))

# launch list shares
rows = list(csv.reader(l for l in open(launches) if not l.startswith("==")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
col = {n: i for i, n in enumerate(rows[hi])}
agg, cnt = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(col):
        continue
    name = r[col["Kernel Name"]].split("(")[0]
    v, u = float(r[col["Metric Value"]]), r[col["Metric Unit"]]
    agg[name] += v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    cnt[name] += 1
total = sum(agg.values())
for k, v in agg.most_common(7):
    print("%-60s %8.3f ms %5.2f%% (%d)" % (k[:60], v, 100 * v / total, cnt[k]))
print("total", total)
open("profiles/%s_launches_bench_200k.csv" % TAG, "w").write(open(launches).read())
b = json.loads(open(bench_json).read().strip().splitlines()[-1])
open("profiles/%s_bench_c4_1gpu.json" % TAG, "w").write(json.dumps(b) + "\n")
# speed of light of the instruction-bound kernel: issue slots of a B200 and the
# instructions a perfectly adaptive sampler of this family would need
winst = float(d["smsp__inst_executed.sum"][1]) / COUNTS
json.dump({"issue_slots_per_s": 148 * 4 * 1.965e9, "warp_inst_per_count": round(winst, 3),
           "thread_inst_per_count": round(thr / COUNTS, 1), "speed_of_light_warp_inst_per_count": round(57.5 / 32, 3),
           "issue_utilisation": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][1]) / 100,
           "source": "148 SMs x 4 schedulers x 1.965 GHz issue slots; %.2f warp-instructions per count executed (ncu, "
                     "profiles/%s_draw_counts_ncu_raw.txt); speed of light = 57.5 thread-instructions per count (Philox 13, "
                     "parameterisation 14 + 4 MUFU, E[X]+1 = 4.5 cdf terms x 4, store and route 3; DESIGN.md 4.2)" % (winst, TAG)},
          open("profiles/issue_bound.json", "w"))
print("bench: value %.4e  ms/step %.2f  roofline %.1f GB/s frac %.4f  e2e %.3e  default_api %.3e  u16 %.3e  u8 %.3e" % (
    b["value"], b["ms_per_step"], b["roofline"]["achieved"], b["roofline"]["frac"], b["e2e"]["value"],
    b["e2e"]["default_api"]["value"], b["e2e"]["narrow_u16"]["value"], b["e2e"]["narrow_u8"]["value"]))

#!/bin/bash
# same-box A/B of library variants (tools/lib_<name>.so): time (3 reps, best) at two depths + executed instructions
names="$@"
run() { env PST_LIB=tools/lib_$1.so python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 3 --scale-mean $2 2>&1 | grep -E "^hybrid|rror"; }
for round in 1 2; do for l in $names; do echo "== $l depth 0 (round $round)"; run $l 0; done; done
for l in $names; do echo "== $l depth 1.5"; run $l 1.5; done
for l in $names; do echo "== inst $l"; env PST_LIB=tools/lib_$l.so ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:draw_counts_kernel -s 1 -c 1 python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 1 2>&1 | grep -E "smsp__|gpu__time"; done

#!/bin/bash
run() { echo "== $*"; env "$@" python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 3 2>&1 | grep -E "^hybrid|rror"; }
inst() { env "$@" ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:draw_counts_hybrid -s 1 -c 1 python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 1 2>&1 | grep -E "smsp__"; }
for l in old new old new; do run PST_LIB=tools/lib_$l.so; done
for l in old new; do echo "== inst $l"; inst PST_LIB=tools/lib_$l.so; done

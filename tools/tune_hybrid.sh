#!/bin/bash
run() { echo "== $*"; env "$@" python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 3 2>&1 | grep -E "^hybrid|rror"; }
for st in 0 8 16 24; do
for kf in 6 8 10; do
run PST_LIB=tools/lib_s$st.so PST_HY_KFIX=$kf
done; done

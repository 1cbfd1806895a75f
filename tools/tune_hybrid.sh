#!/bin/bash
run() { echo "== $*"; env "$@" python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 3 2>&1 | grep -E "^hybrid|rror"; }
run A=1
run PST_LIB=tools/lib_c7.so
run PST_LIB=tools/lib_c6.so
run PST_HY_KFIX=8
run PST_HY_KFIX=12

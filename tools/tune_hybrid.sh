#!/bin/bash
run() { echo "== $*"; env "$@" python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 3 2>&1 | grep -E "^hybrid|rror"; }
run PST_LIB=tools/lib_old.so
run PST_LIB=tools/lib_new.so
run PST_LIB=tools/lib_old.so
run PST_LIB=tools/lib_new.so

#!/bin/bash
# developer tool: sweep the hybrid sampler's profiling knobs on the GPU box
run() { echo "== $*"; env "$@" python tools/sampler_bench.py --cells 100000 --samplers hybrid --reps 3 2>&1 | grep -E "hybrid|rror"; }
run PST_HY_MU_MAX=32 PST_HY_VAR_MAX=400
run PST_HY_MU_MAX=32 PST_HY_VAR_MAX=400 PST_HY_CTAS=16
run PST_HY_MU_MAX=32 PST_HY_VAR_MAX=400 PST_LIB=tools/lib_c6.so
run PST_HY_MU_MAX=32 PST_HY_VAR_MAX=400 PST_LIB=tools/lib_c5.so
run PST_HY_KFIX=10 PST_HY_MU_MAX=32 PST_HY_VAR_MAX=400
run PST_HY_KFIX=12 PST_HY_MU_MAX=32 PST_HY_VAR_MAX=400
run PST_HY_KFIX=10 PST_HY_MU_MAX=24 PST_HY_VAR_MAX=300

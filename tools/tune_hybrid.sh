#!/bin/bash
# developer tool (GPU box): sweep the hybrid sampler's knobs.  Needs a developer build:
#   PST_NVCC_DEFS="-DPST_DEV_KNOBS" python prosstt_b200/build.py --force
run() { echo "== $*"; env "$@" python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 3 2>&1 | grep -E "^hybrid|rror"; }
run PST_HY_KFIX=10
run PST_HY_VAR_MAX=800
run PST_HY_VAR_MAX=1e9
run PST_HY_MU_MAX=24 PST_HY_VAR_MAX=1e9

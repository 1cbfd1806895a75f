#!/bin/bash
# one-GPU validation session under gpurun: GPU tests, smoke, the default bench line, both samplers at four depths
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/v_pytest.txt
python __graft_entry__.py --smoke > gpurun_out/v_smoke.txt 2>&1
python bench.py > gpurun_out/v_bench_c4.json 2> gpurun_out/v_bench.err
for m in -2 0 1.5 3; do echo "== depth $m"; python tools/sampler_bench.py --cells 100000 --reps 3 --scale-mean $m 2>&1 | grep -E "^hybrid|^gamma|rror"; done > gpurun_out/v_sampler.txt 2>&1
cat gpurun_out/v_pytest.txt; tail -2 gpurun_out/v_smoke.txt; tail -n 3 gpurun_out/v_bench.err; cut -c1-120 gpurun_out/v_sampler.txt

#!/bin/bash
# One GPU-box session: tests, deep goodness-of-fit, sampler timings, bench line.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s_pytest.txt
python tools/gof_deep.py > gpurun_out/s_gof.txt 2>&1
for m in 0 1.5 3; do python tools/sampler_bench.py --cells 200000 --scale-mean $m 2>&1 | grep -E "hybrid|gamma|rror" ; done > gpurun_out/s_sampler.txt 2>&1
python bench.py > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:draw_counts_hybrid -s 1 -c 1 python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 1 2>&1 | grep -E "smsp__|gpu__time" > gpurun_out/s_inst.txt
cat gpurun_out/s_pytest.txt gpurun_out/s_sampler.txt gpurun_out/s_inst.txt; tail -c 1500 gpurun_out/s_bench.json

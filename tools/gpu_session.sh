#!/bin/bash
set -x
mkdir -p gpurun_out
for l in pipe pipe3 pipe4 pipe5 pipe6 pipe7; do for m in 0 3; do echo "== $l depth $m"; PST_LIB=tools/lib_$l.so python tools/sampler_bench.py --cells 100000 --samplers gamma_poisson --reps 3 --scale-mean $m 2>&1 | grep -E "^hybrid|^gamma|rror"; done; done > gpurun_out/l_sampler.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/l_pytest.txt
for l in pipe5 pipe6; do echo "== inst $l"; env PST_LIB=tools/lib_$l.so ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active --clock-control none -k regex:draw_counts_mixture_kernel -s 1 -c 1 python tools/sampler_bench.py --cells 50000 --samplers gamma_poisson --reps 1 2>&1 | grep -E "smsp__|gpu__time|sm__inst"; done > gpurun_out/l_inst.txt 2>&1
cat gpurun_out/l_pytest.txt; grep -E "^==|^gamma" gpurun_out/l_sampler.txt | cut -c1-100; cat gpurun_out/l_inst.txt

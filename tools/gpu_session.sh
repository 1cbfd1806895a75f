#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/n_pytest.txt
bash tools/sanitize.sh > gpurun_out/n_sanitize.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:draw_counts_mixture_kernel -s 1 -c 1 -o gpurun_out/prof_r02_mixture -f python tools/sampler_bench.py --cells 50000 --samplers gamma_poisson --reps 1 > gpurun_out/n_ncu.log 2>&1
ncu -i gpurun_out/prof_r02_mixture.ncu-rep --page raw --csv > gpurun_out/n_mixture_raw.csv 2>/dev/null
timeout 400 python tools/gof_deep.py --draws 1e9 > gpurun_out/n_gof.txt 2>&1
cat gpurun_out/n_pytest.txt gpurun_out/n_sanitize.txt; cut -c1-330 gpurun_out/n_gof.txt

#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py --sampler gamma_poisson --no-cpu-baseline --e2e-steps 3 > gpurun_out/r_bench_c4_gamma_poisson.json 2> gpurun_out/r_bench.err
for m in -2 1.5; do echo "== depth $m"; python tools/sampler_bench.py --cells 100000 --samplers gamma_poisson --reps 3 --scale-mean $m 2>&1 | grep -E "^hybrid|^gamma|rror"; done > gpurun_out/r_sampler.txt 2>&1
tail -n 2 gpurun_out/r_bench.err; cat gpurun_out/r_sampler.txt | cut -c1-120; head -c 600 gpurun_out/r_bench_c4_gamma_poisson.json

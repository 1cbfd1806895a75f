#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/e2e_sweep.py > gpurun_out/o_e2e_sweep.txt 2>&1
timeout 300 python tools/gof_deep.py --draws 1e9 --samplers gamma_poisson --seed 20261017 > gpurun_out/o_gof_seed2.txt 2>&1
cat gpurun_out/o_e2e_sweep.txt | grep -v "^tree\|Warn"; cut -c1-60,190-330 gpurun_out/o_gof_seed2.txt

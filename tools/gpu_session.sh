#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s_pytest.txt
python tools/gof_bins.py > gpurun_out/gb1.txt 2>&1
python tools/gof_deep.py > gpurun_out/s_gof.txt 2>&1
python tools/lineage_bench.py > gpurun_out/s_lineage.txt 2>&1
python tools/sampler_bench.py --cells 200000 --samplers hybrid 2>&1 | grep -E "hybrid|rror" > gpurun_out/s_sampler.txt
cat gpurun_out/s_pytest.txt gpurun_out/s_lineage.txt gpurun_out/s_sampler.txt; head -3 gpurun_out/gb1.txt; grep "<<<" gpurun_out/gb1.txt; grep hybrid gpurun_out/s_gof.txt | cut -c1-250

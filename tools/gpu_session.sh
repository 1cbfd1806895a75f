#!/bin/bash
mkdir -p gpurun_out
for l in meta1 meta2 meta1 meta2; do echo "== $l depth 0"; PST_LIB=tools/lib_$l.so python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 3 2>&1 | grep -E "^hybrid|rror"; done > gpurun_out/t_sampler.txt 2>&1
PST_LIB=tools/lib_meta2.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t_pytest.txt
cat gpurun_out/t_sampler.txt | cut -c1-170; cat gpurun_out/t_pytest.txt

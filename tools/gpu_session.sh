#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "fused or full_size or gloo or partition" 2>&1 | tail -4 > gpurun_out/g_pytest.txt
for l in stats7 stats6; do echo "== $l"; PST_LIB=tools/lib_$l.so python tools/stats_bench.py; done > gpurun_out/s_stats.txt 2>&1
cat gpurun_out/g_pytest.txt gpurun_out/s_stats.txt

#!/bin/bash
set -x
mkdir -p gpurun_out
for l in c64 c32 c16; do echo "== $l"; PST_LIB=tools/lib_$l.so python tools/stats_bench.py; done > gpurun_out/s_stats.txt 2>&1
cat gpurun_out/s_stats.txt

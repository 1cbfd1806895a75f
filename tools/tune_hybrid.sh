#!/bin/bash
for i in 1 2; do
echo "== grouped"; python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 3 2>&1 | grep -E "^hybrid|rror"
echo "== not grouped (order param NULL)"; PST_NO_GROUP=1 python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 3 2>&1 | grep -E "^hybrid|rror"
echo "== compiled without order support"; PST_NO_GROUP=1 PST_LIB=tools/lib_noorder.so python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 3 2>&1 | grep -E "^hybrid|rror"
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv

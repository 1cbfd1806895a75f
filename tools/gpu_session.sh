#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/g_pytest.txt
python __graft_entry__.py --smoke > gpurun_out/g_smoke.txt 2>&1
for w in c1 c2 c3; do python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/g_bench_$w.json 2> gpurun_out/g_bench_$w.err; done
python bench.py > gpurun_out/g_bench_c4.json 2> gpurun_out/g_bench_c4.err
cat gpurun_out/g_pytest.txt gpurun_out/g_smoke.txt; tail -n 3 gpurun_out/g_bench_c3.err gpurun_out/g_bench_c4.err

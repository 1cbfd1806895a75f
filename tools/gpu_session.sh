#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/g_pytest.txt
bash tools/sanitize.sh > gpurun_out/g_sanitizer.txt 2>&1
python bench.py --workload c3 --steps 20 --warmup 3 > gpurun_out/g_bench_c3.json 2> gpurun_out/g_bench_c3.err
python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/g_bench_c2.json 2> gpurun_out/g_bench_c2.err
python bench.py --workload c1 --steps 20 --warmup 3 > gpurun_out/g_bench_c1.json 2> gpurun_out/g_bench_c1.err
cat gpurun_out/g_pytest.txt; grep -E "==|SUMMARY|passed|failed" gpurun_out/g_sanitizer.txt

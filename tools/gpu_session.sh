#!/bin/bash
# final round-2 evidence on one B200
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/f_pytest.txt
python __graft_entry__.py --smoke > gpurun_out/f_smoke.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:draw_counts_kernel -s 1 -c 1 -o gpurun_out/prof_r02_final -f python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 1 > gpurun_out/f_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --cells 200000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-strong > gpurun_out/f_ncu_bench.log 2>&1
python bench.py > gpurun_out/f_bench_c4.json 2> gpurun_out/f_bench_c4.err
for w in c1 c2 c3; do python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/f_bench_$w.json 2> gpurun_out/f_bench_$w.err; done
python bench.py --workload c5 --steps 3 --warmup 1 > gpurun_out/f_bench_c5.json 2> gpurun_out/f_bench_c5.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
python tools/gof_deep.py > gpurun_out/f_gof.txt 2>&1
python tools/lineage_bench.py > gpurun_out/f_lineage.txt 2>&1
python tools/stats_bench.py > gpurun_out/f_stats.txt 2>&1
for m in -2 0 1.5 3; do python tools/sampler_bench.py --cells 200000 --scale-mean $m 2>&1 | grep -E "hybrid|gamma|rror" ; done > gpurun_out/f_depth.txt 2>&1
python tools/relmeans_bench.py > gpurun_out/f_relmeans.txt 2>&1
cat gpurun_out/f_pytest.txt gpurun_out/f_smoke.txt gpurun_out/f_lineage.txt gpurun_out/f_stats.txt gpurun_out/f_depth.txt; tail -n 2 gpurun_out/f_bench_*.err

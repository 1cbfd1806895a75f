#!/bin/bash
# 8-GPU session: host bandwidth with all ranks at once, bench c4 (weak + strong + e2e variants), c5 streamed, bit-exactness
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > gpurun_out/m8_topo.txt 2>&1; lscpu | head -25 >> gpurun_out/m8_topo.txt; free -g >> gpurun_out/m8_topo.txt; cat /sys/devices/system/node/online >> gpurun_out/m8_topo.txt
$TR tools/host_bw.py > gpurun_out/m8_hostbw.txt 2>&1
$TR bench.py --gpus 8 > gpurun_out/m8_bench_c4.json 2> gpurun_out/m8_bench_c4.err
$TR bench.py --gpus 8 --workload c5 --steps 3 --warmup 1 > gpurun_out/m8_bench_c5.json 2> gpurun_out/m8_bench_c5.err
$TR tools/multi_gpu_check.py > gpurun_out/m8_check.txt 2>&1
python -m pytest tests -m gpu -x -q -k "requested_device" 2>&1 | tail -3 > gpurun_out/m8_pytest.txt
tail -3 gpurun_out/m8_bench_c4.err gpurun_out/m8_bench_c5.err; cat gpurun_out/m8_check.txt gpurun_out/m8_pytest.txt; grep -E "D2H|threads  4|threads 32" gpurun_out/m8_hostbw.txt | head -40

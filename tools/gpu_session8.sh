#!/bin/bash
# 8-GPU session: bench c4 (weak + strong + e2e variants), c5 streamed, bit-exactness, 2-GPU device test
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 > gpurun_out/m8_bench_c4.json 2> gpurun_out/m8_bench_c4.err
$TR bench.py --gpus 8 --workload c5 --steps 3 --warmup 1 > gpurun_out/m8_bench_c5.json 2> gpurun_out/m8_bench_c5.err
$TR bench.py --gpus 8 --workload c3 --steps 20 --warmup 3 > gpurun_out/m8_bench_c3.json 2> gpurun_out/m8_bench_c3.err
$TR tools/multi_gpu_check.py > gpurun_out/m8_check.txt 2>&1
python -m pytest tests -m gpu -x -q -k "requested_device" 2>&1 | tail -3 > gpurun_out/m8_pytest.txt
python examples/quickstart.py > gpurun_out/m8_quickstart.txt 2>&1
tail -n 3 gpurun_out/m8_bench_c4.err gpurun_out/m8_bench_c5.err gpurun_out/m8_bench_c3.err; cat gpurun_out/m8_check.txt gpurun_out/m8_pytest.txt gpurun_out/m8_quickstart.txt

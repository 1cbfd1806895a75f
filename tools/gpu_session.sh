#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s_pytest.txt
bash tools/ab.sh atomic long longballot > gpurun_out/s_ab.txt 2>&1
cat gpurun_out/s_pytest.txt gpurun_out/s_ab.txt

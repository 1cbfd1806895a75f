#!/bin/bash
# 8-GPU session: bench c4 (weak + strong + e2e transports), bit-exactness across ranks
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 > gpurun_out/p8_bench_c4.json 2> gpurun_out/p8_bench_c4.err
$TR tools/multi_gpu_check.py > gpurun_out/p8_check.txt 2>&1
tail -n 3 gpurun_out/p8_bench_c4.err; tail -5 gpurun_out/p8_check.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/p8_bench_c4.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value %.3e strong %.3e e2e %.3e (%s)  direct %.3e  via_u8 %.3e  default_api %.3e (first call %.3e)  u16 %.3e u8 %.3e" % (
    d["value"], d.get("strong", {}).get("value", 0), e["value"], e.get("transport"), e["int32_direct"]["value"], e["int32_via_u8"]["value"],
    e["default_api"]["value"], e["default_api"]["first_call"]["value"], e["narrow_u16"]["value"], e["narrow_u8"]["value"]))
PY

#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/host_bw.py > gpurun_out/i_hostbw.txt 2>&1
python bench.py > gpurun_out/i_bench_c4.json 2> gpurun_out/i_bench.err
tail -n 3 gpurun_out/i_bench.err
cat gpurun_out/i_hostbw.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/i_bench_c4.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value %.3e  e2e %.3e (%s)  direct %.3e  via_u8 %.3e  default_api %.3e  u16 %.3e u8 %.3e" % (
    d["value"], e["value"], e.get("transport"), e["int32_direct"]["value"], e["int32_via_u8"]["value"],
    e["default_api"]["value"], e["narrow_u16"]["value"], e["narrow_u8"]["value"]))
PY

#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/q_pytest.txt
python __graft_entry__.py --smoke > gpurun_out/q_smoke.txt 2>&1
python bench.py > gpurun_out/q_bench_c4.json 2> gpurun_out/q_bench.err
for l in one two; do for m in 0 3; do echo "== $l depth $m"; PST_LIB=tools/lib_$l.so python tools/sampler_bench.py --cells 100000 --samplers gamma_poisson --reps 3 --scale-mean $m 2>&1 | grep -E "^hybrid|^gamma|rror"; done; done > gpurun_out/q_sampler.txt 2>&1
cat gpurun_out/q_pytest.txt; tail -2 gpurun_out/q_smoke.txt; tail -n 3 gpurun_out/q_bench.err; grep -E "^==|^gamma" gpurun_out/q_sampler.txt | cut -c1-160
python - <<'PY'
import json
d = json.loads(open("gpurun_out/q_bench_c4.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value %.3e  e2e %.3e (%s)  direct %.3e  via_u8 %.3e  default_api %.3e (first call %.3e)  u16 %.3e u8 %.3e" % (
    d["value"], e["value"], e.get("transport"), e["int32_direct"]["value"], e["int32_via_u8"]["value"],
    e["default_api"]["value"], e["default_api"]["first_call"]["value"], e["narrow_u16"]["value"], e["narrow_u8"]["value"]))
PY

#!/bin/bash
# round-2 run Z: loose/full BA shim tests + the full default bench (all legs, CPU baselines)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_shim.py -m gpu -q -x > gpurun_out/z_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/z_pytest.log
tail -12 gpurun_out/z_pytest.log
timeout 1500 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc $?"; tail -3 gpurun_out/z_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/z_bench.json').read().strip().splitlines()[-1])
print("C2", d["value"], d["e2e"]["value"], d["ms_per_step"], "C4", d["c4"]["value"], d["c4"]["e2e"]["value"], "C5", d["c5"]["value"])
lb=d["localba"]; print("C3", lb["value"], lb["ms_per_solve"], {k:v for k,v in lb["batched"].items() if k!="roofline"})
print("cpu", d["cpu_baseline"]["value"], d["c4"]["cpu_baseline"]["value"], lb["cpu_baseline"]["value"], d["c5"]["cpu_baseline"]["value"])
PY

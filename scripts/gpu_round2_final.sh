#!/bin/bash
# round-2 final check: what the driver runs at round end - the GPU suite, smoke(), the default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/final_pytest.log
tail -5 gpurun_out/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 1500 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc $?"; tail -2 gpurun_out/final_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print("C2", d["value"], d["e2e"]["value"], d["ms_per_step"], "C4", d["c4"]["value"], d["c4"]["e2e"]["value"], "C5", d["c5"]["value"])
lb=d["localba"]; print("C3", lb["value"], lb["ms_per_solve"], {k:v for k,v in lb["batched"].items() if k!="roofline"})
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel"], "clocks", d["clocks"])
PY

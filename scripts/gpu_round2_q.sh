#!/bin/bash
mkdir -p gpurun_out
./scripts/latency_probe > gpurun_out/q_latency.txt 2>&1; cat gpurun_out/q_latency.txt
timeout 900 python -m pytest tests/test_frontend_gpu.py tests/test_ba_gpu.py -m gpu -q -x > gpurun_out/q_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/q_pytest.log
tail -4 gpurun_out/q_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 --no-ba > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc $?"; tail -3 gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
print("C2", d["value"], d["e2e"]["value"], "C4", d["c4"]["value"], d["c4"]["e2e"]["value"])
print({k: v["ms_per_step"] for k, v in d["c4"]["roofline"]["per_kernel"].items()})
PY

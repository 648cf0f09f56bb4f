#!/bin/bash
# round-2 run R: KLT iteration diet (cached search taps, 32-bit redux path, float convergence tests) + batched-BA time split
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_frontend_gpu.py -m gpu -q -x > gpurun_out/r_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r_pytest.log
tail -4 gpurun_out/r_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 --no-ba > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r_bench.json').read().strip().splitlines()[-1])
print("C2", d["value"], d["e2e"]["value"], d["roofline"]["kernel_time_shares"], d["ms_per_step"], "C4", d["c4"]["value"], d["c4"]["e2e"]["value"])
print({k: v["ms_per_step"] for k, v in d["c4"]["roofline"]["per_kernel"].items()})
PY
timeout 600 python scripts/ba_batch_probe.py > gpurun_out/r_ba_batch_probe.log 2>&1; cat gpurun_out/r_ba_batch_probe.log

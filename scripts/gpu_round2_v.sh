#!/bin/bash
# round-2 run V: KLT interior template (interpolate, then differentiate), owner-mode Schur with whole-chunk staging, SAD fix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_frontend_gpu.py -m gpu -q -x > gpurun_out/v_pytest_fe.log 2>&1; echo "pytest rc $?" >> gpurun_out/v_pytest_fe.log
tail -6 gpurun_out/v_pytest_fe.log
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_ba_sharded.py -m gpu -q -x > gpurun_out/v_pytest_ba.log 2>&1; echo "pytest rc $?" >> gpurun_out/v_pytest_ba.log
tail -4 gpurun_out/v_pytest_ba.log
timeout 600 python scripts/ba_batch_probe.py 128 296 > gpurun_out/v_ba_batch_probe.log 2>&1; cat gpurun_out/v_ba_batch_probe.log
OV2_BA_TRACE=1 timeout 300 python scripts/ba_batch_probe.py 128 2>&1 | grep "ba trace" | tail -1 | cut -c1-600
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 --no-ba > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc $?"; tail -3 gpurun_out/v_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[-1])
print("C2", d["value"], d["e2e"]["value"], d["roofline"]["kernel_time_shares"], d["ms_per_step"], "C4", d["c4"]["value"], d["c4"]["e2e"]["value"])
print({k: v["ms_per_step"] for k, v in d["c4"]["roofline"]["per_kernel"].items()})
PY

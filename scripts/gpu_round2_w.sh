#!/bin/bash
# round-2 run W: owner-mode Schur with unrolled chunk staging + DMMA landmark syrk; bench legs of local BA
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_ba_sharded.py tests/test_host_shim.py -m gpu -q -x > gpurun_out/w_pytest_ba.log 2>&1; echo "pytest rc $?" >> gpurun_out/w_pytest_ba.log
tail -6 gpurun_out/w_pytest_ba.log
timeout 600 python scripts/ba_batch_probe.py 128 296 > gpurun_out/w_ba_batch_probe.log 2>&1; cat gpurun_out/w_ba_batch_probe.log
OV2_BA_TRACE=1 timeout 300 python scripts/ba_batch_probe.py 128 2>&1 | grep "ba trace" | tail -1 | cut -c1-600
OV2_BA_SCHUR_SMEM=3 OV2_BA_TRACE=1 timeout 300 python scripts/ba_one.py 2>&1 | grep "ba trace" | tail -1 | cut -c1-600
timeout 900 python bench.py --steps 10 --warmup 3 --no-c5 > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; echo "bench rc $?"; tail -3 gpurun_out/w_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w_bench.json').read().strip().splitlines()[-1])
print("C2", d["value"], d["e2e"]["value"], "C4", d["c4"]["value"], d["c4"]["e2e"]["value"])
lb=d["localba"]; print("C3", lb["value"], lb["ms_per_solve"], {k:v for k,v in lb["batched"].items() if k!="roofline"})
PY

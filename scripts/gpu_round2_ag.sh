#!/bin/bash
# round-2 run AG: CLAHE apply without conversion-unit instructions: parity + C4 timing + ncu of the C4 streaming kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_frontend_gpu.py -m gpu -q -x > gpurun_out/ag_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/ag_pytest.log
tail -4 gpurun_out/ag_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 --no-ba > gpurun_out/ag_bench.json 2> gpurun_out/ag_bench.err; echo "bench rc $?"; tail -2 gpurun_out/ag_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ag_bench.json').read().strip().splitlines()[-1])
print("C2", d["value"], d["e2e"]["value"], d["ms_per_step"], "C4", d["c4"]["value"], d["c4"]["e2e"]["value"], d["c4"]["ms_per_step"])
print({k: v["ms_per_step"] for k, v in d["c4"]["roofline"]["per_kernel"].items()})
PY
FULL="ncu --set full --clock-control none --import-source on"
timeout 600 $FULL -k "regex:clahe_apply|fast_sweep|fast_cells" -s 3 -c 3 -o gpurun_out/ag_misc python bench.py --kernels-only --only c4 --c4-batch 32 --steps 1 --warmup 1 > gpurun_out/ag_ncu.log 2>&1; tail -2 gpurun_out/ag_ncu.log

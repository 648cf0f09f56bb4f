#!/bin/bash
# round-2 run AA: wavefront FAST sweep - parity + C2 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_frontend_gpu.py tests/test_host_shim.py -m gpu -q -x > gpurun_out/aa_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/aa_pytest.log
tail -6 gpurun_out/aa_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 --no-ba --no-c4 > gpurun_out/aa_bench.json 2> gpurun_out/aa_bench.err; echo "bench rc $?"; tail -3 gpurun_out/aa_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/aa_bench.json').read().strip().splitlines()[-1])
print("C2", d["value"], d["e2e"]["value"], d["roofline"]["kernel_time_shares"], d["ms_per_step"])
PY

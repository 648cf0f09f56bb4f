#!/bin/bash
# round-2 run AI: three-keypoints-per-warp KLT (opt-in): parity + timing against the default kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_frontend_gpu.py -m gpu -q -x -k "klt or c4_resolution" > gpurun_out/ai_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/ai_pytest.log
tail -12 gpurun_out/ai_pytest.log
for m in 1 3; do
OV2_KLT_MODE=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 --no-ba --no-c4 > gpurun_out/ai_bench_$m.json 2> gpurun_out/ai_bench_$m.err; echo "mode $m bench rc $?"; tail -2 gpurun_out/ai_bench_$m.err
python - <<PY
import json
d=json.loads(open('gpurun_out/ai_bench_$m.json').read().strip().splitlines()[-1])
print("mode $m C2", d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_time_shares"], d["roofline"]["avg_launch_ms"])
PY
done

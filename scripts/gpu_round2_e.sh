#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_ba_sharded.py -m gpu -q -x > gpurun_out/e_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/e_pytest.log
tail -6 gpurun_out/e_pytest.log
timeout 300 python scripts/ba_trace.py > gpurun_out/e_ba_trace.log 2>&1; grep "ba trace" gpurun_out/e_ba_trace.log | awk 'NR%3==0'
timeout 400 python scripts/ba_sweep.py > gpurun_out/e_ba_sweep.log 2>&1; grep -v '"ncopy": "1"' gpurun_out/e_ba_sweep.log | tail -40
timeout 400 python -m pytest tests/test_frontend_gpu.py -m gpu -q > gpurun_out/e_pytest_fe.log 2>&1; tail -3 gpurun_out/e_pytest_fe.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 --no-ba > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc $?"

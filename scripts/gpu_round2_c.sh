#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_sharded.py "tests/test_frontend_gpu.py::test_c4_full_chain_stereo_1280x720" -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c_pytest.log
tail -5 gpurun_out/c_pytest.log
timeout 300 python scripts/ba_trace.py > gpurun_out/c_ba_trace.log 2>&1; tail -20 gpurun_out/c_ba_trace.log
timeout 300 python scripts/parity_stats.py > gpurun_out/c_parity_stats.json 2> gpurun_out/c_parity_stats.err; cat gpurun_out/c_parity_stats.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 --no-ba > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc $?"; tail -c 600 gpurun_out/c_bench.json

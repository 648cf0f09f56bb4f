#!/bin/bash
# GPU run B (round 2): persistent BA kernel - parity tests, C3/C5 numbers at several group sizes, 2-rank p2p in one process
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_ba_sharded.py -m gpu -x -q > gpurun_out/b_pytest_ba.log 2>&1; echo "pytest rc $?" >> gpurun_out/b_pytest_ba.log
tail -5 gpurun_out/b_pytest_ba.log
timeout 600 python -m pytest tests/test_frontend_gpu.py tests/test_pnp_gpu.py tests/test_host_shim.py -m gpu -q > gpurun_out/b_pytest_fe.log 2>&1; echo "pytest rc $?" >> gpurun_out/b_pytest_fe.log
tail -3 gpurun_out/b_pytest_fe.log
timeout 300 python scripts/ba_sweep.py > gpurun_out/b_ba_sweep.log 2>&1
tail -30 gpurun_out/b_ba_sweep.log

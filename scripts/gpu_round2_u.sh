#!/bin/bash
# round-2 run U: owner-mode Schur + sub-group back-substitution (BA), batched getLineMinSAD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_ba_sharded.py tests/test_host_shim.py -m gpu -q -x > gpurun_out/u_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/u_pytest.log
tail -15 gpurun_out/u_pytest.log
timeout 300 python -m pytest tests/test_frontend_gpu.py -m gpu -q -x -k "line_min_sad" > gpurun_out/u_pytest_sad.log 2>&1; echo "pytest rc $?" >> gpurun_out/u_pytest_sad.log
tail -8 gpurun_out/u_pytest_sad.log
timeout 600 python scripts/ba_batch_probe.py 128 296 > gpurun_out/u_ba_batch_probe.log 2>&1; cat gpurun_out/u_ba_batch_probe.log
OV2_BA_TRACE=1 timeout 300 python scripts/ba_batch_probe.py 128 2>&1 | grep "ba trace" | tail -1 | cut -c1-600
timeout 300 python scripts/ba_trace.py > gpurun_out/u_ba_trace.log 2>&1; grep "ba trace\|^C" gpurun_out/u_ba_trace.log | awk '/^C/{name=$0} /ba trace/{c[name]++; if(c[name]==3) print name" :: "$0}' | cut -c1-520

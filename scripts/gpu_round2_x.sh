#!/bin/bash
# round-2 run X: full GPU suite + thread-overlap probe of the batched BA path
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/x_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/x_pytest.log
tail -6 gpurun_out/x_pytest.log
timeout 300 python scripts/ba_threads_probe.py 296 > gpurun_out/x_threads.log 2>&1; cat gpurun_out/x_threads.log
timeout 300 python scripts/ba_threads_probe.py 148 > gpurun_out/x_threads148.log 2>&1; head -3 gpurun_out/x_threads148.log

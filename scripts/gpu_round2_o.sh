#!/bin/bash
# round-2 run O: look-ahead Gauss-Jordan pivots, shuffle-broadcast Cholesky block
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_ba_sharded.py tests/test_host_shim.py tests/test_pnp_gpu.py -m gpu -q -x > gpurun_out/o_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/o_pytest.log
tail -4 gpurun_out/o_pytest.log
timeout 300 python scripts/ba_trace.py > gpurun_out/o_ba_trace.log 2>&1; grep "ba trace\|^C" gpurun_out/o_ba_trace.log | awk '/^C/{name=$0} /ba trace/{c[name]++; if(c[name]==3) print name" :: "$0}' | grep "C5\|ctas 148" | cut -c1-420
timeout 400 python scripts/ba_sweep.py > gpurun_out/o_ba_sweep.log 2>&1; grep -v '"ncopy": "1"' gpurun_out/o_ba_sweep.log | grep '"ctas": null\|legacy' | cut -c1-120

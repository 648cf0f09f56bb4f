#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_sharded.py -m gpu -q > gpurun_out/f_pytest_p2p.log 2>&1; echo "pytest rc $?" >> gpurun_out/f_pytest_p2p.log
tail -4 gpurun_out/f_pytest_p2p.log
timeout 300 python scripts/ba_trace.py > gpurun_out/f_ba_trace.log 2>&1; grep "ba trace\|^C" gpurun_out/f_ba_trace.log | awk '/^C/{name=$0} /ba trace/{c[name]++; if(c[name]==3) print name" :: "$0}' | grep "coop 0"
timeout 400 python scripts/ba_sweep.py > gpurun_out/f_ba_sweep.log 2>&1; grep -v '"ncopy": "1"' gpurun_out/f_ba_sweep.log | tail -40 | cut -c1-120
timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_frontend_gpu.py tests/test_host_shim.py -m gpu -q > gpurun_out/f_pytest.log 2>&1; tail -4 gpurun_out/f_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 --no-ba > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc $?"

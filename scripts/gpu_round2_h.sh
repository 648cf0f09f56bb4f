#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/h_pytest.log
tail -4 gpurun_out/h_pytest.log
timeout 300 python scripts/ba_trace.py > gpurun_out/h_ba_trace.log 2>&1; grep "ba trace\|^C" gpurun_out/h_ba_trace.log | awk '/^C/{name=$0} /ba trace/{c[name]++; if(c[name]==3) print name" :: "$0}' | grep "smem 0\|C5" | cut -c1-420
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc $?"; tail -3 gpurun_out/h_bench.err

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_ba_sharded.py "tests/test_frontend_gpu.py::test_c4_full_chain_stereo_1280x720" "tests/test_frontend_gpu.py::test_single_scale_detector_1280x720_cell35" "tests/test_frontend_gpu.py::test_single_scale_detector_bit_exact" -m gpu -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/d_pytest.log
tail -6 gpurun_out/d_pytest.log
timeout 300 python scripts/ba_trace.py > gpurun_out/d_ba_trace.log 2>&1; grep "ba trace" gpurun_out/d_ba_trace.log | awk 'NR%3==0'
timeout 400 python scripts/ba_sweep.py > gpurun_out/d_ba_sweep.log 2>&1; grep -v '"ncopy": "1"' gpurun_out/d_ba_sweep.log | tail -40
timeout 300 python scripts/parity_stats.py > gpurun_out/d_parity_stats.json 2> gpurun_out/d_parity_stats.err; cat gpurun_out/d_parity_stats.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pyr_levels -s 2 -c 1 -o gpurun_out/d_pyr python bench.py --kernels-only --batch 64 --steps 1 --warmup 1 > gpurun_out/d_ncu_pyr.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 --no-ba > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc $?"

#!/bin/bash
# round-2 run AD: source-level ncu capture of fast_cells_kernel (+ BA optional-mode tests)
mkdir -p gpurun_out
FULL="ncu --set full --clock-control none --import-source on"
timeout 600 $FULL -k regex:fast_cells -s 1 -c 1 -o gpurun_out/ad_fastcells python bench.py --kernels-only --batch 64 --steps 1 --warmup 1 > gpurun_out/ad_ncu.log 2>&1; tail -2 gpurun_out/ad_ncu.log
timeout 900 python -m pytest tests/test_ba_gpu.py -m gpu -q -x -k optional > gpurun_out/ad_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/ad_pytest.log
tail -4 gpurun_out/ad_pytest.log

#!/bin/bash
# round-2 run S: BRIEF-32 mode parity + source-level ncu capture of fb_klt_kernel (where do the instructions go)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frontend_gpu.py -m gpu -q -x -k "describe" > gpurun_out/s_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/s_pytest.log
tail -6 gpurun_out/s_pytest.log
FULL="ncu --set full --clock-control none --import-source on"
timeout 600 $FULL -k regex:klt -s 4 -c 1 -o gpurun_out/s_klt python bench.py --kernels-only --batch 64 --steps 1 --warmup 1 > gpurun_out/s_ncu_klt.log 2>&1
tail -3 gpurun_out/s_ncu_klt.log
ls -la gpurun_out/s_*

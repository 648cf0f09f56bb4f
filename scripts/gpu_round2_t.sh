#!/bin/bash
# round-2 run T: batched-BA host path (one memset / one D2H / threaded packing) + source-level ncu capture of fb_klt_kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_ba_sharded.py tests/test_host_shim.py -m gpu -q -x > gpurun_out/t_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/t_pytest.log
tail -4 gpurun_out/t_pytest.log
timeout 600 python scripts/ba_batch_probe.py 128 296 > gpurun_out/t_ba_batch_probe.log 2>&1; cat gpurun_out/t_ba_batch_probe.log
OV2_BA_TRACE=1 timeout 300 python scripts/ba_batch_probe.py 128 2>&1 | grep "ba trace" | tail -2 | cut -c1-600
FULL="ncu --set full --clock-control none --import-source on"
timeout 600 $FULL -k regex:klt -s 1 -c 1 -o gpurun_out/t_klt python bench.py --kernels-only --batch 64 --steps 1 --warmup 1 > gpurun_out/t_ncu_klt.log 2>&1
tail -3 gpurun_out/t_ncu_klt.log
ls -la gpurun_out/t_*

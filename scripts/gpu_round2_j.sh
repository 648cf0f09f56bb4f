#!/bin/bash
# round-2 run J: full GPU suite, full bench, launch lists of the three legs, full ncu captures of the top kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j_pytest.log
tail -5 gpurun_out/j_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc $?"; tail -3 gpurun_out/j_bench.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU -c 600 --log-file gpurun_out/j_launches_c2.csv python bench.py --kernels-only --steps 2 --warmup 1 > gpurun_out/j_ncu_c2.log 2>&1
timeout 300 $NCU -c 900 --log-file gpurun_out/j_launches_c4.csv python bench.py --kernels-only --only c4 --steps 2 --warmup 1 > gpurun_out/j_ncu_c4.log 2>&1
timeout 300 $NCU -c 200 --log-file gpurun_out/j_launches_c5.csv python bench.py --kernels-only --only c5 --steps 2 --warmup 1 > gpurun_out/j_ncu_c5.log 2>&1
FULL="ncu --set full --clock-control none --import-source on"
timeout 300 $FULL -k regex:pyr_levels -s 2 -c 1 -o gpurun_out/j_pyr python bench.py --kernels-only --batch 64 --steps 1 --warmup 1 > gpurun_out/j_ncu_pyr.log 2>&1
timeout 400 $FULL -k regex:klt -s 4 -c 2 -o gpurun_out/j_klt python bench.py --kernels-only --batch 64 --steps 1 --warmup 1 > gpurun_out/j_ncu_klt.log 2>&1
timeout 400 $FULL -k regex:ba_lm_kernel -s 1 -c 1 -o gpurun_out/j_balm_c5 python bench.py --kernels-only --only c5 --steps 1 --warmup 1 > gpurun_out/j_ncu_balm_c5.log 2>&1
timeout 400 $FULL -k regex:ba_lm_kernel -s 1 -c 1 -o gpurun_out/j_balm_c3 python bench.py --kernels-only --only c5 --c5-small --steps 1 --warmup 1 > gpurun_out/j_ncu_balm_c3.log 2>&1
timeout 400 $FULL -k "regex:ss_sweep|ss_response|clahe_apply|clahe_lut" -s 8 -c 4 -o gpurun_out/j_c4 python bench.py --kernels-only --only c4 --c4-batch 32 --steps 1 --warmup 1 > gpurun_out/j_ncu_c4full.log 2>&1
ls -la gpurun_out/j_*

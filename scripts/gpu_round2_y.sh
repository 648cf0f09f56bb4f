#!/bin/bash
# round-2 run Y: ncu evidence for the round-2b kernels (owner-mode BA batch, KLT, BRIEF-32, row search) + launch lists
mkdir -p gpurun_out
FULL="ncu --set full --clock-control none --import-source on"
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 600 $FULL -k regex:ba_lm_kernel -s 2 -c 1 -o gpurun_out/y_balm_batch python scripts/ba_batch_probe.py 296 > gpurun_out/y_ncu_balm.log 2>&1; tail -2 gpurun_out/y_ncu_balm.log
timeout 600 $FULL -k regex:klt -s 1 -c 1 -o gpurun_out/y_klt python bench.py --kernels-only --batch 64 --steps 1 --warmup 1 > gpurun_out/y_ncu_klt.log 2>&1; tail -2 gpurun_out/y_ncu_klt.log
timeout 600 $FULL -k "regex:describe_box|line_min_sad" -s 2 -c 2 -o gpurun_out/y_newops python scripts/profile_new_ops.py > gpurun_out/y_ncu_newops.log 2>&1; tail -2 gpurun_out/y_ncu_newops.log
timeout 300 python scripts/profile_new_ops.py > gpurun_out/y_newops_times.log 2>&1; tail -1 gpurun_out/y_newops_times.log
timeout 300 $NCU -c 600 --log-file gpurun_out/y_launches_c2.csv python bench.py --kernels-only --steps 2 --warmup 1 > gpurun_out/y_l_c2.log 2>&1
timeout 300 $NCU -c 900 --log-file gpurun_out/y_launches_c4.csv python bench.py --kernels-only --only c4 --steps 2 --warmup 1 > gpurun_out/y_l_c4.log 2>&1
ls -la gpurun_out/y_*

#!/bin/bash
# GPU run A (round 2): parity suite, full bench line (C2 + C4 + C5 + C3), launch lists of the C4 and C5 legs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc $?" >> gpurun_out/a_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches_c4.csv \
    python bench.py --kernels-only --only c4 --c4-batch 32 --c4-unique 4 --steps 1 --warmup 1 > gpurun_out/a_ncu_c4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches_c5.csv \
    python bench.py --kernels-only --only c5 --steps 1 --warmup 1 > gpurun_out/a_ncu_c5.log 2>&1
tail -3 gpurun_out/a_pytest.log; tail -c 1500 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err

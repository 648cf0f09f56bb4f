#!/bin/bash
# round-2 run K: BA changes (register Cholesky block, batched RMW) + restructured KLT parity tests + the captures run J missed
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/k_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/k_pytest.log
tail -5 gpurun_out/k_pytest.log
timeout 300 python scripts/ba_trace.py > gpurun_out/k_ba_trace.log 2>&1; grep "ba trace\|^C" gpurun_out/k_ba_trace.log | awk '/^C/{name=$0} /ba trace/{c[name]++; if(c[name]==3) print name" :: "$0}' | cut -c1-420
timeout 400 python scripts/ba_sweep.py > gpurun_out/k_ba_sweep.log 2>&1; grep -v '"ncopy": "1"' gpurun_out/k_ba_sweep.log | tail -24 | cut -c1-110
FULL="ncu --set full --clock-control none --import-source on"
timeout 400 $FULL -k regex:fb_klt -s 1 -c 1 -o gpurun_out/k_klt python bench.py --kernels-only --batch 64 --steps 1 --warmup 1 > gpurun_out/k_ncu_klt.log 2>&1
timeout 400 $FULL -k regex:fast_cells -s 1 -c 1 -o gpurun_out/k_fastcells python bench.py --kernels-only --batch 64 --steps 1 --warmup 1 > gpurun_out/k_ncu_fastcells.log 2>&1
timeout 400 $FULL -k regex:ss_response -s 1 -c 1 -o gpurun_out/k_ssresp python bench.py --kernels-only --only c4 --c4-batch 32 --steps 1 --warmup 1 > gpurun_out/k_ncu_ssresp.log 2>&1
timeout 400 $FULL -k regex:ss_sweep -s 1 -c 1 -o gpurun_out/k_sssweep python bench.py --kernels-only --only c4 --c4-batch 32 --steps 1 --warmup 1 > gpurun_out/k_ncu_sssweep.log 2>&1
timeout 400 $FULL -k regex:subpix -s 1 -c 1 -o gpurun_out/k_subpix python bench.py --kernels-only --only c4 --c4-batch 32 --steps 1 --warmup 1 > gpurun_out/k_ncu_subpix.log 2>&1
ls -la gpurun_out/k_*

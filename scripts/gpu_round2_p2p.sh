#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/ba_p2p_trace.py > gpurun_out/p2p_trace_$N.log 2>&1; echo rc $?
grep "ba trace\|world" gpurun_out/p2p_trace_$N.log | tail -8 | cut -c1-520

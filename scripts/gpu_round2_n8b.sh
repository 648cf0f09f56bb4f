#!/bin/bash
# 8-GPU box, round 2b: torchrun bench at N=8, 4, 2 (C2/C4 weak scaling, C5 sharded over peer memory)
mkdir -p gpurun_out
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/nb${N}_bench.json 2> gpurun_out/nb${N}_bench.err; echo "bench N=$N rc $?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/nb${N}_bench.err | tail -3
python - <<PY
import json
d=json.loads(open('gpurun_out/nb${N}_bench.json').read().strip().splitlines()[-1])
print("N=${N} C2", d["value"], d["e2e"]["value"], "C4", d["c4"]["value"], d["c4"]["e2e"]["value"])
c5=d["c5"]; print("C5", c5["value"], c5["ms_per_solve"], c5.get("iters"), c5.get("same_iterations_as_cpu"), c5.get("max_pose_diff_vs_cpu"))
PY
done

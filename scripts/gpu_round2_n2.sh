#!/bin/bash
# 2-GPU run: torchrun bench (C2/C4 weak scaling with NUMA pinning, C5 sharded over peer memory)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n2_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench rc $?"; tail -5 gpurun_out/n2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n2_bench.json').read().strip().splitlines()[-1])
print("C2", d["value"], d["e2e"]["value"], d["e2e"].get("numa"), "C4", d["c4"]["value"], d["c4"]["e2e"]["value"])
c5=d["c5"]; print("C5", c5["value"], c5["ms_per_solve"], c5.get("iters"), c5.get("same_iterations_as_cpu"), c5.get("max_pose_diff_vs_cpu"), c5.get("collective"))
PY

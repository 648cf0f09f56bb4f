#!/bin/bash
# round-2 run AF: e2e chunk-count sweep (C2)
mkdir -p gpurun_out
for c in 4 6 12 16; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-c5 --no-ba --no-c4 --e2e-chunks $c > gpurun_out/af_bench_$c.json 2> gpurun_out/af_bench_$c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/af_bench_$c.json').read().strip().splitlines()[-1])
print("chunks", $c, "C2 e2e", d["e2e"]["value"], d["e2e"]["pipelined_repeats"], "stepped", d["e2e"]["stepped_value"])
PY
done
python scripts/pcie_bw.py 2>&1 | tail -4

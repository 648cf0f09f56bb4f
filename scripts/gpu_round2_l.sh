#!/bin/bash
# round-2 run L: sliding-window single-scale response kernel + explicit global fp64 reductions in the BA kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/l_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/l_pytest.log
tail -5 gpurun_out/l_pytest.log
timeout 300 python scripts/ba_trace.py > gpurun_out/l_ba_trace.log 2>&1; grep "ba trace\|^C" gpurun_out/l_ba_trace.log | awk '/^C/{name=$0} /ba trace/{c[name]++; if(c[name]==3) print name" :: "$0}' | grep "C5\|ctas 148" | cut -c1-420
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc $?"; tail -3 gpurun_out/l_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/l_bench.json').read().strip().splitlines()[-1])
print("C2", d["value"], d["e2e"]["value"], "C4", d["c4"]["value"], d["c4"]["e2e"]["value"], "C5", d["c5"]["value"], "C3", d["localba"]["value"], d["localba"]["batched"]["value"])
print({k: v["ms_per_step"] for k, v in d["c4"]["roofline"]["per_kernel"].items()})
PY

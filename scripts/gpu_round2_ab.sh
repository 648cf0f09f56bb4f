#!/bin/bash
# round-2 run AB: full GPU suite (legacy sharded path with an empty shard, sticky capacity flag), smoke(), default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/ab_pytest.log
tail -8 gpurun_out/ab_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/ab_smoke.log 2>&1; tail -2 gpurun_out/ab_smoke.log
timeout 1500 python bench.py > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err; echo "bench rc $?"; tail -3 gpurun_out/ab_bench.err
timeout 900 python bench.py --impl reference > gpurun_out/ab_bench_ref.json 2> gpurun_out/ab_bench_ref.err; echo "ref rc $?"; tail -c 600 gpurun_out/ab_bench_ref.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ab_bench.json').read().strip().splitlines()[-1])
print("C2", d["value"], d["e2e"]["value"], d["ms_per_step"], "C4", d["c4"]["value"], d["c4"]["e2e"]["value"], "C5", d["c5"]["value"])
lb=d["localba"]; print("C3", lb["value"], lb["ms_per_solve"], {k:v for k,v in lb["batched"].items() if k!="roofline"})
PY

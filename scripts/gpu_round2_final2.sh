#!/bin/bash
# round-2 last check: the whole GPU suite + smoke() with the match operator in; timing of ov2_match_to_map
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/final2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/final2_pytest.log
tail -4 gpurun_out/final2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final2_smoke.log 2>&1; tail -1 gpurun_out/final2_smoke.log
timeout 300 python - <<'PY' > gpurun_out/final2_match_time.log 2>&1
import time, numpy as np, sys
sys.path.insert(0, '.')
from ov2slam_b200 import api, synth
ctx = api.Context(0)
sc = synth.make_match_scene(4, 2000, 1600)
for _ in range(3): api.match_to_map(ctx, sc)
ctx.profile(True); r = api.match_to_map(ctx, sc); rep = ctx.profile_report(); ctx.profile(False)
t0 = time.perf_counter()
for _ in range(20): api.match_to_map(ctx, sc)
dt = (time.perf_counter() - t0) / 20
print({k: round(v[0] * 1e3, 1) for k, v in rep.items()}, "us kernel;", round(dt * 1e6, 1), "us per call (host arrays in/out, python binding); matched", int((r[0] >= 0).sum()), "of 1600 candidates")
PY
cat gpurun_out/final2_match_time.log

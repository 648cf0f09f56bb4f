#!/bin/bash
# round-2 run AC: two-pivot Gauss-Jordan (n <= 48): parity + C3 timing with and without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_ba_sharded.py tests/test_host_shim.py -m gpu -q -x > gpurun_out/ac_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/ac_pytest.log
tail -6 gpurun_out/ac_pytest.log
for g in 1 0; do
echo "OV2_BA_GJ2=$g"
OV2_BA_GJ2=$g OV2_BA_TRACE=1 timeout 300 python scripts/ba_one.py 2>&1 | grep "ba trace" | tail -1 | cut -c1-420
OV2_BA_GJ2=$g timeout 300 python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
from ov2slam_b200 import api, synth
ctx = api.Context(0); opt = api.Optimizer(ctx)
pb0 = synth.make_ba_problem(3, 10, 2000, 8000)
clone = lambda d=pb0: {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in d.items()}
for _ in range(5): opt.local_ba(clone())
pbs = [clone() for _ in range(100)]
ctx.sync(); t0 = time.perf_counter()
for pb in pbs: r, _ = opt.local_ba(pb)
ctx.sync(); dt = time.perf_counter() - t0
print("C3 single: %.1f solves/s, %.1f us/solve, iters %d+%d, final cost %.9g" % (100 / dt, 1e4 * dt, r["iters_robust"], r["iters_refine"], r["final_cost"]))
PY
done

"""Where a batched C3 launch spends its time: kernel time (CUDA events around the launch, ov2_profile) against the wall
time of ov2_localba_solve_batch (host packing + H2D + kernel + D2H + unpacking), per batch size and CTAs per window."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ov2slam_b200 import api, synth  # noqa: E402

ctx = api.Context(0)
clone = lambda d: {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in d.items()}
base = [synth.make_ba_problem(100 + i, 10, 2000, 8000) for i in range(8)]
KS = [int(a) for a in sys.argv[1:]] or [64, 128, 148, 296, 592]
for K in KS:
    for ctas in ((None,) if len(sys.argv) > 1 else (None, "1", "2", "4")):
        if ctas is None:
            os.environ.pop("OV2_BA_CTAS", None)
        else:
            os.environ["OV2_BA_CTAS"] = ctas
        try:
            mk = lambda: [clone(base[i % 8]) for i in range(K)]
            api.local_ba_batch(ctx, mk())
            sets = [mk() for _ in range(3)]
            ctx.sync()
            t0 = time.perf_counter()
            for st in sets:
                api.local_ba_batch(ctx, st)
            ctx.sync()
            dt = (time.perf_counter() - t0) / 3
            ctx.profile(True)
            api.local_ba_batch(ctx, mk())
            rep = ctx.profile_report()
            ctx.profile(False)
            kms = sum(v[0] for v in rep.values())
            print(json.dumps({"batch": K, "ctas": ctas, "solves_per_s": round(K / dt, 1), "ms_per_launch": round(1e3 * dt, 3),
                              "kernel_ms": round(kms, 3), "kernel_only_solves_per_s": round(K / (kms * 1e-3), 1)}), flush=True)
        except Exception as e:
            print(json.dumps({"batch": K, "ctas": ctas, "error": str(e)[:200]}), flush=True)

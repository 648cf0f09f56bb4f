"""Do two host threads (two contexts) overlap their batched local-BA calls?  Prints, per call, entry / exit times of the ABI
call and the wall time of 1-thread and 2-thread runs."""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ov2slam_b200 import api, synth  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 296
clone = lambda d: {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in d.items()}
base = [synth.make_ba_problem(100 + i, 10, 2000, 8000) for i in range(8)]
mk = lambda: [clone(base[i % 8]) for i in range(K)]
ctxs = [api.Context(0), api.Context(0)]
for c in ctxs:
    for _ in range(2):
        api.local_ba_batch(c, mk())
reps = 4
sets = [[mk() for _ in range(reps)] for _ in range(2)]
log = []
orig = api.load().ov2_localba_solve_batch


def work(ti):
    for st in sets[ti]:
        t0 = time.perf_counter()
        api.local_ba_batch(ctxs[ti], st)
        log.append((ti, t0, time.perf_counter()))


for c in ctxs:
    c.sync()
t0 = time.perf_counter()
work(0)
t1 = time.perf_counter()
print("1 thread: %.2f ms per launch, %.0f solves/s" % (1e3 * (t1 - t0) / reps, K * reps / (t1 - t0)))
log.clear()
sets = [[mk() for _ in range(reps)] for _ in range(2)]
th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
t0 = time.perf_counter()
for t in th:
    t.start()
for t in th:
    t.join()
t1 = time.perf_counter()
print("2 threads: %.2f ms per launch, %.0f solves/s" % (1e3 * (t1 - t0) / (2 * reps), 2 * K * reps / (t1 - t0)))
for ti, a, b in sorted(log, key=lambda r: r[1]):
    print("thread %d call [%.2f, %.2f] ms" % (ti, 1e3 * (a - t0), 1e3 * (b - t0)))

import sys; sys.path.insert(0, '.')
import numpy as np
from ov2slam_b200 import api, synth
ctx = api.Context(0)
pb0 = synth.make_ba_problem(3, 10, 2000, 8000)
opt = api.Optimizer(ctx)
for _ in range(3):
    pb = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in pb0.items()}
    print(opt.local_ba(pb)[0])

import sys, os; sys.path.insert(0, '.')
import numpy as np
from ov2slam_b200 import api, synth
from oracle import ba_ref as B
os.environ["OV2_BA_DEBUG"] = "1"
pb = synth.make_ba_problem(5, 6, 100, 300)
pb["pose_const"][:] = 1
ref = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in pb.items()}
log = []
r = B.local_ba(ref, log=log)
for e in log: print('[ref] it %d x_cost %.10g cand %.10g mcc %.10g radius %.4g rho %.4g step %.6g' % (e['it'], e['x_cost'], e['cand_cost'], e['mcc'], e['radius'], e['rho'], np.linalg.norm(e['step'])))
print({k: v for k, v in r.items() if k not in ('flags', 'summaries')})
ctx = api.Context(0)
gpu = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in pb.items()}
g, flags = api.Optimizer(ctx).local_ba(gpu)
print(g)
d = np.abs(gpu["lm_invdepth"] - ref["lm_invdepth"])
idx = np.argsort(-d)[:8]
for l in idx:
    obs = np.nonzero(pb["obs_lm"] == l)[0]
    print('lm', l, 'diff', d[l], 'init', pb["lm_invdepth"][l], 'ref', ref["lm_invdepth"][l], 'gpu', gpu["lm_invdepth"][l], 'nobs', len(obs), 'flags ref', r["flags"][obs], 'gpu', flags[obs])

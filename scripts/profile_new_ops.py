"""Driver for ncu: the round-2 operators that no bench leg launches - the BRIEF-32 box-sum descriptor and the batched stereo
row search - on a batch of 64 frames 1280x720 (1024 points per frame)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ov2slam_b200 import api, synth  # noqa: E402

ctx = api.Context(0)
B, w, h, n = 64, 1280, 720, 1024
base = np.stack([synth.make_frame(10 + i, w, h) for i in range(4)])
left = np.ascontiguousarray(np.tile(base, (B // 4, 1, 1)))
right = np.ascontiguousarray(np.roll(left, -11, axis=2))
pl, pr = api.Pyramid(ctx, B, w, h, 3), api.Pyramid(ctx, B, w, h, 3)
pl.build(left)
pr.build(right)
rng = np.random.default_rng(0)
pts = (rng.random((B * n, 2)) * [w - 80, h - 80] + 40).astype(np.float32)
fe, ft = api.FeatureExtractor(ctx), api.FeatureTracker(ctx)
pairs = np.clip(np.rint(rng.normal(0, 48 / 5.0, (256, 4))), -24, 24).astype(np.int8)
desc = np.empty((B * n, 32), np.uint8)
valid = np.empty(B * n, np.uint8)
xp = np.empty(B * n, np.float32)
er = np.empty(B * n, np.float32)
for rep in range(3):
    fe.describe_config(fe.DESC_BRIEF32, pairs)
    fe.describe_brief(pl, pts, desc, valid, per_frame=n)
    fe.describe_config(fe.DESC_ORB_FALLBACK)
    fe.describe_brief(pl, pts, desc, valid, per_frame=n)
    ft.line_min_sad(pl, pr, 3, pts / 8.0, 7, True, xp, er, per_frame=n)
ctx.profile(True)
fe.describe_config(fe.DESC_BRIEF32, pairs)
fe.describe_brief(pl, pts, desc, valid, per_frame=n)
fe.describe_config(fe.DESC_ORB_FALLBACK)
fe.describe_brief(pl, pts, desc, valid, per_frame=n)
ft.line_min_sad(pl, pr, 3, pts / 8.0, 7, True, xp, er, per_frame=n)
print({k: (round(v[0], 4), v[1]) for k, v in ctx.profile_report().items()}, "valid", int(valid.sum()), "priors found", int((xp >= 0).sum()))

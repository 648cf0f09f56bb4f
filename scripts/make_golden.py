"""Generate tests/golden/frontend_golden.npz from the REAL OpenCV (cv2) call sequence the
reference uses (oracle/image_ref.py *_cv2 functions) on seeded synthetic inputs.  Inputs are
re-created from the seeds by ov2slam_b200.synth, so only outputs are stored.

    python scripts/make_golden.py      (needs cv2; run in the build container)
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import image_ref as R  # noqa: E402
from ov2slam_b200 import synth  # noqa: E402

assert R.HAVE_CV2
import cv2  # noqa: E402

W, H, SEED = 640, 480, 424242
out = {"cv2_version": np.array(cv2.__version__), "w": W, "h": H, "seed": SEED}
prev, cur, flow = synth.make_pair(SEED, W, H)
out["img_sha"] = np.array(hashlib.sha256(prev.tobytes() + cur.tobytes()).hexdigest())
pyr = R.build_pyramid_cv2(cur, 3)
for l in range(1, 4):
    out[f"pyr{l}"] = pyr[l]
rng = np.random.default_rng(1)
for cs in (50, 35, 16):
    cur_kps = (rng.random((25, 2)) * [W, H]).astype(np.float32)
    for tag, k in (("empty", np.zeros((0, 2), np.float32)), ("kps", cur_kps)):
        sp, ip, th = R.detect_grid_fast_cv2(prev, cs, k, 10)
        out[f"fast_{cs}_{tag}_in"] = k
        out[f"fast_{cs}_{tag}_int"] = ip
        out[f"fast_{cs}_{tag}_subpix"] = sp
        out[f"fast_{cs}_{tag}_th"] = th
sp, ip, _ = R.detect_grid_fast_cv2(prev, 16, np.zeros((0, 2), np.float32), 10)
pts = np.concatenate([sp, (rng.random((200, 2)) * [W, H]).astype(np.float32)])
d, v = R.describe_cv2(prev, pts)
out["desc_pts"], out["desc"], out["desc_valid"] = pts, d, v
kps = np.concatenate([sp[:300], (rng.random((60, 2)) * [W, H]).astype(np.float32)])
is3d, pri = synth.make_priors(SEED, kps, flow)
out["klt_kps"], out["klt_pri"] = kps, pri
for lvl in (0, 1, 3):
    t, s = R.fb_klt_cv2(prev, cur, kps, pri, 9, lvl)
    out[f"klt_{lvl}_tracked"], out[f"klt_{lvl}_status"] = t, s
dst = ROOT / "tests" / "golden" / "frontend_golden.npz"
np.savez_compressed(dst, **out)
print(dst, dst.stat().st_size, "bytes")

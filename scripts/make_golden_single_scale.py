"""Generate tests/golden/single_scale_golden.npz: FeatureExtractor::detectSingleScale outputs from the REAL
OpenCV (cv2) call sequence (oracle/image_ref.py::detect_single_scale_cv2) on seeded synthetic inputs.
Inputs are re-created from the seeds by ov2slam_b200.synth, so only outputs are stored.

    python scripts/make_golden_single_scale.py      (needs cv2; run in the build container)
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

W, H, SEED = 640, 480, 515151
out = {"cv2_version": np.array(cv2.__version__), "w": W, "h": H, "seed": SEED}
im = synth.make_frame(SEED, W, H)
out["img_sha"] = np.array(hashlib.sha256(im.tobytes()).hexdigest())
rng = np.random.default_rng(2)
for cs in (50, 35):
    kps = (rng.random((30, 2)) * [W, H]).astype(np.float32)
    for tag, k, roi, q in (("empty", np.zeros((0, 2), np.float32), (0, 0, W, H), 0.001),
                           ("kps_roi", kps, (20, 24, W - 45, H - 50), 0.0005)):
        sp, ip, qn = R.detect_single_scale_cv2(im, cs, k, roi, q)
        out[f"ss_{cs}_{tag}_in"] = k
        out[f"ss_{cs}_{tag}_roi"] = np.array(roi, np.int32)
        out[f"ss_{cs}_{tag}_q"] = np.array([q, qn], np.float64)
        out[f"ss_{cs}_{tag}_int"] = ip
        out[f"ss_{cs}_{tag}_subpix"] = sp
dst = ROOT / "tests" / "golden" / "single_scale_golden.npz"
np.savez_compressed(dst, **out)
print(dst, dst.stat().st_size, "bytes")

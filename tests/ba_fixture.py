"""A localBA window built on the observation graph / geometry of Ceres' own BAL test problem (tests/golden/bal16_structure.npz,
scripts/make_bal_fixture.py): anchored inverse-depth flattening exactly as Optimizer::localBA's set-up does it
(/root/reference/src/optimizer.cpp:258-290: the lowest-index observing keyframe anchors a landmark), pixel measurements
re-synthesised through OV2SLAM's pinhole residual with seeded noise and gross outliers.  Test infrastructure only."""
from pathlib import Path

import numpy as np


def _quat_to_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def bal_window(seed=0, px_noise=0.5, outlier_frac=0.05, nconst=2, max_pts=None):
    g = np.load(Path(__file__).parent / "golden" / "bal16_structure.npz")
    pose_true = g["pose"].astype(np.float64)
    pts = g["points"].astype(np.float64)
    oc, op = g["obs_cam"].astype(np.int64), g["obs_pt"].astype(np.int64)
    ncam = len(pose_true)
    rng = np.random.default_rng(seed)
    f = float(np.mean(g["focal"]))
    K = np.array([f, f, 512.0, 384.0])
    R = np.stack([_quat_to_rot(p[3:]) for p in pose_true])
    t = pose_true[:, :3]
    pc = np.einsum("nji,nj->ni", R[oc], pts[op] - t[oc])           # Rwc^T (X - twc)
    ok = pc[:, 2] > 0.05
    oc, op, pc = oc[ok], op[ok], pc[ok]
    px = np.stack([K[0] * pc[:, 0] / pc[:, 2] + K[2], K[1] * pc[:, 1] / pc[:, 2] + K[3]], 1)
    order = np.lexsort((oc, op))                                   # by landmark, then ascending keyframe
    oc, op, pc, px = oc[order], op[order], pc[order], px[order]
    first = np.r_[True, op[1:] != op[:-1]]
    counts = np.bincount(op, minlength=len(pts))
    keep_lm = np.nonzero(counts >= 2)[0]
    if max_pts is not None:
        keep_lm = keep_lm[:max_pts]
    remap = -np.ones(len(pts), np.int64)
    remap[keep_lm] = np.arange(len(keep_lm))
    sel = remap[op] >= 0
    oc, op, pc, px, first = oc[sel], remap[op[sel]], pc[sel], px[sel], first[sel]
    anchor_px = px[first].astype(np.float32).astype(np.float64)    # cv::Point2f in the reference
    anchor_cam = oc[first].astype(np.int32)
    invd_true = 1.0 / pc[first, 2]
    rest = ~first
    obs_px = px[rest] + rng.normal(0, px_noise, (int(rest.sum()), 2))
    bad = rng.random(len(obs_px)) < outlier_frac
    ang = rng.uniform(0, 2 * np.pi, int(bad.sum()))
    obs_px[bad] += rng.uniform(10, 50, int(bad.sum()))[:, None] * np.stack([np.cos(ang), np.sin(ang)], 1)
    pose = pose_true.copy()
    const = np.zeros(ncam, np.uint8)
    const[:nconst] = 1
    scale = float(np.median(pc[:, 2]))
    for c in range(nconst, ncam):                                  # perturb the optimised keyframes (about 0.5 deg, 0.2 % of the depth)
        w = rng.normal(0, np.deg2rad(0.5), 3)
        th = np.linalg.norm(w)
        dq = np.r_[np.sin(th / 2) * w / th, np.cos(th / 2)]
        x1, y1, z1, w1 = dq
        x2, y2, z2, w2 = pose[c, 3:]
        pose[c, 3:] = [w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                       w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2]
        pose[c, :3] += rng.normal(0, 0.002 * scale, 3)
    return dict(K=K, pose=pose, pose_const=const, lm_anchor_cam=anchor_cam, lm_anchor_px=np.ascontiguousarray(anchor_px),
                lm_invdepth=invd_true * (1 + rng.normal(0, 0.05, len(invd_true))), obs_cam=oc[rest].astype(np.int32),
                obs_lm=op[rest].astype(np.int32), obs_px=np.ascontiguousarray(obs_px.astype(np.float32).astype(np.float64)),
                truth_pose=pose_true, truth_invdepth=invd_true)

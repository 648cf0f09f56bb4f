"""tests/golden/bal16_structure.npz: the observation GRAPH, camera poses and 3-D points of Ceres' own bundle-adjustment test
problem (Thirdparty/ceres-solver/data/problem-16-22106-pre.txt, used by its generated_bundle_adjustment_tests: 16 cameras,
22 106 points, 83 718 observations) converted to OV2SLAM's conventions - a STRUCTURE fixture: BAL's Snavely
camera (per-camera focal length + radial distortion, looking down -z) is not OV2SLAM's residual, so the pixel measurements
are re-synthesised through the anchored inverse-depth model by tests/ba_fixture.py; what is kept from the real data set is
what a synthetic generator cannot give: the track-length distribution (2 .. 16 views), the uneven camera-to-camera
covisibility and the scene geometry.

    python scripts/make_bal_fixture.py        (reads /root/reference, writes tests/golden/)
"""
import os

import numpy as np

SRC = "/root/reference/Thirdparty/ceres-solver/data/problem-16-22106-pre.txt"


def rodrigues(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def rot_to_quat_xyzw(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    q /= np.linalg.norm(q)
    return q if q[3] >= 0 else -q


if __name__ == "__main__":
    with open(SRC) as f:
        ncam, npts, nobs = (int(v) for v in f.readline().split())
        obs = np.array([f.readline().split() for _ in range(nobs)], dtype=np.float64)
        rest = np.array([float(f.readline()) for _ in range(9 * ncam + 3 * npts)])
    cams = rest[:9 * ncam].reshape(ncam, 9)
    pts = rest[9 * ncam:].reshape(npts, 3)
    S = np.diag([1.0, -1.0, -1.0])                      # BAL looks down -z with y up; OV2SLAM: +z forward, y down
    pose = np.zeros((ncam, 7))
    for c in range(ncam):
        Rcw = S @ rodrigues(cams[c, :3])
        tcw = S @ cams[c, 3:6]
        Rwc = Rcw.T
        pose[c, :3] = -Rwc @ tcw
        pose[c, 3:] = rot_to_quat_xyzw(Rwc)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "bal16_structure.npz")
    np.savez_compressed(out, pose=pose, points=pts.astype(np.float32), obs_cam=obs[:, 0].astype(np.uint8), obs_pt=obs[:, 1].astype(np.int32),
                        focal=cams[:, 6].astype(np.float32))
    print("wrote", out, ncam, npts, nobs, os.path.getsize(out))

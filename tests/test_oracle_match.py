"""Oracle of Mapper::matchToMap (oracle/match_ref.py): the distortion model against cv2.projectPoints, the flow on
hand-built cases."""
import numpy as np
import pytest

from oracle import match_ref as M
from ov2slam_b200 import synth


def test_projection_matches_cv2_project_points():
    """CameraCalibration::projectCamToImageDist (camera_calibration.cpp:254-281) goes through cv::projectPoints on the
    float-rounded normalised point: the restatement must give the same float pixel."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    K = (458.654, 457.296, 367.215, 248.375)
    dist = (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)
    Kcv = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float64)
    for _ in range(3000):
        p = np.array([rng.uniform(-2, 2), rng.uniform(-1.5, 1.5), rng.uniform(0.5, 8)])
        x, y = M.project_dist(p, K, dist)
        pt = np.array([[np.float32(p[0] / p[2]), np.float32(p[1] / p[2]), np.float32(1.0)]], np.float32)
        ref, _ = cv2.projectPoints(pt, np.zeros(3, np.float32), np.zeros(3, np.float32), Kcv, np.array(dist, np.float64))
        rx, ry = ref.reshape(2)
        assert x == np.float32(rx) and y == np.float32(ry)
        ux, uy = M.project_dist(p, K, None)
        assert ux == np.float32(K[0] * (p[0] / p[2]) + K[2]) and uy == np.float32(K[1] * (p[1] / p[2]) + K[3])


def _tiny_scene():
    """One frame at the origin, a 2 x 2 grid of 50-px cells, three keypoints, four map points."""
    I = np.concatenate([np.eye(3).reshape(-1), np.zeros(3)])
    K = np.array([100.0, 100.0, 50.0, 50.0])
    d = lambda seed: np.random.default_rng(seed).integers(0, 256, 32, dtype=np.uint8)
    desc = np.stack([d(0), d(1), d(2), d(0), d(0)])          # map points 3 (candidate) and 0 share a descriptor; extra copy for mp 4
    desc[3, 0] ^= 1                                           # candidate 3: one bit away from map point 0
    desc[4, 0] ^= 3                                           # candidate 4: two bits away
    mask = lambda *ks: np.array([sum(1 << k for k in ks), 0, 0, 0], np.uint64)
    return dict(K=K, dist=None, img_w=100, img_h=100, ncellsize=50, nbwcells=2, Tcw=I,
                cell_ptr=np.array([0, 0, 0, 0, 3], np.int32), cell_kp=np.array([0, 1, 2], np.int32),
                kp_px=np.array([[60.0, 60.0], [62.0, 61.0], [90.0, 90.0]], np.float32), kp_lm=np.array([0, 1, 2], np.int32),
                mp_xyz=np.array([[0.1, 0.1, 1.0], [0.12, 0.11, 1.0], [0.4, 0.4, 1.0], [0.1, 0.1, 1.0], [0.1, 0.1, 1.0]]),
                mp_desc_ptr=np.array([0, 1, 2, 3, 4, 5], np.int32), desc=desc,
                mp_kfmask=np.stack([mask(0), mask(1), mask(2), mask(3), mask(4)]),
                mp_obs_ptr=np.array([0, 0, 0, 0, 0, 0], np.int32), obs_kf=np.zeros(0, np.int32), obs_px=np.zeros((0, 2), np.float32),
                kf_Tcw=np.tile(I, (5, 1)), cand_mp=np.array([3, 4], np.int32),
                dmaxpxdist=np.float32(4.0), fdistratio=np.float32(0.2), view_th=np.float32(0.5))


def test_flow_on_a_hand_built_scene():
    sc = _tiny_scene()
    bk, bd, km, kd = M.match_to_map(sc)
    # both candidates project to (60, 60) and match keypoint 0 (1 and 2 bits); keypoint 1 is 2.2 px away but its descriptor is
    # unrelated (~128 bits > 51.2); the per-keypoint pass keeps the closer candidate (index 0 of cand_mp)
    assert list(bk) == [0, 0] and list(bd) == [1.0, 2.0]
    assert list(km) == [0, -1, -1] and kd[0] == 1.0 and kd[1] == 1024.0
    # observed together with the keypoint's map point -> not a candidate pair
    sc2 = dict(sc)
    sc2["mp_kfmask"] = sc["mp_kfmask"].copy()
    sc2["mp_kfmask"][3, 0] |= np.uint64(1)
    bk2, _, km2, _ = M.match_to_map(sc2)
    assert list(bk2) == [-1, 0] and km2[0] == 1
    # second best within the 0.9 ratio -> rejected: make keypoint 1's map point a near copy too
    sc3 = dict(sc)
    sc3["desc"] = sc["desc"].copy()
    sc3["desc"][1] = sc["desc"][0]
    bk3, _, _, _ = M.match_to_map(sc3)
    assert list(bk3) == [-1, -1]
    # behind the camera / outside the image / beyond the pixel radius
    sc4 = dict(sc)
    sc4["mp_xyz"] = sc["mp_xyz"].copy()
    sc4["mp_xyz"][3] = [0.1, 0.1, -1.0]
    sc4["mp_xyz"][4] = [0.2, 0.1, 1.0]                        # projects to (70, 60): 10 px from keypoint 0
    bk4, _, _, _ = M.match_to_map(sc4)
    assert list(bk4) == [-1, -1]


def test_synthetic_scene_has_matches_and_rejections():
    sc = synth.make_match_scene(3, 300, 160)
    bk, bd, km, kd = M.match_to_map(sc)
    assert 20 < (bk >= 0).sum() < 150 and (km >= 0).sum() > 15
    assert np.all(bd[bk >= 0] <= 51.2) and np.all(kd[km < 0] == 1024.0)

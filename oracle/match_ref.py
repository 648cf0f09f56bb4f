"""CPU ORACLE for Mapper::matchToMap (local-map matching).  TEST INFRASTRUCTURE ONLY.

Last "next" row of the scope table: /root/reference/src/mapper.cpp:576-774 restated on the FLATTENED
arrays ov2_match_to_map takes (the map walking that produces them is the host shim's job):

  projection                 Frame::projWorldToCam + CameraCalibration::projectCamToImageDist
                             (src/frame.cpp:811-828, src/camera_calibration.cpp:254-281): pinhole, or pinhole +
                             cv::projectPoints on the float-rounded normalised point for a radial-tangential model
                             (restated from OpenCV's cvProjectPoints2Internal; pinned against cv2.projectPoints by
                             tests/test_oracle_match.py); fisheye is not restated
  culling                    depth < 0.1, |z / norm| < cos(max half fov), outside the image (:618-637)
  candidate keypoints        Frame::getSurroundingKeypoints(pt) (src/frame.cpp:624-650): the FOUR cells
                             (r-1 .. r) x (c-1 .. c) around the projection, cell lists in insertion order
  per candidate              pixel distance, "never observed together" (keyframe sets disjoint), mean co-projection
                             error of the map point into the keyframes that observe the keypoint's map point,
                             minimal Hamming distance over all descriptor pairs (src/map_point.cpp:236-252) (:654-724)
  best / second, 0.9 ratio   (:726-744)
  per keypoint               the candidate with the smallest distance, later candidates win ties (:758-772)

The reference iterates an unordered_set of map-point ids; the order of `cand_mp` is used here (and by the kernel).
Rotations are applied as 3 x 3 matrices (the reference goes through Sophus' quaternion product: last-bit differences in
the camera-frame point).

PINNING STATUS: the distortion model is pinned against cv2; the flow against the reference's own Mapper::matchToMap (src/mapper.cpp
compiled in place with the reference's Frame / MapManager / MapPoint, oracle/ref_build/build_map_ref.py) on maps built from flattened
scenes with an undistorted calibration: the same pairs for the same candidate order (tests/test_oracle_vs_reference_map.py); hand-built
cases in tests/test_oracle_match.py.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def project_dist(campt, K, dist):
    """CameraCalibration::projectCamToImageDist for the pinhole model: float32 pixel (x, y)."""
    fx, fy, cx, cy = (float(v) for v in K)
    invz = 1.0 / float(campt[2])
    x, y = float(campt[0]) * invz, float(campt[1]) * invz
    if dist is None:
        return f32(fx * x + cx), f32(fy * y + cy)
    # cv::Point3f(x, y, 1.) -> cv::projectPoints with zero rotation / translation
    x, y = float(f32(x)), float(f32(y))
    k1, k2, p1, p2, k3 = (float(v) for v in dist)
    r2 = x * x + y * y
    r4 = r2 * r2
    r6 = r4 * r2
    a1 = 2 * x * y
    a2 = r2 + 2 * x * x
    a3 = r2 + 2 * y * y
    cdist = 1 + k1 * r2 + k2 * r4 + k3 * r6
    xd = x * cdist + p1 * a1 + p2 * a2
    yd = y * cdist + p1 * a3 + p2 * a1
    return f32(xd * fx + cx), f32(yd * fy + cy)


def _norm2(dx, dy):
    """cv::norm(cv::Point2f): sqrt in double of the float components."""
    return math.sqrt(float(dx) * float(dx) + float(dy) * float(dy))


def _hamming(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def match_to_map(sc: dict):
    """Returns (best_kp int32[ncand], best_dist float32[ncand], kp_match int32[nkps], kp_dist float32[nkps])."""
    K, dist = sc["K"], sc.get("dist")
    R, t = np.asarray(sc["Tcw"][:9], np.float64).reshape(3, 3), np.asarray(sc["Tcw"][9:], np.float64)
    cs, nbw = int(sc["ncellsize"]), int(sc["nbwcells"])
    ncells = len(sc["cell_ptr"]) - 1
    w, h = int(sc["img_w"]), int(sc["img_h"])
    dmax, ratio, view_th = f32(sc["dmaxpxdist"]), f32(sc["fdistratio"]), f32(sc["view_th"])
    kf_T = np.asarray(sc["kf_Tcw"], np.float64).reshape(-1, 12)
    ncand = len(sc["cand_mp"])
    best_kp = -np.ones(ncand, np.int32)
    best_dist = np.zeros(ncand, np.float32)
    mindist = f32(float(f32(32 * ratio)) * 8.0)
    for ci, mp in enumerate(sc["cand_mp"]):
        wpt = np.asarray(sc["mp_xyz"][mp], np.float64)
        campt = R @ wpt + t
        if campt[2] < 0.1:
            continue
        view_angle = f32(campt[2] / math.sqrt(float(campt @ campt)))
        if abs(view_angle) < view_th:
            continue
        px, py = project_dist(campt, K, dist)
        if not (px >= 0 and py >= 0 and px < w and py < h):
            continue
        rkp, ckp = int(math.floor(float(f32(py / f32(cs))))), int(math.floor(float(f32(px / f32(cs)))))
        bestid = secid = -1
        bestdist = secdist = mindist
        for r in (rkp - 1, rkp):
            for c in (ckp - 1, ckp):
                idx = r * nbw + c
                if r < 0 or c < 0 or idx >= ncells:
                    continue
                for j in sc["cell_kp"][sc["cell_ptr"][idx]:sc["cell_ptr"][idx + 1]]:
                    lm = int(sc["kp_lm"][j])
                    if lm < 0:
                        continue
                    kx, ky = sc["kp_px"][j]
                    pxdist = f32(_norm2(f32(px - kx), f32(py - ky)))
                    if pxdist > dmax:
                        continue
                    d0, d1 = sc["mp_desc_ptr"][lm], sc["mp_desc_ptr"][lm + 1]
                    if d1 == d0:
                        continue
                    if np.any(np.bitwise_and(sc["mp_kfmask"][mp], sc["mp_kfmask"][lm])):
                        continue
                    coproj, nb = f32(0.0), 0
                    for o in range(sc["mp_obs_ptr"][lm], sc["mp_obs_ptr"][lm + 1]):
                        T = kf_T[sc["obs_kf"][o]]
                        cpt = T[:9].reshape(3, 3) @ wpt + T[9:]
                        qx, qy = project_dist(cpt, K, dist)
                        ox, oy = sc["obs_px"][o]
                        coproj = f32(float(coproj) + _norm2(f32(ox - qx), f32(oy - qy)))
                        nb += 1
                    if nb > 0 and f32(coproj / f32(nb)) > dmax:
                        continue
                    dd = f32(1000.0)
                    for a in sc["desc"][sc["mp_desc_ptr"][mp]:sc["mp_desc_ptr"][mp + 1]]:
                        for b in sc["desc"][d0:d1]:
                            v = f32(_hamming(a, b))
                            if v < dd:
                                dd = v
                    if dd <= bestdist:
                        secdist, secid = bestdist, bestid
                        bestdist, bestid = dd, int(j)
                    elif dd <= secdist:
                        secdist, secid = dd, int(j)
        if bestid != -1 and secid != -1 and 0.9 * float(secdist) < float(bestdist):
            bestid = -1
        if bestid >= 0:
            best_kp[ci] = bestid
            best_dist[ci] = bestdist
    nk = len(sc["kp_px"])
    kp_match = -np.ones(nk, np.int32)
    kp_dist = np.full(nk, 1024.0, np.float32)
    for ci in range(ncand):
        j = best_kp[ci]
        if j >= 0 and best_dist[ci] <= kp_dist[j]:
            kp_dist[j] = best_dist[ci]
            kp_match[j] = ci
    return best_kp, best_dist, kp_match, kp_dist

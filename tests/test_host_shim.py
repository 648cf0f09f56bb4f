"""The drop-in C++ classes (ov2slam_b200/host/) behave like the Python binding of the same C ABI:
CPU part = they compile and link; GPU part = same keypoints / descriptors / tracks."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from ov2slam_b200 import api, build, synth
from oracle import image_ref as R_img

ROOT = Path(__file__).resolve().parents[1]


def test_host_shims_compile_and_link():
    exe = build.build_host_shims()
    assert exe.exists()


@pytest.mark.gpu
def test_host_shims_match_python_binding(ctx, tmp_path):
    exe = build.build_host_shims()
    w, h = 640, 480
    prev, cur, _ = synth.make_pair(77, w, h)
    (tmp_path / "p.raw").write_bytes(prev.tobytes())
    (tmp_path / "c.raw").write_bytes(cur.tobytes())
    out = subprocess.run([str(exe), str(tmp_path / "p.raw"), str(tmp_path / "c.raw"), str(w), str(h)],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    tok = out.stdout.split()
    got = dict(nkps=int(tok[1]), th=int(tok[3]), kx=float(tok[5]), ky=float(tok[6]), ndesc=int(tok[8]), dsum=int(tok[10]),
               ngood=int(tok[12]))
    pp = api.Pyramid(ctx, 1, w, h, 3)
    cp = api.Pyramid(ctx, 1, w, h, 3)
    pp.build(prev[None])
    cp.build(cur[None])
    fe = api.FeatureExtractor(ctx, nfast_th=10)
    pts, _ = fe.detect_grid_fast_frame(pp, 0, 50, np.zeros((0, 2), np.float32))
    d = np.zeros((len(pts), 32), np.uint8)
    v = np.zeros(len(pts), np.uint8)
    fe.describe_brief(pp, pts, d, v)
    pri = pts.copy()
    st = np.zeros(len(pts), np.uint8)
    api.FeatureTracker(ctx, 30, 0.01).fb_klt_tracking(pp, cp, 9, 3, 30.0, 0.5, pts, pri, st)
    assert got["nkps"] == len(pts) and got["th"] == fe.nfast_th_
    assert abs(got["kx"] - float(pts[:, 0].astype(np.float64).sum())) < 1e-2
    assert got["ndesc"] == int(v.sum())
    assert got["dsum"] == int((d[v.astype(bool)].astype(np.int64) * np.arange(1, 33)).sum())
    assert got["ngood"] == int(st.sum())


_rect_subpix_template = R_img.get_rect_subpix_u8_ref
_line_min_sad_ref = R_img.line_min_sad_ref


def test_get_line_min_sad_matches_the_reference_function():
    """The drop-in FeatureTracker::getLineMinSAD (host C++) against a restatement of the reference's function with
    OpenCV's 8-bit getRectSubPix template: identical priors and errors - sub-pixel points, border window shrink, both
    scan directions (round 1 rounded the point, scanned the other way and dropped the shrink rule: ADVICE r1).  With
    cv2's own getRectSubPix (an IPP build: +-1 grey level on ~0.7 % of the pixels) the errors agree to 0.05 and the
    priors almost always."""
    exe = build.build_host_shims()
    w, h = 320, 240
    left = synth.make_frame(5, w, h, nrect=60)
    right = np.ascontiguousarray(np.roll(left, -7, axis=1))
    rng = np.random.default_rng(3)
    pts = [(rng.uniform(0, w - 1), rng.uniform(0, h - 1), 7, int(rng.integers(0, 2))) for _ in range(40)]
    pts += [(2.3, 100.7, 7, 1), (w - 2.2, 50.1, 7, 0), (150.5, 1.4, 7, 1), (150.5, h - 1.6, 7, 0), (60.0, 60.0, 8, 1), (0.4, 0.3, 7, 1)]
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        (open(f"{d}/l.raw", "wb")).write(left.tobytes())
        (open(f"{d}/r.raw", "wb")).write(right.tobytes())
        args = [str(exe), "--sad", f"{d}/l.raw", f"{d}/r.raw", str(w), str(h), str(len(pts))]
        for x, y, ws, gl in pts:
            args += [repr(float(np.float32(x))), repr(float(np.float32(y))), str(ws), str(gl)]
        out = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    got = [tuple(float(v) for v in line.split()) for line in out.stdout.strip().splitlines()]
    assert len(got) == len(pts)
    try:
        import cv2
        cv_sub = lambda im, ws, cx, cy: cv2.getRectSubPix(im, (ws, ws), (float(cx), float(cy)))
    except Exception:
        cv_sub = None
    agree = 0
    for (x, y, ws, gl), (gx, ge) in zip(pts, got):
        rx, re = _line_min_sad_ref(left, right, (x, y), ws, bool(gl), _rect_subpix_template)
        if re is None:
            assert gx == -1.0
            continue
        assert np.float32(gx) == np.float32(rx) and np.float32(ge) == np.float32(re), ((x, y, ws, gl), (gx, ge), (rx, re))
        if cv_sub is not None:
            cx, ce = _line_min_sad_ref(left, right, (x, y), ws, bool(gl), cv_sub)
            assert abs(ge - ce) <= 0.05
            agree += np.float32(gx) == np.float32(cx)
    if cv_sub is not None:
        assert agree >= 0.8 * len(pts)


def _write_window(path, pb):
    import struct
    stereo = pb.get("obs_type") is not None
    ncam, npts, nobs = len(pb["pose"]), len(pb["lm_invdepth"]), len(pb["obs_cam"])
    with open(path, "wb") as f:
        f.write(struct.pack("4i", ncam, npts, nobs, int(stereo)))
        f.write(np.asarray(pb["K"], np.float64).tobytes())
        f.write(np.asarray(pb["Kr"] if stereo else pb["K"], np.float64).tobytes())
        f.write(np.asarray(pb["Trl"] if stereo else [0, 0, 0, 0, 0, 0, 1], np.float64).tobytes())
        for k, dt in (("pose", np.float64), ("pose_const", np.uint8), ("lm_anchor_cam", np.int32), ("lm_anchor_px", np.float64),
                      ("lm_invdepth", np.float64), ("obs_cam", np.int32), ("obs_lm", np.int32), ("obs_px", np.float64)):
            f.write(np.ascontiguousarray(pb[k], dt).tobytes())
        f.write(np.ascontiguousarray(pb["obs_type"] if stereo else np.zeros(nobs), np.uint8).tobytes())


def _drop_landmarks_seen_only_by_constant_keyframes(pb):
    """Optimizer::localBA collects its landmarks from the 3D keypoints of the OPTIMISED keyframes (optimizer.cpp:178-180): a
    landmark whose anchor and observers are all constant keyframes is not part of the reference's window."""
    const = pb["pose_const"].astype(bool)
    npts = len(pb["lm_invdepth"])
    seen = ~const[pb["lm_anchor_cam"]]
    np.logical_or.at(seen, pb["obs_lm"], ~const[pb["obs_cam"]])
    keep_l = np.nonzero(seen)[0]
    remap = -np.ones(npts, np.int64)
    remap[keep_l] = np.arange(len(keep_l))
    keep_o = np.nonzero(seen[pb["obs_lm"]])[0]
    out = dict(pb)
    for k in ("lm_anchor_cam", "lm_anchor_px", "lm_invdepth"):
        out[k] = np.ascontiguousarray(pb[k][keep_l])
    for k in ("obs_cam", "obs_px") + (("obs_type",) if pb.get("obs_type") is not None else ()):
        out[k] = np.ascontiguousarray(pb[k][keep_o])
    out["obs_lm"] = np.ascontiguousarray(remap[pb["obs_lm"][keep_o]].astype(np.int32))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("stereo,stop", [(False, False), (True, False), (False, True)])
def test_optimizer_localba_shim_updates_the_map_like_the_flat_solve(ctx, tmp_path, stereo, stop):
    """The drop-in Optimizer::localBA (host/optimizer_localba_gpu.cpp, compiled against the stand-in map classes) on a
    synthetic map: window selection over the covisibility map, anchored inverse-depth flattening (mono / stereo), GPU solve,
    write-back - the keyframe poses and landmark inverse depths left in the map equal ov2_localba_solve on the flat window;
    a stop request raised before the call skips the refinement and is consumed (bstop_localba_ was never cleared)."""
    exe = build.build_optimizer_shim()
    pb = _drop_landmarks_seen_only_by_constant_keyframes(synth.make_ba_problem(61 + stereo, 9, 600, 2400, stereo=stereo))
    _write_window(tmp_path / "w.bin", pb)
    out = subprocess.run([str(exe), str(tmp_path / "w.bin"), str(tmp_path / "r.bin")] + (["stop"] if stop else []),
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    ncam, npts = len(pb["pose"]), len(pb["lm_invdepth"])
    raw = np.fromfile(tmp_path / "r.bin", np.uint8)
    pose = raw[:ncam * 56].view(np.float64).reshape(ncam, 7)
    invd = raw[ncam * 56:ncam * 56 + npts * 8].view(np.float64)
    ref = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in pb.items()}
    if stop:
        api.request_stop_local_ba(ctx, True)
    res, flags = api.Optimizer(ctx).local_ba(ref)
    api.request_stop_local_ba(ctx, False)
    if stop:
        assert res["iters_refine"] == 0
    # the shim stores unit quaternions through Sophus (sign / normalisation): compare rotations up to sign
    for c in range(ncam):
        assert np.abs(pose[c, :3] - ref["pose"][c, :3]).max() <= 1e-7
        q, r = pose[c, 3:], ref["pose"][c, 3:] / np.linalg.norm(ref["pose"][c, 3:])
        assert min(np.abs(q - r).max(), np.abs(q + r).max()) <= 1e-7
    alive = invd >= 0
    assert alive.mean() > 0.9
    assert np.abs(invd[alive] - ref["lm_invdepth"][alive]).max() <= 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("mode,stereo", [("loose", False), ("loose", True), ("full", False), ("full", True)])
def test_optimizer_loose_and_full_ba_shims(ctx, tmp_path, mode, stereo):
    """8f-4: the drop-in Optimizer::looseBA / fullBA (optimizer.cpp:900-1671 / 1674-2331; same residual blocks as localBA over a
    keyframe RANGE: the first two (mono) / one (stereo) keyframes constant, all others optimised, one 5-iteration solve with
    function_tolerance 1e-4 for looseBA, up to 100 + 100 iterations at Ceres' default 1e-6 for fullBA) on a synthetic map:
    the map they leave behind equals ov2_localba_solve on the flat window with those options; looseBA also moves the current
    frame rigidly with the loop keyframe (checked inside the self-test)."""
    exe = build.build_optimizer_shim()
    pb = synth.make_ba_problem(71 + stereo, 9, 600, 2400, stereo=stereo)
    pb["pose_const"] = np.zeros_like(pb["pose_const"])
    pb["pose_const"][:1 if stereo else 2] = 1
    pb = _drop_landmarks_seen_only_by_constant_keyframes(pb)
    _write_window(tmp_path / "w.bin", pb)
    out = subprocess.run([str(exe), str(tmp_path / "w.bin"), str(tmp_path / "r.bin"), mode], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    ncam, npts = len(pb["pose"]), len(pb["lm_invdepth"])
    raw = np.fromfile(tmp_path / "r.bin", np.uint8)
    pose = raw[:ncam * 56].view(np.float64).reshape(ncam, 7)
    invd = raw[ncam * 56:ncam * 56 + npts * 8].view(np.float64)
    ref = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in pb.items()}
    if mode == "loose":
        res, flags = api.Optimizer(ctx).local_ba(ref, max_iters_robust=5, max_iters_refine=0, function_tolerance=1e-4, apply_l2_after_robust=0)
        assert res["iters_refine"] == 0
    else:
        res, flags = api.Optimizer(ctx).local_ba(ref, max_iters_robust=100, max_iters_refine=100, function_tolerance=1e-6)
    assert res["final_cost"] < res["initial_cost"]
    for c in range(ncam):
        assert np.abs(pose[c, :3] - ref["pose"][c, :3]).max() <= 1e-7
        q, r = pose[c, 3:], ref["pose"][c, 3:] / np.linalg.norm(ref["pose"][c, 3:])
        assert min(np.abs(q - r).max(), np.abs(q + r).max()) <= 1e-7
    alive = invd >= 0
    assert alive.mean() > 0.9
    assert np.abs(invd[alive] - ref["lm_invdepth"][alive]).max() <= 1e-6 * np.abs(ref["lm_invdepth"][alive]).max() + 1e-7


# ---------------------------------------------------------------------------------------------- Mapper::matchToMap drop-in
def _write_match_scene(path, sc):
    hd = np.array([len(sc["kp_px"]), len(sc["mp_xyz"]), len(sc["desc"]), len(sc["obs_kf"]), len(sc["kf_Tcw"]), len(sc["cand_mp"]),
                   len(sc["cell_ptr"]) - 1, sc["nbwcells"], sc["ncellsize"], sc["img_w"], sc["img_h"], sc["dist"] is not None], np.int32)
    dist = np.zeros(5) if sc["dist"] is None else np.asarray(sc["dist"], np.float64)
    with open(path, "wb") as f:
        f.write(hd.tobytes())
        for a, t in ((sc["K"], np.float64), (dist, np.float64), (sc["Tcw"], np.float64), (sc["kf_Tcw"], np.float64), (sc["mp_xyz"], np.float64),
                     (sc["kp_px"], np.float32), (sc["obs_px"], np.float32), (sc["cell_ptr"], np.int32), (sc["cell_kp"], np.int32),
                     (sc["kp_lm"], np.int32), (sc["mp_desc_ptr"], np.int32), (sc["mp_obs_ptr"], np.int32), (sc["obs_kf"], np.int32),
                     (sc["cand_mp"], np.int32), (sc["desc"], np.uint8)):
            f.write(np.ascontiguousarray(a, t).tobytes())


def _read_match_result(path):
    raw = np.fromfile(path, np.int32)
    n = int(raw[0])
    order = raw[1:1 + n]
    m = int(raw[1 + n])
    pairs = raw[2 + n:2 + n + 2 * m].reshape(m, 2)
    return order, {(int(a), int(b)) for a, b in pairs}, int(raw[2 + n + 2 * m])


def _read_match_dump(path):
    raw = np.fromfile(path, np.uint8)
    hd = raw[:48].view(np.int32)
    nkps, nmps, ndesc, nobs, nkfs, ncand, ncells, nbw, cs, w, h, has_dist = (int(v) for v in hd)
    fl = raw[48:60].view(np.float32)
    off = [60]

    def take(n, t):
        b = n * np.dtype(t).itemsize
        a = raw[off[0]:off[0] + b].view(t).copy()
        off[0] += b
        return a
    sc = dict(img_w=w, img_h=h, ncellsize=cs, nbwcells=nbw, dmaxpxdist=fl[0], fdistratio=fl[1], view_th=fl[2])
    sc["K"] = take(4, np.float64)
    d = take(5, np.float64)
    sc["dist"] = d if has_dist else None
    sc["Tcw"] = take(12, np.float64)
    sc["kf_Tcw"] = take(12 * nkfs, np.float64).reshape(nkfs, 12)
    sc["mp_xyz"] = take(3 * nmps, np.float64).reshape(nmps, 3)
    sc["kp_px"] = take(2 * nkps, np.float32).reshape(nkps, 2)
    sc["obs_px"] = take(2 * nobs, np.float32).reshape(nobs, 2)
    sc["cell_ptr"] = take(ncells + 1, np.int32)
    sc["cell_kp"] = take(int(sc["cell_ptr"][-1]), np.int32)
    sc["kp_lm"] = take(nkps, np.int32)
    sc["mp_desc_ptr"] = take(nmps + 1, np.int32)
    sc["mp_obs_ptr"] = take(nmps + 1, np.int32)
    sc["obs_kf"] = take(nobs, np.int32)
    sc["cand_mp"] = take(ncand, np.int32)
    sc["mp_kfmask"] = take(4 * nmps, np.uint64).reshape(nmps, 4)
    sc["desc"] = take(32 * ndesc, np.uint8).reshape(ndesc, 32)
    assert off[0] == len(raw)
    return sc


def _match_pairs_by_identity(sc, kp_match):
    """(keypoint, map point) pairs named by what they ARE (pixel + own map point's position; candidate's position), so that two
    flattenings of the same map with different index orders can be compared."""
    out = set()
    for j, c in enumerate(kp_match):
        if c >= 0:
            out.add((tuple(sc["kp_px"][j]), tuple(sc["mp_xyz"][sc["kp_lm"][j]]), tuple(sc["mp_xyz"][sc["cand_mp"][c]])))
    return out


@pytest.mark.parametrize("distorted", [True, False])
def test_mapper_match_shim_flattens_the_map_like_the_scene(tmp_path, distorted):
    """The drop-in Mapper::matchToMap (host/mapper_match_gpu.cpp, against the stand-in map classes) on a synthetic map, with the
    GPU call replaced by a recorder (tests/helpers/match_dump.c, LD_PRELOAD): the flat problem it builds from the frame grid, the
    keyframes and the map points gives, through the oracle, the matches of the scene the map was built from (candidates in the
    order the shim walked the id set); map points without descriptor are filtered, keypoints whose map point is gone are dropped
    from the map (mapper.cpp:667-672), the search radius doubles for a frame with fewer than 30 3-D keypoints."""
    from oracle import match_ref as M
    exe = build.build_mapper_shim()
    rec = tmp_path / "librec.so"
    subprocess.check_call(["gcc", "-shared", "-fPIC", "-O1", "-o", str(rec), str(ROOT / "tests" / "helpers" / "match_dump.c")])
    sc = synth.make_match_scene(5 + distorted, nkps=500, ncand=260, nkfs=9, distorted=distorted)
    _write_match_scene(tmp_path / "s.bin", sc)
    env = dict(os.environ, LD_PRELOAD=str(rec), OV2_MATCH_DUMP=str(tmp_path / "d.bin"))
    out = subprocess.run([str(exe), str(tmp_path / "s.bin"), str(tmp_path / "r.bin")], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    order, pairs, ndropped = _read_match_result(tmp_path / "r.bin")
    assert not pairs and ndropped == int((sc["kp_lm"] < 0).sum())
    dump = _read_match_dump(tmp_path / "d.bin")
    assert dump["dmaxpxdist"] == sc["dmaxpxdist"] and dump["fdistratio"] == sc["fdistratio"]
    assert abs(float(dump["view_th"]) - float(sc["view_th"])) <= 1e-6
    ndesc_of = np.diff(sc["mp_desc_ptr"])
    assert len(dump["cand_mp"]) == int((ndesc_of[sc["cand_mp"]] > 0).sum())
    ref = dict(sc, cand_mp=np.asarray(order, np.int32))
    want = _match_pairs_by_identity(ref, M.match_to_map(ref)[2])
    got = _match_pairs_by_identity(dump, M.match_to_map(dump)[2])
    assert len(want) > 40
    # poses travel through unit quaternions (as in the reference's map): last-bit differences in the projections
    assert len(want ^ got) <= max(1, len(want) // 100)


@pytest.mark.gpu
def test_mapper_match_shim_matches_python_binding(ctx, tmp_path):
    """The same drop-in end to end on the GPU: the (keypoint id -> map-point id) map it returns equals ov2_match_to_map on the scene
    the map was built from, candidates in the order the shim walked the id set."""
    exe = build.build_mapper_shim()
    sc = synth.make_match_scene(9, nkps=1200, ncand=700, nkfs=12, distorted=True)
    _write_match_scene(tmp_path / "s.bin", sc)
    out = subprocess.run([str(exe), str(tmp_path / "s.bin"), str(tmp_path / "r.bin")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    order, pairs, _ = _read_match_result(tmp_path / "r.bin")
    ref = dict(sc, cand_mp=np.asarray(order, np.int32))
    _, _, kp_match, _ = api.match_to_map(ctx, ref)
    want = {(int(sc["kp_lm"][j]), int(order[c])) for j, c in enumerate(kp_match) if c >= 0}
    assert len(want) > 100
    assert len(want ^ pairs) <= max(1, len(want) // 100)

"""Draft for round 2 (NOT a bench.py line): per-eye front-end step at BASELINE.json configs[3] geometry -
1280x720 frames, parameters_files/accurate (CLAHE on, single-scale detector, cell 35) - timed on the GPU
through the Python binding with device-resident inputs, next to the cv2 reference sequence on the host.

    python scripts/bench_c4.py [--batch 64] [--steps 10] [--cpu-frames 16] [--no-gpu]

One step = CLAHE(prev), CLAHE(cur), pyramid(prev), pyramid(cur), fb-KLT (1024 keypoints / frame, nbpyrlvl 3),
detectSingleScale (+ cornerSubPix) with the tracked points as existing keypoints, descriptors of the new points.
The GPU arm has not run yet (written after the round-1 GPU budget was spent); the CPU arm runs anywhere cv2 does,
but its detector stage walks the cells in a Python loop (oracle/image_ref.py) - it is a functional reference, NOT
a fair CPU baseline for this detector (the C++ reference does the same work in compiled code)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ov2slam_b200 import synth  # noqa: E402

W, H, CELL, NKP, QUALITY = 1280, 720, 35, 1024, 0.001


def make_inputs(n, seed=4000):
    prevs, curs, kps, pris = [], [], [], []
    for i in range(n):
        p, c, flow = synth.make_pair(seed + i, W, H)
        rng = np.random.default_rng(seed + i)
        k = np.stack([rng.uniform(12, W - 12, NKP), rng.uniform(12, H - 12, NKP)], 1).astype(np.float32)
        _, pri = synth.make_priors(seed + i, k, flow, 0.0)
        prevs.append(p); curs.append(c); kps.append(k); pris.append(pri)
    return np.stack(prevs), np.stack(curs), np.stack(kps), np.stack(pris)


def cpu_arm(nframes):
    import cv2
    from oracle import image_ref as R
    cv2.setNumThreads(1)
    prevs, curs, kps, pris = make_inputs(nframes)
    clahe = cv2.createCLAHE(3.0, (W // 50, H // 50))
    t0 = time.perf_counter()
    for f in range(nframes):
        p, c = clahe.apply(prevs[f]), clahe.apply(curs[f])
        cv2.buildOpticalFlowPyramid(p, (9, 9), 3)
        cv2.buildOpticalFlowPyramid(c, (9, 9), 3)
        tr, st = R.fb_klt_cv2(p, c, kps[f], pris[f], 9, 3)
        new, _, _ = R.detect_single_scale_cv2(c, CELL, tr[st.astype(bool)], (0, 0, W, H), QUALITY)
        R.describe_cv2(c, new)
    dt = time.perf_counter() - t0
    return nframes / dt


def gpu_arm(batch, steps):
    import torch
    from ov2slam_b200 import api
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = api.Context(0, stream=stream.cuda_stream)
    prevs, curs, kps, pris = make_inputs(batch)
    d_prev, d_cur = torch.from_numpy(prevs).to(dev), torch.from_numpy(curs).to(dev)
    d_pe, d_ce = torch.empty_like(d_prev), torch.empty_like(d_cur)
    d_kps = torch.from_numpy(kps.reshape(-1, 2)).to(dev)
    d_pri0 = torch.from_numpy(pris.reshape(-1, 2)).to(dev)
    d_pri = d_pri0.clone()
    d_st = torch.empty(batch * NKP, dtype=torch.uint8, device=dev)
    ncell = (H // CELL) * (W // CELL)
    d_new = torch.empty((batch * ncell, 2), dtype=torch.float32, device=dev)
    d_cnt = torch.empty(batch, dtype=torch.int32, device=dev)
    d_q = torch.full((batch,), QUALITY, dtype=torch.float64, device=dev)
    d_off = torch.arange(0, (batch + 1) * NKP, NKP, dtype=torch.int32, device=dev)
    d_desc = torch.empty((batch * ncell, 32), dtype=torch.uint8, device=dev)
    d_val = torch.empty(batch * ncell, dtype=torch.uint8, device=dev)
    pp, pc = api.Pyramid(ctx, batch, W, H, 3), api.Pyramid(ctx, batch, W, H, 3)
    ft, fe = api.FeatureTracker(ctx, 30, 0.01), api.FeatureExtractor(ctx, nmaxdist=CELL, dmaxquality=QUALITY)

    def step():
        d_pri.copy_(d_pri0)
        d_q.fill_(QUALITY)
        api.clahe(ctx, d_prev, d_pe, W, H, batch)
        api.clahe(ctx, d_cur, d_ce, W, H, batch)
        pp.build(d_pe)
        pc.build(d_ce)
        ft.fb_klt_tracking(pp, pc, 9, 3, 30.0, 0.5, d_kps, d_pri, d_st, per_frame=NKP)
        # existing keypoints = all tracked priors (the status filter is a host-side compaction in the reference)
        fe.detect_single_scale(pc, CELL, 0, batch, d_q, d_new, d_cnt, d_off, d_pri)
        fe.describe_brief(pc, d_new, d_desc, d_val, n=batch * ncell, per_frame=ncell)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    return batch * steps / (e0.elapsed_time(e1) / 1e3)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--cpu-frames", type=int, default=8)
    ap.add_argument("--no-gpu", action="store_true")
    a = ap.parse_args()
    out = {"workload": "C4 per-eye front-end draft: 1280x720, CLAHE, single-scale detector cell 35, 1024 keypoints/frame",
           "cpu_frames_per_s_single_core": cpu_arm(a.cpu_frames)}
    if not a.no_gpu:
        out["gpu_frames_per_s_resident"] = gpu_arm(a.batch, a.steps)
    print(json.dumps(out))

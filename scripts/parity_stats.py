"""Observed parity margins of the float-tolerance operators (fb-KLT, cornerSubPix) on the GPU vs cv2 and vs the numpy
restatement: fraction of bit-equal results, max / percentiles of |delta| - the numbers DESIGN.md quotes next to the
test tolerances.  Prints one JSON line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ov2slam_b200 import api, synth  # noqa: E402
from oracle import image_ref as R  # noqa: E402

ctx = api.Context(0)
out = {}
w, h = 640, 480
dk, ds, dr = [], [], []
neq = ntot = 0
st_mis = st_mis_ref = 0
for seed in range(40, 52):
    prev, cur, flow = synth.make_pair(seed, w, h)
    pp, cp = api.Pyramid(ctx, 1, w, h, 3), api.Pyramid(ctx, 1, w, h, 3)
    pp.build(prev[None]); cp.build(cur[None])
    rng = np.random.default_rng(seed)
    kps = np.stack([rng.uniform(3, w - 3, 1000), rng.uniform(3, h - 3, 1000)], 1).astype(np.float32)
    for lvl in (0, 1, 3):
        _, pri = synth.make_priors(seed + lvl, kps, flow, 0.6 if lvl else 1.0)
        got = pri.copy(); st = np.zeros(len(kps), np.uint8)
        api.FeatureTracker(ctx, 30, 0.01).fb_klt_tracking(pp, cp, 9, lvl, 30.0, 0.5, kps, got, st)
        ref, rs = (R.fb_klt_cv2 if R.HAVE_CV2 else R.fb_klt_ref)(prev, cur, kps, pri, 9, lvl)
        st_mis += int((st != rs).sum())
        ok = (st == 1) & (rs == 1)
        d = np.abs(got[ok] - ref[ok]).max(axis=1)
        dk.append(d); neq += int((d == 0).sum()); ntot += len(d)
        if R.HAVE_CV2 and seed < 43:
            # the exact-integer restatement (oracle/image_ref.py::fb_klt_ref): what the kernel implements
            sub = slice(0, 400)
            ref2, rs2 = R.fb_klt_ref(prev, cur, kps[sub], pri[sub], 9, lvl)
            st_mis_ref += int((st[sub] != rs2).sum())
            ok2 = (st[sub] == 1) & (rs2 == 1)
            dr.append(np.abs(got[sub][ok2] - ref2[ok2]).max(axis=1))
    fe = api.FeatureExtractor(ctx, nfast_th=10)
    pts, ipts = fe.detect_grid_fast_frame(pp, 0, 16, np.zeros((0, 2), np.float32))
    rsp = (R.corner_subpix_cv2 if R.HAVE_CV2 else R.corner_subpix_ref)(prev, ipts.astype(np.float32))
    ds.append(np.abs(pts - rsp).max(axis=1))
    pp.close(); cp.close()
dk = np.concatenate(dk); ds = np.concatenate(ds)
out["klt"] = {"tracks": int(ntot), "bit_equal_frac": neq / max(ntot, 1), "status_mismatches": st_mis, "max": float(dk.max()),
              "p99": float(np.percentile(dk, 99)), "p999": float(np.percentile(dk, 99.9)), "n_gt_1e-5": int((dk > 1e-5).sum()),
              "n_gt_1e-4": int((dk > 1e-4).sum()), "reference": "cv2" if R.HAVE_CV2 else "numpy restatement"}
if dr:
    dr = np.concatenate(dr)
    out["klt_vs_exact_integer_restatement"] = {"tracks": int(len(dr)), "bit_equal_frac": float((dr == 0).mean()), "max": float(dr.max()),
                                                "status_mismatches": st_mis_ref}
out["subpix"] = {"points": int(len(ds)), "bit_equal_frac": float((ds == 0).mean()), "max": float(ds.max()), "p99": float(np.percentile(ds, 99)),
                 "n_gt_1e-5": int((ds > 1e-5).sum()), "n_gt_1e-4": int((ds > 1e-4).sum())}
print(json.dumps(out))

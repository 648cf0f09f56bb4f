"""Generate tests/golden/ba_stereo_150k.npz: the numpy oracle's (oracle/ba_ref.py) result on a stereo localBA window with
150 000 residual blocks (50 KF x 10 000 landmarks x 70 000 left-camera observations + their right-camera twins + the
anchor-frame right-camera blocks), BASELINE.json configs[4] size.  The oracle takes about a minute here, so the GPU parity
test (tests/test_ba_gpu.py::test_localba_stereo_150k_blocks_matches_golden) compares against this committed fixture.

    python scripts/make_golden_ba.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ov2slam_b200 import synth  # noqa: E402
from oracle import ba_ref as B  # noqa: E402

SEED, NCAM, NPTS, NOBS = 45, 50, 10000, 70000

if __name__ == "__main__":
    pb = synth.make_ba_problem(SEED, NCAM, NPTS, NOBS, stereo=True)
    assert len(pb["obs_cam"]) == NPTS + 2 * NOBS == 150000
    r = B.local_ba(pb)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ba_stereo_150k.npz")
    np.savez_compressed(out, seed=SEED, ncam=NCAM, npts=NPTS, nobs=NOBS, pose=pb["pose"], lm_invdepth=pb["lm_invdepth"],
                        flags=np.packbits(r["flags"] & 1), flags2=np.packbits((r["flags"] >> 1) & 1),
                        iters=np.array([r["iters_robust"], r["iters_refine"]]), costs=np.array([r["initial_cost"], r["final_cost"]]),
                        n_outliers=np.array([r["n_outliers_first"], r["n_outliers_second"]]),
                        termination=np.array([{"CONVERGENCE": 0, "NO_CONVERGENCE": 1, "FAILURE": 2}[r["termination"]]]))
    print("wrote", out, {k: v for k, v in r.items() if k not in ("flags", "summaries")})

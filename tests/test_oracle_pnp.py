"""CPU tests of the oracle for MultiViewGeometry::ceresPnP ("next" row 2): Jacobian by central
differences, ground-truth recovery, the two-stage outlier flow, agreement with scipy on the L2 cost.
Parity against the real Ceres is unpinned (it cannot be built here) - see oracle/pnp_ref.py."""
import numpy as np
import pytest

from oracle import ba_ref, pnp_ref

K = (458.0, 457.0, 367.0, 248.0)


def _scene(seed, n=200, noise=0.0, outliers=0):
    rng = np.random.default_rng(seed)
    axis = rng.standard_normal(3)
    axis /= np.linalg.norm(axis)
    ang = 0.3 * rng.random()
    q = np.concatenate([axis * np.sin(ang / 2), [np.cos(ang / 2)]])
    pose = np.concatenate([rng.standard_normal(3) * 0.5, q])
    R = ba_ref.quat_to_rot(q)
    pc = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(2, 10, n)], 1)
    wpts = pc @ R.T + pose[:3]                                  # Twc * pc
    px = np.stack([K[0] * pc[:, 0] / pc[:, 2] + K[2], K[1] * pc[:, 1] / pc[:, 2] + K[3]], 1)
    px += rng.standard_normal(px.shape) * noise
    bad = rng.choice(n, outliers, replace=False) if outliers else np.zeros(0, int)
    px[bad] += rng.uniform(15, 60, (len(bad), 2)) * rng.choice([-1, 1], (len(bad), 2))
    return pose, wpts, px, np.sort(bad)


def _perturb(pose, seed, t=0.05, r=0.02):
    rng = np.random.default_rng(seed)
    return ba_ref.pose_plus(pose, np.concatenate([rng.standard_normal(3) * t, rng.standard_normal(3) * r]))


def test_pnp_jacobian_matches_central_differences():
    pose, wpts, px, _ = _scene(1, 40)
    pose = _perturb(pose, 2)
    ev = pnp_ref.evaluate(pose, wpts, px, K, None, True)
    h = 1e-6
    for k in range(6):
        d = np.zeros(6)
        d[k] = h
        rp = pnp_ref.evaluate(ba_ref.pose_plus(pose, d), wpts, px, K, None, False)["r"]
        rm = pnp_ref.evaluate(ba_ref.pose_plus(pose, -d), wpts, px, K, None, False)["r"]
        num = (rp - rm) / (2 * h)
        assert np.abs(num - ev["J"][:, :, k]).max() <= 1e-5 * max(1.0, np.abs(num).max())


def test_pnp_scale_is_the_square_root_information():
    pose, wpts, px, _ = _scene(3, 20, noise=1.0)
    e0 = pnp_ref.evaluate(pose, wpts, px, K, None, True)
    e1 = pnp_ref.evaluate(pose, wpts, px, K, np.full(20, 2), True)      # sigma = 2^2: residual and Jacobian / 4
    assert np.allclose(e1["r"], e0["r"] / 4) and np.allclose(e1["J"], e0["J"] / 4)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pnp_recovers_ground_truth_without_noise(seed):
    pose, wpts, px, _ = _scene(10 + seed)
    ok, est, out = pnp_ref.ceres_pnp(px, wpts, _perturb(pose, seed), K, nmaxiter=5)
    assert ok and len(out) == 0
    assert np.abs(est[:3] - pose[:3]).max() <= 1e-6
    assert min(np.abs(est[3:] - pose[3:]).max(), np.abs(est[3:] + pose[3:]).max()) <= 1e-7


def test_pnp_two_stage_outlier_flow():
    pose, wpts, px, bad = _scene(20, n=300, noise=0.5, outliers=45)
    ok, est, out = pnp_ref.ceres_pnp(px, wpts, _perturb(pose, 5), K, nmaxiter=5, apply_l2_after_robust=True)
    assert ok
    assert set(bad) <= set(out) and len(out) <= len(bad) + 6          # gross outliers flagged, few 2-sigma inliers
    assert np.abs(est[:3] - pose[:3]).max() <= 5e-3
    # without the refinement stage the robust estimate is kept, same outlier list
    ok2, est2, out2 = pnp_ref.ceres_pnp(px, wpts, _perturb(pose, 5), K, nmaxiter=5, apply_l2_after_robust=False)
    assert ok2 and np.array_equal(out, out2)
    assert np.abs(est2[:3] - pose[:3]).max() <= 2e-2


def test_pnp_all_outliers_returns_false_and_keeps_the_pose():
    pose, wpts, px, _ = _scene(30, n=30)
    px = px + 500.0
    start = _perturb(pose, 1)
    ok, est, out = pnp_ref.ceres_pnp(px, wpts, start, K, nmaxiter=5)
    assert not ok and len(out) == 30 and np.array_equal(est, start)


def test_pnp_l2_minimiser_agrees_with_scipy():
    scipy_opt = pytest.importorskip("scipy.optimize")
    pose, wpts, px, _ = _scene(40, n=120, noise=0.7)
    start = _perturb(pose, 3)
    est, summ, _ = pnp_ref.ceres_solve_pose(start, wpts, px, K, np.ones(120, bool), 50, None, function_tolerance=1e-14)

    def f(d):
        return pnp_ref.evaluate(ba_ref.pose_plus(start, d), wpts, px, K, None, False)["r"].ravel()

    sol = scipy_opt.least_squares(f, np.zeros(6), method="lm", xtol=1e-14, ftol=1e-14, gtol=1e-14)
    ref = ba_ref.pose_plus(start, sol.x)
    assert abs(summ["final_cost"] - sol.cost) <= 1e-9 * sol.cost
    assert np.abs(est[:3] - ref[:3]).max() <= 1e-7

"""CPU ORACLE for Optimizer::localBA's solve path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (ov2slam_b200/) never does.

numpy float64 restatement of what the reference executes between problem setup and write-back:

  residual + analytic Jacobians   /root/reference/src/ceres_parametrization.cpp:361-473
                                  (DirectLeftSE3::ReprojectionErrorKSE3AnchInvDepth::Evaluate)
  pose update                     include/ceres_parametrization/.../se3left_parametrization.hpp:41-60
                                  (T+ = Sophus::SE3d::exp(delta) * T; Sophus se3.hpp:763-784,
                                   so3.hpp:585-621, quaternion product so3.hpp:329-343)
  robust loss + corrector         Ceres 2.0.0 loss_function.cc:48-62 (HuberLoss),
                                  corrector.cc:42-156, residual_block.cc:161-197
  trust-region loop               trust_region_minimizer.cc:67-134 (+ :244-301, :377-448, :706-826)
  LM strategy                     levenberg_marquardt_strategy.cc:66-160
  step quality                    trust_region_step_evaluator.cc:52-68
  Schur elimination / back-sub    schur_eliminator_impl.h:179-377, schur_complement_solver.cc:118-176
  unused / constant blocks        program.cc:305-387 (RemoveFixedBlocks)
  two-stage solve + outlier scan  /root/reference/src/optimizer.cpp:436-479, :492-594, :603-627, :637-735

PINNING STATUS.  Pinned against the reference's own code (tests/test_oracle_vs_reference_source.py, test_oracle_vs_reference_ceres.py):
  * residuals, Jacobians, chi2 / depth flags and the pose update against /root/reference/src/ceres_parametrization.cpp compiled in
    place with the reference tree's own Sophus headers (oracle/ref_build/build_ref.py), 1.3e-13 relative;
  * the WHOLE two-stage solve - every iteration's cost, cost change, step norm, relative decrease, radius, gradient max-norm and
    accept / reject decision, termination, both outlier scans, final state - against the Ceres 2.0 vendored in the reference tree,
    compiled in place and driven with the Ceres calls Optimizer::localBA makes (oracle/ref_build/build_ceres_ref.py, ceres_ba_ref.cpp):
    mono, stereo, C3-size, BAL-structure and ill-conditioned windows with rejected steps; final states equal to ~1e-15.
What that real-Ceres build stands on: a stand-in linear-algebra header instead of Eigen (this container has none; products,
Cholesky and QR are summed in index order, so results equal real Eigen's only to rounding), DENSE_SCHUR where the shipped
configurations select SPARSE_SCHUR (`use_sparse_schur: 1`: the same eliminator code, the reduced system factorised densely instead of
sparsely), aborting stubs for three source files off that path, and no wall-clock cap.  Also pinned piecewise (tests/test_oracle_ba.py):
  * HuberLoss and the Corrector against Ceres' own known answers (loss_function_test.cc:92-103,
    corrector_test.cc:57-135),
  * the LM radius rules against levenberg_marquardt_strategy_test.cc,
  * the analytic Jacobians against central differences,
  * the Schur path against a dense solve of the full normal equations,
  * convergence to ground truth on noise-free synthetic windows.
"""
from __future__ import annotations

import numpy as np

DBL_MAX = np.finfo(np.float64).max
SOPHUS_EPS = 1e-10  # Sophus::Constants<double>::epsilon()


# ------------------------------------------------------------------ SE3 helpers (xyzw storage)
def quat_normalize(q):
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def quat_mul(a, b):
    """Sophus SO3 product formula (so3.hpp:338-342); a, b = [x, y, z, w]."""
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def quat_to_rot(q):
    """Rotation matrix of a unit quaternion [x, y, z, w] (Eigen toRotationMatrix)."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - (tyy + tzz)
    R[..., 0, 1] = txy - twz
    R[..., 0, 2] = txz + twy
    R[..., 1, 0] = txy + twz
    R[..., 1, 1] = 1 - (txx + tzz)
    R[..., 1, 2] = tyz - twx
    R[..., 2, 0] = txz - twy
    R[..., 2, 1] = tyz + twx
    R[..., 2, 2] = 1 - (txx + tyy)
    return R


def hat(v):
    z = np.zeros(v.shape[:-1])
    return np.stack([np.stack([z, -v[..., 2], v[..., 1]], -1),
                     np.stack([v[..., 2], z, -v[..., 0]], -1),
                     np.stack([-v[..., 1], v[..., 0], z], -1)], -2)


def se3_exp(delta):
    """Sophus::SE3d::exp for delta = [upsilon(3), omega(3)] -> (t[3], q[xyzw]); vectorised."""
    ups, om = delta[..., :3], delta[..., 3:]
    th2 = (om * om).sum(-1)
    small = th2 < SOPHUS_EPS * SOPHUS_EPS
    th = np.where(small, 0.0, np.sqrt(th2))
    ths = np.where(small, 1.0, th)
    imag = np.where(small, 0.5 - th2 / 48.0 + th2 * th2 / 3840.0, np.sin(0.5 * ths) / ths)
    real = np.where(small, 1.0 - th2 / 8.0 + th2 * th2 / 384.0, np.cos(0.5 * ths))
    q = np.concatenate([imag[..., None] * om, real[..., None]], -1)
    Om = hat(om)
    Om2 = Om @ Om
    I = np.broadcast_to(np.eye(3), Om.shape)
    smallV = th < SOPHUS_EPS
    a = np.where(smallV, 0.0, (1 - np.cos(ths)) / (ths * ths))
    b = np.where(smallV, 0.0, (ths - np.sin(ths)) / (ths * ths * ths))
    V = np.where(smallV[..., None, None], quat_to_rot(q), I + a[..., None, None] * Om + b[..., None, None] * Om2)
    t = (V @ ups[..., None])[..., 0]
    return t, q


def pose_plus(pose, delta):
    """SE3LeftParameterization::Plus on [t, q(xyzw)] poses: exp(delta) * T."""
    t, q = pose[..., :3], quat_normalize(pose[..., 3:])          # SE3d(q, t) normalises q
    et, eq = se3_exp(delta)
    qn = quat_normalize(quat_mul(eq, q))                          # SO3 product normalises
    tn = et + (quat_to_rot(eq) @ t[..., None])[..., 0]
    return np.concatenate([tn, qn], -1)


# ------------------------------------------------------------------ loss (Ceres)
def huber(s, a):
    """HuberLoss(a)::Evaluate: rho = [rho, rho', rho''] at s = |r|^2 (loss_function.cc:48-62)."""
    b = a * a
    s = np.asarray(s, np.float64)
    out = np.empty((3,) + s.shape)
    big = s > b
    r = np.sqrt(np.where(big, s, 1.0))
    out[0] = np.where(big, 2 * a * r - b, s)
    out[1] = np.where(big, np.maximum(np.finfo(float).tiny, a / r), 1.0)
    out[2] = np.where(big, -out[1] / (2 * np.where(big, s, 1.0)), 0.0)
    return out


def corrector(sq_norm, rho):
    """Corrector::Corrector (corrector.cc:42-111): returns (sqrt_rho1, residual_scaling, alpha_sq_norm)."""
    sqrt_rho1 = np.sqrt(rho[1])
    if sq_norm == 0.0 or rho[2] <= 0.0:
        return sqrt_rho1, sqrt_rho1, 0.0
    D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1]
    alpha = 1.0 - np.sqrt(D)
    return sqrt_rho1, sqrt_rho1 / (1 - alpha), alpha / sq_norm


# ------------------------------------------------------------------ LM strategy scalars
def lm_step_accepted(radius, step_quality, max_radius=1e16):
    radius = radius / max(1.0 / 3.0, 1.0 - (2.0 * step_quality - 1.0) ** 3)
    return min(max_radius, radius)


def lm_step_rejected(radius, decrease_factor):
    return radius / decrease_factor, decrease_factor * 2.0


# ------------------------------------------------------------------ residuals
def evaluate(pb, pose, invd, idx, need_jac=True):
    """All residual blocks `idx` at (pose, invd).  Returns dict with r (n,2), chi2, depth_pos and,
    if need_jac, Ja (n,2,6) anchor, Jo (n,2,6) observer, Jl (n,2) inverse depth.

    obs_type (optional, default all 0):
      0  left-camera observation in another keyframe    ReprojectionErrorKSE3AnchInvDepth        (:361-473)
      1  right-camera observation in another keyframe   ReprojectionErrorRightCamKSE3AnchInvDepth (:579-712)
      2  right-camera observation in the anchor frame   ReprojectionErrorRightAnchCamKSE3AnchInvDepth (:476-577)
         (depends on the inverse depth only: calibrations and the extrinsic are constant blocks)
    """
    fx, fy, cx, cy = pb["K"]
    lm = pb["obs_lm"][idx]
    ca = pb["lm_anchor_cam"][lm]
    co = pb["obs_cam"][idx]
    typ = pb["obs_type"][idx] if pb.get("obs_type") is not None else np.zeros(len(idx), np.uint8)
    n = len(lm)
    ta, qa = pose[ca, :3], quat_normalize(pose[ca, 3:])
    to, qo = pose[co, :3], quat_normalize(pose[co, 3:])
    Rwa, Rwc = quat_to_rot(qa), quat_to_rot(qo)
    zanch = 1.0 / invd[lm]
    ua = pb["lm_anchor_px"][lm]
    # invK * [u, v, 1]
    bearing = np.stack([(ua[:, 0] - cx) / fx, (ua[:, 1] - cy) / fy, np.ones(n)], -1)
    anchpt = zanch[:, None] * bearing
    rot_ap = (Rwa @ anchpt[..., None])[..., 0]
    wpt = rot_ap + ta
    Rcw = np.swapaxes(Rwc, -1, -2)
    lcam = (Rcw @ (wpt - to)[..., None])[..., 0]
    right = typ > 0
    campt = lcam.copy()
    kfx, kfy, kcx, kcy = (np.full(n, v) for v in (fx, fy, cx, cy))
    M = Rcw.copy()                      # d(campt)/d(wpt)
    if right.any():
        rfx, rfy, rcx, rcy = pb["Kr"]
        trl = np.asarray(pb["Trl"], np.float64)
        Rrl = quat_to_rot(quat_normalize(trl[3:]))
        src = np.where((typ == 2)[:, None], anchpt, lcam)          # anchor-frame right obs: Trl * anchpt
        rc = (Rrl @ src[..., None])[..., 0] + trl[:3]
        campt = np.where(right[:, None], rc, campt)
        kfx, kfy = np.where(right, rfx, kfx), np.where(right, rfy, kfy)
        kcx, kcy = np.where(right, rcx, kcx), np.where(right, rcy, kcy)
        M = np.where((typ == 1)[:, None, None], Rrl @ Rcw, M)
        M = np.where((typ == 2)[:, None, None], np.broadcast_to(Rrl, M.shape), M)
    linvz = 1.0 / campt[:, 2]
    pred = np.stack([kfx * campt[:, 0] * linvz + kcx, kfy * campt[:, 1] * linvz + kcy], -1)
    r = pred - pb["obs_px"][idx]
    out = dict(r=r, chi2=(r * r).sum(-1), depth_pos=campt[:, 2] > 0, lm=lm, ca=ca, co=co, typ=typ)
    if need_jac:
        linvz2 = linvz * linvz
        Jc = np.zeros((n, 2, 3))
        Jc[:, 0, 0] = linvz * kfx
        Jc[:, 0, 2] = -campt[:, 0] * linvz2 * kfx
        Jc[:, 1, 1] = linvz * kfy
        Jc[:, 1, 2] = -campt[:, 1] * linvz2 * kfy
        JR = Jc @ M
        sk = hat(wpt)
        JRs = JR @ sk
        pose_dep = (typ != 2)[:, None, None]
        out["Ja"] = np.where(pose_dep, np.concatenate([JR, -JRs], -1), 0.0)
        out["Jo"] = np.where(pose_dep, np.concatenate([-JR, JRs], -1), 0.0)
        Jlam = -zanch[:, None] * np.where((typ == 2)[:, None], anchpt, rot_ap)
        out["Jl"] = (JR @ Jlam[..., None])[..., 0]
    return out


# ------------------------------------------------------------------ one ceres::Solve
class SolveSummary(dict):
    pass


def ceres_solve(pb, pose, invd, active, max_iters, huber_a, function_tolerance=1e-3, log=None):
    """ceres::Solve on the residual blocks `active` (bool[nobs]) with Ceres 2.0 defaults as set by
    optimizer.cpp:436-470 (LEVENBERG_MARQUARDT, *_SCHUR, jacobi scaling, eta 0.1, no time cap).
    huber_a = None -> trivial loss.  Returns (pose, invd, summary, last_eval) where last_eval holds
    chi2 / depth_pos of the LAST Evaluate() call on each active residual block (optimizer.cpp
    reads these mutable members afterwards)."""
    pose = pose.copy()
    invd = invd.copy()
    idx = np.nonzero(active)[0]
    nobs = len(idx)
    ncam, npts = len(pose), len(invd)
    summ = SolveSummary(iterations=0, successful=0, unsuccessful=0, termination="NO_CONVERGENCE",
                        initial_cost=0.0, final_cost=0.0, trace=[])
    last = dict(chi2=np.zeros(len(active)), depth_pos=np.ones(len(active), bool))
    if nobs == 0:
        summ["termination"] = "CONVERGENCE"
        return pose, invd, summ, last
    lm_i = pb["obs_lm"][idx]
    ca_i = pb["lm_anchor_cam"][lm_i]
    co_i = pb["obs_cam"][idx]
    const = pb["pose_const"].astype(bool)
    # Program::RemoveFixedBlocks: constant blocks and blocks no residual uses drop out
    typ_i = pb["obs_type"][idx] if pb.get("obs_type") is not None else np.zeros(nobs, np.uint8)
    pose_dep = typ_i != 2                  # anchor-frame right-camera residuals depend on lambda only
    cam_used = np.zeros(ncam, bool)
    cam_used[ca_i[pose_dep]] = True
    cam_used[co_i[pose_dep]] = True
    cam_var = cam_used & ~const
    lm_var = np.zeros(npts, bool)
    lm_var[lm_i] = True
    cam_slot = -np.ones(ncam, np.int64)
    cam_slot[cam_var] = np.arange(cam_var.sum())
    ncv = int(cam_var.sum())
    lm_slot = -np.ones(npts, np.int64)
    lm_slot[lm_var] = np.arange(lm_var.sum())
    nlv = int(lm_var.sum())
    sa, so, sl = cam_slot[ca_i], cam_slot[co_i], lm_slot[lm_i]
    sa = np.where(pose_dep, sa, -1)
    so = np.where(pose_dep, so, -1)
    ma, mo = sa >= 0, so >= 0

    def cost_and_corr(ev):
        s = ev["chi2"]
        if huber_a is None:
            return 0.5 * s.sum(), np.ones_like(s)
        rho = huber(s, huber_a)
        return 0.5 * rho[0].sum(), np.sqrt(rho[1])   # alpha = 0 branch always (rho'' <= 0)

    def record(ev):
        last["chi2"][idx] = ev["chi2"]
        last["depth_pos"][idx] = ev["depth_pos"]

    def eval_jac(pose, invd):
        ev = evaluate(pb, pose, invd, idx, True)
        record(ev)
        cost, w = cost_and_corr(ev)
        r = ev["r"] * w[:, None]
        Ja = ev["Ja"] * w[:, None, None]
        Jo = ev["Jo"] * w[:, None, None]
        Jl = ev["Jl"] * w[:, None]
        return cost, r, Ja, Jo, Jl

    def colnorm2(Ja, Jo, Jl):
        cn_c = np.zeros((ncv, 6))
        np.add.at(cn_c, sa[ma], (Ja[ma] ** 2).sum(1))
        np.add.at(cn_c, so[mo], (Jo[mo] ** 2).sum(1))
        cn_l = np.zeros(nlv)
        np.add.at(cn_l, sl, (Jl ** 2).sum(1))
        return cn_c, cn_l

    def x_norm(pose, invd):
        return np.sqrt((pose[cam_var] ** 2).sum() + (invd[lm_var] ** 2).sum())

    def grad_max_norm(pose, invd, g_c, g_l):
        # |x - Plus(x, -gradient)|_inf in the ambient space (trust_region_minimizer.cc:283-299)
        m = 0.0
        if ncv:
            p0 = pose[cam_var]
            p0n = np.concatenate([p0[:, :3], p0[:, 3:]], -1)
            p1 = pose_plus(p0, -g_c)
            m = max(m, np.abs(p0n - p1).max())
        if nlv:
            m = max(m, np.abs(g_l).max())
        return m

    # ---- iteration 0
    x_cost, r, Ja, Jo, Jl = eval_jac(pose, invd)
    cn_c, cn_l = colnorm2(Ja, Jo, Jl)
    sc_c = 1.0 / (1.0 + np.sqrt(cn_c))
    sc_l = 1.0 / (1.0 + np.sqrt(cn_l))

    def scale(Ja, Jo, Jl):
        scp = np.vstack([sc_c, np.zeros((1, 6))])          # slot ncv = "constant camera": zero columns
        Ja = Ja * scp[np.where(ma, sa, ncv)][:, None, :]
        Jo = Jo * scp[np.where(mo, so, ncv)][:, None, :]
        Jl = Jl * sc_l[sl][:, None]
        return Ja, Jo, Jl

    def gradient(r, Ja, Jo, Jl):
        g_c = np.zeros((ncv, 6))
        np.add.at(g_c, sa[ma], np.einsum("nij,ni->nj", Ja[ma], r[ma]))
        np.add.at(g_c, so[mo], np.einsum("nij,ni->nj", Jo[mo], r[mo]))
        g_l = np.zeros(nlv)
        np.add.at(g_l, sl, (Jl * r).sum(1))
        return g_c, g_l

    g_c, g_l = gradient(r, Ja, Jo, Jl)         # unscaled Jacobian
    gmax = grad_max_norm(pose, invd, g_c, g_l)
    Ja, Jo, Jl = scale(Ja, Jo, Jl)
    summ["initial_cost"] = x_cost
    minimum_cost = DBL_MAX
    xnorm = -1.0
    radius, decrease_factor, reuse_diag = 1e4, 2.0, False
    diag_c = diag_l = None
    best_pose, best_invd = pose.copy(), invd.copy()
    step_successful = True
    num_invalid = 0
    iteration = 0
    while True:
        # FinalizeIterationAndCheckIfMinimizerCanContinue
        if step_successful:
            summ["successful"] += 1
            if x_cost < minimum_cost:
                minimum_cost = x_cost
                best_pose, best_invd = pose.copy(), invd.copy()
        else:
            summ["unsuccessful"] += 1
        summ["trace"].append(dict(it=iteration, cost=x_cost, radius=radius, ok=step_successful))
        if iteration >= max_iters:
            summ["termination"] = "NO_CONVERGENCE"
            break
        if step_successful and gmax <= 1e-10:
            summ["termination"] = "CONVERGENCE"
            break
        if radius <= 1e-32:
            summ["termination"] = "CONVERGENCE"
            break
        iteration += 1
        # ---- ComputeTrustRegionStep (LevenbergMarquardtStrategy::ComputeStep + Schur solver)
        if not reuse_diag:
            dc, dl = colnorm2(Ja, Jo, Jl)
            diag_c = np.clip(dc, 1e-6, 1e32)
            diag_l = np.clip(dl, 1e-6, 1e32)
        D_c = np.sqrt(diag_c / radius)
        D_l = np.sqrt(diag_l / radius)
        reuse_diag = True
        ete = np.zeros(nlv)
        np.add.at(ete, sl, (Jl * Jl).sum(1))
        ete += D_l ** 2
        ge = np.zeros(nlv)
        np.add.at(ge, sl, (Jl * r).sum(1))
        S = np.zeros((ncv * 6, ncv * 6))
        rhs = np.zeros(ncv * 6)
        # F'F and F'r
        blocks = []   # (slot array, J (n,2,6), mask)
        for s_, J_, m_ in ((sa, Ja, ma), (so, Jo, mo)):
            blocks.append((s_, J_, m_))
        for s_, J_, m_ in blocks:
            FtF = np.einsum("nij,nik->njk", J_[m_], J_[m_])
            for k, b in zip(s_[m_], FtF):
                S[6 * k:6 * k + 6, 6 * k:6 * k + 6] += b
            Ftr = np.einsum("nij,ni->nj", J_[m_], r[m_])
            for k, b in zip(s_[m_], Ftr):
                rhs[6 * k:6 * k + 6] += b
        both = ma & mo
        cross = np.einsum("nij,nik->njk", Ja[both], Jo[both])
        for ka, ko, b in zip(sa[both], so[both], cross):
            S[6 * ka:6 * ka + 6, 6 * ko:6 * ko + 6] += b
            S[6 * ko:6 * ko + 6, 6 * ka:6 * ka + 6] += b.T
        for k in range(ncv):
            S[6 * k:6 * k + 6, 6 * k:6 * k + 6] += np.diag(D_c[k] ** 2)
        # E'F per landmark: dense (nlv, ncv*6) is fine at oracle sizes
        EtF = np.zeros((nlv, ncv * 6))
        for s_, J_, m_ in blocks:
            v = np.einsum("ni,nij->nj", Jl[m_], J_[m_])
            for l, k, b in zip(sl[m_], s_[m_], v):
                EtF[l, 6 * k:6 * k + 6] += b
        S -= EtF.T @ (EtF / ete[:, None])
        rhs -= EtF.T @ (ge / ete)
        step_valid = True
        try:
            L = np.linalg.cholesky(S) if ncv else None
            z = np.linalg.solve(L.T, np.linalg.solve(L, rhs)) if ncv else np.zeros(0)
        except np.linalg.LinAlgError:
            step_valid = False
            z = np.zeros(ncv * 6)
        y = (ge - EtF @ z) / ete
        if not (np.isfinite(z).all() and np.isfinite(y).all()):
            step_valid = False
        step_c = -z.reshape(ncv, 6)
        step_l = -y
        model_cost_change = 0.0
        if step_valid:
            stp = np.vstack([step_c, np.zeros((1, 6))])
            Jstep = (np.einsum("nij,nj->ni", Ja, stp[np.where(ma, sa, ncv)])
                     + np.einsum("nij,nj->ni", Jo, stp[np.where(mo, so, ncv)])
                     + Jl * step_l[sl][:, None])
            model_cost_change = -(Jstep * (r + Jstep / 2.0)).sum()
            step_valid = model_cost_change > 0.0
        if log is not None and not step_valid:
            log.append(dict(ctl=True, it=iteration, x_cost=x_cost, cand_cost=0.0, mcc=model_cost_change, step2=0.0, candx2=0.0,
                            gmax=gmax, invalid=True))
        if not step_valid:
            num_invalid += 1
            if num_invalid >= 5:
                summ["termination"] = "FAILURE"
                break
            radius, decrease_factor = lm_step_rejected(radius, decrease_factor)
            reuse_diag = True
            step_successful = False
            continue
        num_invalid = 0
        delta_c = step_c * sc_c
        delta_l = step_l * sc_l
        # ---- ComputeCandidatePointAndEvaluateCost
        cand_pose = pose.copy()
        cand_invd = invd.copy()
        if ncv:
            cand_pose[cam_var] = pose_plus(pose[cam_var], delta_c)
        cand_invd[lm_var] = invd[lm_var] + delta_l
        ev = evaluate(pb, cand_pose, cand_invd, idx, False)
        record(ev)
        cand_cost, _ = cost_and_corr(ev)
        if not np.isfinite(cand_cost):
            cand_cost = DBL_MAX
        # ---- ParameterToleranceReached / FunctionToleranceReached
        step_norm = np.sqrt(((pose[cam_var] - cand_pose[cam_var]) ** 2).sum() +
                            ((invd[lm_var] - cand_invd[lm_var]) ** 2).sum())
        if log is not None:
            # everything the trust-region controller consumes this iteration (tests/test_host_logic.py replays the device
            # controller, csrc/ba_lm_ctl.cuh, on these records)
            log.append(dict(ctl=True, it=iteration, x_cost=x_cost, cand_cost=cand_cost, mcc=model_cost_change, step2=step_norm ** 2,
                            candx2=x_norm(cand_pose, cand_invd) ** 2, gmax=gmax, invalid=False))
        if step_norm <= 1e-8 * (xnorm + 1e-8):
            summ["termination"] = "CONVERGENCE"
            break
        cost_change = x_cost - cand_cost
        if abs(cost_change) <= function_tolerance * x_cost:
            summ["termination"] = "CONVERGENCE"
            break
        # ---- IsStepSuccessful (max_consecutive_nonmonotonic_steps = 0 => plain ratio)
        rel = -DBL_MAX if cand_cost >= DBL_MAX else (x_cost - cand_cost) / model_cost_change
        if log is not None:
            log.append(dict(it=iteration, x_cost=x_cost, cand_cost=cand_cost, mcc=model_cost_change, rho=rel,
                            radius=radius, step=np.concatenate([delta_c.ravel(), delta_l])))
        if rel > 1e-3:
            pose, invd = cand_pose, cand_invd
            xnorm = x_norm(pose, invd)
            x_cost, r, Ja, Jo, Jl = eval_jac(pose, invd)
            g_c, g_l = gradient(r, Ja, Jo, Jl)
            gmax = grad_max_norm(pose, invd, g_c, g_l)
            Ja, Jo, Jl = scale(Ja, Jo, Jl)
            step_successful = True
            radius = lm_step_accepted(radius, rel)
            decrease_factor = 2.0
            reuse_diag = False
        else:
            step_successful = False
            radius, decrease_factor = lm_step_rejected(radius, decrease_factor)
            reuse_diag = True
    summ["iterations"] = iteration
    summ["final_cost"] = minimum_cost if minimum_cost < DBL_MAX else x_cost
    return best_pose, best_invd, summ, last


# ------------------------------------------------------------------ Optimizer::localBA solve part
def local_ba(pb, max_iters_robust=5, max_iters_refine=10, huber_th=5.9915, function_tolerance=1e-3,
             use_robust=True, apply_l2_after_robust=True, log=None, refine_trivial_loss=None):
    """Solve section of Optimizer::localBA (optimizer.cpp:436-735), mono residual blocks.
    Updates pb["pose"], pb["lm_invdepth"] in place.  Returns a result dict incl. `flags`
    (uint8[nobs]: bit0 = outlier after solve #1, bit1 = after solve #2)."""
    th_f = np.float32(huber_th)                     # const float mono_th (optimizer.cpp:47)
    a = float(np.sqrt(th_f))                         # HuberLoss(std::sqrt(mono_th)): float sqrt
    th = float(th_f)
    nobs = len(pb["obs_cam"])
    active = np.ones(nobs, bool)
    flags = np.zeros(nobs, np.uint8)
    pose, invd, s1, last = ceres_solve(pb, pb["pose"], pb["lm_invdepth"], active, max_iters_robust,
                                       a if use_robust else None, function_tolerance, log)
    bad1 = active & ((last["chi2"] > th) | ~last["depth_pos"])
    flags[bad1] |= 1
    res = dict(iters_robust=s1["iterations"], iters_refine=0, initial_cost=s1["initial_cost"],
               final_cost=s1["final_cost"], n_outliers_first=int(bad1.sum()), n_outliers_second=0,
               termination=s1["termination"], summaries=[s1])
    if apply_l2_after_robust and use_robust and bad1.any():
        active = active & ~bad1                      # problem.RemoveResidualBlock (optimizer.cpp:518-521)
        # mono windows keep the Huber loss in the refinement: the wrapper is only reset to the trivial
        # loss when both the left-camera and the other-frame right-camera residual lists are still
        # non-empty after the first outlier scan (optimizer.cpp:606-608)
        if refine_trivial_loss is None:
            typ = pb["obs_type"] if pb.get("obs_type") is not None else np.zeros(nobs, np.uint8)
            refine_trivial_loss = bool((active & (typ == 0)).any() and (active & (typ == 1)).any())
        pose, invd, s2, last2 = ceres_solve(pb, pose, invd, active, max_iters_refine,
                                            None if refine_trivial_loss else a, function_tolerance, log)
        bad2 = active & ((last2["chi2"] > th) | ~last2["depth_pos"])
        flags[bad2] |= 2
        res.update(iters_refine=s2["iterations"], initial_cost=s2["initial_cost"], final_cost=s2["final_cost"],
                   n_outliers_second=int(bad2.sum()), termination=s2["termination"])
        res["summaries"].append(s2)
    pb["pose"][...] = pose
    pb["lm_invdepth"][...] = invd
    res["flags"] = flags
    return res

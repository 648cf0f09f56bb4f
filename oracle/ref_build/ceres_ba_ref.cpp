// TEST INFRASTRUCTURE (oracle/): Optimizer::localBA's solve on a flat window, run by the REAL Ceres 2.0 (the sources vendored in the
// reference tree, compiled in place by build_ceres_ref.py) on the REFERENCE'S OWN cost functions (src/ceres_parametrization.cpp).
// The driver does what /root/reference/src/optimizer.cpp does between problem set-up and write-back, with the same Ceres calls:
//   parameter blocks        calibration(s) and stereo extrinsic constant (:95-125), keyframe poses with SE3LeftParameterization, some
//                           constant (:169-183, :239-245), one inverse-depth block per landmark (:266); elimination ordering: landmarks
//                           in group 0, everything else in group 1 (:96-123, :172, :267)
//   residual blocks         the three anchored inverse-depth cost functions behind ONE LossFunctionWrapper(HuberLoss(sqrt(mono_th))) (:49, :270-390)
//   options                 DENSE_SCHUR (the `use_sparse_schur: 0` configuration; SPARSE_SCHUR differs only in how the reduced system is
//                           factorised), LEVENBERG_MARQUARDT, 1 thread, max_num_iterations 5, function_tolerance 1e-3 (:436-466);
//                           the wall-clock cap (max_solver_time_in_seconds) is lifted: it makes the reference itself non-deterministic
//   two-stage flow          Solve, outlier scan on the cost functions' chi2err_ / isdepthpositive_ (the values of their LAST Evaluate,
//                           :492-594), RemoveResidualBlock, loss reset to L2 only when left- and right-camera lists are both non-empty
//                           (:606-608), second Solve with 10 iterations, second scan (:637-735)
// tests/test_oracle_vs_reference_ceres.py compares oracle/ba_ref.py::local_ba with this, iteration by iteration.
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "ceres_parametrization.hpp"

namespace {

struct Block {
    ceres::CostFunction* f;
    ceres::ResidualBlockId id;
    int obs;
    int type;
    bool alive;
};

bool is_bad(const Block& b, float th) {
    double chi2;
    bool dpos;
    if (b.type == 0) { auto* e = static_cast<DirectLeftSE3::ReprojectionErrorKSE3AnchInvDepth*>(b.f); chi2 = e->chi2err_; dpos = e->isdepthpositive_; }
    else if (b.type == 1) { auto* e = static_cast<DirectLeftSE3::ReprojectionErrorRightCamKSE3AnchInvDepth*>(b.f); chi2 = e->chi2err_; dpos = e->isdepthpositive_; }
    else { auto* e = static_cast<DirectLeftSE3::ReprojectionErrorRightAnchCamKSE3AnchInvDepth*>(b.f); chi2 = e->chi2err_; dpos = e->isdepthpositive_; }
    return chi2 > th || !dpos;
}

// one row per entry of Solver::Summary::iterations
void dump_iterations(const ceres::Solver::Summary& s, double* log, int maxlog, int* nlog) {
    int n = 0;
    for (const auto& it : s.iterations) {
        if (n >= maxlog) break;
        double* r = log + 9 * n++;
        r[0] = it.iteration; r[1] = it.step_is_valid; r[2] = it.step_is_successful; r[3] = it.cost; r[4] = it.cost_change;
        r[5] = it.gradient_max_norm; r[6] = it.step_norm; r[7] = it.relative_decrease; r[8] = it.trust_region_radius;
    }
    *nlog = n;
}

void dump_summary(const ceres::Solver::Summary& s, double* out) {
    out[0] = s.initial_cost; out[1] = s.final_cost; out[2] = (double)s.termination_type; out[3] = s.num_successful_steps;
    out[4] = s.num_unsuccessful_steps; out[5] = (double)s.iterations.size(); out[6] = s.num_residual_blocks_reduced; out[7] = s.num_parameter_blocks_reduced;
}

}  // namespace

extern "C" int ov2ref_local_ba(int ncam, int npts, int nobs, const double* K, const double* Kr, const double* Trl, double* pose,
                               const uint8_t* pose_const, const int32_t* lm_anchor_cam, const double* lm_anchor_px, double* lm_invdepth,
                               const int32_t* obs_cam, const int32_t* obs_lm, const double* obs_px, const uint8_t* obs_type, int max_iters_robust,
                               int max_iters_refine, double mono_th_in, double function_tolerance, int use_robust, int apply_l2_after_robust,
                               uint8_t* flags, double* summaries /* [2][8] */, double* iter_log /* [2][maxlog][9] */, int maxlog, int* nlog /* [2] */) {
    const float mono_th = (float)mono_th_in;                          // const float mono_th = 5.9915 (optimizer.cpp:47)
    ceres::Problem problem;
    auto* loss_function = new ceres::LossFunctionWrapper(new ceres::HuberLoss(std::sqrt(mono_th)), ceres::TAKE_OWNERSHIP);
    if (!use_robust) loss_function->Reset(nullptr, ceres::TAKE_OWNERSHIP);
    auto* ordering = new ceres::ParameterBlockOrdering;

    std::array<double, 4> calib = {K[0], K[1], K[2], K[3]}, rcalib = {0, 0, 0, 0};
    std::array<double, 7> extrin = {0, 0, 0, 0, 0, 0, 1};
    problem.AddParameterBlock(calib.data(), 4);
    ordering->AddElementToGroup(calib.data(), 1);
    problem.SetParameterBlockConstant(calib.data());
    const bool stereo = obs_type != nullptr && Kr != nullptr && Trl != nullptr;
    if (stereo) {
        for (int i = 0; i < 4; ++i) rcalib[i] = Kr[i];
        for (int i = 0; i < 7; ++i) extrin[i] = Trl[i];
        problem.AddParameterBlock(rcalib.data(), 4);
        ordering->AddElementToGroup(rcalib.data(), 1);
        problem.SetParameterBlockConstant(rcalib.data());
        problem.AddParameterBlock(extrin.data(), 7, new SE3LeftParameterization());
        ordering->AddElementToGroup(extrin.data(), 1);
        problem.SetParameterBlockConstant(extrin.data());
    }
    std::vector<std::array<double, 7>> poses(ncam);
    for (int c = 0; c < ncam; ++c) {
        for (int i = 0; i < 7; ++i) poses[c][i] = pose[7 * c + i];
        problem.AddParameterBlock(poses[c].data(), 7, new SE3LeftParameterization());
        ordering->AddElementToGroup(poses[c].data(), 1);
        if (pose_const[c]) problem.SetParameterBlockConstant(poses[c].data());
    }
    std::vector<std::array<double, 1>> invd(npts);
    for (int l = 0; l < npts; ++l) {
        invd[l][0] = lm_invdepth[l];
        problem.AddParameterBlock(invd[l].data(), 1);
        ordering->AddElementToGroup(invd[l].data(), 0);
    }
    std::vector<Block> blocks;
    blocks.reserve(nobs);
    for (int i = 0; i < nobs; ++i) {
        const int l = obs_lm[i], ca = lm_anchor_cam[l], co = obs_cam[i], t = stereo ? obs_type[i] : 0;
        const double u = obs_px[2 * i], v = obs_px[2 * i + 1], ua = lm_anchor_px[2 * l], va = lm_anchor_px[2 * l + 1];
        Block b{nullptr, nullptr, i, t, true};
        if (t == 0) {
            b.f = new DirectLeftSE3::ReprojectionErrorKSE3AnchInvDepth(u, v, ua, va, 1.);
            b.id = problem.AddResidualBlock(b.f, loss_function, calib.data(), poses[ca].data(), poses[co].data(), invd[l].data());
        } else if (t == 1) {
            b.f = new DirectLeftSE3::ReprojectionErrorRightCamKSE3AnchInvDepth(u, v, ua, va, 1.);
            b.id = problem.AddResidualBlock(b.f, loss_function, calib.data(), rcalib.data(), poses[ca].data(), poses[co].data(), extrin.data(), invd[l].data());
        } else {
            b.f = new DirectLeftSE3::ReprojectionErrorRightAnchCamKSE3AnchInvDepth(u, v, ua, va, 1.);
            b.id = problem.AddResidualBlock(b.f, loss_function, calib.data(), rcalib.data(), extrin.data(), invd[l].data());
        }
        blocks.push_back(b);
    }

    ceres::Solver::Options options;
    options.linear_solver_ordering.reset(ordering);
    options.linear_solver_type = ceres::DENSE_SCHUR;
    options.trust_region_strategy_type = ceres::LEVENBERG_MARQUARDT;
    options.num_threads = 1;
    options.max_num_iterations = max_iters_robust;
    options.function_tolerance = function_tolerance;
    options.max_solver_time_in_seconds = 1e9;
    options.minimizer_progress_to_stdout = false;
    options.logging_type = ceres::SILENT;

    memset(flags, 0, (size_t)nobs);
    memset(summaries, 0, sizeof(double) * 16);
    nlog[0] = nlog[1] = 0;
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    dump_summary(summary, summaries);
    dump_iterations(summary, iter_log, maxlog, &nlog[0]);

    size_t nbbad = 0;
    for (auto& b : blocks) {
        if (!is_bad(b, mono_th)) continue;
        flags[b.obs] |= 1;
        nbbad++;
        if (apply_l2_after_robust) { problem.RemoveResidualBlock(b.id); b.alive = false; }       // (:518-521; the cost function dies with it)
    }
    int solves = 1;
    if (apply_l2_after_robust && use_robust && nbbad > 0) {
        bool any_left = false, any_right = false;
        for (const auto& b : blocks) if (b.alive) { any_left |= b.type == 0; any_right |= b.type == 1; }
        if (any_left && any_right) loss_function->Reset(nullptr, ceres::TAKE_OWNERSHIP);          // (:606-608)
        options.max_num_iterations = max_iters_refine;
        options.function_tolerance = function_tolerance;          // the reference re-states 1e-3 here (:611), the value it also uses for the first solve
        ceres::Solve(options, &problem, &summary);
        dump_summary(summary, summaries + 8);
        dump_iterations(summary, iter_log + 9 * (size_t)maxlog, maxlog, &nlog[1]);
        for (const auto& b : blocks)
            if (b.alive && is_bad(b, mono_th)) flags[b.obs] |= 2;
        solves = 2;
    }
    for (int c = 0; c < ncam; ++c) for (int i = 0; i < 7; ++i) pose[7 * c + i] = poses[c][i];
    for (int l = 0; l < npts; ++l) lm_invdepth[l] = invd[l][0];
    return solves;
}

// MultiViewGeometry::ceresPnP (/root/reference/src/multi_view_geometry.cpp:492-588) with the same Ceres calls: one pose block with
// SE3LeftParameterization, ReprojectionErrorSE3 blocks (sigma = 2^scale) behind LossFunctionWrapper(HuberLoss(sqrt(chi2th))), DENSE_QR,
// Levenberg-Marquardt, nmaxiter iterations, function_tolerance 1e-3; outlier scan on the last Evaluate; all outliers -> false without
// touching Twc; optional second solve with the trivial loss on the inliers.  The 5 ms wall-clock cap is lifted.
// Returns -1 when every block is an outlier, else summary.IsSolutionUsable().
extern "C" int ov2ref_ceres_pnp(int n, const double* unpx, const double* wpts, const int32_t* scales, double* Twc, int nmaxiter, double chi2th_in,
                                int use_robust, int apply_l2_after_robust, const double* Kin, uint8_t* outlier, double* summaries /* [2][8] */) {
    const float chi2th = (float)chi2th_in;
    const float fx = (float)Kin[0], fy = (float)Kin[1], cx = (float)Kin[2], cy = (float)Kin[3];      // float arguments (:497)
    ceres::Problem problem;
    const double chi2thrtsq = std::sqrt(chi2th);
    auto* loss_function = new ceres::LossFunctionWrapper(new ceres::HuberLoss(chi2thrtsq), ceres::TAKE_OWNERSHIP);
    if (!use_robust) loss_function->Reset(NULL, ceres::TAKE_OWNERSHIP);
    std::array<double, 7> posepar;
    for (int i = 0; i < 7; ++i) posepar[i] = Twc[i];
    {   // PoseParametersBlock(0, Twc) stores Sophus' unit quaternion
        const double q = std::sqrt(posepar[3] * posepar[3] + posepar[4] * posepar[4] + posepar[5] * posepar[5] + posepar[6] * posepar[6]);
        for (int i = 3; i < 7; ++i) posepar[i] /= q;
    }
    problem.AddParameterBlock(posepar.data(), 7, new SE3LeftParameterization());
    std::vector<DirectLeftSE3::ReprojectionErrorSE3*> verrors;
    std::vector<ceres::ResidualBlockId> vrids;
    for (int i = 0; i < n; ++i) {
        auto* f = new DirectLeftSE3::ReprojectionErrorSE3(unpx[2 * i], unpx[2 * i + 1], fx, fy, cx, cy,
                                                          Eigen::Vector3d(wpts[3 * i], wpts[3 * i + 1], wpts[3 * i + 2]), std::pow(2., scales ? scales[i] : 0));
        vrids.push_back(problem.AddResidualBlock(f, loss_function, posepar.data()));
        verrors.push_back(f);
    }
    ceres::Solver::Options options;
    options.linear_solver_type = ceres::DENSE_QR;
    options.trust_region_strategy_type = ceres::LEVENBERG_MARQUARDT;
    options.num_threads = 1;
    options.max_num_iterations = nmaxiter;
    options.max_solver_time_in_seconds = 1e9;
    options.function_tolerance = 1.e-3;
    options.minimizer_progress_to_stdout = false;
    options.logging_type = ceres::SILENT;
    memset(summaries, 0, sizeof(double) * 16);
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    dump_summary(summary, summaries);
    int nbbad = 0;
    for (int i = 0; i < n; ++i) {
        outlier[i] = 0;
        if (verrors[i]->chi2err_ > chi2th || !verrors[i]->isdepthpositive_) {
            if (apply_l2_after_robust) problem.RemoveResidualBlock(vrids[i]);
            outlier[i] = 1;
            nbbad++;
        }
    }
    if (nbbad == n) return -1;
    if (apply_l2_after_robust && nbbad > 0) {
        loss_function->Reset(NULL, ceres::TAKE_OWNERSHIP);
        ceres::Solve(options, &problem, &summary);
        dump_summary(summary, summaries + 8);
    }
    for (int i = 0; i < 7; ++i) Twc[i] = posepar[i];
    return summary.IsSolutionUsable() ? 1 : 0;
}

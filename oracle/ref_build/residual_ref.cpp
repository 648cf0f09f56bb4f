// TEST INFRASTRUCTURE (oracle/): C entry points over the REFERENCE's own residual classes.  This file is compiled together with
// /root/reference/src/ceres_parametrization.cpp (from where it lies; never copied) against the stand-in linear-algebra headers in
// mini/ into oracle/_ref/libov2ref_residuals.so, so that tests can check the restatements (oracle/ba_ref.py, oracle/pnp_ref.py and,
// through them, the CUDA kernels) against what the reference's source computes: residuals, chi2 / depth flags and every Jacobian
// block of the three anchored inverse-depth cost functions, the PnP cost function, and SE3LeftParameterization::Plus.
// Pose blocks are [tx ty tz qx qy qz qw] as PoseParametersBlock stores them (se3_param_block.hpp:39-46).
#include "ceres_parametrization.hpp"

extern "C" {

// DirectLeftSE3::ReprojectionErrorKSE3AnchInvDepth (ceres_parametrization.cpp:361-473).  Jacobians row-major: Jk 2x4, Ja / Jo 2x7, Jl 2x1.
int ref_eval_anch_invdepth(const double* K, const double* Twanch, const double* Twc, double invdepth, double u, double v, double ua, double va,
                           double sigma, double* res, double* Jk, double* Ja, double* Jo, double* Jl, double* chi2, int* depth_positive) {
    DirectLeftSE3::ReprojectionErrorKSE3AnchInvDepth f(u, v, ua, va, sigma);
    const double* params[4] = {K, Twanch, Twc, &invdepth};
    double* jac[4] = {Jk, Ja, Jo, Jl};
    const bool ok = f.Evaluate(params, res, jac);
    *chi2 = f.chi2err_;
    *depth_positive = f.isdepthpositive_ ? 1 : 0;
    return ok ? 1 : 0;
}

// ...RightAnchCam... (:476-577): right-camera observation in the anchor keyframe.  Jkl / Jkr 2x4, Jrl 2x7, Jl 2x1.
int ref_eval_right_anch(const double* Kl, const double* Kr, const double* Trl, double invdepth, double ur, double vr, double ua, double va,
                        double sigma, double* res, double* Jkl, double* Jkr, double* Jrl, double* Jl, double* chi2, int* depth_positive) {
    DirectLeftSE3::ReprojectionErrorRightAnchCamKSE3AnchInvDepth f(ur, vr, ua, va, sigma);
    const double* params[4] = {Kl, Kr, Trl, &invdepth};
    double* jac[4] = {Jkl, Jkr, Jrl, Jl};
    const bool ok = f.Evaluate(params, res, jac);
    *chi2 = f.chi2err_;
    *depth_positive = f.isdepthpositive_ ? 1 : 0;
    return ok ? 1 : 0;
}

// ...RightCam... (:579-712): right-camera observation in another keyframe.
int ref_eval_right_cam(const double* Kl, const double* Kr, const double* Twanch, const double* Twc, const double* Trl, double invdepth,
                       double ur, double vr, double ua, double va, double sigma, double* res, double* Jkl, double* Jkr, double* Ja,
                       double* Jo, double* Jrl, double* Jl, double* chi2, int* depth_positive) {
    DirectLeftSE3::ReprojectionErrorRightCamKSE3AnchInvDepth f(ur, vr, ua, va, sigma);
    const double* params[6] = {Kl, Kr, Twanch, Twc, Trl, &invdepth};
    double* jac[6] = {Jkl, Jkr, Ja, Jo, Jrl, Jl};
    const bool ok = f.Evaluate(params, res, jac);
    *chi2 = f.chi2err_;
    *depth_positive = f.isdepthpositive_ ? 1 : 0;
    return ok ? 1 : 0;
}

// DirectLeftSE3::ReprojectionErrorSE3 (:301-358), the ceresPnP residual.  J 2x7.
int ref_eval_pnp(const double* K, const double* Twc, const double* wpt, double u, double v, double sigma, double* res, double* J, double* chi2,
                 int* depth_positive) {
    DirectLeftSE3::ReprojectionErrorSE3 f(u, v, K[0], K[1], K[2], K[3], Eigen::Vector3d(wpt[0], wpt[1], wpt[2]), sigma);
    const double* params[1] = {Twc};
    double* jac[1] = {J};
    const bool ok = f.Evaluate(params, res, jac);
    *chi2 = f.chi2err_;
    *depth_positive = f.isdepthpositive_ ? 1 : 0;
    return ok ? 1 : 0;
}

// SE3LeftParameterization::Plus / ComputeJacobian (se3left_parametrization.hpp:41-69).  J 7x6 row-major.
int ref_se3_plus(const double* x, const double* delta, double* x_plus_delta, double* J) {
    SE3LeftParameterization p;
    const bool ok = p.Plus(x, delta, x_plus_delta);
    if (J) p.ComputeJacobian(x, J);
    return ok && p.GlobalSize() == 7 && p.LocalSize() == 6 ? 1 : 0;
}

}  // extern "C"

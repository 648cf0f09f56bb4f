// Optimizer::localBA on the GPU: the reference's declaration (include/optimizer.hpp:33-67) is kept,
// this translation unit supplies Optimizer::localBA(Frame&, bool) so estimator.cpp:91 links
// unchanged; the other Optimizer members (looseBA, fullBA, pose graphs, stop flags) stay in the
// reference's own optimizer.cpp built with -DOV2_EXTERNAL_LOCALBA (one #ifndef around its localBA
// body).
//
// Compile-checked and self-tested in this container against stand-in headers that mirror the reference's class
// interfaces (host/standin/ref/: optimizer.hpp, map_manager.hpp, sophus/se3.hpp, Eigen/Dense) - the reference's own
// headers need ROS, PCL, OpenCV, Eigen and Sophus, none of which exist here;
// host/optimizer_selftest.cpp builds a synthetic map, runs this localBA and tests/test_host_shim.py compares the
// map it leaves behind with the flat solve of the same window.  On a box that builds the reference, compile this file
// with the reference's include path instead of the stand-ins.
// Mono and stereo windows (inverse-depth parametrisation, buse_inv_depth_ = 1: every shipped
// configuration); the XYZ parametrisation (buse_inv_depth_ = 0) is not flattened and falls through
// with a diagnostic.
//
// What stays on the host is exactly the part SURVEY.md 8a row G scopes out of acceleration: walking
// the covisibility graph / hash maps to FLATTEN the window into the SoA arrays ov2_localba_solve
// takes, and applying the result to the map.  The numerical work (Ceres in the reference,
// optimizer.cpp:436-735) is one ABI call = one kernel launch.
#include <atomic>
#include <chrono>
#include <iostream>
#include <map>
#include <thread>
#include <unordered_map>
#include <vector>

#include "optimizer.hpp"          // the reference's, unchanged
#include "../../include/ov2b200.h"

namespace {

ov2_ctx* ba_context() {
    static ov2_ctx* ctx = nullptr;                   // estimator thread only (optim_mutex_, estimator.cpp:85)
    if (!ctx && ov2_create(0, &ctx) != OV2_OK) ctx = nullptr;
    return ctx;
}

// Flat window: camera index <-> keyframe id, landmark index <-> map point id.
struct Window {
    std::vector<int> kfids;  std::unordered_map<int, int> cam_of_kf;
    std::vector<std::shared_ptr<Frame>> kfs;
    std::vector<double> pose;  std::vector<uint8_t> pose_const;
    std::vector<int> lmids;  std::vector<std::shared_ptr<MapPoint>> lms;
    std::vector<int32_t> lm_anchor_cam;  std::vector<double> lm_anchor_px, lm_invdepth;
    std::vector<int32_t> obs_cam, obs_lm;  std::vector<double> obs_px;  std::vector<uint8_t> obs_type;

    // residual type as include/ov2b200.h: 0 left camera / other keyframe, 1 right camera / other keyframe,
    // 2 right camera / anchor keyframe (obs_cam = anchor)
    void add_obs(int cam, int lm, double u, double v, uint8_t type) {
        obs_cam.push_back(cam); obs_lm.push_back(lm); obs_px.push_back(u); obs_px.push_back(v); obs_type.push_back(type);
    }

    int add_camera(int kfid, const std::shared_ptr<Frame>& kf, bool constant) {
        auto it = cam_of_kf.find(kfid);
        if (it != cam_of_kf.end()) return it->second;
        const int c = (int)kfids.size();
        cam_of_kf.emplace(kfid, c);
        kfids.push_back(kfid);
        kfs.push_back(kf);
        const Sophus::SE3d T = kf->getTwc();
        const Eigen::Quaterniond q = T.unit_quaternion();
        const Eigen::Vector3d t = T.translation();
        const double p[7] = {t.x(), t.y(), t.z(), q.x(), q.y(), q.z(), q.w()};   // se3_param_block.hpp:39-46
        pose.insert(pose.end(), p, p + 7);
        pose_const.push_back(constant ? 1 : 0);
        return c;
    }
};

}  // namespace

void Optimizer::localBA(Frame &newframe, const bool buse_robust_cost)
{
    // Early-outs of the reference (optimizer.cpp:61-63): poor tracking -> no BA.
    const int nmincov = pslamstate_->nmin_covscore_;
    if ((int)newframe.nb3dkps_ < nmincov) return;
    ov2_ctx* ctx = ba_context();
    if (!ctx) { std::cerr << "[ov2b200] localBA: no CUDA device (no CPU fallback)\n"; return; }
    const bool stereo = pslamstate_->stereo_;
    if (!pslamstate_->buse_inv_depth_) {
        std::cerr << "[ov2b200] localBA: only the anchored inverse-depth parametrisation (buse_inv_depth: 1) is built\n";
        return;
    }

    // ---- 1. window: covisible keyframes, newest first; optimise while the covisibility score holds,
    //         everything older than the first weak keyframe is constant (optimizer.cpp:128-188)
    Window win;
    std::map<int, int> cov = newframe.getCovisibleKfMap();
    cov.emplace(newframe.kfid_, newframe.nb3dkps_);
    std::vector<int> lm_order;            // landmark ids seen from optimised keyframes, de-duplicated
    std::unordered_map<int, int> lm_seen;
    bool freeze = false;
    const int newest = cov.rbegin()->first;
    for (auto it = cov.rbegin(); it != cov.rend(); ++it) {
        const int kfid = it->first;
        const int score = kfid > newframe.kfid_ ? (int)newframe.nbkps_ : it->second;
        auto kf = pmap_->getKeyframe(kfid);
        if (!kf) { newframe.removeCovisibleKf(kfid); continue; }
        const bool optimise = score >= nmincov && !freeze && kfid > 0;
        if (!optimise) freeze = true;
        win.add_camera(kfid, kf, !optimise);
        if (optimise)
            for (const auto& kp : kf->getKeypoints3d())
                if (lm_seen.emplace(kp.lmid_, 1).second) lm_order.push_back(kp.lmid_);
    }

    // ---- 2. landmarks: first valid observer (ascending kf id) anchors the inverse depth, every
    //         further observer contributes one residual block (optimizer.cpp:191-392)
    std::vector<std::pair<int, int>> bad_obs;     // (kfid, lmid) to drop from the map afterwards
    for (int lmid : lm_order) {
        auto lm = pmap_->getMapPoint(lmid);
        if (!lm || lm->isBad()) continue;
        int anchor_cam = -1;
        const int l = (int)win.lmids.size();
        for (int kfid : lm->getKfObsSet()) {
            if (kfid > newest) continue;
            auto cit = win.cam_of_kf.find(kfid);
            std::shared_ptr<Frame> kf = cit != win.cam_of_kf.end() ? win.kfs[cit->second] : pmap_->getKeyframe(kfid);
            if (!kf) { pmap_->removeMapPointObs(kfid, lmid); continue; }
            const int cam = win.add_camera(kfid, kf, /*constant=*/cit == win.cam_of_kf.end());
            const auto kp = kf->getKeypointById(lmid);
            if (kp.lmid_ != lmid) { pmap_->removeMapPointObs(lmid, kfid); continue; }
            // kp.scale_ is 0 for every keypoint the reference's front-end creates (map_manager.cpp adds
            // keypoints without a scale), so the residuals' information matrix 2^-scale * I is the identity
            if (anchor_cam < 0) {
                anchor_cam = cam;
                win.lmids.push_back(lmid);
                win.lms.push_back(lm);
                win.lm_anchor_cam.push_back(cam);
                win.lm_anchor_px.push_back(kp.unpx_.x);
                win.lm_anchor_px.push_back(kp.unpx_.y);
                win.lm_invdepth.push_back(1.0 / (kf->getTcw() * lm->getPoint()).z());
                // the anchor's own right-camera observation constrains the inverse depth alone (optimizer.cpp:268-284)
                if (stereo && kp.is_stereo_) win.add_obs(cam, l, kp.runpx_.x, kp.runpx_.y, 2);
                continue;
            }
            win.add_obs(cam, l, kp.unpx_.x, kp.unpx_.y, 0);            // appended landmark by landmark: already sorted
            if (stereo && kp.is_stereo_) win.add_obs(cam, l, kp.runpx_.x, kp.runpx_.y, 1);   // (optimizer.cpp:295-324)
        }
    }
    // ---- 3. gauge: at least two constant keyframes in mono, one in stereo (optimizer.cpp:65-69, 396-407).
    //         The reference walks its unordered_map of local keyframes (unspecified order); ascending
    //         keyframe id is used here.
    size_t nconst = 0;
    for (uint8_t c : win.pose_const) nconst += c;
    {
        const size_t nmincst = stereo ? 1 : 2;
        std::map<int, int> by_id(win.cam_of_kf.begin(), win.cam_of_kf.end());
        for (auto it = by_id.begin(); nconst < nmincst && it != by_id.end(); ++it)
            if (!win.pose_const[it->second]) { win.pose_const[it->second] = 1; nconst++; }
    }
    if (win.obs_cam.empty()) return;
    // ov2_localba_solve optimises at most 64 keyframes per window (MAX_VAR_CAMS): beyond that the OLDEST optimised
    // keyframes are held constant (the reference has no such limit; windows that large do not occur with the shipped
    // nmin_covscore / covisibility settings - said once on stderr if it ever happens)
    {
        size_t nvar = win.pose_const.size() - nconst;
        if (nvar > 64) {
            static bool told = false;
            if (!told) { std::cerr << "[ov2b200] localBA: " << nvar << " optimised keyframes, the oldest " << nvar - 64 << " are held constant\n"; told = true; }
            std::map<int, int> by_id(win.cam_of_kf.begin(), win.cam_of_kf.end());
            for (auto it = by_id.begin(); nvar > 64 && it != by_id.end(); ++it)
                if (!win.pose_const[it->second]) { win.pose_const[it->second] = 1; nvar--; }
        }
    }

    // ---- 4. solve on the GPU (replaces ceres::Solve x2 + the two outlier scans)
    auto cal = newframe.pcalib_leftcam_;
    const double K[4] = {cal->fx_, cal->fy_, cal->cx_, cal->cy_};
    ov2_ba_problem pb;
    pb.ncam = (int)win.kfids.size(); pb.npts = (int)win.lmids.size(); pb.nobs = (int)win.obs_cam.size();
    pb.K = K; pb.pose = win.pose.data(); pb.pose_const = win.pose_const.data();
    pb.lm_anchor_cam = win.lm_anchor_cam.data(); pb.lm_anchor_px = win.lm_anchor_px.data();
    pb.lm_invdepth = win.lm_invdepth.data();
    pb.obs_cam = win.obs_cam.data(); pb.obs_lm = win.obs_lm.data(); pb.obs_px = win.obs_px.data();
    double Kr[4] = {0, 0, 0, 0}, Trl[7] = {0, 0, 0, 0, 0, 0, 1};
    pb.obs_type = nullptr; pb.Kr = nullptr; pb.Trl = nullptr;
    if (stereo) {
        auto calr = newframe.pcalib_rightcam_;
        Kr[0] = calr->fx_; Kr[1] = calr->fy_; Kr[2] = calr->cx_; Kr[3] = calr->cy_;
        const Sophus::SE3d T = calr->getExtrinsic().inverse();          // Trl (optimizer.cpp:112-114)
        const Eigen::Quaterniond q = T.unit_quaternion();
        const Eigen::Vector3d t = T.translation();
        const double e[7] = {t.x(), t.y(), t.z(), q.x(), q.y(), q.z(), q.w()};
        for (int i = 0; i < 7; ++i) Trl[i] = e[i];
        pb.obs_type = win.obs_type.data(); pb.Kr = Kr; pb.Trl = Trl;
    }
    ov2_ba_opts op;
    op.max_iters_robust = 5; op.max_iters_refine = 10;                // optimizer.cpp:462, :610
    op.huber_th = pslamstate_->robust_mono_th_; op.function_tolerance = 1.e-3;
    op.use_robust = buse_robust_cost ? 1 : 0;
    op.apply_l2_after_robust = pslamstate_->apply_l2_after_robust_ ? 1 : 0;
    op.refine_loss = -1;                                              // as optimizer.cpp:606-608 decides
    ov2_ba_result res;
    std::vector<uint8_t> flags(win.obs_cam.size(), 0);
    // The reference polls stopLocalBA() once, between the robust solve and the refinement (optimizer.cpp:603-604); the
    // stop request arrives from the mapper thread (estimator.cpp:228-232) while the solve is running.  The solve kernel
    // polls a flag at exactly that point (ov2_localba_request_stop); a watcher forwards bstop_localba_ to it, so the
    // reference's signalStopLocalBA() needs no edit.
    ov2_localba_request_stop(ctx, stopLocalBA() ? 1 : 0);
    std::atomic<bool> solve_done{false};
    std::thread watcher([&]() {
        while (!solve_done.load(std::memory_order_acquire)) {
            if (bstop_localba_) { ov2_localba_request_stop(ctx, 1); return; }
            std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
    });
    const ov2_status solve_rc = ov2_localba_solve(ctx, &pb, &op, &res, flags.data());
    solve_done.store(true, std::memory_order_release);
    watcher.join();
    ov2_localba_request_stop(ctx, 0);
    if (solve_rc != OV2_OK) {
        std::cerr << "[ov2b200] localBA: " << ov2_last_error(ctx) << "\n";
        bstop_localba_ = false;
        return;                                                        // map untouched
    }

    // ---- 5. apply to the map under map_mutex_ (optimizer.cpp:741-897)
    std::lock_guard<std::mutex> lock(pmap_->map_mutex_);
    std::unordered_map<int, int> suspicious;                           // lmid -> had a rejected observation
    // rejected right-camera observations only lose their stereo half (removeStereoKeypointById,
    // optimizer.cpp:742-750); rejected left-camera observations are removed from the map (:752-763)
    for (size_t i = 0; i < flags.size(); ++i) {
        if (!flags[i] || win.obs_type[i] == 0) continue;
        win.kfs[win.obs_cam[i]]->removeStereoKeypointById(win.lmids[win.obs_lm[i]]);
        suspicious.emplace(win.lmids[win.obs_lm[i]], 1);
    }
    for (size_t i = 0; i < flags.size(); ++i) {
        if (!flags[i] || win.obs_type[i] != 0) continue;
        const int kfid = win.kfids[win.obs_cam[i]], lmid = win.lmids[win.obs_lm[i]];
        pmap_->removeMapPointObs(lmid, kfid);
        if (kfid == pmap_->pcurframe_->kfid_) pmap_->removeObsFromCurFrameById(lmid);
        suspicious.emplace(lmid, 1);
    }
    for (size_t c = 0; c < win.kfids.size(); ++c) {
        if (win.pose_const[c]) continue;
        const double* p = &win.pose[7 * c];
        win.kfs[c]->setTwc(Sophus::SE3d(Eigen::Quaterniond(p[6], p[3], p[4], p[5]), Eigen::Vector3d(p[0], p[1], p[2])));
    }
    auto cull_if_weak = [&](int lmid, const std::shared_ptr<MapPoint>& lm) {
        if (lm->isBad()) { pmap_->removeMapPoint(lmid); return true; }
        if (lm->getKfObsSet().size() < 3 && lm->kfid_ < newframe.kfid_ - 3 && !lm->isobs_) { pmap_->removeMapPoint(lmid); return true; }
        return false;
    };
    for (size_t l = 0; l < win.lmids.size(); ++l) {
        const int lmid = win.lmids[l];
        auto lm = win.lms[l];
        if (cull_if_weak(lmid, lm)) { suspicious.erase(lmid); continue; }
        const double invd = win.lm_invdepth[l];
        if (1.0 / invd <= 0.0) { pmap_->removeMapPoint(lmid); suspicious.erase(lmid); continue; }
        auto ait = win.cam_of_kf.find(lm->kfid_);                       // anchor looked up by the map point's own kf id
        if (ait == win.cam_of_kf.end()) { suspicious.emplace(lmid, 1); continue; }
        auto kfa = win.kfs[ait->second];
        const auto kp = kfa->getKeypointById(lmid);
        const Eigen::Vector3d ray(kp.unpx_.x, kp.unpx_.y, 1.0);
        pmap_->updateMapPoint(lmid, kfa->getTwc() * ((1.0 / invd) * kfa->pcalib_leftcam_->iK_ * ray), invd);
    }
    for (const auto& s : suspicious) {
        auto lm = pmap_->getMapPoint(s.first);
        if (lm) cull_if_weak(s.first, lm);
    }
    bstop_localba_ = false;                                            // optimizer.cpp:896: a stop request is consumed by the BA it interrupted
}


// TEST INFRASTRUCTURE (oracle/): the REFERENCE'S OWN Optimizer::localBA (/root/reference/src/optimizer.cpp, compiled where it lies with
// frame.cpp, map_point.cpp, map_manager.cpp, camera_calibration.cpp, multi_view_geometry.cpp and the Ceres 2.0 of the reference tree)
// run end to end on a map that this driver builds through the reference's own map API from a flat window: keyframes are created in
// index order (Frame::addKeypoint, MapManager::addMapPoint for landmarks first seen there, prepareFrame, addKeyframe), landmarks get
// their 3-D position from the anchor keyframe (MapManager::updateMapPoint), covisibility comes from MapManager::updateFrameCovisibility,
// then localBA(newest keyframe, robust) selects its window, sets the Ceres problem up, solves, culls and writes back - all reference code.
// tests/test_oracle_vs_reference_map.py compares the map it leaves behind with the flat solve (oracle/ba_ref.py::local_ba, which
// equals the GPU solve and what the drop-in Optimizer::localBA feeds it).
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "optimizer.hpp"
#include "mapper.hpp"
#include "loop_closer.hpp"

#include <algorithm>

extern "C" void* ov2ref_fe_create(int, int, double, int);
extern "C" void ov2ref_fe_destroy(void*);

extern "C" int ov2ref_run_local_ba(int ncam, int npts, int nobs, const double* K, int width, int height, const double* pose /* [ncam][7] Twc */,
                                   const int32_t* lm_anchor_cam, const double* lm_anchor_px, const double* lm_invdepth, const int32_t* obs_cam,
                                   const int32_t* obs_lm, const double* obs_px, int nmin_covscore, int mode /* 0 localBA, 1 looseBA, 2 fullBA */, double* pose_out, double* xyz_out /* [npts][3] world */,
                                   uint8_t* lm_alive /* [npts] */, int32_t* nobs_left /* [npts] keyframes still observing */) {
    auto params = std::make_shared<SlamParams>();
    SlamParams& S = *params;
    S.debug_ = false; S.log_timings_ = false;
    S.stereo_ = false; S.mono_ = true; S.bdo_stereo_rect_ = false;
    S.buse_inv_depth_ = true; S.apply_l2_after_robust_ = true; S.robust_mono_th_ = 5.9915f;
    S.use_sparse_schur_ = false;                      // DENSE_SCHUR (no sparse library in this build; optimizer.cpp:439-443)
    S.use_dogleg_ = false; S.use_subspace_dogleg_ = false; S.use_nonmonotic_step_ = false;
    S.bforce_realtime_ = false;
    S.nmin_covscore_ = nmin_covscore;
    S.nbmaxkps_ = 100000;
    S.blocalba_is_on_ = false;
    auto calib = std::make_shared<CameraCalibration>("pinhole", K[0], K[1], K[2], K[3], 0., 0., 0., 0., (double)width, (double)height);
    calib->Dcv_.release();                            // undistorted images: undistortImagePoint returns the pixel itself (camera_calibration.cpp:262)
    auto cur = std::make_shared<Frame>(calib, 35);
    auto fe = std::make_shared<FeatureExtractor>(1000, 35, 0.001, 10);
    auto ft = std::make_shared<FeatureTracker>(30, 0.01f, nullptr);
    auto map = std::make_shared<MapManager>(params, cur, fe, ft);
    Optimizer opt(params, map);

    // observations per keyframe, in window order
    std::vector<std::vector<std::pair<int, const double*>>> per_kf(ncam);
    for (int l = 0; l < npts; ++l) per_kf[lm_anchor_cam[l]].push_back({l, lm_anchor_px + 2 * l});
    for (int i = 0; i < nobs; ++i) per_kf[obs_cam[i]].push_back({obs_lm[i], obs_px + 2 * i});
    std::vector<int> lmid_of(npts, -1);
    for (int c = 0; c < ncam; ++c) {
        cur->reset();
        cur->updateFrame(c, 0.05 * c);
        const double* p = pose + 7 * c;
        cur->setTwc(Sophus::SE3d(Eigen::Quaterniond(p[6], p[3], p[4], p[5]), Eigen::Vector3d(p[0], p[1], p[2])));
        for (const auto& ob : per_kf[c]) {
            const int l = ob.first;
            const cv::Point2f px((float)ob.second[0], (float)ob.second[1]);
            if (lmid_of[l] < 0) {
                if (lm_anchor_cam[l] != c) return -2;            // the anchor must be the first keyframe that sees a landmark
                lmid_of[l] = map->nlmid_;
                cur->addKeypoint(px, map->nlmid_);
                map->addMapPoint();
            } else {
                cur->addKeypoint(px, lmid_of[l]);
            }
        }
        map->prepareFrame();
        map->addKeyframe();
    }
    for (int l = 0; l < npts; ++l) {
        auto kf = map->getKeyframe(lm_anchor_cam[l]);
        const Keypoint kp = kf->getKeypointById(lmid_of[l]);
        const double z = 1. / lm_invdepth[l];
        const Eigen::Vector3d campt = z * (kp.bv_ / kp.bv_.z());
        map->updateMapPoint(lmid_of[l], kf->projCamToWorld(campt), lm_invdepth[l]);
    }
    for (int c = 0; c < ncam; ++c) map->updateFrameCovisibility(*map->getKeyframe(c));

    if (mode == 1) opt.looseBA(0, ncam - 1, true);              // optimizer.cpp:900-1671
    else if (mode == 2) opt.fullBA(true);                        // :1674-2331
    else opt.localBA(*map->getKeyframe(ncam - 1), true);

    for (int c = 0; c < ncam; ++c) {
        const Sophus::SE3d T = map->getKeyframe(c)->getTwc();
        const Eigen::Vector3d t = T.translation();
        const Eigen::Quaterniond q = T.unit_quaternion();
        double* o = pose_out + 7 * c;
        o[0] = t.x(); o[1] = t.y(); o[2] = t.z(); o[3] = q.x(); o[4] = q.y(); o[5] = q.z(); o[6] = q.w();
    }
    for (int l = 0; l < npts; ++l) {
        auto plm = map->getMapPoint(lmid_of[l]);
        lm_alive[l] = plm != nullptr;
        nobs_left[l] = plm ? (int)plm->getKfObsSet().size() : 0;
        const Eigen::Vector3d w = plm ? plm->getPoint() : Eigen::Vector3d(0, 0, 0);
        xyz_out[3 * l] = w.x(); xyz_out[3 * l + 1] = w.y(); xyz_out[3 * l + 2] = w.z();
    }
    return 0;
}


// ---- loop closing is outside the scope: mapper.cpp refers to three LoopCloser entry points (src/loop_closer.cpp needs cv::BFMatcher & co)
LoopCloser::LoopCloser(std::shared_ptr<SlamParams>, std::shared_ptr<MapManager>) { fprintf(stderr, "oracle/_ref: LoopCloser is not part of this build\n"); abort(); }
void LoopCloser::run() { abort(); }
void LoopCloser::addNewKf(const std::shared_ptr<Frame>&, const cv::Mat&) { abort(); }

// The REFERENCE'S OWN Mapper::matchToMap (/root/reference/src/mapper.cpp:576-774) on a map built from a flattened matching scene
// (ov2slam_b200.synth.make_match_scene, undistorted): keyframes with their poses and the keypoints that observe the map points, map
// points with position / descriptors / keyframe sets, the frame with its keypoints added in index order (so its grid cells list them
// in that order).  A default-constructed Mapper is used (its other constructor starts the mapper, estimator and loop-closer threads).
// ids: map point m <-> lmid m; a keypoint without map point gets lmid 1000000 + its index; keyframe k <-> kfid 10 + k; the frame is 5000.
// Returns the number of pairs; order_out = the order in which the candidate id set was walked.
extern "C" int ov2ref_match_to_map(int nkps, int nmps, int nkfs, int ncand, const double* K, int width, int height, int ncellsize, const double* Tcw,
                                   const double* kf_Tcw, const double* mp_xyz, const float* kp_px, const int32_t* kp_lm, const int32_t* desc_ptr,
                                   const uint8_t* desc, const int32_t* obs_ptr, const int32_t* obs_kf, const float* obs_px, const int32_t* cand_mp,
                                   float fmaxprojerr, float fdistratio, int32_t* order_out, int32_t* pairs_out) {
    auto params = std::make_shared<SlamParams>();
    params->debug_ = false; params->log_timings_ = false;
    auto calib = std::make_shared<CameraCalibration>("pinhole", K[0], K[1], K[2], K[3], 0., 0., 0., 0., (double)width, (double)height);
    calib->Dcv_.release();
    auto cur = std::make_shared<Frame>(calib, (size_t)ncellsize);
    auto map = std::make_shared<MapManager>(params, cur, nullptr, nullptr);
    auto pose_of = [](const double* T) {               // camera <- world (row-major R, t)  ->  Twc
        Eigen::Matrix3d R;
        R << T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8];
        return Sophus::SE3d(Eigen::Quaterniond(R), Eigen::Vector3d(T[9], T[10], T[11])).inverse();
    };
    std::vector<std::shared_ptr<Frame>> kfs(nkfs);
    for (int k = 0; k < nkfs; ++k) {
        kfs[k] = std::make_shared<Frame>(calib, (size_t)ncellsize);
        kfs[k]->id_ = kfs[k]->kfid_ = 10 + k;
        kfs[k]->setTwc(pose_of(kf_Tcw + 12 * k));
        map->map_pkfs_.emplace(10 + k, kfs[k]);
    }
    for (int m = 0; m < nmps; ++m) {
        auto lm = std::make_shared<MapPoint>(m, 0, true);
        lm->set_kfids_.clear();
        lm->setPoint(Eigen::Vector3d(mp_xyz[3 * m], mp_xyz[3 * m + 1], mp_xyz[3 * m + 2]));
        lm->is3d_ = true;
        for (int d = desc_ptr[m]; d < desc_ptr[m + 1]; ++d) {
            cv::Mat row(1, 32, CV_8U);
            memcpy(row.data, desc + 32 * (size_t)d, 32);
            if (d == desc_ptr[m]) lm->desc_ = row;
            lm->map_kf_desc_.emplace(1000 + d - desc_ptr[m], row);
        }
        for (int o = obs_ptr[m]; o < obs_ptr[m + 1]; ++o) {
            lm->addKfObs(10 + obs_kf[o]);
            Keypoint kp;                                      // straight into the hash map: an observation may lie outside the image, the grid is not read
            kp.lmid_ = m;
            kp.px_ = cv::Point2f(obs_px[2 * o], obs_px[2 * o + 1]);
            kp.unpx_ = kp.px_;
            kfs[obs_kf[o]]->mapkps_[m] = kp;
        }
        map->map_plms_.emplace(m, lm);
    }
    Frame frame(calib, (size_t)ncellsize);
    frame.id_ = frame.kfid_ = 5000;
    frame.setTwc(pose_of(Tcw));
    for (int j = 0; j < nkps; ++j) frame.addKeypoint(cv::Point2f(kp_px[2 * j], kp_px[2 * j + 1]), kp_lm[j] >= 0 ? kp_lm[j] : 1000000 + j);
    std::unordered_set<int> local(cand_mp, cand_mp + ncand);
    int n = 0;
    for (const int id : local) order_out[n++] = id;
    Mapper mapper;
    mapper.pslamstate_ = params;
    mapper.pmap_ = map;
    const std::map<int, int> res = mapper.matchToMap(frame, fmaxprojerr, fdistratio, local);
    n = 0;
    for (const auto& kv : res) { pairs_out[2 * n] = kv.first; pairs_out[2 * n + 1] = kv.second; ++n; }
    return n;
}


// The REFERENCE'S OWN MapManager::stereoMatching (/root/reference/src/map_manager.cpp:367-611, with its FeatureTracker: fbKltTracking and
// getLineMinSAD of src/feature_tracker.cpp) on a synthetic stereo keyframe: keypoints (pixel, 3-D flag, map point or none), a rig, and
// the two image pyramids (levels 0-3, built by the caller with the real library; entry 2 * level of the reference's pyramid vectors).
// out: per keypoint, in Frame::getKeypoints() order: lmid, is_stereo, right pixel.
extern "C" int ov2ref_stereo_matching(int nkps, int rect, const double* K, const double* Kr, const double* Tc0c1, const double* Twc, int width, int height,
                                      int ncellsize, const int32_t* lmid, const int32_t* is3d, const int32_t* has_mp, const float* px, const double* wpt,
                                      const uint8_t* const* left_levels, const uint8_t* const* right_levels, const int32_t* level_rows, const int32_t* level_cols,
                                      int32_t* order_out, int32_t* is_stereo_out, float* rpx_out) {
    auto params = std::make_shared<SlamParams>();
    SlamParams& S = *params;
    S.debug_ = false; S.log_timings_ = false; S.stereo_ = true; S.mono_ = false;
    S.bdo_stereo_rect_ = rect != 0;
    S.nklt_pyr_lvl_ = 3; S.nklt_win_size_ = 9; S.nklt_err_ = 30; S.fmax_fbklt_dist_ = 0.5f;
    auto lcal = std::make_shared<CameraCalibration>("pinhole", K[0], K[1], K[2], K[3], 0., 0., 0., 0., (double)width, (double)height);
    auto rcal = std::make_shared<CameraCalibration>("pinhole", Kr[0], Kr[1], Kr[2], Kr[3], 0., 0., 0., 0., (double)width, (double)height);
    lcal->Dcv_.release();
    rcal->Dcv_.release();
    rcal->setupExtrinsic(Sophus::SE3d(Eigen::Quaterniond(Tc0c1[6], Tc0c1[3], Tc0c1[4], Tc0c1[5]), Eigen::Vector3d(Tc0c1[0], Tc0c1[1], Tc0c1[2])));
    auto cur = std::make_shared<Frame>(lcal, rcal, (size_t)ncellsize);
    auto ft = std::make_shared<FeatureTracker>(30, 0.01f, nullptr);
    auto map = std::make_shared<MapManager>(params, cur, nullptr, ft);
    Frame frame(lcal, rcal, (size_t)ncellsize);
    frame.id_ = frame.kfid_ = 7;
    frame.setTwc(Sophus::SE3d(Eigen::Quaterniond(Twc[6], Twc[3], Twc[4], Twc[5]), Eigen::Vector3d(Twc[0], Twc[1], Twc[2])));
    for (int j = 0; j < nkps; ++j) {
        frame.addKeypoint(cv::Point2f(px[2 * j], px[2 * j + 1]), lmid[j]);
        if (has_mp[j]) {
            auto lm = std::make_shared<MapPoint>(lmid[j], 7, true);
            lm->setPoint(Eigen::Vector3d(wpt[3 * j], wpt[3 * j + 1], wpt[3 * j + 2]));
            lm->is3d_ = true;
            map->map_plms_.emplace(lmid[j], lm);
        }
        if (is3d[j]) frame.turnKeypoint3d(lmid[j]);
    }
    std::vector<cv::Mat> lpyr(8), rpyr(8);
    for (int lv = 0; lv < 4; ++lv) {
        lpyr[2 * lv] = cv::Mat(level_rows[lv], level_cols[lv], CV_8UC1, (void*)left_levels[lv]);
        rpyr[2 * lv] = cv::Mat(level_rows[lv], level_cols[lv], CV_8UC1, (void*)right_levels[lv]);
    }
    int n = 0;
    for (const auto& kp : frame.getKeypoints()) order_out[n++] = kp.lmid_;
    map->stereoMatching(frame, lpyr, rpyr);
    for (int i = 0; i < n; ++i) {
        const Keypoint kp = frame.getKeypointById(order_out[i]);
        is_stereo_out[i] = kp.lmid_ == order_out[i] && kp.is_stereo_ ? 1 : 0;
        rpx_out[2 * i] = kp.rpx_.x; rpx_out[2 * i + 1] = kp.rpx_.y;
    }
    return n;
}


// The REFERENCE'S OWN MapManager::extractKeypoints (/root/reference/src/map_manager.cpp:286-341: describe the tracked keypoints on the raw
// image, detect new ones where the tracks left cells empty, describe those, add them to the frame and the map) - the caller of the
// detector and the descriptor that one front-end step mirrors.  in: tracked keypoints (each gets a map point); detector: 0 FAST grid,
// 1 single scale.  out: every keypoint of the frame afterwards: lmid, pixel, 32-byte descriptor (zero when none), has_desc.
extern "C" int ov2ref_extract_keypoints(const uint8_t* im, const uint8_t* imraw, int rows, int cols, const double* K, int nmaxdist, int detector, int nfast_th,
                                        double dmaxquality, const float* tracked, int ntracked, int cap, int32_t* lmid_out, float* px_out, uint8_t* desc_out,
                                        uint8_t* has_desc_out, int* fast_th_out, double* quality_out) {
    auto params = std::make_shared<SlamParams>();
    SlamParams& S = *params;
    S.debug_ = false; S.log_timings_ = false;
    S.use_brief_ = true; S.use_shi_tomasi_ = false; S.use_fast_ = detector == 0; S.use_singlescale_detector_ = detector == 1;
    S.nmaxdist_ = nmaxdist; S.nbmaxkps_ = 100000;
    auto calib = std::make_shared<CameraCalibration>("pinhole", K[0], K[1], K[2], K[3], 0., 0., 0., 0., (double)cols, (double)rows);
    calib->Dcv_.release();
    auto cur = std::make_shared<Frame>(calib, (size_t)nmaxdist);
    ov2ref_fe_destroy(ov2ref_fe_create(1000, nmaxdist, dmaxquality, nfast_th));       // resets the file-scope detector objects
    auto fe = std::make_shared<FeatureExtractor>(1000, nmaxdist, dmaxquality, nfast_th);
    auto map = std::make_shared<MapManager>(params, cur, fe, nullptr);
    cur->updateFrame(1, 0.05);
    for (int i = 0; i < ntracked; ++i) {
        cur->addKeypoint(cv::Point2f(tracked[2 * i], tracked[2 * i + 1]), map->nlmid_);
        map->addMapPoint();
    }
    cv::Mat mim(rows, cols, CV_8UC1, (void*)im), mraw(rows, cols, CV_8UC1, (void*)imraw);
    map->extractKeypoints(mim, mraw);
    *fast_th_out = fe->nfast_th_;
    *quality_out = fe->dmaxquality_;
    std::vector<Keypoint> kps = cur->getKeypoints();
    std::sort(kps.begin(), kps.end(), [](const Keypoint& a, const Keypoint& b) { return a.lmid_ < b.lmid_; });
    int n = 0;
    for (const auto& kp : kps) {
        if (n >= cap) break;
        lmid_out[n] = kp.lmid_;
        px_out[2 * n] = kp.px_.x; px_out[2 * n + 1] = kp.px_.y;
        has_desc_out[n] = kp.desc_.empty() ? 0 : 1;
        if (has_desc_out[n]) memcpy(desc_out + 32 * (size_t)n, kp.desc_.ptr(0), 32); else memset(desc_out + 32 * (size_t)n, 0, 32);
        ++n;
    }
    return (int)kps.size();
}


// The REFERENCE'S OWN MultiViewGeometry::ceresPnP (/root/reference/src/multi_view_geometry.cpp:492-588), called as
// VisualFrontEnd::computePose does (src/visual_front_end.cpp:791-801).  It caps the solve at 5 ms of wall clock: callers keep the
// problem small enough for that not to bind on this build.  Returns its verdict; outliers_out gets the rejected indices.
extern "C" int ov2ref_real_ceres_pnp(int n, const double* unpx, const double* wpts, const int32_t* scales, double* Twc, int nmaxiter, float chi2th, int use_robust,
                                     int apply_l2, const float* K, int32_t* outliers_out, int* noutliers) {
    std::vector<Eigen::Vector2d, Eigen::aligned_allocator<Eigen::Vector2d>> vunkps;
    std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>> vwpts;
    std::vector<int> vscales, vout;
    for (int i = 0; i < n; ++i) {
        vunkps.push_back(Eigen::Vector2d(unpx[2 * i], unpx[2 * i + 1]));
        vwpts.push_back(Eigen::Vector3d(wpts[3 * i], wpts[3 * i + 1], wpts[3 * i + 2]));
        vscales.push_back(scales ? scales[i] : 0);
    }
    Sophus::SE3d T(Eigen::Quaterniond(Twc[6], Twc[3], Twc[4], Twc[5]), Eigen::Vector3d(Twc[0], Twc[1], Twc[2]));
    const bool ok = MultiViewGeometry::ceresPnP(vunkps, vwpts, vscales, T, nmaxiter, chi2th, use_robust != 0, apply_l2 != 0, K[0], K[1], K[2], K[3], vout);
    const Eigen::Vector3d t = T.translation();
    const Eigen::Quaterniond q = T.unit_quaternion();
    Twc[0] = t.x(); Twc[1] = t.y(); Twc[2] = t.z(); Twc[3] = q.x(); Twc[4] = q.y(); Twc[5] = q.z(); Twc[6] = q.w();
    *noutliers = (int)vout.size();
    for (size_t i = 0; i < vout.size(); ++i) outliers_out[i] = vout[i];
    return ok ? 1 : 0;
}

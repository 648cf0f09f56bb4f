// MapManager::stereoMatching on the GPU (8f-3): the reference's declaration (include/map_manager.hpp:83) is kept, this translation
// unit supplies the body of /root/reference/src/map_manager.cpp:367-611.  Exclude the reference's own body with
// #ifndef OV2_EXTERNAL_STEREOMATCHING.  What runs where:
//   host (here)   prior generation - projection of the keypoint's own map point into the right image (:405-416), or, for
//                 non-rectified rigs, the inverse-distance-weighted depth of the 3-D neighbours in the surrounding grid cells pushed
//                 along the keypoint's bearing (:436-480) - list building, the retry of failed prior tracks on the full pyramid
//                 (:528-534), the epipolar test (row difference or Sampson distance, :579-600) and the map update;
//   device        the rectified-rig prior: FeatureTracker::getLineMinSAD for ALL the keypoints that need it in ONE call
//                 (FeatureTracker::getLineMinSADBatch -> ov2_line_min_sad, csrc/frontend_sad.cu) instead of one call per keypoint
//                 (:417-431); the two forward-backward KLT passes (FeatureTracker::fbKltTracking -> ov2_fb_klt), on the cached
//                 device pyramids of the two images (each uploaded once).
// Order of every list (and therefore of the tracker's inputs and of the map updates) is the reference's.
// Compile-checked and self-tested against the stand-in map classes (host/standin/ref/): host/stereo_selftest.cpp.
#include <cmath>
#include <iostream>
#include <vector>

#include "map_manager.hpp"             // the reference's, unchanged
#include "multi_view_geometry.hpp"
#include "feature_tracker.hpp"         // the drop-in class (host/feature_tracker.hpp)

void MapManager::stereoMatching(Frame &frame, const std::vector<cv::Mat> &vleftpyr, const std::vector<cv::Mat> &vrightpyr)
{
    // Find stereo correspondances with left kps
    auto vleftkps = frame.getKeypoints();
    size_t nbkps = vleftkps.size();

    const size_t nmaxpyrlvl = pslamstate_->nklt_pyr_lvl_ * 2;          // index of the coarsest image in the pyramid vectors
    const int winsize = 7;
    const float uppyrcoef = std::pow(2, pslamstate_->nklt_pyr_lvl_);
    const float downpyrcoef = 1. / uppyrcoef;

    std::vector<int> v3dkpids, vkpids;
    std::vector<cv::Point2f> v3dkps, v3dpriors, vkps, vpriors;
    v3dkpids.reserve(frame.nb3dkps_); v3dkps.reserve(frame.nb3dkps_); v3dpriors.reserve(frame.nb3dkps_);
    vkpids.reserve(nbkps); vkps.reserve(nbkps); vpriors.reserve(nbkps);

    // rectified rigs: the keypoints whose prior comes from the row search, and where their prior sits in vpriors
    std::vector<cv::Point2f> vsadpts;
    std::vector<size_t> vsadslot;
    const bool bsad = pslamstate_->bdo_stereo_rect_;
    if (bsad && (vleftpyr.size() <= nmaxpyrlvl || vrightpyr.size() <= nmaxpyrlvl)) {
        std::cerr << "[ov2b200] stereoMatching: the pyramids have no level " << pslamstate_->nklt_pyr_lvl_ << " for the row search\n";
        return;
    }

    for (size_t i = 0; i < nbkps; i++) {
        auto &kp = vleftkps.at(i);
        cv::Point2f priorpt = kp.px_;

        // If 3D, check if we can find a prior in right image (:405-420)
        if (kp.is3d_) {
            auto plm = getMapPoint(kp.lmid_);
            if (plm != nullptr) {
                cv::Point2f projpt = frame.projWorldToRightImageDist(plm->getPoint());
                if (frame.isInRightImage(projpt)) {
                    v3dkps.push_back(kp.px_);
                    v3dpriors.push_back(projpt);
                    v3dkpids.push_back(kp.lmid_);
                    continue;
                }
            } else {
                removeMapPointObs(kp.lmid_, frame.kfid_);
                continue;
            }
        }

        if (bsad) {
            // prior from the row search on the coarsest level: queued, resolved by one device call below (:422-435)
            vsadpts.push_back(cv::Point2f(kp.px_.x * downpyrcoef, kp.px_.y * downpyrcoef));
            vsadslot.push_back(vkps.size());
        } else {
            // prior from the depth of the 3-D neighbours (:437-480)
            const size_t nbmin3dcokps = 1;
            auto vnearkps = frame.getSurroundingKeypoints(kp);
            size_t nb3dkp = 0;
            double mean_z = 0., weights = 0.;
            for (const auto &cokp : vnearkps) {
                if (!cokp.is3d_) continue;
                auto plm = getMapPoint(cokp.lmid_);
                if (plm == nullptr) continue;
                nb3dkp++;
                const float dx = cokp.unpx_.x - kp.unpx_.x, dy = cokp.unpx_.y - kp.unpx_.y;       // cv::norm(Point2f): float difference,
                const double coef = 1. / std::sqrt((double)dx * dx + (double)dy * dy);             // double sum of squares
                weights += coef;
                mean_z += coef * frame.projWorldToCam(plm->getPoint()).z();
            }
            if (nb3dkp >= nbmin3dcokps) {
                mean_z /= weights;
                const Eigen::Vector3d predcampt = mean_z * (kp.bv_ / kp.bv_.z());
                cv::Point2f projpt = frame.projCamToRightImageDist(predcampt);
                if (frame.isInRightImage(projpt)) {
                    v3dkps.push_back(kp.px_);
                    v3dpriors.push_back(projpt);
                    v3dkpids.push_back(kp.lmid_);
                    continue;
                }
            }
        }
        vkpids.push_back(kp.lmid_);
        vkps.push_back(kp.px_);
        vpriors.push_back(priorpt);
    }

    if (!vsadpts.empty()) {
        std::vector<float> vxprior, vl1err;
        ptracker_->getLineMinSADBatch(vleftpyr, vrightpyr, pslamstate_->nklt_pyr_lvl_, vsadpts, winsize, true, vxprior, vl1err);
        for (size_t k = 0; k < vsadpts.size(); k++) {
            float xprior = vxprior[k];
            xprior *= uppyrcoef;
            const size_t slot = vsadslot[k];
            if (xprior >= 0 && xprior <= vkps[slot].x) vpriors[slot].x = xprior;       // :431-433
        }
    }

    // Storing good tracks
    std::vector<cv::Point2f> vgoodrkps;
    std::vector<int> vgoodids;
    vgoodrkps.reserve(nbkps);
    vgoodids.reserve(nbkps);

    // 1st track 3d kps if using prior (:497-541)
    if (!v3dpriors.empty()) {
        size_t nbpyrlvl = 1;
        int nwinsize = pslamstate_->nklt_win_size_;
        if (vleftpyr.size() < 2 * (nbpyrlvl + 1)) nbpyrlvl = vleftpyr.size() / 2 - 1;
        std::vector<bool> vkpstatus;
        ptracker_->fbKltTracking(vleftpyr, vrightpyr, nwinsize, nbpyrlvl, pslamstate_->nklt_err_, pslamstate_->fmax_fbklt_dist_,
                                 v3dkps, v3dpriors, vkpstatus);
        size_t nbgood = 0;
        const size_t nb3dkps = v3dkps.size();
        for (size_t i = 0; i < nb3dkps; i++) {
            if (vkpstatus.at(i)) {
                vgoodrkps.push_back(v3dpriors.at(i));
                vgoodids.push_back(v3dkpids.at(i));
                nbgood++;
            } else {
                // tracking failed: retry on the full pyramid with the 2-D keypoints
                vkpids.push_back(v3dkpids.at(i));
                vkps.push_back(v3dkps.at(i));
                vpriors.push_back(v3dpriors.at(i));
            }
        }
        if (pslamstate_->debug_)
            std::cout << "\n >>> Stereo KLT Tracking on priors : " << nbgood << " out of " << nb3dkps << " kps tracked!\n";
    }

    // 2nd track other kps if any (:544-572)
    if (!vkps.empty()) {
        std::vector<bool> vkpstatus;
        ptracker_->fbKltTracking(vleftpyr, vrightpyr, pslamstate_->nklt_win_size_, pslamstate_->nklt_pyr_lvl_, pslamstate_->nklt_err_,
                                 pslamstate_->fmax_fbklt_dist_, vkps, vpriors, vkpstatus);
        size_t nbgood = 0;
        const size_t nb2dkps = vkps.size();
        for (size_t i = 0; i < nb2dkps; i++) {
            if (vkpstatus.at(i)) {
                vgoodrkps.push_back(vpriors.at(i));
                vgoodids.push_back(vkpids.at(i));
                nbgood++;
            }
        }
        if (pslamstate_->debug_)
            std::cout << "\n >>> Stereo KLT Tracking w. no priors : " << nbgood << " out of " << nb2dkps << " kps tracked!\n";
    }

    nbkps = vgoodids.size();
    size_t nbgood = 0;
    float epi_err = 0.;
    for (size_t i = 0; i < nbkps; i++) {
        cv::Point2f lunpx = frame.getKeypointById(vgoodids.at(i)).unpx_;
        cv::Point2f runpx = frame.pcalib_rightcam_->undistortImagePoint(vgoodrkps.at(i));
        // Check epipolar consistency (same row for rectified images)
        if (pslamstate_->bdo_stereo_rect_) {
            epi_err = fabs(lunpx.y - runpx.y);
            vgoodrkps.at(i).y = lunpx.y;               // correct the right keypoint onto the same row
        } else {
            epi_err = MultiViewGeometry::computeSampsonDistance(frame.Frl_, lunpx, runpx);
        }
        if (epi_err <= 2.) {
            frame.updateKeypointStereo(vgoodids.at(i), vgoodrkps.at(i));
            nbgood++;
        }
    }
    if (pslamstate_->debug_)
        std::cout << "\n \t>>> Nb of stereo tracks: " << nbgood << " out of " << nbkps << "\n";
}

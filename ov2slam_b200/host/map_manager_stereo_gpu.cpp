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

namespace {

// keypoints queued for one tracker call: id, left pixel, starting point in the right image
struct TrackQueue {
    std::vector<int> ids;
    std::vector<cv::Point2f> left, right;
    void push(int id, const cv::Point2f& l, const cv::Point2f& r) { ids.push_back(id); left.push_back(l); right.push_back(r); }
    bool empty() const { return ids.empty(); }
    size_t size() const { return ids.size(); }
};

// the inverse-distance-weighted depth of the 3-D keypoints around `kp` (the four grid cells Frame::getSurroundingKeypoints visits);
// false when there is none (map_manager.cpp:437-463)
bool neighbour_depth(const MapManager& map, const Frame& frame, const Keypoint& kp, double& depth) {
    double zsum = 0., wsum = 0.;
    size_t used = 0;
    for (const auto& other : frame.getSurroundingKeypoints(kp)) {
        if (!other.is3d_) continue;
        const auto plm = map.getMapPoint(other.lmid_);
        if (plm == nullptr) continue;
        const float dx = other.unpx_.x - kp.unpx_.x, dy = other.unpx_.y - kp.unpx_.y;      // cv::norm(Point2f): float difference,
        const double wgt = 1. / std::sqrt((double)dx * dx + (double)dy * dy);               // double sum of squares
        wsum += wgt;
        zsum += wgt * frame.projWorldToCam(plm->getPoint()).z();
        used++;
    }
    if (used == 0) return false;
    depth = zsum / wsum;
    return true;
}

}  // namespace

void MapManager::stereoMatching(Frame &frame, const std::vector<cv::Mat> &vleftpyr, const std::vector<cv::Mat> &vrightpyr)
{
    const SlamParams& cfg = *pslamstate_;
    const bool rectified = cfg.bdo_stereo_rect_;
    const int coarsest = cfg.nklt_pyr_lvl_;                      // pyramid level of the row search (entry 2 * level of the vectors)
    const float up = std::pow(2, coarsest), down = 1. / up;
    if (rectified && (vleftpyr.size() <= (size_t)(2 * coarsest) || vrightpyr.size() <= (size_t)(2 * coarsest))) {
        std::cerr << "[ov2b200] stereoMatching: the pyramids have no level " << coarsest << " for the row search\n";
        return;
    }

    // ---- 1. every keypoint of the frame goes to one of two queues: `guided` (a prior from geometry exists: tracked on 2 levels
    //         first) or `blind` (full pyramid).  Keypoint order is kept inside each queue.                       (:395-493)
    TrackQueue guided, blind;
    std::vector<cv::Point2f> rowsearch_pts;                       // rectified rigs: coarsest-level pixels awaiting the row search
    std::vector<size_t> rowsearch_slot;                           // ... and the entry of `blind` each one belongs to
    for (const Keypoint& kp : frame.getKeypoints()) {
        if (kp.is3d_) {
            const auto plm = getMapPoint(kp.lmid_);
            if (plm == nullptr) {                                 // dangling observation: drop it, no stereo for it (:411-414)
                removeMapPointObs(kp.lmid_, frame.kfid_);
                continue;
            }
            const cv::Point2f proj = frame.projWorldToRightImageDist(plm->getPoint());
            if (frame.isInRightImage(proj)) { guided.push(kp.lmid_, kp.px_, proj); continue; }
        }
        if (rectified) {
            rowsearch_pts.push_back(cv::Point2f(kp.px_.x * down, kp.px_.y * down));
            rowsearch_slot.push_back(blind.size());
        } else {
            double depth;
            if (neighbour_depth(*this, frame, kp, depth)) {
                const Eigen::Vector3d campt = depth * (kp.bv_ / kp.bv_.z());
                const cv::Point2f proj = frame.projCamToRightImageDist(campt);
                if (frame.isInRightImage(proj)) { guided.push(kp.lmid_, kp.px_, proj); continue; }
            }
        }
        blind.push(kp.lmid_, kp.px_, kp.px_);
    }

    // ---- 2. rectified rigs: ONE device call finds the best column on the coarsest level for all queued keypoints; a hit to the
    //         left of the keypoint becomes the x of its starting point                                            (:417-435)
    if (!rowsearch_pts.empty()) {
        std::vector<float> col, err;
        ptracker_->getLineMinSADBatch(vleftpyr, vrightpyr, coarsest, rowsearch_pts, 7, true, col, err);
        for (size_t k = 0; k < col.size(); ++k) {
            const float x = col[k] * up;
            const size_t slot = rowsearch_slot[k];
            if (x >= 0 && x <= blind.left[slot].x) blind.right[slot].x = x;
        }
    }

    // ---- 3. the two forward-backward KLT passes; a guided track that fails is retried blind from where the tracker left it
    TrackQueue matched;                                            // id, (unused), right pixel
    auto track = [&](TrackQueue& q, int nlevels, TrackQueue* retry, const char* what) {
        if (q.empty()) return;
        std::vector<bool> ok;
        ptracker_->fbKltTracking(vleftpyr, vrightpyr, cfg.nklt_win_size_, nlevels, cfg.nklt_err_, cfg.fmax_fbklt_dist_, q.left, q.right, ok);
        size_t kept = 0;
        for (size_t i = 0; i < q.size(); ++i) {
            if (ok.at(i)) { matched.push(q.ids[i], q.left[i], q.right[i]); kept++; }
            else if (retry) retry->push(q.ids[i], q.left[i], q.right[i]);
        }
        if (cfg.debug_) std::cout << "\n >>> Stereo KLT Tracking " << what << " : " << kept << " out of " << q.size() << " kps tracked!\n";
    };
    int guided_levels = 1;                                         // 2 levels (:499-504)
    if (vleftpyr.size() < 2 * (size_t)(guided_levels + 1)) guided_levels = (int)(vleftpyr.size() / 2) - 1;
    track(guided, guided_levels, &blind, "on priors");
    track(blind, cfg.nklt_pyr_lvl_, nullptr, "w. no priors");

    // ---- 4. epipolar gate and map update (:575-604)
    size_t nstereo = 0;
    for (size_t i = 0; i < matched.size(); ++i) {
        const cv::Point2f lun = frame.getKeypointById(matched.ids[i]).unpx_;
        const cv::Point2f run = frame.pcalib_rightcam_->undistortImagePoint(matched.right[i]);
        float epi;
        if (rectified) {
            epi = fabs(lun.y - run.y);
            matched.right[i].y = lun.y;                            // same row by construction of the rig
        } else {
            epi = MultiViewGeometry::computeSampsonDistance(frame.Frl_, lun, run);
        }
        if (epi <= 2.) {
            frame.updateKeypointStereo(matched.ids[i], matched.right[i]);
            nstereo++;
        }
    }
    if (cfg.debug_) std::cout << "\n \t>>> Nb of stereo tracks: " << nstereo << " out of " << matched.size() << "\n";
}

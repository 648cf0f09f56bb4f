// The reference's include/optimizer.hpp (:33-67) class declaration, member for member: this IS the drop-in boundary
// (estimator.cpp, mapper.cpp, loop_closer.cpp, ov2slam.cpp compile against it unchanged).  Stand-in copy for the
// container's compile check; a real build includes the reference's own header.
#pragma once

#include <deque>
#include <vector>

#include "map_manager.hpp"

class Optimizer {

public:
    Optimizer(std::shared_ptr<SlamParams> pslamstate, std::shared_ptr<MapManager> pmap)
        : pslamstate_(pslamstate), pmap_(pmap), bstop_localba_(false)
    {}

    void localBA(Frame &newframe, const bool buse_robust_cost);

    void looseBA(const int inikfid, const int nkfid, const bool buse_robust_cost);

    void fullBA(const bool buse_robust_cost);

    void signalStopLocalBA();
    bool stopLocalBA();

    bool localPoseGraph(Frame &newframe, int kfloop_id, const Sophus::SE3d &newTwc);

    void structureOnlyBA(const std::vector<int> &vlm2optids);

    bool fullPoseGraph(std::vector<Sophus::SE3d, Eigen::aligned_allocator<Sophus::SE3d>> &vTwc,
        std::vector<Sophus::SE3d, Eigen::aligned_allocator<Sophus::SE3d>> &vTpc,
        std::vector<bool> &viskf);

    std::shared_ptr<SlamParams> pslamstate_;
    std::shared_ptr<MapManager> pmap_;

    bool bstop_localba_;

    std::mutex localba_mutex_;
};

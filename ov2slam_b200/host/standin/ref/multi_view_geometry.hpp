// The two MultiViewGeometry::ceresPnP declarations of the reference's include/multi_view_geometry.hpp (:82-93), the part of
// that class the per-frame pose refinement links against (src/visual_front_end.cpp:791-801).  Stand-in copy for the
// container's compile check; a real build includes the reference's own header (Eigen::Vector2d comes from Eigen there).
#pragma once
#include <vector>

#include <Eigen/Dense>
#include <sophus/se3.hpp>

namespace Eigen {
struct Vector2d {
    double v[2];
    Vector2d() : v{0, 0} {}
    Vector2d(double x, double y) : v{x, y} {}
    double x() const { return v[0]; }
    double y() const { return v[1]; }
};
}  // namespace Eigen

class MultiViewGeometry {
public:
    static bool ceresPnP(const std::vector<Eigen::Vector2d, Eigen::aligned_allocator<Eigen::Vector2d> > &vunkps,
                        const std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d> > &vwpts,
                        Sophus::SE3d &Twc,
                        const int nmaxiter, const float chi2th, const bool buse_robust, const bool bapply_l2_after_robust,
                        const float fx, const float fy, const float cx, const float cy, std::vector<int> &voutliersidx);

    static bool ceresPnP(const std::vector<Eigen::Vector2d, Eigen::aligned_allocator<Eigen::Vector2d> > &vunkps,
                        const std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d> > &vwpts,
                        const std::vector<int> &vscales,
                        Sophus::SE3d &Twc,
                        const int nmaxiter, const float chi2th, const bool buse_robust, const bool bapply_l2_after_robust,
                        const float fx, const float fy, const float cx, const float cy, std::vector<int> &voutliersidx);
};

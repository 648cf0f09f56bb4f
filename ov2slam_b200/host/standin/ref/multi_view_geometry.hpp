// The two MultiViewGeometry::ceresPnP declarations of the reference's include/multi_view_geometry.hpp (:82-93), the part of
// that class the per-frame pose refinement links against (src/visual_front_end.cpp:791-801).  Stand-in copy for the
// container's compile check; a real build includes the reference's own header (Eigen::Vector2d comes from Eigen there).
#pragma once
#include <cmath>
#include <vector>

#include <opencv2/core.hpp>

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
    // multi_view_geometry.cpp:798-821 (float intermediates as there)
    static float computeSampsonDistance(const Eigen::Matrix3d &Frl, const cv::Point2f &leftpt, const cv::Point2f &rightpt) {
        const Eigen::Vector3d l(leftpt.x, leftpt.y, 1.), r(rightpt.x, rightpt.y, 1.);
        const Eigen::Vector3d Fl = Frl * l, Ftr = Frl.transpose() * r;
        float num = (float)(r.x() * Fl.x() + r.y() * Fl.y() + r.z() * Fl.z());
        num *= num;
        const float x1 = (float)Ftr.x(), x2 = (float)Fl.x(), y1 = (float)Ftr.y(), y2 = (float)Fl.y();
        const float den = x1 * x1 + y1 * y1 + x2 * x2 + y2 * y2;
        return std::sqrt(num / den);
    }

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

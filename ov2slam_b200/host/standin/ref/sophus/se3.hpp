// Minimal STAND-IN for Sophus::SE3d (see Eigen/Dense next to it).
#pragma once
#include <Eigen/Dense>

namespace Sophus {

class SE3d {
public:
    SE3d() {}
    SE3d(const Eigen::Quaterniond& q, const Eigen::Vector3d& t) : q_(q), t_(t) { q_.normalize(); }
    const Eigen::Quaterniond& unit_quaternion() const { return q_; }
    const Eigen::Vector3d& translation() const { return t_; }
    Eigen::Matrix3d rotationMatrix() const { return q_.toRotationMatrix(); }
    Eigen::Vector3d operator*(const Eigen::Vector3d& p) const { return rotationMatrix() * p + t_; }
    SE3d operator*(const SE3d& o) const { return SE3d(q_ * o.q_, rotationMatrix() * o.t_ + t_); }
    SE3d inverse() const {
        const Eigen::Quaterniond qi = q_.conjugate();
        const Eigen::Vector3d ti = qi.toRotationMatrix() * t_;
        return SE3d(qi, Eigen::Vector3d(-ti.x(), -ti.y(), -ti.z()));
    }
private:
    Eigen::Quaterniond q_;
    Eigen::Vector3d t_;
};

}  // namespace Sophus

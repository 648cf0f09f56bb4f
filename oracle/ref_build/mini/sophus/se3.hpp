// TEST INFRASTRUCTURE (oracle/): SO3d / SE3d with Sophus' spelling and Sophus 1.1's formulas, restated for the mini linear-algebra
// header next to it (Sophus itself is header-only but needs the real Eigen).  Followed: Thirdparty/Sophus/sophus/so3.hpp
// (:247-290 log, :329-343 product, :362-371 point action, :585-620 exp), se3.hpp (:103-111 Adj, :223-256 log, :763-784 exp).
#pragma once
#include <cmath>

#include <Eigen/Core>

namespace Sophus {

template <class S> struct Constants {
    static S epsilon() { return S(1e-10); }
    static S pi() { return S(3.141592653589793238462643383279502884); }
};

class SO3d {
public:
    SO3d() {}
    template <class Q> explicit SO3d(const Q& q) : q_(q.w(), q.x(), q.y(), q.z()) { q_.normalize(); }
    static Eigen::Matrix3d hat(const Eigen::Vector3d& o) {
        Eigen::Matrix3d m;
        m << 0., -o(2), o(1),
             o(2), 0., -o(0),
             -o(1), o(0), 0.;
        return m;
    }
    static SO3d expAndTheta(const Eigen::Vector3d& omega, double* theta) {
        const double theta_sq = omega.squaredNorm();
        double imag_factor, real_factor;
        if (theta_sq < Constants<double>::epsilon() * Constants<double>::epsilon()) {
            *theta = 0.;
            const double theta_po4 = theta_sq * theta_sq;
            imag_factor = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
            real_factor = 1. - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
        } else {
            *theta = std::sqrt(theta_sq);
            const double half_theta = 0.5 * (*theta);
            imag_factor = std::sin(half_theta) / (*theta);
            real_factor = std::cos(half_theta);
        }
        SO3d r;
        r.q_ = Eigen::Quaterniond(real_factor, imag_factor * omega.x(), imag_factor * omega.y(), imag_factor * omega.z());   // not re-normalised
        return r;
    }
    static SO3d exp(const Eigen::Vector3d& omega) { double th; return expAndTheta(omega, &th); }
    struct TangentAndTheta { Eigen::Vector3d tangent; double theta; };
    TangentAndTheta logAndTheta() const {
        TangentAndTheta J;
        const double squared_n = q_.vec().squaredNorm();
        const double w = q_.w();
        double two_atan_nbyw_by_n;
        if (squared_n < Constants<double>::epsilon() * Constants<double>::epsilon()) {
            const double squared_w = w * w;
            two_atan_nbyw_by_n = 2. / w - (2.0 / 3.0) * squared_n / (w * squared_w);
            J.theta = 2. * squared_n / w;
        } else {
            const double n = std::sqrt(squared_n);
            if (std::fabs(w) < Constants<double>::epsilon()) two_atan_nbyw_by_n = (w > 0. ? 1. : -1.) * Constants<double>::pi() / n;
            else two_atan_nbyw_by_n = 2. * std::atan(n / w) / n;
            J.theta = two_atan_nbyw_by_n * n;
        }
        J.tangent = two_atan_nbyw_by_n * q_.vec();
        return J;
    }
    Eigen::Vector3d log() const { return logAndTheta().tangent; }
    const Eigen::Quaterniond& unit_quaternion() const { return q_; }
    Eigen::Matrix3d matrix() const { return q_.toRotationMatrix(); }
    Eigen::Matrix3d Adj() const { return matrix(); }
    SO3d inverse() const { SO3d r; r.q_ = q_.conjugate(); return r; }
    SO3d operator*(const SO3d& o) const {
        const Eigen::Quaterniond &a = q_, &b = o.q_;
        return SO3d(Eigen::Quaterniond(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                                       a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                                       a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                                       a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x()));      // the product type normalises
    }
    template <class P> Eigen::Vector3d operator*(const Eigen::MatrixBase<P>& pin) const {
        const Eigen::Vector3d p = pin;
        Eigen::Vector3d uv = q_.vec().cross(p);
        uv += uv;
        return p + q_.w() * uv + q_.vec().cross(uv);
    }
private:
    Eigen::Quaterniond q_;
};

class SE3d {
public:
    typedef Eigen::Matrix<double, 6, 1> Tangent;
    SE3d() {}
    SE3d(const SO3d& so3, const Eigen::Vector3d& t) : so3_(so3), t_(t) {}
    template <class Q, class V> SE3d(const Q& q, const Eigen::MatrixBase<V>& t) : so3_(q), t_(t) {}
    static SE3d exp(const Tangent& a) {
        const Eigen::Vector3d omega = a.tail<3>();
        double theta;
        const SO3d so3 = SO3d::expAndTheta(omega, &theta);
        const Eigen::Matrix3d Omega = SO3d::hat(omega);
        const Eigen::Matrix3d Omega_sq = Omega * Omega;
        Eigen::Matrix3d V;
        if (theta < Constants<double>::epsilon()) {
            V = so3.matrix();
        } else {
            const double theta_sq = theta * theta;
            V = (Eigen::Matrix3d::Identity() + (1. - std::cos(theta)) / (theta_sq)*Omega + (theta - std::sin(theta)) / (theta_sq * theta) * Omega_sq);
        }
        return SE3d(so3, V * a.head<3>());
    }
    template <class M> static SE3d exp(const Eigen::MatrixBase<M>& a) { return exp(Tangent(a)); }
    Tangent log() const {
        Tangent upsilon_omega;
        const auto omega_and_theta = so3_.logAndTheta();
        const double theta = omega_and_theta.theta;
        upsilon_omega.tail<3>() = omega_and_theta.tangent;
        const Eigen::Matrix3d Omega = SO3d::hat(omega_and_theta.tangent);
        if (std::fabs(theta) < Constants<double>::epsilon()) {
            const Eigen::Matrix3d V_inv = Eigen::Matrix3d::Identity() - 0.5 * Omega + (1. / 12.) * (Omega * Omega);
            upsilon_omega.head<3>() = V_inv * t_;
        } else {
            const double half_theta = 0.5 * theta;
            const Eigen::Matrix3d V_inv = (Eigen::Matrix3d::Identity() - 0.5 * Omega +
                                           (1. - theta * std::cos(half_theta) / (2. * std::sin(half_theta))) / (theta * theta) * (Omega * Omega));
            upsilon_omega.head<3>() = V_inv * t_;
        }
        return upsilon_omega;
    }
    Eigen::Matrix<double, 6, 6> Adj() const {
        const Eigen::Matrix3d R = so3_.matrix();
        Eigen::Matrix<double, 6, 6> res;
        res.block<3, 3>(0, 0) = R;
        res.block<3, 3>(3, 3) = R;
        res.block<3, 3>(0, 3) = SO3d::hat(t_) * R;
        return res;
    }
    SE3d inverse() const { const SO3d invR = so3_.inverse(); return SE3d(invR, invR * (t_ * -1.)); }
    SE3d operator*(const SE3d& o) const { return SE3d(so3_ * o.so3_, t_ + so3_ * o.t_); }
    template <class P> Eigen::Vector3d operator*(const Eigen::MatrixBase<P>& p) const { return so3_ * p + t_; }
    Eigen::Matrix3d rotationMatrix() const { return so3_.matrix(); }
    const Eigen::Vector3d& translation() const { return t_; }
    const Eigen::Quaterniond& unit_quaternion() const { return so3_.unit_quaternion(); }
    const SO3d& so3() const { return so3_; }
private:
    SO3d so3_;
    Eigen::Vector3d t_;
};

}  // namespace Sophus

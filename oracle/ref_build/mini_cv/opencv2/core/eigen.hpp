// TEST INFRASTRUCTURE (oracle/): cv::eigen2cv / cv::cv2eigen between the stand-in matrix types (see ../core.hpp, mini/Eigen/Core).
#pragma once
#include <Eigen/Core>
#include <opencv2/core.hpp>

namespace cv {
template <class D> inline void eigen2cv(const Eigen::MatrixBase<D>& src, Mat& dst) {
    dst.create(src.rows(), src.cols(), CV_64F);
    for (int r = 0; r < src.rows(); ++r) for (int c = 0; c < src.cols(); ++c) dst.at<double>(r, c) = src.coeff(r, c);
}
template <class D> inline void cv2eigen(const Mat& src, Eigen::MatrixBase<D>& dst) {
    for (int r = 0; r < src.rows; ++r) for (int c = 0; c < src.cols; ++c) dst(r, c) = src.type() == CV_64F ? src.at<double>(r, c) : (double)src.at<float>(r, c);
}
}  // namespace cv

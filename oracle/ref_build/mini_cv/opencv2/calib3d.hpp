// TEST INFRASTRUCTURE (oracle/): calibration functions the reference's camera_calibration.cpp names.  None lies on the checked paths
// (undistorted pinhole cameras, maps set up directly): they are declared so that the file compiles and abort when reached.
#pragma once
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
#include <opencv2/features2d.hpp>

namespace cv {
enum { INTER_LINEAR = 1, BORDER_CONSTANT = 0, CALIB_ZERO_DISPARITY = 1024 };
inline Mat getOptimalNewCameraMatrix(const Mat&, const Mat&, Size, double, Size = Size(), Rect* = nullptr, bool = false) { mini_cv_missing("getOptimalNewCameraMatrix"); return Mat(); }
inline void initUndistortRectifyMap(const Mat&, const Mat&, const Mat&, const Mat&, Size, int, Mat&, Mat&) { mini_cv_missing("initUndistortRectifyMap"); }
inline void undistortPoints(const std::vector<Point2f>&, std::vector<Point2f>&, const Mat&, const Mat&, const Mat& = Mat(), const Mat& = Mat()) { mini_cv_missing("undistortPoints"); }
inline void projectPoints(const std::vector<Point3f>&, const Mat&, const Mat&, const Mat&, const Mat&, std::vector<Point2f>&) { mini_cv_missing("projectPoints"); }
inline void remap(const Mat&, Mat&, const Mat&, const Mat&, int, int = 0, const Scalar& = Scalar()) { mini_cv_missing("remap"); }
inline void stereoRectify(const Mat&, const Mat&, const Mat&, const Mat&, Size, const Mat&, const Mat&, Mat&, Mat&, Mat&, Mat&, Mat&, int = 0, double = -1, Size = Size(), Rect* = nullptr, Rect* = nullptr) { mini_cv_missing("stereoRectify"); }
enum { SOLVEPNP_ITERATIVE = 0, SOLVEPNP_P3P = 2, RANSAC = 8, LMEDS = 4 };
inline void triangulatePoints(const Matx34f&, const Matx34f&, const std::vector<Point2f>&, const std::vector<Point2f>&, Mat&) { mini_cv_missing("triangulatePoints"); }
inline bool solvePnPRansac(const std::vector<Point3f>&, const std::vector<Point2f>&, const Mat&, const Mat&, Mat&, Mat&, bool = false, int = 100, float = 8.f, double = 0.99,
                           Mat& = *(Mat*)nullptr, int = 0) { mini_cv_missing("solvePnPRansac"); return false; }
inline bool solvePnP(const std::vector<Point3f>&, const std::vector<Point2f>&, const Mat&, const Mat&, Mat&, Mat&, bool = false, int = 0) { mini_cv_missing("solvePnP"); return false; }
inline void Rodrigues(const Mat&, Mat&) { mini_cv_missing("Rodrigues"); }
inline Mat findEssentialMat(const std::vector<Point2f>&, const std::vector<Point2f>&, const Mat&, int, double, double, Mat&) { mini_cv_missing("findEssentialMat"); return Mat(); }
inline int recoverPose(const Mat&, const std::vector<Point2f>&, const std::vector<Point2f>&, const Mat&, Mat&, Mat&, Mat&) { mini_cv_missing("recoverPose"); return 0; }
namespace fisheye {
inline void estimateNewCameraMatrixForUndistortRectify(const Mat&, const Mat&, Size, const Mat&, Mat&, double = 0, Size = Size(), double = 1) { mini_cv_missing("fisheye::estimateNewCameraMatrixForUndistortRectify"); }
inline void initUndistortRectifyMap(const Mat&, const Mat&, const Mat&, const Mat&, Size, int, Mat&, Mat&) { mini_cv_missing("fisheye::initUndistortRectifyMap"); }
inline void undistortPoints(const std::vector<Point2f>&, std::vector<Point2f>&, const Mat&, const Mat&, const Mat& = Mat(), const Mat& = Mat()) { mini_cv_missing("fisheye::undistortPoints"); }
inline void projectPoints(const std::vector<Point3f>&, std::vector<Point2f>&, const Mat&, const Mat&, const Mat&, const Mat&, double = 0) { mini_cv_missing("fisheye::projectPoints"); }
inline void stereoRectify(const Mat&, const Mat&, const Mat&, const Mat&, Size, const Mat&, const Mat&, Mat&, Mat&, Mat&, Mat&, Mat&, int, Size = Size(), double = 0, double = 1) { mini_cv_missing("fisheye::stereoRectify"); }
inline void distortPoints(const std::vector<Point2f>&, std::vector<Point2f>&, const Mat&, const Mat&, double = 0) { mini_cv_missing("fisheye::distortPoints"); }
}  // namespace fisheye
}  // namespace cv

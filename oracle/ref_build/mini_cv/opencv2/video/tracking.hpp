// TEST INFRASTRUCTURE (oracle/): see core.hpp - pyramidal Lucas-Kanade forwarded to the real OpenCV through a callback.
#pragma once
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
#include <opencv2/features2d.hpp>

namespace cv {

enum { OPTFLOW_USE_INITIAL_FLOW = 4, OPTFLOW_LK_GET_MIN_EIGENVALS = 8 };

// prev / next are the reference's pyramids (image, derivative, image, derivative ...): entry 0 is the 8-bit image; the real library
// re-derives exactly those pyramids from it (what buildOpticalFlowPyramid produced)
inline void calcOpticalFlowPyrLK(const std::vector<Mat>& prev, const std::vector<Mat>& next, const std::vector<Point2f>& prevPts, std::vector<Point2f>& nextPts,
                                 std::vector<unsigned char>& status, std::vector<float>& err, Size win = Size(21, 21), int maxLevel = 3,
                                 TermCriteria crit = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 30, 0.01), int flags = 0,
                                 double minEigThreshold = 1e-4) {
    const int n = (int)prevPts.size();
    status.assign(n, 0);
    err.assign(n, 0.f);
    if (nextPts.size() != prevPts.size()) nextPts = prevPts;
    if (n == 0) return;
    const Mat &p = prev.at(0), &q = next.at(0);
    mini_cv_callbacks().lk(p.data, q.data, p.rows, p.cols, p.step, q.step, &prevPts[0].x, &nextPts[0].x, n, status.data(), err.data(), win.width, maxLevel,
                           crit.maxCount, crit.epsilon, flags, minEigThreshold);
}

// the pyramid vector as the reference keeps it: 2 * (maxLevel + 1) entries (image, derivative, ...); only entry 0 carries pixels here,
// the tracker callback re-derives the levels from it
inline int buildOpticalFlowPyramid(const Mat& img, std::vector<Mat>& pyr, Size, int maxLevel, bool = true, int = 4, int = 0, bool = true) {
    pyr.assign(2 * (size_t)(maxLevel + 1), Mat());
    pyr[0] = img;
    return maxLevel;
}

}  // namespace cv

// TEST INFRASTRUCTURE (oracle/): see core.hpp - image-processing calls forwarded to the real OpenCV through callbacks.
#pragma once
#include "core.hpp"

namespace cv {

class CLAHE { public: virtual ~CLAHE() {} virtual void apply(InputArray, OutputArray) = 0; };

inline void circle(Mat& img, Point center, int radius, const Scalar& color, int thickness = 1) {
    mini_cv_callbacks().circle(img.data, img.rows, img.cols, img.step, img.type(), center.x, center.y, radius, color.val[0], thickness);
}
inline void cornerSubPix(const Mat& im, std::vector<Point2f>& pts, Size win, Size zero, TermCriteria crit) {
    if (pts.empty()) return;
    mini_cv_callbacks().corner_subpix(im.data, im.rows, im.cols, im.step, &pts[0].x, (int)pts.size(), win.width, zero.width, crit.maxCount, crit.epsilon);
}
inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigma) {
    dst.create(src.rows, src.cols, CV_8U);
    const unsigned char* parent = src.parent_data ? src.parent_data : src.data;
    const size_t off = (size_t)(src.data - parent);
    mini_cv_callbacks().gaussian_blur(parent, src.parent_data ? src.parent_rows : src.rows, src.parent_data ? src.parent_cols : src.cols, src.step,
                                      (int)(off % src.step), (int)(off / src.step), src.rows, src.cols, dst.data, ksize.width, sigma);
}
inline void cornerMinEigenVal(const Mat& src, Mat& dst, int block, int ksize = 3) {
    dst.create(src.rows, src.cols, CV_32F);
    mini_cv_callbacks().corner_min_eigen(src.data, src.rows, src.cols, src.step, (float*)dst.data, block, ksize);
}
inline void getRectSubPix(const Mat& im, Size patch, Point2f center, Mat& out) {
    out.create(patch.height, patch.width, CV_8U);
    mini_cv_callbacks().get_rect_subpix(im.data, im.rows, im.cols, im.step, patch.width, patch.height, center.x, center.y, out.data);
}
inline void minMaxLoc(const Mat& m, double* minv, double* maxv, Point* minloc, Point* maxloc) {
    double a, b;
    int p[2], q[2];
    mini_cv_callbacks().min_max_loc((const float*)m.data, m.rows, m.cols, m.step, &a, &b, p, q);
    if (minv) *minv = a;
    if (maxv) *maxv = b;
    if (minloc) *minloc = Point(p[0], p[1]);
    if (maxloc) *maxloc = Point(q[0], q[1]);
}

}  // namespace cv

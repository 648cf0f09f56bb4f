// TEST INFRASTRUCTURE (oracle/): OpenCV's core types with OpenCV's spelling, just large enough to compile the reference's OWN
// front-end source (/root/reference/src/feature_extractor.cpp with the headers it includes) in this container, which has the OpenCV
// library only as the Python module cv2.  Containers and glue are written here; every ARITHMETIC call (FAST, cornerSubPix, GaussianBlur,
// cornerMinEigenVal, minMaxLoc, circle, ORB::compute) is forwarded through a table of callbacks that tests fill with the real cv2
// functions (oracle/ref_build/cv_callbacks.py) - so the reference's control flow runs on the real OpenCV's numbers.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#define OV2_STANDIN_OPENCV 1      // the drop-in classes take their stand-in branch on these containers
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
#define CV_64FC1 6

typedef unsigned char uchar;

namespace cv {

inline int cvRound(double v) { return (int)std::lrint(v); }
template <class T> inline T saturate_cast(float v) { return (T)v; }
template <> inline int saturate_cast<int>(float v) { return cvRound(v); }
template <class T> inline T saturate_cast(double v) { return (T)v; }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <class T> inline T saturate_cast(int v) { return (T)v; }

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename T2> operator Point_<T2>() const { return Point_<T2>(saturate_cast<T2>(x), saturate_cast<T2>(y)); }
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
    bool operator!=(const Point_& o) const { return !(*this == o); }
    Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); }
    Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); }
    Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
};
template <typename T> inline Point_<T> operator*(const Point_<T>& a, float s) { return Point_<T>(saturate_cast<T>(a.x * s), saturate_cast<T>(a.y * s)); }
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
template <typename T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }

template <typename T> struct Size_ { T width, height; Size_() : width(0), height(0) {} Size_(T w, T h) : width(w), height(h) {} };
typedef Size_<int> Size;
template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    Rect_(const Point_<T>& a, const Point_<T>& b) : x(std::min(a.x, b.x)), y(std::min(a.y, b.y)), width(std::max(a.x, b.x) - std::min(a.x, b.x)), height(std::max(a.y, b.y) - std::min(a.y, b.y)) {}
    friend std::ostream& operator<<(std::ostream& os, const Rect_& r) { return os << "[" << r.width << " x " << r.height << " from (" << r.x << ", " << r.y << ")]"; }
    bool empty() const { return width <= 0 || height <= 0; }
};
typedef Rect_<int> Rect;
struct Scalar {
    double val[4];
    Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) : val{v0, v1, v2, v3} {}
    double operator[](int i) const { return val[i]; }
    double& operator[](int i) { return val[i]; }
};
struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(-2147483647 - 1, 2147483647); }
};
struct TermCriteria {
    enum { COUNT = 1, MAX_ITER = 1, EPS = 2 };
    int type, maxCount;
    double epsilon;
    TermCriteria(int t = 0, int c = 0, double e = 0) : type(t), maxCount(c), epsilon(e) {}
};
class FileStorage;      // named by the reference's SlamParams constructor (YAML loading: not compiled here)
template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

// single-channel 8-bit / float / double matrix with shared, reference-counted storage and ROI views
class Mat {
public:
    int rows = 0, cols = 0, flags = 0;
    size_t step = 0;
    unsigned char* data = nullptr;
    // the matrix a ROI view was cut from (OpenCV keeps datastart / dataend for this: filters read the PARENT's pixels beyond a ROI's edge)
    unsigned char* parent_data = nullptr;
    int parent_rows = 0, parent_cols = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); setTo(s); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : rows(r), cols(c), flags(type), step(st ? st : (size_t)c * esz(type)), data((unsigned char*)ext) {}
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == flags && step == (size_t)c * esz(type)) return;
        rows = r; cols = c; flags = type; step = (size_t)c * esz(type);
        buf_.reset(new unsigned char[(size_t)r * step + 8], std::default_delete<unsigned char[]>());
        data = buf_.get();
        memset(data, 0, (size_t)r * step);
    }
    static size_t esz(int type) { return type == CV_8U ? 1 : (type == CV_32F ? 4 : 8); }
    int type() const { return flags; }
    size_t elemSize() const { return esz(flags); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    // OpenCV's Mat::size is a member OBJECT that is callable and streamable
    struct MatSize {
        const Mat* m;
        Size operator()() const { return Size(m->cols, m->rows); }
        friend std::ostream& operator<<(std::ostream& os, const MatSize& s) { return os << s.m->cols << " x " << s.m->rows; }
    };
    MatSize size{this};
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), flags(o.flags), step(o.step), data(o.data), parent_data(o.parent_data), parent_rows(o.parent_rows),
                        parent_cols(o.parent_cols), buf_(o.buf_) {}
    Mat& operator=(const Mat& o) {
        rows = o.rows; cols = o.cols; flags = o.flags; step = o.step; data = o.data; parent_data = o.parent_data; parent_rows = o.parent_rows;
        parent_cols = o.parent_cols; buf_ = o.buf_;
        return *this;
    }
    void release() { *this = Mat(); }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, flags);
        for (int r = 0; r < rows; ++r) memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& o) const { o = clone(); }
    unsigned char* ptr(int r = 0) { return data + (size_t)r * step; }
    const unsigned char* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <class T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (size_t)i * step); }
    template <class T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step); }
    Mat col(int c) const { return (*this)(Rect(c, 0, 1, rows)); }
    Mat& operator/=(double d) {
        for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) { if (flags == CV_32F) at<float>(i, j) = (float)(at<float>(i, j) / d); else if (flags == CV_64F) at<double>(i, j) /= d; }
        return *this;
    }
    static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < r && i < c; ++i) { if (type == CV_32F) m.at<float>(i, i) = 1.f; else if (type == CV_64F) m.at<double>(i, i) = 1.0; else m.at<unsigned char>(i, i) = 1; } return m; }
    template <class T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
    template <class T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
    Mat operator()(const Rect& r) const {
        Mat m;
        m.rows = r.height; m.cols = r.width; m.flags = flags; m.step = step;
        m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
        m.buf_ = buf_;
        m.parent_data = parent_data ? parent_data : data;
        m.parent_rows = parent_data ? parent_rows : rows;
        m.parent_cols = parent_data ? parent_cols : cols;
        return m;
    }
    Mat operator()(const Range& rr, const Range& cr) const {
        const int r0 = rr.start < 0 ? 0 : rr.start, r1 = rr.end > rows ? rows : rr.end, c0 = cr.start < 0 ? 0 : cr.start, c1 = cr.end > cols ? cols : cr.end;
        return (*this)(Rect(c0, r0, c1 - c0, r1 - r0));
    }
    Mat row(int r) const { return (*this)(Rect(0, r, cols, 1)); }
    Mat& operator=(const Scalar& s) { return setTo(s); }            // fills (a ROI of) the matrix
    friend std::ostream& operator<<(std::ostream& os, const Mat& m) {
        for (int r = 0; r < m.rows; ++r) {
            for (int c = 0; c < m.cols; ++c) os << (c ? ", " : (r ? "\n " : "[")) << (m.flags == CV_8U ? (double)m.at<unsigned char>(r, c) : (m.flags == CV_32F ? (double)m.at<float>(r, c) : m.at<double>(r, c)));
        }
        return os << "]";
    }
    Mat& setTo(const Scalar& s) {
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                if (flags == CV_8U) at<unsigned char>(r, c) = (unsigned char)s.val[0];
                else if (flags == CV_32F) at<float>(r, c) = (float)s.val[0];
                else at<double>(r, c) = s.val[0];
            }
        return *this;
    }
    static Mat ones(int r, int c, int type) { return Mat(r, c, type, Scalar(1)); }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    Mat mul(const Mat& o) const {                     // per-element product (float matrices on this path)
        Mat m(rows, cols, flags);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                if (flags == CV_32F) m.at<float>(r, c) = at<float>(r, c) * o.at<float>(r, c);
                else if (flags == CV_64F) m.at<double>(r, c) = at<double>(r, c) * o.at<double>(r, c);
                else m.at<unsigned char>(r, c) = (unsigned char)(at<unsigned char>(r, c) * o.at<unsigned char>(r, c));
            }
        return m;
    }
private:
    std::shared_ptr<unsigned char> buf_;
};
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
inline Mat operator*(double s, const Mat& m) {
    Mat r = m.clone();
    for (int i = 0; i < r.rows; ++i)
        for (int j = 0; j < r.cols; ++j) {
            if (r.type() == CV_64F) r.at<double>(i, j) *= s;
            else if (r.type() == CV_32F) r.at<float>(i, j) = (float)(r.at<float>(i, j) * s);
            else r.at<unsigned char>(i, j) = (unsigned char)(r.at<unsigned char>(i, j) * s);
        }
    return r;
}
struct Matx34f {
    float val[12];
    Matx34f() : val{} {}
    Matx34f(float a, float b, float c, float d, float e, float f, float g, float h, float i, float j, float k, float l) : val{a, b, c, d, e, f, g, h, i, j, k, l} {}
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;

struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
    KeyPoint() {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    static void convert(const std::vector<Point2f>& pts, std::vector<KeyPoint>& kps, float size = 1, float response = 1, int octave = 0, int class_id = -1) {
        kps.resize(pts.size());
        for (size_t i = 0; i < pts.size(); ++i) kps[i] = KeyPoint(pts[i], size, -1, response, octave, class_id);
    }
    static void convert(const std::vector<KeyPoint>& kps, std::vector<Point2f>& pts) {
        pts.resize(kps.size());
        for (size_t i = 0; i < kps.size(); ++i) pts[i] = kps[i].pt;
    }
};

// the reference's loops are written for cv::parallel_for_; their DEFINED result is the sequential one
template <class F> inline void parallel_for_(const Range& range, F f, double = -1.) { f(range); }

// ---- callbacks into the real OpenCV (cv2), set by the test harness
struct MiniCvCallbacks {
    void (*circle)(unsigned char* data, int rows, int cols, size_t step, int type, int cx, int cy, int radius, double color, int thickness);
    int (*fast_detect)(int threshold, const unsigned char* img, int rows, int cols, size_t step, const unsigned char* mask, size_t mask_step, int mask_type,
                       float* out_xy_resp, int cap);
    void (*corner_subpix)(const unsigned char* img, int rows, int cols, size_t step, float* pts_xy, int n, int win, int zero, int max_iter, double eps);
    // (px, py) = position of the ROI inside its parent; parent == img and px = py = 0 for a matrix that is not a ROI
    void (*gaussian_blur)(const unsigned char* parent, int parent_rows, int parent_cols, size_t step, int px, int py, int rows, int cols, unsigned char* out,
                          int ksize, double sigma);
    void (*corner_min_eigen)(const unsigned char* img, int rows, int cols, size_t step, float* out, int block, int ksize);
    void (*min_max_loc)(const float* m, int rows, int cols, size_t step, double* minv, double* maxv, int* minxy, int* maxxy);
    int (*orb_compute)(const unsigned char* img, int rows, int cols, size_t step, float* pts_xy, int n, unsigned char* desc_out);
    void (*lk)(const unsigned char* prev, const unsigned char* next, int rows, int cols, size_t prev_step, size_t next_step, const float* prev_pts, float* next_pts,
               int n, unsigned char* status, float* err, int win, int max_level, int max_iter, double eps, int flags, double min_eig);
    void (*get_rect_subpix)(const unsigned char* img, int rows, int cols, size_t step, int pw, int ph, float cx, float cy, unsigned char* out);
};
MiniCvCallbacks& mini_cv_callbacks();
inline double norm(const Mat& a, const Mat& b, int type) {                    // 8-bit matrices, L1: an exact integer sum
    if ((type != NORM_L1 && type != NORM_HAMMING) || a.type() != CV_8U) { fprintf(stderr, "oracle mini OpenCV: only the 8-bit L1 / Hamming norms are provided\n"); abort(); }
    long s = 0;
    if (type == NORM_HAMMING) {
        for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) s += __builtin_popcount((unsigned)(a.at<unsigned char>(r, c) ^ b.at<unsigned char>(r, c)));
        return (double)s;
    }
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) s += std::abs((int)a.at<unsigned char>(r, c) - (int)b.at<unsigned char>(r, c));
    return (double)s;
}
inline void mini_cv_missing(const char* what) { fprintf(stderr, "oracle mini OpenCV: %s is not provided\n", what); abort(); }

}  // namespace cv

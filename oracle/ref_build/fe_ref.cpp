// TEST INFRASTRUCTURE (oracle/): C entry points over the REFERENCE'S OWN FeatureExtractor (/root/reference/src/feature_extractor.cpp,
// compiled where it lies against the stand-in OpenCV containers of mini_cv/, whose arithmetic calls are forwarded to the real OpenCV
// through callbacks - see mini_cv/opencv2/core.hpp and cv_callbacks.py).  The reference's control flow (grid walk, float mask and
// its discs, libstdc++'s std::sort on the responses, threshold / quality adaptation, second detections, the descriptor re-alignment)
// runs as written; tests compare oracle/image_ref.py with it.
#include <cstdint>
#include <vector>

#include "feature_extractor.hpp"
#include "feature_tracker.hpp"

extern cv::Ptr<cv::FastFeatureDetector> pfast_;          // file-scope detector objects of feature_extractor.cpp (:67-73): one per
extern cv::Ptr<cv::DescriptorExtractor> pbrief_;         // process in the reference; re-created per extractor here so that tests are independent

cv::MiniCvCallbacks& cv::mini_cv_callbacks() {
    static cv::MiniCvCallbacks c = {};
    return c;
}

extern "C" {

void ov2ref_cv_set_callbacks(const cv::MiniCvCallbacks* c) { cv::mini_cv_callbacks() = *c; }

void* ov2ref_fe_create(int nmaxpts, int nmaxdist, double dmaxquality, int nfast_th) {
    pfast_.reset();
    pbrief_.reset();
    return new FeatureExtractor((size_t)nmaxpts, (size_t)nmaxdist, dmaxquality, nfast_th);
}
void ov2ref_fe_destroy(void* fe) { delete (FeatureExtractor*)fe; }

static std::vector<cv::Point2f> points_of(const float* xy, int n) {
    std::vector<cv::Point2f> v(n);
    for (int i = 0; i < n; ++i) v[i] = cv::Point2f(xy[2 * i], xy[2 * i + 1]);
    return v;
}
static int store(const std::vector<cv::Point2f>& v, float* out, int cap) {
    const int n = (int)v.size() < cap ? (int)v.size() : cap;
    for (int i = 0; i < n; ++i) { out[2 * i] = v[i].x; out[2 * i + 1] = v[i].y; }
    return (int)v.size();
}

int ov2ref_fe_detect_grid_fast(void* fe, const uint8_t* img, int rows, int cols, int ncellsize, const float* cur, int ncur, const int* roi, float* out,
                               int cap, int* fast_th_out) {
    FeatureExtractor* f = (FeatureExtractor*)fe;
    cv::Mat im(rows, cols, CV_8UC1, (void*)img);
    const auto v = f->detectGridFAST(im, ncellsize, points_of(cur, ncur), cv::Rect(roi[0], roi[1], roi[2], roi[3]));
    *fast_th_out = f->nfast_th_;
    return store(v, out, cap);
}

int ov2ref_fe_detect_single_scale(void* fe, const uint8_t* img, int rows, int cols, int ncellsize, const float* cur, int ncur, const int* roi, float* out,
                                  int cap, double* dmaxquality_out) {
    FeatureExtractor* f = (FeatureExtractor*)fe;
    cv::Mat im(rows, cols, CV_8UC1, (void*)img);
    const auto v = f->detectSingleScale(im, ncellsize, points_of(cur, ncur), cv::Rect(roi[0], roi[1], roi[2], roi[3]));
    *dmaxquality_out = f->dmaxquality_;
    return store(v, out, cap);
}

// desc: n x 32 bytes, valid[i] = 0 where the reference returns an empty Mat
int ov2ref_fe_describe(void* fe, const uint8_t* img, int rows, int cols, const float* pts, int n, uint8_t* desc, uint8_t* valid) {
    FeatureExtractor* f = (FeatureExtractor*)fe;
    cv::Mat im(rows, cols, CV_8UC1, (void*)img);
    const std::vector<cv::Mat> v = f->describeBRIEF(im, points_of(pts, n));
    if ((int)v.size() != n) return -1;
    for (int i = 0; i < n; ++i) {
        valid[i] = v[i].empty() ? 0 : 1;
        if (valid[i]) memcpy(desc + 32 * (size_t)i, v[i].ptr(0), 32);
        else memset(desc + 32 * (size_t)i, 0, 32);
    }
    return n;
}

// ---- FeatureTracker (/root/reference/src/feature_tracker.cpp): forward-backward KLT and the stereo row search
// pyramids as the reference keeps them: 2 * (nlevels + 1) entries, of which the arithmetic reads entry 0
int ov2ref_ft_fb_klt(const uint8_t* prev, const uint8_t* cur, int rows, int cols, int npyr_entries, int nwinsize, int nbpyrlvl, float ferr, float fmax_fbklt_dist,
                     int nmax_iter, float fmax_px_precision, const float* kps, float* priors_inout, int n, uint8_t* status_out) {
    FeatureTracker ft(nmax_iter, fmax_px_precision, nullptr);
    std::vector<cv::Mat> pp(npyr_entries), cp(npyr_entries);
    pp[0] = cv::Mat(rows, cols, CV_8UC1, (void*)prev);
    cp[0] = cv::Mat(rows, cols, CV_8UC1, (void*)cur);
    std::vector<cv::Point2f> vkps = points_of(kps, n), vpriors = points_of(priors_inout, n);
    std::vector<bool> st;
    ft.fbKltTracking(pp, cp, nwinsize, nbpyrlvl, ferr, fmax_fbklt_dist, vkps, vpriors, st);
    for (int i = 0; i < n; ++i) { status_out[i] = i < (int)st.size() && st[i] ? 1 : 0; priors_inout[2 * i] = vpriors[i].x; priors_inout[2 * i + 1] = vpriors[i].y; }
    return (int)st.size();
}

void ov2ref_ft_line_min_sad(const uint8_t* iml, const uint8_t* imr, int rows, int cols, float x, float y, int nwinsize, int goleft, float* xprior, float* l1err) {
    FeatureTracker ft(30, 0.01f, nullptr);
    cv::Mat l(rows, cols, CV_8UC1, (void*)iml), r(rows, cols, CV_8UC1, (void*)imr);
    *l1err = 255.f;
    ft.getLineMinSAD(l, r, cv::Point2f(x, y), nwinsize, *xprior, *l1err, goleft != 0);
}

}  // extern "C"

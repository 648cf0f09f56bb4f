// FeatureTracker on the GPU: host shim over the C ABI (include/ov2b200.h).
// Behaviour mirrored: /root/reference/src/feature_tracker.cpp:35-137 (fbKltTracking), :216-221
// (inBorder).  No OpenCV calls on the hot path; no CPU fallback: a CUDA failure leaves every
// keypoint marked lost (status false) and prints the ABI error once, it never silently tracks on
// the CPU.
#include "feature_tracker.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "../../include/ov2b200.h"

struct FeatureTracker::ThreadState {
    ov2_ctx* ctx = nullptr;
    ov2_pyr* prev = nullptr;
    ov2_pyr* cur = nullptr;
    int w = 0, h = 0, nlev = 0;
    ~ThreadState() {
        if (prev) ov2_pyr_destroy(prev);
        if (cur) ov2_pyr_destroy(cur);
        if (ctx) ov2_destroy(ctx);
    }
};

namespace {
std::mutex g_mu;
std::unordered_map<std::thread::id, FeatureTracker::ThreadState*>* g_states = nullptr;
int env_device() {
    const char* e = getenv("OV2_DEVICE");
    return e ? atoi(e) : 0;
}
}  // namespace

FeatureTracker::FeatureTracker(int nmax_iter, float fmax_px_precision, cv::Ptr<cv::CLAHE> pclahe)
    : klt_convg_crit_(cv::TermCriteria::COUNT + cv::TermCriteria::EPS, nmax_iter, fmax_px_precision), pclahe_(pclahe) {}

FeatureTracker::~FeatureTracker() {}

FeatureTracker::ThreadState* FeatureTracker::state() const {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_states) g_states = new std::unordered_map<std::thread::id, ThreadState*>();
    auto it = g_states->find(std::this_thread::get_id());
    if (it != g_states->end()) return it->second;
    ThreadState* s = new ThreadState();
    if (ov2_create(env_device(), &s->ctx) != OV2_OK) {
        fprintf(stderr, "[ov2b200] FeatureTracker: no CUDA device - the GPU front-end has no CPU fallback\n");
        s->ctx = nullptr;
    }
    (*g_states)[std::this_thread::get_id()] = s;
    return s;
}

void FeatureTracker::fbKltTracking(const std::vector<cv::Mat> &vprevpyr, const std::vector<cv::Mat> &vcurpyr, int nwinsize,
        int nbpyrlvl, float ferr, float fmax_fbklt_dist, std::vector<cv::Point2f> &vkps,
        std::vector<cv::Point2f> &vpriorkps, std::vector<bool> &vkpstatus) const
{
    if (vkps.empty()) return;                               // feature_tracker.cpp:43-46
    const size_t nbkps = vkps.size();
    vkpstatus.reserve(vkpstatus.size() + nbkps);
    // the pyramid vector interleaves image / derivative Mats (2 per level); the 8-bit level 0 is all
    // the device path needs - the lower levels and the Scharr planes are recomputed on the GPU
    if ((int)vprevpyr.size() < 2 * (nbpyrlvl + 1)) nbpyrlvl = (int)vprevpyr.size() / 2 - 1;   // :50-52
    if (nbpyrlvl < 0) nbpyrlvl = 0;
    ThreadState* s = state();
    auto fail_all = [&]() { for (size_t i = 0; i < nbkps; ++i) vkpstatus.push_back(false); };
    if (!s->ctx || vprevpyr.empty() || vcurpyr.empty()) { fail_all(); return; }
    const cv::Mat& p0 = vprevpyr[0];
    const cv::Mat& c0 = vcurpyr[0];
    const int nlev_extra = (int)vprevpyr.size() / 2 - 1;
    if (!s->prev || s->w != p0.cols || s->h != p0.rows || s->nlev != nlev_extra) {
        if (s->prev) ov2_pyr_destroy(s->prev);
        if (s->cur) ov2_pyr_destroy(s->cur);
        s->prev = s->cur = nullptr;
        if (ov2_pyr_create(s->ctx, 1, p0.cols, p0.rows, nlev_extra, &s->prev) != OV2_OK ||
            ov2_pyr_create(s->ctx, 1, p0.cols, p0.rows, nlev_extra, &s->cur) != OV2_OK) { fail_all(); return; }
        s->w = p0.cols; s->h = p0.rows; s->nlev = nlev_extra;
    }
    // level-0 Mats are ROIs of border-padded buffers: honour their row step
    if (ov2_pyr_build(s->ctx, s->prev, p0.data, p0.step, p0.step * p0.rows, 0, 1) != OV2_OK ||
        ov2_pyr_build(s->ctx, s->cur, c0.data, c0.step, c0.step * c0.rows, 0, 1) != OV2_OK) {
        fprintf(stderr, "[ov2b200] fbKltTracking: %s\n", ov2_last_error(s->ctx));
        fail_all();
        return;
    }
    ov2_klt_params prm;
    prm.win = nwinsize;
    prm.max_iter = klt_convg_crit_.maxCount;
    prm.eps = (float)klt_convg_crit_.epsilon;
    prm.ferr = ferr;
    prm.fb_dist = fmax_fbklt_dist;
    std::vector<uint8_t> status(nbkps, 0);
    static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "Point2f layout");
    ov2_status st = ov2_fb_klt(s->ctx, s->prev, s->cur, &prm, (int)nbkps, nullptr, 0, (int)nbkps, nullptr, nbpyrlvl,
                               reinterpret_cast<const float*>(vkps.data()), reinterpret_cast<float*>(vpriorkps.data()),
                               status.data());
    if (st != OV2_OK) {
        fprintf(stderr, "[ov2b200] fbKltTracking: %s\n", ov2_last_error(s->ctx));
        fail_all();
        return;
    }
    for (size_t i = 0; i < nbkps; ++i) vkpstatus.push_back(status[i] != 0);
}

// Outside the hot-path scope: line search for rectified stereo priors.
void FeatureTracker::getLineMinSAD(const cv::Mat &iml, const cv::Mat &imr, const cv::Point2f &pt, const int nwinsize,
        float &xprior, float &l1err, bool bgoleft) const
{
    xprior = -1;
    if (nwinsize % 2 == 0) return;
    const float x = std::round(pt.x), y = std::round(pt.y);
    const int half = nwinsize / 2;
    if (x - half < 0) return;
    const int px = (int)x, py = (int)y;
    if (py - half < 0 || py + half >= iml.rows || px - half < 0 || px + half >= iml.cols) return;
    const int nwinsizesq = nwinsize * nwinsize;
    float minsad = 255.f;
    int best = -1;
    const int xstart = bgoleft ? half : px, xend = bgoleft ? px + 1 : imr.cols - half;
    for (int c = xstart; c < xend; ++c) {
        if (c - half < 0 || c + half >= imr.cols) continue;
        long sad = 0;
        for (int dy = -half; dy <= half; ++dy)
            for (int dx = -half; dx <= half; ++dx)
                sad += std::abs((int)iml.ptr(py + dy)[px + dx] - (int)imr.ptr(py + dy)[c + dx]);
        const float v = (float)sad / nwinsizesq;
        if (v < minsad) { minsad = v; best = c; }
    }
    if (best >= 0) xprior = (float)best;
    l1err = minsad;
}

bool FeatureTracker::inBorder(const cv::Point2f &pt, const cv::Mat &im) const
{
    const float BORDER_SIZE = 1.f;
    return BORDER_SIZE <= pt.x && pt.x < im.cols - BORDER_SIZE && BORDER_SIZE <= pt.y && pt.y < im.rows - BORDER_SIZE;
}

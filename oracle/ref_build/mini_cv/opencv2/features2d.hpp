// TEST INFRASTRUCTURE (oracle/): see core.hpp - the detector / descriptor classes the reference's feature extractor instantiates.
#pragma once
#include "core.hpp"

namespace cv {

class Feature2D {
public:
    virtual ~Feature2D() {}
    virtual void detect(const Mat&, std::vector<KeyPoint>&, const Mat& = Mat()) { mini_cv_missing("Feature2D::detect"); }
    virtual void compute(const Mat&, std::vector<KeyPoint>&, Mat&) { mini_cv_missing("Feature2D::compute"); }
};
typedef Feature2D DescriptorExtractor;
typedef Feature2D FeatureDetector;

class FastFeatureDetector : public Feature2D {
public:
    static Ptr<FastFeatureDetector> create(int threshold = 10, bool nonmax = true, int = 2) { auto p = std::make_shared<FastFeatureDetector>(); p->th_ = threshold; (void)nonmax; return p; }
    void setThreshold(int t) { th_ = t; }
    int getThreshold() const { return th_; }
    void detect(const Mat& im, std::vector<KeyPoint>& kps, const Mat& mask = Mat()) override {
        std::vector<float> out(3 * (size_t)im.rows * im.cols + 3);
        const int n = mini_cv_callbacks().fast_detect(th_, im.data, im.rows, im.cols, im.step, mask.data, mask.step, mask.type(), out.data(), im.rows * im.cols);
        kps.resize(n);
        for (int i = 0; i < n; ++i) { kps[i] = KeyPoint(Point2f(out[3 * i], out[3 * i + 1]), 7.f, -1, out[3 * i + 2]); }
    }
private:
    int th_ = 10;
};

class GFTTDetector : public Feature2D {      // detectGFTT: no shipped configuration uses it
public:
    static Ptr<GFTTDetector> create(int = 1000, double = 0.01, double = 1, int = 3, bool = false, double = 0.04) { return std::make_shared<GFTTDetector>(); }
    void setMaxFeatures(int) {}
    void setQualityLevel(double) {}
    void setMinDistance(double) {}
    void detect(const Mat&, std::vector<KeyPoint>&, const Mat& = Mat()) override { mini_cv_missing("GFTTDetector::detect"); }
};

class ORB : public Feature2D {
public:
    static Ptr<ORB> create(int nfeatures = 500, float scale = 1.2f, int nlevels = 8) {
        if (nfeatures != 500 || scale != 1.f || nlevels != 0) mini_cv_missing("ORB::create with parameters other than (500, 1., 0)");
        return std::make_shared<ORB>();
    }
    // descriptors of the keypoints ORB keeps (those too close to the border are REMOVED from kps, as OpenCV does)
    void compute(const Mat& im, std::vector<KeyPoint>& kps, Mat& descs) override {
        const int n = (int)kps.size();
        std::vector<float> pts(2 * (size_t)n + 2);
        for (int i = 0; i < n; ++i) { pts[2 * i] = kps[i].pt.x; pts[2 * i + 1] = kps[i].pt.y; }
        std::vector<unsigned char> d(32 * (size_t)n + 32);
        const int m = mini_cv_callbacks().orb_compute(im.data, im.rows, im.cols, im.step, pts.data(), n, d.data());
        std::vector<KeyPoint> kept(m);
        for (int i = 0; i < m; ++i) { kept[i] = KeyPoint(Point2f(pts[2 * i], pts[2 * i + 1]), 1.f); }
        kps.swap(kept);
        descs = Mat(m, 32, CV_8U);
        if (m) memcpy(descs.data, d.data(), 32 * (size_t)m);
    }
};

struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>&, int) { mini_cv_missing("KeyPointsFilter::retainBest"); }
    static void runByPixelsMask(std::vector<KeyPoint>&, const Mat&) { mini_cv_missing("KeyPointsFilter::runByPixelsMask"); }
};

}  // namespace cv

// Drop-in replacement of the reference's include/feature_extractor.hpp (:31-55): same class, same
// public members and signatures, so map_manager.cpp compiles and links against it unchanged.
// detectGridFAST, detectSingleScale and describeBRIEF run on the GPU through the C ABI
// (include/ov2b200.h).  detectGFTT (use_shi_tomasi: 1, no reference configuration enables it) is
// declared so callers link, and reports loudly that it is not built.
#pragma once

#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>

#include <vector>

struct ov2_ctx;
struct ov2_pyr;

class FeatureExtractor {

public:
    FeatureExtractor() {};
    FeatureExtractor(size_t nmaxpts, size_t nmaxdist, double dmaxquality, int nfast_th);

    std::vector<cv::Point2f> detectGFTT(const cv::Mat &im, const std::vector<cv::Point2f> &vcurkps,
                                        const cv::Mat &roi, int nbmax=-1) const;

    std::vector<cv::Point2f> detectGridFAST(const cv::Mat &im, const int ncellsize,
        const std::vector<cv::Point2f> &vcurkps, const cv::Rect &roi);

    std::vector<cv::Mat> describeBRIEF(const cv::Mat &im, const std::vector<cv::Point2f> &vpts) const;

    std::vector<cv::Point2f> detectSingleScale(const cv::Mat &im, const int ncellsize,
            const std::vector<cv::Point2f> &vcurkps, const cv::Rect &roi);

    void setMask(const cv::Mat &im, const std::vector<cv::Point2f> &vpts,  const int dist, cv::Mat &mask) const;

    size_t nmaxpts_, nmaxdist_, nmindist_;
    double dmaxquality_, dminquality_;
    int nfast_th_;

    std::vector<int> vumax_;
};

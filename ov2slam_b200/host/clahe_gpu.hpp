// GPU CLAHE behind the cv::CLAHE interface (row C of the scope table).  The reference creates ONE cv::CLAHE
// (/root/reference/src/ov2slam.cpp:82-87: cv::createCLAHE(fclahe_val, Size(W/50, H/50))), hands it to FeatureTracker and calls
// pclahe_->apply(raw, equalised) per image from the front-end thread (src/visual_front_end.cpp:1158-1160) and from the mapper thread
// for the right image (src/mapper.cpp:75-76).  Replacing that one createCLAHE call by ov2shim::createCLAHE(...) routes both through
// ov2_clahe (csrc/frontend_clahe.cu: bit-exact with cv::CLAHE, tests/test_frontend_gpu.py); nothing else changes.
#pragma once
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>

namespace ov2shim {
cv::Ptr<cv::CLAHE> createCLAHE(double clipLimit, cv::Size tileGridSize);
}

// Drives the cv::CLAHE adapter (host/clahe_gpu.cpp): a source that is a ROI of a wider buffer into a fresh destination, the same
// image in place, and an empty image.  Writes the two outputs; tests/test_host_shim.py compares them with the mock's rule (CPU) .
//   clahe_selftest width height out.bin
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "clahe_gpu.hpp"

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int w = atoi(argv[1]), h = atoi(argv[2]), pad = 11;
    std::vector<unsigned char> buf((size_t)(w + pad) * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w + pad; ++x) buf[(size_t)y * (w + pad) + x] = (unsigned char)((x * 7 + y * 13 + (x * y) % 5) & 255);
    cv::Mat roi(h, w, CV_8UC1, buf.data(), (size_t)(w + pad));
    cv::Ptr<cv::CLAHE> clahe = ov2shim::createCLAHE(3.0, cv::Size(w / 50, h / 50));
    cv::Mat out;
    clahe->apply(roi, out);
    if (out.empty() || out.rows != h || out.cols != w) return 3;
    cv::Mat inplace(h, w, CV_8UC1);
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) inplace.ptr(y)[x] = roi.ptr(y)[x];
    clahe->apply(inplace, inplace);
    cv::Mat none, none_out;
    clahe->apply(none, none_out);
    if (!none_out.empty()) return 4;
    FILE* g = fopen(argv[3], "wb");
    if (!g) return 5;
    for (int y = 0; y < h; ++y) fwrite(out.ptr(y), 1, w, g);
    for (int y = 0; y < h; ++y) fwrite(inplace.ptr(y), 1, w, g);
    fclose(g);
    printf("clahe adapter: %d x %d ok\n", w, h);
    return 0;
}

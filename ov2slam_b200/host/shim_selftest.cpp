// Links the drop-in classes against libov2b200.so and drives them the way the reference's callers
// do (map_manager.cpp:286-341 extractKeypoints, visual_front_end.cpp:196-251 kltTracking) on a
// synthetic image.  Prints a short digest that tests/test_host_shim.py compares with the Python
// binding's result for the same input (run on the GPU box).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "feature_extractor.hpp"
#include "feature_tracker.hpp"

// --sad left.raw right.raw W H n x0 y0 win goleft ... : FeatureTracker::getLineMinSAD on a list of points (host code only:
// runs without a GPU); prints "xprior l1err" per point.
static int sad_mode(int argc, char** argv) {
    if (argc < 7) return 2;
    const int W = atoi(argv[4]), H = atoi(argv[5]), n = atoi(argv[6]);
    cv::Mat l(H, W, CV_8UC1), r(H, W, CV_8UC1);
    FILE* f = fopen(argv[2], "rb"); if (!f || fread(l.data, 1, (size_t)W * H, f) != (size_t)W * H) return 3; fclose(f);
    f = fopen(argv[3], "rb"); if (!f || fread(r.data, 1, (size_t)W * H, f) != (size_t)W * H) return 3; fclose(f);
    FeatureTracker ft(30, 0.01f, cv::Ptr<cv::CLAHE>());
    for (int k = 0; k < n; ++k) {
        const char* const* a = argv + 7 + 4 * k;
        float xp = 0.f, err = -1.f;
        ft.getLineMinSAD(l, r, cv::Point2f((float)atof(a[0]), (float)atof(a[1])), atoi(a[2]), xp, err, atoi(a[3]) != 0);
        printf("%.9g %.9g\n", xp, err);
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && std::string(argv[1]) == "--sad") return sad_mode(argc, argv);
    if (argc < 4) { fprintf(stderr, "usage: shim_selftest prev.raw cur.raw W H\n"); return 2; }
    const int W = atoi(argv[3]), H = argc > 4 ? atoi(argv[4]) : 480;
    cv::Mat prev(H, W, CV_8UC1), cur(H, W, CV_8UC1);
    FILE* f = fopen(argv[1], "rb"); if (!f || fread(prev.data, 1, (size_t)W * H, f) != (size_t)W * H) return 3; fclose(f);
    f = fopen(argv[2], "rb"); if (!f || fread(cur.data, 1, (size_t)W * H, f) != (size_t)W * H) return 3; fclose(f);
    FeatureExtractor fe(0, 50, 0.0, 10);
    std::vector<cv::Point2f> kps = fe.detectGridFAST(prev, 50, std::vector<cv::Point2f>(), cv::Rect());
    std::vector<cv::Mat> descs = fe.describeBRIEF(prev, kps);
    unsigned long dsum = 0; int ndesc = 0;
    for (auto& d : descs) if (!d.empty()) { ndesc++; for (int k = 0; k < 32; ++k) dsum += d.data[k] * (unsigned long)(k + 1); }
    FeatureTracker ft(30, 0.01f, cv::Ptr<cv::CLAHE>());
    // an 8-entry pyramid vector as cv::buildOpticalFlowPyramid would give (only entry 0 is read)
    std::vector<cv::Mat> pp(8), cp(8);
    pp[0] = prev; cp[0] = cur;
    std::vector<cv::Point2f> pri = kps;
    std::vector<bool> status;
    ft.fbKltTracking(pp, cp, 9, 3, 30.f, 0.5f, kps, pri, status);
    int ngood = 0; double sx = 0, sy = 0;
    for (size_t i = 0; i < status.size(); ++i) if (status[i]) { ngood++; sx += pri[i].x - kps[i].x; sy += pri[i].y - kps[i].y; }
    double kx = 0, ky = 0; for (auto& p : kps) { kx += p.x; ky += p.y; }
    printf("nkps %zu th %d ksum %.4f %.4f ndesc %d dsum %lu ngood %d flow %.4f %.4f\n", kps.size(), fe.nfast_th_, kx, ky, ndesc, dsum,
           ngood, ngood ? sx / ngood : 0.0, ngood ? sy / ngood : 0.0);
    return 0;
}

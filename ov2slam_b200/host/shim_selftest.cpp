// Links the drop-in classes against libov2b200.so and drives them the way the reference's callers
// do (map_manager.cpp:286-341 extractKeypoints, visual_front_end.cpp:196-251 kltTracking) on a
// synthetic image.  Prints a short digest that tests/test_host_shim.py compares with the Python
// binding's result for the same input (run on the GPU box).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
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

// --bench frames.raw W H N reps : the CLASS path as the reference's live loop drives it, one frame per call
// (VisualFrontEnd::trackMono -> kltTracking's two fbKltTracking calls, src/visual_front_end.cpp:196,242; then
// MapManager::extractKeypoints: describeBRIEF, detectGridFAST, describeBRIEF, src/map_manager.cpp:286-341), prev / cur
// pyramid vectors swapped and the current one refilled in place every frame (:1168-1172).  Prints frames/s.
static int bench_mode(int argc, char** argv) {
    if (argc < 7) return 2;
    const int W = atoi(argv[3]), H = atoi(argv[4]), N = atoi(argv[5]), reps = atoi(argv[6]);
    std::vector<unsigned char> all((size_t)W * H * N);
    FILE* f = fopen(argv[2], "rb"); if (!f || fread(all.data(), 1, all.size(), f) != all.size()) return 3; fclose(f);
    FeatureExtractor fe(0, 50, 0.0, 10);
    FeatureTracker ft(30, 0.01f, cv::Ptr<cv::CLAHE>());
    std::vector<cv::Mat> prevpyr(8), curpyr(8);
    prevpyr[0] = cv::Mat(H, W, CV_8UC1); curpyr[0] = cv::Mat(H, W, CV_8UC1);
    memcpy(curpyr[0].data, all.data(), (size_t)W * H);
    std::vector<cv::Point2f> kps = fe.detectGridFAST(curpyr[0], 50, std::vector<cv::Point2f>(), cv::Rect());
    long tracked = 0, detected = 0;
    struct timespec t0, t1;
    int frames = 0;
    for (int rep = -1; rep < reps; ++rep) {                // rep -1: warm-up pass
        if (rep == 0) { clock_gettime(CLOCK_MONOTONIC, &t0); frames = 0; tracked = detected = 0; }
        for (int k = 1; k < N; ++k) {
            prevpyr.swap(curpyr);                                                      // visual_front_end.cpp:1168-1170
            memcpy(curpyr[0].data, all.data() + (size_t)k * W * H, (size_t)W * H);     // buildOpticalFlowPyramid refills cur in place
            // kltTracking: keypoints with a 3D prior first (nbpyrlvl 1), the rest with the full pyramid
            std::vector<cv::Point2f> k3, p3, k2, p2;
            for (size_t i = 0; i < kps.size(); ++i) ((i % 5) < 3 ? k3 : k2).push_back(kps[i]);
            p3 = k3; p2 = k2;
            std::vector<bool> s3, s2;
            ft.fbKltTracking(prevpyr, curpyr, 9, 1, 30.f, 0.5f, k3, p3, s3);
            ft.fbKltTracking(prevpyr, curpyr, 9, 3, 30.f, 0.5f, k2, p2, s2);
            std::vector<cv::Point2f> alive;
            for (size_t i = 0; i < s3.size(); ++i) if (s3[i]) alive.push_back(p3[i]);
            for (size_t i = 0; i < s2.size(); ++i) if (s2[i]) alive.push_back(p2[i]);
            tracked += (long)alive.size();
            // createKeyframe -> extractKeypoints on the current image
            std::vector<cv::Mat> d0 = fe.describeBRIEF(curpyr[0], alive);
            std::vector<cv::Point2f> fresh = fe.detectGridFAST(curpyr[0], 50, alive, cv::Rect());
            std::vector<cv::Mat> d1 = fe.describeBRIEF(curpyr[0], fresh);
            detected += (long)fresh.size();
            kps = alive;
            kps.insert(kps.end(), fresh.begin(), fresh.end());
            frames++;
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double dt = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    printf("class_path frames %d seconds %.6f fps %.2f tracked_per_frame %.1f detected_per_frame %.1f\n", frames, dt, frames / dt,
           (double)tracked / frames, (double)detected / frames);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && std::string(argv[1]) == "--sad") return sad_mode(argc, argv);
    if (argc > 1 && std::string(argv[1]) == "--bench") return bench_mode(argc, argv);
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

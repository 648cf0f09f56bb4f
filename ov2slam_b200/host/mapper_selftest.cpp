// Drives the drop-in Mapper::matchToMap (host/mapper_match_gpu.cpp) on a SYNTHETIC map: reads a flattened matching scene written
// by tests/test_host_shim.py (ov2slam_b200.synth.make_match_scene), builds the frame (grid, keypoints), the keyframes and the map
// points (descriptors, keyframe sets, per-keyframe keypoints) from it with the stand-in map classes
// (host/standin/ref/map_manager.hpp), calls matchToMap, and writes the order in which the local-map id set was walked (it
// decides ties, and an unordered_set's order is the library's business) and the resulting (keypoint lmid -> map-point lmid) pairs.
// The test replays the same candidate order through the Python binding of ov2_match_to_map and compares.
//
//   mapper_selftest scene.bin result.bin
// scene.bin: i32 nkps, nmps, ndesc, nobs, nkfs, ncand, ncells, nbwcells, ncellsize, img_w, img_h, has_dist;
//            f64 K[4], dist[5], Tcw[12], kf_Tcw[nkfs][12], mp_xyz[nmps][3]; f32 kp_px[nkps][2], obs_px[nobs][2];
//            i32 cell_ptr[ncells+1], cell_kp[nkps], kp_lm[nkps], desc_ptr[nmps+1], obs_ptr[nmps+1], obs_kf[nobs], cand_mp[ncand];
//            u8 desc[ndesc][32]
// result.bin: i32 ncand, order[ncand]; i32 npairs, pairs[npairs][2]; i32 n_removed_obs
// ids: map point m <-> lmid m; a keypoint without map point gets lmid 1000000 + its index (no such map point exists: the
// shim must drop the observation, mapper.cpp:667-672); keyframe k <-> kfid 10 + k; the frame is keyframe 5000.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mapper.hpp"

template <typename T> static bool rd(FILE* f, std::vector<T>& v, size_t n) { v.resize(n); return n == 0 || fread(v.data(), sizeof(T), n, f) == n; }

// camera <- world (row-major R, then t)  ->  Twc as quaternion + translation (what the map stores)
static Sophus::SE3d twc_of(const double* T) {
    const double* R = T;
    double q[4];   // w x y z
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        const double s = std::sqrt(tr + 1.0) * 2;
        q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        const double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
        q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        const double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
        q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
    } else {
        const double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
        q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
    }
    const Sophus::SE3d Tcw(Eigen::Quaterniond(q[0], q[1], q[2], q[3]), Eigen::Vector3d(T[9], T[10], T[11]));
    return Tcw.inverse();
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: mapper_selftest scene.bin result.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    int32_t hd[12];
    if (fread(hd, 4, 12, f) != 12) return 3;
    const int nkps = hd[0], nmps = hd[1], ndesc = hd[2], nobs = hd[3], nkfs = hd[4], ncand = hd[5], ncells = hd[6];
    std::vector<double> K, dist, Tcw, kfT, xyz;
    std::vector<float> kp_px, obs_px;
    std::vector<int32_t> cell_ptr, cell_kp, kp_lm, desc_ptr, obs_ptr, obs_kf, cand_mp;
    std::vector<uint8_t> desc;
    if (!rd(f, K, 4) || !rd(f, dist, 5) || !rd(f, Tcw, 12) || !rd(f, kfT, 12 * (size_t)nkfs) || !rd(f, xyz, 3 * (size_t)nmps) ||
        !rd(f, kp_px, 2 * (size_t)nkps) || !rd(f, obs_px, 2 * (size_t)nobs) || !rd(f, cell_ptr, ncells + 1) || !rd(f, cell_kp, nkps) ||
        !rd(f, kp_lm, nkps) || !rd(f, desc_ptr, nmps + 1) || !rd(f, obs_ptr, nmps + 1) || !rd(f, obs_kf, nobs) || !rd(f, cand_mp, ncand) ||
        !rd(f, desc, 32 * (size_t)ndesc))
        return 3;
    fclose(f);
    auto params = std::make_shared<SlamParams>();
    auto map = std::make_shared<MapManager>();
    auto cal = std::make_shared<CameraCalibration>();
    cal->setK(K[0], K[1], K[2], K[3]);
    cal->img_w_ = hd[9]; cal->img_h_ = hd[10];
    static unsigned char dcv_marker[5];
    if (hd[11]) {
        cal->k1_ = dist[0]; cal->k2_ = dist[1]; cal->p1_ = dist[2]; cal->p2_ = dist[3];
        cal->Dcv_ = cv::Mat(1, 5, 0, dcv_marker);                       // non-empty: the distorted branch
    }
    std::vector<std::shared_ptr<Frame>> kfs(nkfs);
    for (int k = 0; k < nkfs; ++k) {
        kfs[k] = std::make_shared<Frame>();
        kfs[k]->id_ = kfs[k]->kfid_ = 10 + k;
        kfs[k]->pcalib_leftcam_ = cal;
        kfs[k]->setTwc(twc_of(&kfT[12 * k]));
        map->map_pkfs_[10 + k] = kfs[k];
    }
    for (int m = 0; m < nmps; ++m) {
        auto lm = std::make_shared<MapPoint>();
        lm->lmid_ = m; lm->is3d_ = true;
        lm->ptxyz_ = Eigen::Vector3d(xyz[3 * m], xyz[3 * m + 1], xyz[3 * m + 2]);
        for (int d = desc_ptr[m]; d < desc_ptr[m + 1]; ++d) {
            cv::Mat row(1, 32, 0);
            memcpy(row.ptr(0), &desc[32 * (size_t)d], 32);
            if (d == desc_ptr[m]) lm->desc_ = row;
            lm->map_kf_desc_[d - desc_ptr[m]] = row;
        }
        for (int o = obs_ptr[m]; o < obs_ptr[m + 1]; ++o) {
            lm->set_kfids_.insert(10 + obs_kf[o]);
            Keypoint kp;
            kp.lmid_ = m; kp.px_ = cv::Point2f(obs_px[2 * o], obs_px[2 * o + 1]);
            kfs[obs_kf[o]]->mapkps_[m] = kp;
        }
        map->map_plms_[m] = lm;
    }
    Frame frame;
    frame.id_ = frame.kfid_ = 5000;
    frame.pcalib_leftcam_ = cal;
    frame.setTwc(twc_of(Tcw.data()));
    frame.ncellsize_ = hd[8]; frame.nbwcells_ = hd[7]; frame.nbhcells_ = ncells / hd[7];
    frame.vgridkps_.resize(ncells);
    frame.nb3dkps_ = 0;                                                 // < 30: the radius doubles (mapper.cpp:597-600)
    auto kpid = [&](int j) { return kp_lm[j] >= 0 ? kp_lm[j] : 1000000 + j; };
    for (int j = 0; j < nkps; ++j) {
        Keypoint kp;
        kp.lmid_ = kpid(j); kp.px_ = cv::Point2f(kp_px[2 * j], kp_px[2 * j + 1]);
        frame.mapkps_[kp.lmid_] = kp;
    }
    for (int c = 0; c < ncells; ++c)
        for (int i = cell_ptr[c]; i < cell_ptr[c + 1]; ++i) frame.vgridkps_[c].push_back(kpid(cell_kp[i]));
    std::unordered_set<int> local(cand_mp.begin(), cand_mp.end());
    std::vector<int32_t> order(local.begin(), local.end());

    Mapper mapper(params, map);
    const float fmaxprojerr = 2.f, fdistratio = 0.2f;
    const std::map<int, int> res = mapper.matchToMap(frame, fmaxprojerr, fdistratio, local);

    FILE* g = fopen(argv[2], "wb");
    if (!g) return 4;
    int32_t n = (int32_t)order.size();
    fwrite(&n, 4, 1, g); fwrite(order.data(), 4, order.size(), g);
    n = (int32_t)res.size();
    fwrite(&n, 4, 1, g);
    for (const auto& kv : res) { int32_t pr[2] = {kv.first, kv.second}; fwrite(pr, 4, 2, g); }
    n = (int32_t)map->removed_obs_.size();
    fwrite(&n, 4, 1, g);
    fclose(g);
    printf("matchToMap: %d candidates, %d keypoints, %zu matches, %d observations dropped\n", ncand, nkps, res.size(), n);
    return 0;
}

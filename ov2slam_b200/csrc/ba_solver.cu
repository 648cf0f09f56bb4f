// L: local bundle adjustment on the GPU (float64).
//
// Reference behaviour replaced: the solve sections of Optimizer::localBA
// (/root/reference/src/optimizer.cpp:436-479, :492-594, :603-627, :637-735), i.e. Ceres 2.0's
// TrustRegionMinimizer (trust_region_minimizer.cc:67-134) + LevenbergMarquardtStrategy
// (levenberg_marquardt_strategy.cc:66-160) + Schur linear solver (schur_eliminator_impl.h:179-377)
// on ReprojectionErrorKSE3AnchInvDepth residual blocks (src/ceres_parametrization.cpp:361-473)
// with SE3LeftParameterization (se3left_parametrization.hpp:41-60) and HuberLoss(sqrt(5.9915)).
//
// Per LM iteration (device kernels, LM controller scalars on the host):
//   ba_eval_kernel<true>   one thread per residual block: r, analytic Jacobians (anchor 2x6,
//                          observer 2x6, inverse depth 2x1), robustified by sqrt(rho') (the
//                          Corrector's alpha = 0 branch is the only one Huber reaches), cost
//   ba_schur_kernel        one warp per landmark (observations are CSR by landmark): E'E (scalar),
//                          E'r, F'F blocks, E'F rows, Schur complement and reduced rhs accumulated
//                          into the (6 Ncv)^2 camera system with fp64 atomics (RED.ADD.F64)
//   ba_reduced_solve_kernel one CTA: Jacobi scaling / LM damping of the camera diagonal, Cholesky
//                          (U'U) in shared memory when it fits, two triangular solves, camera Plus
//   ba_backsub_kernel      one warp per landmark: y_e, candidate inverse depth, model cost change
//   ba_eval_kernel<false>  candidate cost (and the chi2 / depth flags the reference reads afterwards)
// Jacobi scaling is applied algebraically (damping_i = clamp(s_i^2 c_i)/(radius s_i^2) on the
// unscaled system), which is the same linear system Ceres solves after ScaleColumns.
#include "ov2_common.cuh"
#include "ge_warp.cuh"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace {

// launch grids of per-observation / per-landmark kernels: an empty shard still launches one (idle) block - a zero-block
// grid is cudaErrorInvalidConfiguration
static inline int grid1(int n, int per_block) { const int g = (n + per_block - 1) / per_block; return g < 1 ? 1 : g; }

constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_VAR_CAMS = 64;
constexpr int MAX_N = 6 * MAX_VAR_CAMS;
constexpr double SOPHUS_EPS = 1e-10;

enum { SC_COST = 0, SC_CAND_COST, SC_MCC, SC_STEP2, SC_CANDX2, SC_GMAX_LM, SC_GMAX_CAM, SC_CHOL_FAIL, SC_NBAD, SC_NLEFT, SC_NRIGHT, SC_COUNT = 16 };

struct BaDev {
    int ncam, npts, nobs, ncv, n;
    int ncopy; long long copy_stride;   // privatised accumulation copies (atomic contention), in doubles
    double fx, fy, cx, cy;
    double rfx, rfy, rcx, rcy;          // right calibration (stereo residual blocks)
    double Rrl[9], trl[3];              // constant right-from-left extrinsic
    const uint8_t* obs_type;            // NULL: all left-camera blocks
    double huber_a, huber_b;
    int use_huber;
    const uint8_t* pose_const;
    const int32_t* lm_anchor_cam;
    const double* lm_anchor_px;
    const int32_t* obs_cam;
    const int32_t* obs_lm;
    const double* obs_px;
    const int32_t* lm_ptr;
    uint8_t* active;
    int32_t* cam_slot;
    uint8_t* cam_used;
    double* Jr; double* Ja; double* Jo; double* Jl;
    double* chi2; uint8_t* dpos;
    double* cn_cam; double* sc_cam; double* sc_lm;
    double* S; double* rhs; double* z; double* gcam;
    double* ete; double* ge;
    double* scal;
    uint8_t* flags;
};

// ------------------------------------------------------------------ small SE3 helpers
__device__ __forceinline__ void quat_to_rot(const double* q, double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

__device__ __forceinline__ void load_pose(const double* p, double t[3], double q[4]) {
    t[0] = p[0]; t[1] = p[1]; t[2] = p[2];
    const double n = 1.0 / sqrt(p[3] * p[3] + p[4] * p[4] + p[5] * p[5] + p[6] * p[6]);  // SE3d(q, t) normalises
    q[0] = p[3] * n; q[1] = p[4] * n; q[2] = p[5] * n; q[3] = p[6] * n;
}

// SE3LeftParameterization::Plus: out = Sophus::SE3d::exp(delta) * (q, t)
__device__ __noinline__ void pose_plus(const double* pose, const double* d, double* out) {
    double t[3], q[4];
    load_pose(pose, t, q);
    const double ox = d[3], oy = d[4], oz = d[5];
    const double th2 = ox * ox + oy * oy + oz * oz;
    double imag, real, theta;
    if (th2 < SOPHUS_EPS * SOPHUS_EPS) {
        theta = 0.0;
        const double th4 = th2 * th2;
        imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
        real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
    } else {
        theta = sqrt(th2);
        double sh_, ch_;
        sincos(0.5 * theta, &sh_, &ch_);
        imag = sh_ / theta;
        real = ch_;
    }
    const double e[4] = {imag * ox, imag * oy, imag * oz, real};
    double Re[9];
    quat_to_rot(e, Re);
    double V[9];
    if (theta < SOPHUS_EPS) {
        for (int i = 0; i < 9; ++i) V[i] = Re[i];
    } else {
        const double a = (1.0 - cos(theta)) / th2;
        const double b = (theta - sin(theta)) / (th2 * theta);
        const double O[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
        double O2[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
        for (int i = 0; i < 9; ++i) V[i] = a * O[i] + b * O2[i];
        V[0] += 1.0; V[4] += 1.0; V[8] += 1.0;
    }
    const double et[3] = {V[0] * d[0] + V[1] * d[1] + V[2] * d[2], V[3] * d[0] + V[4] * d[1] + V[5] * d[2],
                          V[6] * d[0] + V[7] * d[1] + V[8] * d[2]};
    // quaternion product (so3.hpp:338-342), then normalisation
    double r[4];
    r[3] = e[3] * q[3] - e[0] * q[0] - e[1] * q[1] - e[2] * q[2];
    r[0] = e[3] * q[0] + e[0] * q[3] + e[1] * q[2] - e[2] * q[1];
    r[1] = e[3] * q[1] + e[1] * q[3] + e[2] * q[0] - e[0] * q[2];
    r[2] = e[3] * q[2] + e[2] * q[3] + e[0] * q[1] - e[1] * q[0];
    const double n = 1.0 / sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
    out[0] = et[0] + Re[0] * t[0] + Re[1] * t[1] + Re[2] * t[2];
    out[1] = et[1] + Re[3] * t[0] + Re[4] * t[1] + Re[5] * t[2];
    out[2] = et[2] + Re[6] * t[0] + Re[7] * t[1] + Re[8] * t[2];
    out[3] = r[0] * n; out[4] = r[1] * n; out[5] = r[2] * n; out[6] = r[3] * n;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ void atomic_max_pos(double* addr, double v) {
    // for non-negative doubles the bit pattern orders like an unsigned integer
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// ------------------------------------------------------------------ residual blocks
template <bool JAC>
__global__ void __launch_bounds__(128) ba_eval_kernel(BaDev D, const double* __restrict__ pose,
                                                      const double* __restrict__ invd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double cost = 0.0;
    if (i < D.nobs && D.active[i]) {
        const int lm = D.obs_lm[i];
        const int ca = D.lm_anchor_cam[lm], co = D.obs_cam[i];
        double ta[3], qa[4], to[3], qo[4], Rwa[9], Rwc[9];
        load_pose(pose + 7 * ca, ta, qa);
        load_pose(pose + 7 * co, to, qo);
        quat_to_rot(qa, Rwa);
        quat_to_rot(qo, Rwc);
        const double zanch = 1.0 / invd[lm];
        const double bx = (D.lm_anchor_px[2 * lm] - D.cx) / D.fx, by = (D.lm_anchor_px[2 * lm + 1] - D.cy) / D.fy;
        const double ap[3] = {zanch * bx, zanch * by, zanch};
        const double rp[3] = {Rwa[0] * ap[0] + Rwa[1] * ap[1] + Rwa[2] * ap[2], Rwa[3] * ap[0] + Rwa[4] * ap[1] + Rwa[5] * ap[2],
                              Rwa[6] * ap[0] + Rwa[7] * ap[1] + Rwa[8] * ap[2]};  // Rwanch * anchpt
        const double wp[3] = {rp[0] + ta[0], rp[1] + ta[1], rp[2] + ta[2]};
        const double dv[3] = {wp[0] - to[0], wp[1] - to[1], wp[2] - to[2]};
        // Rcw = Rwc^T
        const double lc[3] = {Rwc[0] * dv[0] + Rwc[3] * dv[1] + Rwc[6] * dv[2], Rwc[1] * dv[0] + Rwc[4] * dv[1] + Rwc[7] * dv[2],
                              Rwc[2] * dv[0] + Rwc[5] * dv[1] + Rwc[8] * dv[2]};
        // residual type (stereo windows): 0 left camera, 1 right camera in another keyframe
        // (ceres_parametrization.cpp:579-712), 2 right camera in the anchor frame itself (:476-577, a
        // function of the inverse depth only)
        const int typ = D.obs_type ? (int)D.obs_type[i] : 0;
        double cp[3] = {lc[0], lc[1], lc[2]};
        double kfx = D.fx, kfy = D.fy, kcx = D.cx, kcy = D.cy;
        if (typ != 0) {
            const double* sv = typ == 2 ? ap : lc;
            for (int k = 0; k < 3; ++k) cp[k] = D.Rrl[3 * k] * sv[0] + D.Rrl[3 * k + 1] * sv[1] + D.Rrl[3 * k + 2] * sv[2] + D.trl[k];
            kfx = D.rfx; kfy = D.rfy; kcx = D.rcx; kcy = D.rcy;
        }
        const double linvz = 1.0 / cp[2];
        double r0 = kfx * cp[0] * linvz + kcx - D.obs_px[2 * i];
        double r1 = kfy * cp[1] * linvz + kcy - D.obs_px[2 * i + 1];
        const double s = r0 * r0 + r1 * r1;
        D.chi2[i] = s;
        D.dpos[i] = cp[2] > 0.0 ? 1 : 0;
        double w = 1.0;
        if (D.use_huber && s > D.huber_b) {
            const double rs = sqrt(s);
            const double rho1 = fmax(DBL_MIN, D.huber_a / rs);
            cost = 0.5 * (2.0 * D.huber_a * rs - D.huber_b);
            w = sqrt(rho1);
        } else {
            cost = 0.5 * s;
        }
        if (JAC) {
            const double linvz2 = linvz * linvz;
            const double jc[6] = {linvz * kfx, 0.0, -cp[0] * linvz2 * kfx, 0.0, linvz * kfy, -cp[1] * linvz2 * kfy};
            // M = d(camera point)/d(world point): Rcw (left), Rrl Rcw (right, other frame), Rrl (right, anchor frame)
            double M[9];
            if (typ == 0) {
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) M[3 * a + b] = Rwc[3 * b + a];
            } else if (typ == 1) {
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b)
                        M[3 * a + b] = D.Rrl[3 * a] * Rwc[3 * b] + D.Rrl[3 * a + 1] * Rwc[3 * b + 1] + D.Rrl[3 * a + 2] * Rwc[3 * b + 2];
            } else {
                for (int k = 0; k < 9; ++k) M[k] = D.Rrl[k];
            }
            double JR[6];  // J_cam * M  (2x3)
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 3; ++b)
                    JR[3 * a + b] = jc[3 * a] * M[b] + jc[3 * a + 1] * M[3 + b] + jc[3 * a + 2] * M[6 + b];
            // JR * hat(wpt)
            double JS[6];
            for (int a = 0; a < 2; ++a) {
                const double j0 = JR[3 * a], j1 = JR[3 * a + 1], j2 = JR[3 * a + 2];
                JS[3 * a] = j1 * wp[2] - j2 * wp[1];
                JS[3 * a + 1] = j2 * wp[0] - j0 * wp[2];
                JS[3 * a + 2] = j0 * wp[1] - j1 * wp[0];
            }
            double* Ja = D.Ja + 12 * (size_t)i;
            double* Jo = D.Jo + 12 * (size_t)i;
            const double wp_ = typ == 2 ? 0.0 : w;   // anchor-frame right-camera block: no pose block at all
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 3; ++b) {
                    Ja[6 * a + b] = wp_ * JR[3 * a + b];
                    Ja[6 * a + 3 + b] = -wp_ * JS[3 * a + b];
                    Jo[6 * a + b] = -wp_ * JR[3 * a + b];
                    Jo[6 * a + 3 + b] = wp_ * JS[3 * a + b];
                }
            // J_lambda = -zanch * Rwanch * anchpt   (type 2: -zanch * anchpt, the point never leaves the anchor frame)
            const double* lp = typ == 2 ? ap : rp;
            const double jl[3] = {-zanch * lp[0], -zanch * lp[1], -zanch * lp[2]};
            D.Jl[2 * (size_t)i] = w * (JR[0] * jl[0] + JR[1] * jl[1] + JR[2] * jl[2]);
            D.Jl[2 * (size_t)i + 1] = w * (JR[3] * jl[0] + JR[4] * jl[1] + JR[5] * jl[2]);
            D.Jr[2 * (size_t)i] = w * r0;
            D.Jr[2 * (size_t)i + 1] = w * r1;
        }
    }
    cost = warp_sum(cost);
    if ((threadIdx.x & 31) == 0 && cost != 0.0) atomicAdd(D.scal + (JAC ? SC_COST : SC_CAND_COST), cost);
}

// ------------------------------------------------------------------ Schur elimination (warp / landmark)
constexpr int SCHUR_WARPS = 4;

__global__ void __launch_bounds__(SCHUR_WARPS * 32) ba_schur_kernel(BaDev D, double radius, int first_iter) {
    __shared__ double s_etf[SCHUR_WARPS][MAX_VAR_CAMS + 1][6];
    __shared__ int s_slot[SCHUR_WARPS][MAX_VAR_CAMS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l = blockIdx.x * SCHUR_WARPS + warp;
    if (l >= D.npts) return;
    const int p0 = D.lm_ptr[l], p1 = D.lm_ptr[l + 1];
    const int n = D.n;
    // privatised accumulation copy of this CTA: the (6 Ncv)^2 system has few addresses, thousands of
    // fp64 REDs per address serialise in L2; ncopy copies divide that contention, the reduced-solve
    // kernel sums them when it loads the system
    const long long coff = (long long)(blockIdx.x % D.ncopy) * D.copy_stride;
    double* const cS = D.S + coff; double* const cRhs = D.rhs + coff; double* const cG = D.gcam + coff; double* const cCn = D.cn_cam + coff;
    // E'E, E'r
    double cnl = 0.0, ge = 0.0;
    int nact = 0;
    for (int p = p0 + lane; p < p1; p += 32) {
        if (!D.active[p]) continue;
        const double a = D.Jl[2 * (size_t)p], b = D.Jl[2 * (size_t)p + 1];
        cnl += a * a + b * b;
        ge += a * D.Jr[2 * (size_t)p] + b * D.Jr[2 * (size_t)p + 1];
        nact++;
    }
    cnl = warp_sum(cnl);
    ge = warp_sum(ge);
    nact = __reduce_add_sync(FULL, nact);
    if (nact == 0) {   // unused parameter block: dropped from the program (program.cc:305-387)
        if (lane == 0) { D.ete[l] = 0.0; D.ge[l] = 0.0; }
        return;
    }
    double sc = D.sc_lm[l];
    if (first_iter) {
        sc = 1.0 / (1.0 + sqrt(cnl));
        if (lane == 0) D.sc_lm[l] = sc;
    }
    const double diag = fmin(fmax(cnl * sc * sc, 1e-6), 1e32);
    const double ete = cnl + diag / (radius * sc * sc);
    const double inv_ete = 1.0 / ete;
    if (lane == 0) {
        D.ete[l] = ete;
        D.ge[l] = ge;
        atomic_max_pos(D.scal + SC_GMAX_LM, fabs(ge));   // per-rank maximum; summed over ranks when sharded (upper bound)
    }
    // touching variable cameras: entry 0 = anchor, then one per active observation
    const int sa = D.cam_slot[D.lm_anchor_cam[l]];
    int m = 0;   // number of entries (uniform across the warp)
    if (sa >= 0) {
        if (lane < 6) s_etf[warp][0][lane] = 0.0;
        if (lane == 0) s_slot[warp][0] = sa;
        m = 1;
    }
    __syncwarp();
    // per observation (serial over the landmark's observations, lanes over matrix entries)
    for (int p = p0; p < p1; ++p) {
        if (!D.active[p]) continue;
        const bool lm_only = D.obs_type && D.obs_type[p] == 2;   // e-block-only row (schur_eliminator_impl.h:196-217)
        if (lm_only) continue;
        const int so = D.cam_slot[D.obs_cam[p]];
        const double* Ja = D.Ja + 12 * (size_t)p;
        const double* Jo = D.Jo + 12 * (size_t)p;
        const double jl0 = D.Jl[2 * (size_t)p], jl1 = D.Jl[2 * (size_t)p + 1];
        const double r0 = D.Jr[2 * (size_t)p], r1 = D.Jr[2 * (size_t)p + 1];
        if (sa >= 0) {
            // anchor block: F'F (upper triangle of the 6x6), F'r, column norms, E'F
            for (int e = lane; e < 36; e += 32) {
                const int a = e / 6, b = e - 6 * a;
                if (a <= b) atomicAdd(cS + (size_t)(6 * sa + a) * n + 6 * sa + b, Ja[a] * Ja[b] + Ja[6 + a] * Ja[6 + b]);
            }
            if (lane < 6) {
                atomicAdd(cG + 6 * sa + lane, Ja[lane] * r0 + Ja[6 + lane] * r1);
                atomicAdd(cCn + 6 * sa + lane, Ja[lane] * Ja[lane] + Ja[6 + lane] * Ja[6 + lane]);
                s_etf[warp][0][lane] += jl0 * Ja[lane] + jl1 * Ja[6 + lane];
            }
        }
        if (so >= 0) {
            for (int e = lane; e < 36; e += 32) {
                const int a = e / 6, b = e - 6 * a;
                if (a <= b) atomicAdd(cS + (size_t)(6 * so + a) * n + 6 * so + b, Jo[a] * Jo[b] + Jo[6 + a] * Jo[6 + b]);
            }
            // E'F row of this camera: a stereo keyframe contributes two residual blocks (left and right
            // camera) to the same pose block, so look the slot up before appending a new entry
            int idx = -1;
            for (int q = lane; q < m; q += 32)
                if (s_slot[warp][q] == so) idx = q;
            idx = __reduce_max_sync(FULL, idx);
            const bool fresh = idx < 0;
            if (fresh) idx = m;
            if (lane < 6) {
                atomicAdd(cG + 6 * so + lane, Jo[lane] * r0 + Jo[6 + lane] * r1);
                atomicAdd(cCn + 6 * so + lane, Jo[lane] * Jo[lane] + Jo[6 + lane] * Jo[6 + lane]);
                const double e = jl0 * Jo[lane] + jl1 * Jo[6 + lane];
                s_etf[warp][idx][lane] = fresh ? e : s_etf[warp][idx][lane] + e;
            }
            if (lane == 0 && fresh) s_slot[warp][idx] = so;
            if (sa >= 0) {
                // cross block Ja' Jo into the upper block (min slot, max slot)
                for (int e = lane; e < 36; e += 32) {
                    const int a = e / 6, b = e - 6 * a;   // a: anchor column, b: observer column
                    const double v = Ja[a] * Jo[b] + Ja[6 + a] * Jo[6 + b];
                    if (sa < so) atomicAdd(cS + (size_t)(6 * sa + a) * n + 6 * so + b, v);
                    else atomicAdd(cS + (size_t)(6 * so + b) * n + 6 * sa + a, v);
                }
            }
            if (fresh) m++;
        }
        __syncwarp();
    }
    // Schur complement: S[i,j] -= EtF_i' EtF_j / ete (upper blocks), rhs_i = F'r - EtF_i ge / ete
    const int npair = m * m;
    for (int e = lane; e < npair * 36; e += 32) {
        const int pr = e / 36, q = e - 36 * pr;
        const int i = pr / m, j = pr - m * i;
        const int si = s_slot[warp][i], sj = s_slot[warp][j];
        if (si > sj || (si == sj && i > j)) continue;   // upper block triangle; (i,i) once
        const int a = q / 6, b = q - 6 * a;
        if (si == sj && a > b) continue;
        const double v = s_etf[warp][i][a] * s_etf[warp][j][b] * inv_ete;   // slots are unique in the list
        atomicAdd(cS + (size_t)(6 * si + a) * n + 6 * sj + b, -v);
    }
    for (int e = lane; e < m * 6; e += 32) {
        const int i = e / 6, a = e - 6 * i;
        atomicAdd(cRhs + 6 * s_slot[warp][i] + a, -s_etf[warp][i][a] * ge * inv_ete);
    }
}

// ------------------------------------------------------------------ fold of the privatised copies
// The Schur kernel accumulates into ncopy private copies of [rhs | F'r | column norms | S] to spread the
// fp64 atomics; this grid-wide pass sums them into copy 0.  (Inside the single-CTA reduced solve the same
// fold cost ~30 % of that kernel's time: one SM cannot keep enough L2 loads in flight.)
__global__ void __launch_bounds__(256) ba_fold_kernel(BaDev D) {
    const int blk = 3 * D.n + D.n * D.n;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= blk) return;
    double acc = D.rhs[e];
#pragma unroll 8
    for (int k = 1; k < D.ncopy; ++k) acc += D.rhs[e + (long long)k * D.copy_stride];
    D.rhs[e] = acc;
}

// ------------------------------------------------------------------ reduced camera system (one CTA)
// FP64 tensor-core tile product (DMMA): D(8x8) = A(8x4) B(4x8) + C.  Fragment layout (PTX ISA,
// mma.m8n8k4 .f64): lane = 4 g + t; A: (row g, col t); B: (row t, col g); C/D: (row g, cols 2t, 2t+1).
__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int CH_NB = 32;   // block size of the blocked Cholesky (mode 4)

template <int MODE>   // 0: Cholesky in global/L2, 1: Cholesky in shared memory, 2 / 3: Gauss-Jordan in registers (n <= 63 / n <= 96),
                      // 4: blocked Cholesky, trailing update on the FP64 tensor cores (n > 96)
__global__ void __launch_bounds__(MODE >= 2 ? 512 : 1024) ba_reduced_solve_kernel(BaDev D, double radius, int first_iter) {
    extern __shared__ double sA[];
    __shared__ int s_fail;
    const int n = D.n, tid = threadIdx.x, nt = blockDim.x;
    constexpr bool SMEM = MODE == 1;
    double* A;
    if (SMEM) A = sA; else A = D.S;
    __shared__ double s_w[MAX_N];
    double* w = s_w;   // rhs -> solution vector (shared: the triangular solves are latency chains)
    if (tid == 0) s_fail = 0;
    // Jacobi scaling (iteration 0), LM damping, rhs = F'r - (Schur part already accumulated)
    for (int i = tid; i < n; i += nt) {
        const double cn = D.cn_cam[i];
        double sc = D.sc_cam[i];
        if (first_iter) {
            sc = 1.0 / (1.0 + sqrt(cn));
            D.sc_cam[i] = sc;
        }
        const double diag = fmin(fmax(cn * sc * sc, 1e-6), 1e32);
        D.S[(size_t)i * n + i] += diag / (radius * sc * sc);
        w[i] = D.gcam[i] + D.rhs[i];
    }
    __syncthreads();
    if (MODE == 2 || MODE == 3) {
        // Gauss-Jordan elimination on the augmented system [S | b] held in REGISTERS: warp w owns rows
        // w*RPW .. w*RPW+RPW-1, lane l owns columns l + 32 k.  Each pivot step the owners publish pivot
        // row j and column j through a double-buffered shared-memory line, then everyone updates its
        // registers: ONE barrier per pivot and no substitution passes afterwards (a Cholesky + two
        // triangular solves is 3n dependent steps).  The block has only ceil(n / RPW) warps: ncu showed
        // the 32-warp version issue-bound on per-warp overhead (331 k warp-instructions for a 48 x 48
        // system), not latency-bound.  Pivots equal those of the LDL' / Cholesky factorisation, so
        // "pivot <= 0" is the failure test Ceres' LLT applies.  No pivoting: S is SPD after LM damping.
        constexpr int RPW = MODE == 2 ? 4 : 6;   // rows per warp     (n <= 63 -> <= 16 warps | n <= 96 -> 16 warps)
        constexpr int CPL = MODE == 2 ? 2 : 4;   // columns per lane  (n + 1 <= 64 | n + 1 <= 128)
        __shared__ double s_row[2][MAX_N + 8];
        __shared__ double s_col[2][MAX_N + 8];
        const int wp = tid >> 5, ln = tid & 31;
        double a[RPW][CPL];
#pragma unroll
        for (int i = 0; i < RPW; ++i)
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int r = wp * RPW + i, c = ln + 32 * k;
                double v = 0.0;
                if (r < n && c <= n) v = c == n ? w[r] : (r <= c ? D.S[(size_t)r * n + c] : D.S[(size_t)c * n + r]);
                a[i][k] = v;
            }
        for (int j = 0; j < n; ++j) {
            const int b = j & 1, jw = j / RPW, ji = j - jw * RPW, jl = j & 31, jk = j >> 5;
            if (wp == jw) {
#pragma unroll
                for (int i = 0; i < RPW; ++i)
                    if (i == ji) {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) if (ln + 32 * k <= n) s_row[b][ln + 32 * k] = a[i][k];
                    }
            }
            if (ln == jl) {
#pragma unroll
                for (int k = 0; k < CPL; ++k)
                    if (k == jk) {
#pragma unroll
                        for (int i = 0; i < RPW; ++i) if (wp * RPW + i < n) s_col[b][wp * RPW + i] = a[i][k];
                    }
            }
            __syncthreads();
            const double p = s_row[b][j];
            if (!(p > 0.0) || !isfinite(p)) { if (tid == 0) s_fail = 1; break; }   // uniform
            double ip = (double)__frcp_rn((float)p);
            ip = ip * (2.0 - p * ip);
            ip = ip * (2.0 - p * ip);
            double cj[RPW];
#pragma unroll
            for (int i = 0; i < RPW; ++i) cj[i] = s_col[b][min(wp * RPW + i, n - 1)];
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int c = ln + 32 * k;
                const double ajc = c <= n ? s_row[b][c] * ip : 0.0;
#pragma unroll
                for (int i = 0; i < RPW; ++i) {
                    const int r = wp * RPW + i;
                    a[i][k] = r == j ? ajc : a[i][k] - cj[i] * ajc;
                }
            }
        }
        __syncthreads();
        if (s_fail) {
            if (tid == 0) D.scal[SC_CHOL_FAIL] = 1.0;
            return;
        }
        if (ln == (n & 31)) {
#pragma unroll
            for (int k = 0; k < CPL; ++k)
                if (k == (n >> 5)) {
#pragma unroll
                    for (int i = 0; i < RPW; ++i) if (wp * RPW + i < n) w[wp * RPW + i] = a[i][k];
                }
        }
        __syncthreads();
    } else if (MODE == 5) {
        // ONE warp, Gauss-Jordan on the augmented system in shared memory (ge_warp.cuh): no block barrier,
        // one __syncwarp per pivot.  Launched with 32 threads; selected with OV2_BA_SOLVER=5 (candidate
        // replacement of modes 2/3, not yet measured on a B200).
        const int P = gewarp::pitch(n);
        for (int e = tid; e < n * n; e += nt) {
            const int r = e / n, c = e - r * n;
            sA[r * P + c] = r <= c ? D.S[(size_t)r * n + c] : D.S[(size_t)c * n + r];
        }
        for (int i = tid; i < n; i += nt) sA[i * P + n] = w[i];
        __syncwarp();
        bool ok = true;
        for (int j = 0; j < n && ok; ++j) {
            ok = gewarp::step(tid, sA, n, P, j);            // uniform: every lane reads the same pivot
            __syncwarp();
        }
        if (!ok) {
            if (tid == 0) D.scal[SC_CHOL_FAIL] = 1.0;
            return;
        }
        gewarp::finish(tid, sA, n, P, w);
        __syncwarp();
    } else if (MODE == 4) {
        // Blocked right-looking Cholesky S = U'U (upper triangle, in global/L2), block size 32, on the
        // augmented system [S | b] so the forward substitution U'y = b comes out of the panel step:
        //   (1) warp 0 factors the 32 x 32 diagonal block in REGISTERS (lane = column, shuffles only);
        //   (2) one thread per trailing column solves U11' x = a  (panel row block U12, and y_k for b);
        //   (3) A22 -= U12' U12 is a true contraction (k = 32): 16 x 16 warp tiles on the FP64 tensor
        //       cores (mma.sync m8n8k4 -> DMMA), operands staged in shared memory (sP);
        // then the backward substitution runs block by block (GEMV by all warps + a 32-step register
        // solve by warp 0).  9 block steps for the C5 window (n = 288) instead of 288 column steps,
        // each of which cost an L2 round trip in mode 0.
        const int PW = ((n + 15) & ~15) + 8;                 // sP row pitch (doubles): bank-spread for the fragment loads
        double* sP = sA;                                     // [32][PW]  current row panel U12
        __shared__ double sU[CH_NB][CH_NB + 1];              // diagonal block U11
        __shared__ double s_idiag_all[MAX_N + CH_NB];        // 1 / U[j][j] (identity padding of the last block included)
        __shared__ double s_t[CH_NB];
        const int warp = tid >> 5, lane = tid & 31, nwarp = nt >> 5;
        for (int kb = 0; kb < n; kb += CH_NB) {
            const int nb = min(CH_NB, n - kb), c1 = kb + nb, m = n - c1;
            double* s_idiag = s_idiag_all + kb;
            // ---- (1) diagonal block: load, factor in registers
            for (int e = tid; e < CH_NB * CH_NB; e += nt) {
                const int i = e >> 5, j = e & 31;
                sU[i][j] = (i < nb && j < nb && i <= j) ? A[(size_t)(kb + i) * n + kb + j] : (i == j ? 1.0 : 0.0);
            }
            __syncthreads();
            if (warp == 0) {
                // lane = column c of the block; row j is final when step j starts.  The trailing rows are
                // updated in shared memory (row pitch 33: conflict-free), 31 - j independent FMAs per step.
                bool bad = false;
                for (int j = 0; j < CH_NB; ++j) {
                    const double d = sU[j][j];
                    // pivots of the Jacobi-scaled, damped system are O(1); outside the float range the
                    // factorisation is reported as failed (Ceres' LLT would return a useless step there)
                    bad = bad || !(d > 1e-30) || !(d < 1e30);
                    double is = (double)rsqrtf((float)d);                    // MUFU seed + 2 Newton steps in double:
                    is = is * (1.5 - 0.5 * d * is * is);                     // no library call on the 32-step
                    is = is * (1.5 - 0.5 * d * is * is);                     // critical path, error < 1 ulp
                    const double ujc = lane >= j ? sU[j][lane] * is : 0.0;   // row j of U (diagonal: d / sqrt(d))
                    __syncwarp();
                    sU[j][lane] = ujc;
                    if (lane == j) s_idiag[j] = is;
                    __syncwarp();
#pragma unroll 4
                    for (int r = j + 1; r < CH_NB; ++r)
                        if (lane >= r) sU[r][lane] -= sU[j][r] * ujc;
                    __syncwarp();
                }
                if (bad && lane == 0) s_fail = 1;
            }
            __syncthreads();
            if (s_fail) break;                                               // uniform
            for (int e = tid; e < CH_NB * CH_NB; e += nt) {
                const int i = e >> 5, j = e & 31;
                if (i < nb && j < nb && i <= j) A[(size_t)(kb + i) * n + kb + j] = sU[i][j];
            }
            // ---- (2) panel: thread per column c in (c1 .. n-1) and the rhs column (c == n)
            const int m16 = (m + 15) & ~15;
            for (int cc = tid; cc <= m16; cc += nt) {
                if (cc >= m && cc < m16) {                                   // zero padding of the operand panel
#pragma unroll
                    for (int i = 0; i < CH_NB; ++i) sP[i * PW + cc] = 0.0;
                    continue;
                }
                const bool is_rhs = cc == m16;
                if (!is_rhs && cc >= m) continue;
                double a[CH_NB];
#pragma unroll
                for (int i = 0; i < CH_NB; ++i)
                    a[i] = i < nb ? (is_rhs ? w[kb + i] : A[(size_t)(kb + i) * n + c1 + cc]) : 0.0;
#pragma unroll
                for (int k = 0; k < CH_NB; ++k) {
                    const double x = a[k] * s_idiag[k];
                    a[k] = x;
#pragma unroll
                    for (int i = k + 1; i < CH_NB; ++i) a[i] -= sU[k][i] * x;
                }
                if (is_rhs) {
#pragma unroll
                    for (int i = 0; i < CH_NB; ++i) if (i < nb) { w[kb + i] = a[i]; s_t[i] = a[i]; }
                } else {
#pragma unroll
                    for (int i = 0; i < CH_NB; ++i) {
                        sP[i * PW + cc] = a[i];
                        if (i < nb) A[(size_t)(kb + i) * n + c1 + cc] = a[i];
                    }
                }
            }
            __syncthreads();
            // ---- (3) trailing update.  rhs: w[c1 + r] -= sum_k U12[k][r] y_k
            for (int r = tid; r < m; r += nt) {
                double acc = 0.0;
#pragma unroll 8
                for (int k = 0; k < CH_NB; ++k) acc += sP[k * PW + r] * s_t[k];
                w[c1 + r] -= acc;
            }
            // A22[r][c] -= sum_k U12[k][r] U12[k][c], r <= c: 16 x 16 tiles per warp, 4 DMMA accumulators
            {
                const int mt = m16 >> 4, g = lane >> 2, t = lane & 3;
                for (int idx = warp; idx < mt * mt; idx += nwarp) {
                    const int tr = idx / mt, tc = idx - tr * mt;
                    if (tr > tc) continue;
                    const int r0 = tr * 16, q0 = tc * 16;
                    double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
                    for (int k0 = 0; k0 < CH_NB; k0 += 4) {
                        const double* pk = sP + (k0 + t) * PW;
                        const double a0 = pk[r0 + g], a1 = pk[r0 + 8 + g];
                        const double b0 = pk[q0 + g], b1 = pk[q0 + 8 + g];
                        dmma_884(acc[0][0][0], acc[0][0][1], a0, b0);
                        dmma_884(acc[0][1][0], acc[0][1][1], a0, b1);
                        dmma_884(acc[1][0][0], acc[1][0][1], a1, b0);
                        dmma_884(acc[1][1][0], acc[1][1][1], a1, b1);
                    }
#pragma unroll
                    for (int hi = 0; hi < 2; ++hi)
#pragma unroll
                        for (int hj = 0; hj < 2; ++hj)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int r = r0 + 8 * hi + g, c = q0 + 8 * hj + 2 * t + e;
                                if (r <= c && c < m) A[(size_t)(c1 + r) * n + c1 + c] -= acc[hi][hj][e];
                            }
                }
            }
            __syncthreads();
        }
        __syncthreads();
        if (s_fail) {
            if (tid == 0) D.scal[SC_CHOL_FAIL] = 1.0;
            return;
        }
        // ---- backward substitution U z = y, last block first
        for (int kb = ((n - 1) / CH_NB) * CH_NB; kb >= 0; kb -= CH_NB) {
            const int nb = min(CH_NB, n - kb), c1 = kb + nb;
            for (int e = tid; e < CH_NB * CH_NB; e += nt) {
                const int i = e >> 5, j = e & 31;
                sU[i][j] = (i < nb && j < nb && i <= j) ? A[(size_t)(kb + i) * n + kb + j] : (i == j ? 1.0 : 0.0);
            }
            for (int i = warp; i < nb; i += nwarp) {                         // t_i = y_i - U12[i][:] z_rest
                double acc = 0.0;
                for (int c = c1 + lane; c < n; c += 32) acc += A[(size_t)(kb + i) * n + c] * w[c];
                acc = warp_sum(acc);
                if (lane == 0) s_t[i] = w[kb + i] - acc;
            }
            __syncthreads();
            if (warp == 0) {
                double ti = lane < nb ? s_t[lane] : 0.0;
                const double idg = lane < nb ? s_idiag_all[kb + lane] : 1.0;
#pragma unroll
                for (int j = CH_NB - 1; j >= 0; --j) {
                    const double zj = __shfl_sync(FULL, ti * idg, j);
                    if (lane < j) ti -= sU[lane][j] * zj;
                    if (lane == j && j < nb) w[kb + j] = zj;
                }
            }
            __syncthreads();
        }
    } else {
    if (SMEM) {
        for (int e = tid; e < n * n; e += nt) sA[e] = D.S[e];
        __syncthreads();
    }
    // Cholesky A = U'U on the upper triangle, right-looking, ONE barrier per column: row j is only
    // ever read in step j, so its scaling by 1/sqrt(d_j) is folded into the trailing update and into
    // the triangular solves (U[j][c] = A[j][c] * s_ipiv[j], U[j][j] = 1 / s_ipiv[j]).
    __shared__ double s_ipiv[MAX_N];
    const int warp = tid >> 5, lane = tid & 31, nwarp = nt >> 5;
    for (int j = 0; j < n; ++j) {
        const double d = A[(size_t)j * n + j];          // final since the previous barrier
        if (!(d > 0.0) || !isfinite(d)) { if (tid == 0) s_fail = 1; break; }   // uniform
        // 1/d sits on the per-column critical path: float reciprocal + two Newton steps in double
        // (relative error ~2^-92 before rounding) instead of the ~300-cycle IEEE division routine
        double inv_d = (double)__frcp_rn((float)d);
        inv_d = inv_d * (2.0 - d * inv_d);
        inv_d = inv_d * (2.0 - d * inv_d);
        if (tid == 0) s_ipiv[j] = rsqrt(d);
        // trailing update A[r][c] -= A[j][r] A[j][c] / d for j < r <= c: one warp per row, lanes over columns
        for (int r = j + 1 + warp; r < n; r += nwarp) {
            const double ajr = A[(size_t)j * n + r] * inv_d;
            for (int c = r + lane; c < n; c += 32) A[(size_t)r * n + c] -= ajr * A[(size_t)j * n + c];
        }
        __syncthreads();
    }
    __syncthreads();
    if (s_fail) {
        if (tid == 0) D.scal[SC_CHOL_FAIL] = 1.0;
        return;
    }
    // triangular solves U' y = b, U z = y by warp 0 alone (n sequential steps each; __syncwarp is far
    // cheaper than a block barrier); w[] is global, A may be shared or global
    if (tid < 32) {
        for (int j = 0; j < n; ++j) {
            const double ip = s_ipiv[j];
            const double yj = w[j] * ip;                 // / U[j][j]
            __syncwarp();
            if (tid == 0) w[j] = yj;
            for (int c = j + 1 + tid; c < n; c += 32) w[c] -= A[(size_t)j * n + c] * ip * yj;
            __syncwarp();
        }
        for (int j = n - 1; j >= 0; --j) {
            const double zj = w[j] * s_ipiv[j];
            __syncwarp();
            if (tid == 0) w[j] = zj;
            for (int r = tid; r < j; r += 32) w[r] -= A[(size_t)r * n + j] * s_ipiv[r] * zj;
            __syncwarp();
        }
    }
    __syncthreads();
    }   // Cholesky modes
    for (int i = tid; i < n; i += nt) D.z[i] = w[i];   // the back-substitution kernel reads z from global
}

// ------------------------------------------------------------------ back-substitution (warp / landmark)
// The camera-side Plus() work rides along as two extra CTAs (candidate poses; gradient projection):
// SE3 exp is a long chain of fp64 software routines (sqrt, division, sincos), ~20 us of pure latency
// for a handful of threads - here it overlaps the landmark CTAs instead of extending the
// single-CTA reduced solve.
__global__ void __launch_bounds__(128) ba_backsub_kernel(BaDev D, const double* __restrict__ pose, double* __restrict__ cand_pose,
                                                         const double* __restrict__ invd, double* __restrict__ cand_invd,
                                                         int nlm_blocks, int add_cam_norms, int want_gmax) {
    if ((int)blockIdx.x >= nlm_blocks) {
        const int which = blockIdx.x - nlm_blocks;      // 0: candidate = Plus(x, -z), 1: |x - Plus(x, -g)|_inf
        if (which == 1 && !want_gmax) return;
        double st2 = 0.0, cx2 = 0.0, gm = 0.0;
        for (int c = threadIdx.x; c < D.ncam; c += blockDim.x) {
            const int s = D.cam_slot[c];
            if (s < 0) continue;
            double d[6], out[7];
            const double* v = which == 0 ? D.z : D.gcam;
            for (int k = 0; k < 6; ++k) d[k] = -v[6 * s + k];
            pose_plus(pose + 7 * c, d, out);
            for (int k = 0; k < 7; ++k) {
                const double df = pose[7 * c + k] - out[k];
                if (which == 0) { cand_pose[7 * c + k] = out[k]; st2 += df * df; cx2 += out[k] * out[k]; }
                else gm = fmax(gm, fabs(df));
            }
        }
        st2 = warp_sum(st2);
        cx2 = warp_sum(cx2);
        for (int o = 16; o > 0; o >>= 1) gm = fmax(gm, __shfl_xor_sync(FULL, gm, o));
        if ((threadIdx.x & 31) == 0) {
            // sharded solve: every rank solves the same reduced system; only rank 0 contributes the camera norms
            if (which == 0 && add_cam_norms && st2 != 0.0) atomicAdd(D.scal + SC_STEP2, st2);
            if (which == 0 && add_cam_norms && cx2 != 0.0) atomicAdd(D.scal + SC_CANDX2, cx2);
            if (which == 1) atomic_max_pos(D.scal + SC_GMAX_CAM, gm);
        }
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l = blockIdx.x * 4 + warp;
    if (l >= D.npts) return;
    const double ete = D.ete[l];
    if (ete == 0.0) return;   // landmark not in the program: candidate stays equal
    const int p0 = D.lm_ptr[l], p1 = D.lm_ptr[l + 1];
    const int sa = D.cam_slot[D.lm_anchor_cam[l]];
    double za[6] = {0, 0, 0, 0, 0, 0};
    if (sa >= 0)
        for (int k = 0; k < 6; ++k) za[k] = D.z[6 * sa + k];
    // y_e = (E'r - sum_obs Jl' (Ja z_a + Jo z_o)) / ete
    double acc = 0.0;
    for (int p = p0 + lane; p < p1; p += 32) {
        if (!D.active[p]) continue;
        const int so = (D.obs_type && D.obs_type[p] == 2) ? -1 : D.cam_slot[D.obs_cam[p]];
        const double* Ja = D.Ja + 12 * (size_t)p;
        const double* Jo = D.Jo + 12 * (size_t)p;
        double f0 = 0.0, f1 = 0.0;
        if (sa >= 0)
            for (int k = 0; k < 6; ++k) { f0 += Ja[k] * za[k]; f1 += Ja[6 + k] * za[k]; }
        if (so >= 0)
            for (int k = 0; k < 6; ++k) { const double zk = D.z[6 * so + k]; f0 += Jo[k] * zk; f1 += Jo[6 + k] * zk; }
        acc += D.Jl[2 * (size_t)p] * f0 + D.Jl[2 * (size_t)p + 1] * f1;
    }
    acc = warp_sum(acc);
    const double y = (D.ge[l] - acc) / ete;
    const double dl = -y;
    const double cl = invd[l] + dl;
    // model cost change: -(J delta)'(r + J delta / 2), delta = -[z; y]
    double mcc = 0.0;
    for (int p = p0 + lane; p < p1; p += 32) {
        if (!D.active[p]) continue;
        const int so = (D.obs_type && D.obs_type[p] == 2) ? -1 : D.cam_slot[D.obs_cam[p]];
        const double* Ja = D.Ja + 12 * (size_t)p;
        const double* Jo = D.Jo + 12 * (size_t)p;
        double f0 = D.Jl[2 * (size_t)p] * dl, f1 = D.Jl[2 * (size_t)p + 1] * dl;
        if (sa >= 0)
            for (int k = 0; k < 6; ++k) { f0 -= Ja[k] * za[k]; f1 -= Ja[6 + k] * za[k]; }
        if (so >= 0)
            for (int k = 0; k < 6; ++k) { const double zk = D.z[6 * so + k]; f0 -= Jo[k] * zk; f1 -= Jo[6 + k] * zk; }
        mcc -= f0 * (D.Jr[2 * (size_t)p] + 0.5 * f0) + f1 * (D.Jr[2 * (size_t)p + 1] + 0.5 * f1);
    }
    mcc = warp_sum(mcc);
    if (lane == 0) {
        cand_invd[l] = cl;
        atomicAdd(D.scal + SC_MCC, mcc);
        atomicAdd(D.scal + SC_STEP2, dl * dl);
        atomicAdd(D.scal + SC_CANDX2, cl * cl);
    }
}

// ------------------------------------------------------------------ outlier scan (optimizer.cpp:500-530)
__global__ void ba_flag_kernel(BaDev D, double th, int bit, int deactivate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0, left = 0, right = 0;
    if (i < D.nobs && D.active[i]) {
        bad = (D.chi2[i] > th) || !D.dpos[i];
        if (bad) {
            D.flags[i] |= (uint8_t)bit;
            if (deactivate) D.active[i] = 0;   // problem.RemoveResidualBlock
        } else {
            const int typ = D.obs_type ? (int)D.obs_type[i] : 0;
            left = typ == 0;
            right = typ == 1;
        }
    }
    const int nb = __reduce_add_sync(FULL, bad), nl = __reduce_add_sync(FULL, left), nr = __reduce_add_sync(FULL, right);
    if ((threadIdx.x & 31) == 0) {
        if (nb) atomicAdd(D.scal + SC_NBAD, (double)nb);
        if (nl) atomicAdd(D.scal + SC_NLEFT, (double)nl);     // surviving vreprojerr_kfid_lmid entries
        if (nr) atomicAdd(D.scal + SC_NRIGHT, (double)nr);    // surviving vright_reprojerr_kfid_lmid entries
    }
}

__global__ void ba_cam_used_kernel(BaDev D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D.nobs && D.active[i] && !(D.obs_type && D.obs_type[i] == 2)) {
        D.cam_used[D.obs_cam[i]] = 1;
        D.cam_used[D.lm_anchor_cam[D.obs_lm[i]]] = 1;
    }
}

struct Summary { int iterations; double initial_cost, final_cost; int termination; };

// multi-GPU sharding (landmarks partitioned over ranks): sum-allreduce hook, NULL = single GPU
struct Shard { ov2_allreduce_fn fn; void* user; int rank; };

}  // namespace

// One ceres::Solve on the currently active residual blocks.
static ov2_status ba_ceres_solve(ov2_ctx* ctx, BaDev& D, double*& pose, double*& cand_pose, double*& invd,
                                 double*& cand_invd, int max_iters, double function_tolerance, Summary* out, const Shard* sh) {
    cudaStream_t st = ctx->stream;
    const int nobs = D.nobs;
    // Program::RemoveFixedBlocks: cameras that are constant or touch no active residual drop out
    OV2_CUDA(ctx, cudaMemsetAsync(D.cam_used, 0, D.ncam, st));
    OV2_LAUNCH(ctx, "ba_cam_used_kernel", ba_cam_used_kernel<<<grid1(nobs, 256), 256, 0, st>>>(D));
    std::vector<uint8_t> used(D.ncam), cst(D.ncam);
    if (sh && sh->fn) {
        // a camera is in the program if ANY rank has an active residual touching it
        std::vector<uint8_t> lu(D.ncam);
        OV2_CUDA(ctx, cudaMemcpyAsync(lu.data(), D.cam_used, D.ncam, cudaMemcpyDeviceToHost, st));
        OV2_CUDA(ctx, cudaStreamSynchronize(st));
        std::vector<double> ud(D.ncam);
        for (int c = 0; c < D.ncam; ++c) ud[c] = lu[c];
        OV2_CUDA(ctx, cudaMemcpyAsync(D.scal, ud.data(), sizeof(double) * D.ncam, cudaMemcpyHostToDevice, st));
        if (sh->fn(sh->user, D.scal, (size_t)D.ncam, (void*)st) != 0) return ov2_fail(ctx, OV2_ERR_CUDA, "allreduce callback failed");
        OV2_CUDA(ctx, cudaMemcpyAsync(ud.data(), D.scal, sizeof(double) * D.ncam, cudaMemcpyDeviceToHost, st));
        OV2_CUDA(ctx, cudaStreamSynchronize(st));
        for (int c = 0; c < D.ncam; ++c) lu[c] = ud[c] > 0.0;
        OV2_CUDA(ctx, cudaMemcpyAsync(D.cam_used, lu.data(), D.ncam, cudaMemcpyHostToDevice, st));
    }
    OV2_CUDA(ctx, cudaMemcpyAsync(used.data(), D.cam_used, D.ncam, cudaMemcpyDeviceToHost, st));
    OV2_CUDA(ctx, cudaMemcpyAsync(cst.data(), D.pose_const, D.ncam, cudaMemcpyDeviceToHost, st));
    OV2_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<int32_t> slot(D.ncam, -1);
    int ncv = 0;
    bool any = false;
    for (int c = 0; c < D.ncam; ++c) {
        any = any || used[c];
        if (used[c] && !cst[c]) slot[c] = ncv++;
    }
    out->iterations = 0;
    out->initial_cost = out->final_cost = 0.0;
    out->termination = 0;
    if (!any) return OV2_OK;   // no residual blocks (on any rank, when sharded)
    if (ncv > MAX_VAR_CAMS) return ov2_fail(ctx, OV2_ERR_CAPACITY, "ov2_localba_solve: more than 64 optimised keyframes");
    D.ncv = ncv;
    D.n = 6 * ncv;
    const int n = D.n;
    OV2_CUDA(ctx, cudaMemcpyAsync(D.cam_slot, slot.data(), sizeof(int32_t) * D.ncam, cudaMemcpyHostToDevice, st));
    D.rhs = D.scal + SC_COUNT;
    D.gcam = D.rhs + n;
    D.cn_cam = D.gcam + n;
    D.S = D.cn_cam + n;
    D.copy_stride = 3 * (long long)n + (long long)n * n;
    {
        // as many copies as fit the scratch block (sized for one MAX_N system), at most 8; the sharded
        // path keeps one copy because the block is what gets all-reduced
        long long cap = 3LL * MAX_N + (long long)MAX_N * MAX_N;
        long long k = D.copy_stride > 0 ? cap / D.copy_stride : 1;
        D.ncopy = (int)(k < 1 ? 1 : (k > 8 ? 8 : k));
        if (sh && sh->fn) D.ncopy = 1;
        if (getenv("OV2_BA_NCOPY")) { int e = atoi(getenv("OV2_BA_NCOPY")); if (e >= 1 && e <= D.ncopy) D.ncopy = e; }
    }
    // reduced-system solver: register Gauss-Jordan for small windows, blocked tensor-core Cholesky beyond
    // (C5: 288 x 288); the unblocked Cholesky paths (0: global/L2, 1: shared memory) stay selectable
    int solve_mode = n <= 63 ? 2 : (n <= 96 ? 3 : 4);
    if (getenv("OV2_BA_SOLVER")) {
        const int e = atoi(getenv("OV2_BA_SOLVER"));   // 0 / 1 / 4 force a Cholesky path (tests, comparisons)
        if (e == 0 || (e == 4 && n > 0) || (e == 1 && (size_t)n * n * sizeof(double) <= 200 * 1024) || (e == 5 && n > 0 && n <= 96)) solve_mode = e;
    }
    const size_t smem_need = solve_mode == 1 ? (size_t)n * n * sizeof(double)             // modes 2/3 live in registers
                           : (solve_mode == 4 ? (size_t)CH_NB * (((n + 15) & ~15) + 8) * sizeof(double)
                           : (solve_mode == 5 ? (size_t)n * gewarp::pitch(n) * sizeof(double) : 0));
    int solve_threads = solve_mode == 2 ? 32 * div_up(n, 4) : (solve_mode == 3 ? 32 * div_up(n, 6) : (solve_mode == 4 ? 512 : (solve_mode == 5 ? 32 : 1024)));
    if (solve_threads < 32) solve_threads = 32;   // n == 0: every pose constant, only landmarks move
    if (smem_need > 24 * 1024) {   // static (up to ~15 KB) + dynamic beyond 48 KB needs the opt-in
        if (solve_mode == 1)
            OV2_CUDA(ctx, cudaFuncSetAttribute(ba_reduced_solve_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need));
        else if (solve_mode == 5)
            OV2_CUDA(ctx, cudaFuncSetAttribute(ba_reduced_solve_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need));
        else
            OV2_CUDA(ctx, cudaFuncSetAttribute(ba_reduced_solve_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need));
    }
    // candidate buffers start equal to x: parameter blocks that are not in this solve's program
    // (constant / unused) must keep their current value through pointer swaps
    OV2_CUDA(ctx, cudaMemcpyAsync(cand_pose, pose, sizeof(double) * 7 * D.ncam, cudaMemcpyDeviceToDevice, st));
    OV2_CUDA(ctx, cudaMemcpyAsync(cand_invd, invd, sizeof(double) * D.npts, cudaMemcpyDeviceToDevice, st));
    // One accumulation buffer [scal | rhs | gcam | cn_cam | S] -> one memset per LM iteration.
    // Per iteration: memset, (Jacobian evaluation if x is new), Schur, reduced solve, back-substitution,
    // candidate cost, ONE readback of the scalars; the controller below replays Ceres' decisions.
    const size_t accum_bytes = sizeof(double) * ((size_t)SC_COUNT + (size_t)D.ncopy * (size_t)D.copy_stride);
    // pinned readback target: a pageable one makes every per-iteration D2H a staged, blocking copy
    if (!ctx->ba_hscal) OV2_CUDA(ctx, cudaHostAlloc((void**)&ctx->ba_hscal, sizeof(double) * 2 * SC_COUNT, cudaHostAllocDefault));
    double* h = ctx->ba_hscal;
    double x_cost = 0.0, minimum_cost = DBL_MAX, xnorm = -1.0, gmax = DBL_MAX;
    double radius = 1e4, decrease_factor = 2.0;
    bool step_successful = true, x_is_new = true, cost_known = false;
    int iteration = 0, num_invalid = 0, first_iter = 1;
    for (;;) {
        // ---- FinalizeIterationAndCheckIfMinimizerCanContinue (x_cost of a just-accepted point equals
        //      its candidate cost; the re-evaluated value replaces it at the next readback)
        if (step_successful && cost_known && x_cost < minimum_cost) minimum_cost = x_cost;
        if (iteration >= max_iters) { out->termination = 1; break; }
        if (radius <= 1e-32) { out->termination = 0; break; }
        iteration++;
        // ---- ComputeTrustRegionStep
        OV2_CUDA(ctx, cudaMemsetAsync(D.scal, 0, accum_bytes, st));
        if (x_is_new)
            OV2_LAUNCH(ctx, "ba_eval_kernel<jac>", ba_eval_kernel<true><<<grid1(nobs, 128), 128, 0, st>>>(D, pose, invd));
        OV2_LAUNCH(ctx, "ba_schur_kernel", ba_schur_kernel<<<grid1(D.npts, SCHUR_WARPS), SCHUR_WARPS * 32, 0, st>>>(D, radius, first_iter));
        if (sh && sh->fn) {
            // the ONE bulk collective per LM iteration: [cost, gmax, rhs, F'r, column norms, S] summed over
            // ranks (NVLink / NVSwitch via NCCL in the caller); every rank then solves the same system
            if (sh->fn(sh->user, D.scal, (size_t)SC_COUNT + 3 * (size_t)n + (size_t)n * n, (void*)st) != 0)
                return ov2_fail(ctx, OV2_ERR_CUDA, "allreduce callback failed");
        }
        if (D.ncopy > 1 && n > 0)
            OV2_LAUNCH(ctx, "ba_fold_kernel", ba_fold_kernel<<<div_up(3 * n + n * n, 256), 256, 0, st>>>(D));
        if (solve_mode == 2)
            OV2_LAUNCH(ctx, "ba_reduced_solve_kernel", ba_reduced_solve_kernel<2><<<1, solve_threads, 0, st>>>(D, radius, first_iter));
        else if (solve_mode == 3)
            OV2_LAUNCH(ctx, "ba_reduced_solve_kernel", ba_reduced_solve_kernel<3><<<1, solve_threads, 0, st>>>(D, radius, first_iter));
        else if (solve_mode == 5)
            OV2_LAUNCH(ctx, "ba_reduced_solve_kernel", ba_reduced_solve_kernel<5><<<1, solve_threads, smem_need, st>>>(D, radius, first_iter));
        else if (solve_mode == 4)
            OV2_LAUNCH(ctx, "ba_reduced_solve_kernel", ba_reduced_solve_kernel<4><<<1, solve_threads, smem_need, st>>>(D, radius, first_iter));
        else if (solve_mode == 1)
            OV2_LAUNCH(ctx, "ba_reduced_solve_kernel", ba_reduced_solve_kernel<1><<<1, solve_threads, smem_need, st>>>(D, radius, first_iter));
        else
            OV2_LAUNCH(ctx, "ba_reduced_solve_kernel", ba_reduced_solve_kernel<0><<<1, solve_threads, 0, st>>>(D, radius, first_iter));
        {
            const int nlm_blocks = div_up(D.npts, 4);
            OV2_LAUNCH(ctx, "ba_backsub_kernel",
                       ba_backsub_kernel<<<nlm_blocks + 2, 128, 0, st>>>(D, pose, cand_pose, invd, cand_invd, nlm_blocks,
                                                                         (!sh || sh->rank == 0) ? 1 : 0, x_is_new ? 1 : 0));
        }
        OV2_LAUNCH(ctx, "ba_eval_kernel<cost>", ba_eval_kernel<false><<<grid1(nobs, 128), 128, 0, st>>>(D, cand_pose, cand_invd));
        first_iter = 0;
        if (sh && sh->fn) {
            // tiny second collective: candidate cost, model cost change, landmark step / candidate norms
            if (sh->fn(sh->user, D.scal + SC_CAND_COST, 4, (void*)st) != 0) return ov2_fail(ctx, OV2_ERR_CUDA, "allreduce callback failed");
        }
        OV2_CUDA(ctx, cudaMemcpyAsync(h, D.scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, st));
        OV2_CUDA(ctx, cudaStreamSynchronize(st));
        if (x_is_new) {
            x_cost = h[SC_COST];          // cost of the (re-)evaluation at x, as Ceres uses it
            cost_known = true;
            if (iteration == 1) out->initial_cost = x_cost;
            if (x_cost < minimum_cost) minimum_cost = x_cost;
            // GradientToleranceReached() for the point this system was assembled at (its gradient
            // F'r / E'r is a by-product of the Schur pass): Ceres would have stopped before this
            // iteration, so the iteration does not count.
            gmax = fmax(h[SC_GMAX_LM], h[SC_GMAX_CAM]);
            if (gmax <= 1e-10) { out->termination = 0; iteration--; break; }
        }
        x_is_new = false;
        const double model_cost_change = h[SC_MCC];
        double cand_cost = h[SC_CAND_COST];
        if (getenv("OV2_BA_DEBUG"))
            fprintf(stderr, "[ba] it %d x_cost %.10g cand %.10g mcc %.10g radius %.4g step %.6g gmax %.4g chol_fail %g ncv %d\n",
                    iteration, x_cost, cand_cost, model_cost_change, radius, sqrt(h[SC_STEP2]), gmax, h[SC_CHOL_FAIL], D.ncv);
        bool step_valid = h[SC_CHOL_FAIL] == 0.0 && isfinite(model_cost_change) && model_cost_change > 0.0;
        if (!step_valid) {
            // HandleInvalidStep
            if (++num_invalid >= 5) { out->termination = 2; break; }
            radius /= decrease_factor;
            decrease_factor *= 2.0;
            step_successful = false;
            continue;
        }
        num_invalid = 0;
        if (!isfinite(cand_cost)) cand_cost = DBL_MAX;
        // ---- ParameterToleranceReached / FunctionToleranceReached (candidate NOT adopted on exit)
        const double step_norm = sqrt(h[SC_STEP2]);
        if (step_norm <= 1e-8 * (xnorm + 1e-8)) { out->termination = 0; break; }
        const double cost_change = x_cost - cand_cost;
        if (fabs(cost_change) <= function_tolerance * x_cost) { out->termination = 0; break; }
        // ---- IsStepSuccessful
        const double rel = cand_cost >= DBL_MAX ? -DBL_MAX : cost_change / model_cost_change;
        if (rel > 1e-3) {
            // HandleSuccessfulStep: x = candidate; residuals + Jacobians are re-evaluated there at the
            // start of the next iteration (Ceres does it now; the values are the same)
            double* t = pose; pose = cand_pose; cand_pose = t;
            t = invd; invd = cand_invd; cand_invd = t;
            xnorm = sqrt(h[SC_CANDX2]);
            OV2_CUDA(ctx, cudaMemcpyAsync(cand_pose, pose, sizeof(double) * 7 * D.ncam, cudaMemcpyDeviceToDevice, st));
            OV2_CUDA(ctx, cudaMemcpyAsync(cand_invd, invd, sizeof(double) * D.npts, cudaMemcpyDeviceToDevice, st));
            x_cost = cand_cost;
            x_is_new = true;
            step_successful = true;
            radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rel - 1.0, 3));
            radius = fmin(1e16, radius);
            decrease_factor = 2.0;
        } else {
            step_successful = false;
            radius /= decrease_factor;
            decrease_factor *= 2.0;
        }
    }
    out->iterations = iteration;
    out->final_cost = minimum_cost < DBL_MAX ? minimum_cost : x_cost;
    return OV2_OK;
}

static ov2_status localba_impl(ov2_ctx* ctx, const ov2_ba_problem* pb, const ov2_ba_opts* opts,
                               ov2_ba_result* res, uint8_t* outlier_out, const Shard* sh) {
    // a shard of a sharded solve may be empty (no landmarks / no observations): it still takes part in every collective
    if (!ctx || !pb || !opts || !res || pb->ncam <= 0 || pb->npts < 0 || (pb->npts == 0 && !sh) || pb->nobs < 0)
        return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_localba_solve: bad arguments");
    memset(res, 0, sizeof(*res));
    if (pb->nobs == 0 && !sh) return OV2_OK;
    ov2_status st = ov2_begin(ctx);
    if (st != OV2_OK) return st;
    const int ncam = pb->ncam, npts = pb->npts, nobs = pb->nobs;
    // host-side: K, CSR pointers (observations must be sorted by landmark)
    double Kh[4];
    std::vector<int32_t> obs_lm_h(nobs), lm_ptr(npts + 1, 0);
    if (ov2_is_device_ptr(pb->K)) OV2_CUDA(ctx, cudaMemcpy(Kh, pb->K, sizeof(Kh), cudaMemcpyDeviceToHost));
    else memcpy(Kh, pb->K, sizeof(Kh));
    if (ov2_is_device_ptr(pb->obs_lm)) OV2_CUDA(ctx, cudaMemcpy(obs_lm_h.data(), pb->obs_lm, sizeof(int32_t) * nobs, cudaMemcpyDeviceToHost));
    else memcpy(obs_lm_h.data(), pb->obs_lm, sizeof(int32_t) * nobs);
    for (int i = 0; i < nobs; ++i) {
        const int l = obs_lm_h[i];
        if (l < 0 || l >= npts || (i > 0 && l < obs_lm_h[i - 1]))
            return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_localba_solve: obs_lm must be sorted ascending and in range");
        lm_ptr[l + 1]++;
    }
    for (int l = 0; l < npts; ++l) lm_ptr[l + 1] += lm_ptr[l];

    BaDev D;
    memset(&D, 0, sizeof(D));
    D.ncam = ncam; D.npts = npts; D.nobs = nobs;
    D.fx = Kh[0]; D.fy = Kh[1]; D.cx = Kh[2]; D.cy = Kh[3];
    const float th_f = (float)opts->huber_th;                 // const float mono_th (optimizer.cpp:47)
    D.huber_a = (double)sqrtf(th_f);                          // HuberLoss(std::sqrt(mono_th))
    D.huber_b = D.huber_a * D.huber_a;
    D.use_huber = opts->use_robust ? 1 : 0;
    D.rfx = D.fx; D.rfy = D.fy; D.rcx = D.cx; D.rcy = D.cy;
    for (int k = 0; k < 9; ++k) D.Rrl[k] = (k % 4 == 0) ? 1.0 : 0.0;
    D.trl[0] = D.trl[1] = D.trl[2] = 0.0;
    if (pb->obs_type) {
        if (!pb->Kr || !pb->Trl) return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_localba_solve: obs_type given without Kr / Trl");
        double kr[4], t7[7];
        if (ov2_is_device_ptr(pb->Kr)) OV2_CUDA(ctx, cudaMemcpy(kr, pb->Kr, sizeof(kr), cudaMemcpyDeviceToHost)); else memcpy(kr, pb->Kr, sizeof(kr));
        if (ov2_is_device_ptr(pb->Trl)) OV2_CUDA(ctx, cudaMemcpy(t7, pb->Trl, sizeof(t7), cudaMemcpyDeviceToHost)); else memcpy(t7, pb->Trl, sizeof(t7));
        D.rfx = kr[0]; D.rfy = kr[1]; D.rcx = kr[2]; D.rcy = kr[3];
        const double qn = sqrt(t7[3] * t7[3] + t7[4] * t7[4] + t7[5] * t7[5] + t7[6] * t7[6]);
        const double x = t7[3] / qn, y = t7[4] / qn, z = t7[5] / qn, w = t7[6] / qn;
        const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                             2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                             2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
        for (int k = 0; k < 9; ++k) D.Rrl[k] = R[k];
        D.trl[0] = t7[0]; D.trl[1] = t7[1]; D.trl[2] = t7[2];
    }
    const void* d = nullptr;
    void* o = nullptr;
    const bool all_host = !ov2_is_device_ptr(pb->pose) && !ov2_is_device_ptr(pb->lm_invdepth) && !ov2_is_device_ptr(pb->pose_const) &&
                          !ov2_is_device_ptr(pb->lm_anchor_cam) && !ov2_is_device_ptr(pb->lm_anchor_px) &&
                          !ov2_is_device_ptr(pb->obs_cam) && !ov2_is_device_ptr(pb->obs_lm) && !ov2_is_device_ptr(pb->obs_px) &&
                          (!outlier_out || !ov2_is_device_ptr(outlier_out)) && (!pb->obs_type || !ov2_is_device_ptr(pb->obs_type));
    // Host problem description: pack everything into one pinned staging block -> ONE H2D copy
    // (a dozen small pageable copies cost more than the solve's kernels at C3 size).
    size_t off_apx = 0, off_opx = 0, off_pose = 0, off_invd = 0, off_lac = 0, off_oc = 0, off_ol = 0, off_lp = 0, off_pc = 0, off_ty = 0, pack_bytes = 0;
    char* dpack = nullptr;
    if (all_host) {
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o_ = off; off += (bytes + 15) & ~(size_t)15; return o_; };
        off_pose = take(sizeof(double) * 7 * ncam); off_invd = take(sizeof(double) * npts);
        off_apx = take(sizeof(double) * 2 * npts); off_opx = take(sizeof(double) * 2 * nobs);
        off_lac = take(sizeof(int32_t) * npts); off_oc = take(sizeof(int32_t) * nobs); off_ol = take(sizeof(int32_t) * nobs);
        off_lp = take(sizeof(int32_t) * (npts + 1)); off_pc = take((size_t)ncam);
        off_ty = take(pb->obs_type ? (size_t)nobs : 0);
        pack_bytes = off;
        if (ctx->ba_ws_cap < pack_bytes) {
            if (ctx->ba_ws) cudaFreeHost(ctx->ba_ws);
            ctx->ba_ws = nullptr;
            ctx->ba_ws_cap = 0;
            OV2_CUDA(ctx, cudaHostAlloc(&ctx->ba_ws, pack_bytes * 2, cudaHostAllocDefault));
            ctx->ba_ws_cap = pack_bytes * 2;
        }
        char* hp = (char*)ctx->ba_ws;
        memcpy(hp + off_pose, pb->pose, sizeof(double) * 7 * ncam);
        memcpy(hp + off_invd, pb->lm_invdepth, sizeof(double) * npts);
        memcpy(hp + off_apx, pb->lm_anchor_px, sizeof(double) * 2 * npts);
        memcpy(hp + off_opx, pb->obs_px, sizeof(double) * 2 * nobs);
        memcpy(hp + off_lac, pb->lm_anchor_cam, sizeof(int32_t) * npts);
        memcpy(hp + off_oc, pb->obs_cam, sizeof(int32_t) * nobs);
        memcpy(hp + off_ol, pb->obs_lm, sizeof(int32_t) * nobs);
        memcpy(hp + off_lp, lm_ptr.data(), sizeof(int32_t) * (npts + 1));
        memcpy(hp + off_pc, pb->pose_const, (size_t)ncam);
        if (pb->obs_type) memcpy(hp + off_ty, pb->obs_type, (size_t)nobs);
        if ((st = ov2_scratch(ctx, pack_bytes, &o)) != OV2_OK) return st;
        dpack = (char*)o;
        OV2_CUDA(ctx, cudaMemcpyAsync(dpack, hp, pack_bytes, cudaMemcpyHostToDevice, ctx->stream));
        D.pose_const = (const uint8_t*)(dpack + off_pc);
        D.lm_anchor_cam = (const int32_t*)(dpack + off_lac);
        D.lm_anchor_px = (const double*)(dpack + off_apx);
        D.obs_cam = (const int32_t*)(dpack + off_oc);
        D.obs_lm = (const int32_t*)(dpack + off_ol);
        D.obs_px = (const double*)(dpack + off_opx);
        D.obs_type = pb->obs_type ? (const uint8_t*)(dpack + off_ty) : nullptr;
    } else {
#define IN(field, bytes) do { if ((st = ov2_stage_in(ctx, pb->field, (bytes), &d)) != OV2_OK) return st; } while (0)
        IN(pose_const, (size_t)ncam); D.pose_const = (const uint8_t*)d;
        IN(lm_anchor_cam, sizeof(int32_t) * (size_t)npts); D.lm_anchor_cam = (const int32_t*)d;
        IN(lm_anchor_px, sizeof(double) * 2 * (size_t)npts); D.lm_anchor_px = (const double*)d;
        IN(obs_cam, sizeof(int32_t) * (size_t)nobs); D.obs_cam = (const int32_t*)d;
        IN(obs_lm, sizeof(int32_t) * (size_t)nobs); D.obs_lm = (const int32_t*)d;
        IN(obs_px, sizeof(double) * 2 * (size_t)nobs); D.obs_px = (const double*)d;
        D.obs_type = nullptr;
        if (pb->obs_type) { IN(obs_type, (size_t)nobs); D.obs_type = (const uint8_t*)d; }
#undef IN
    }
    double *pose = nullptr, *cand_pose = nullptr, *invd = nullptr, *cand_invd = nullptr;
#define SCR(ptr, type, count) do { if ((st = ov2_scratch(ctx, sizeof(type) * (size_t)(count), &o)) != OV2_OK) return st; ptr = (type*)o; } while (0)
    SCR(pose, double, 7 * ncam); SCR(cand_pose, double, 7 * ncam);
    SCR(invd, double, npts); SCR(cand_invd, double, npts);
    int32_t* d_lmptr = nullptr;
    SCR(d_lmptr, int32_t, npts + 1);
    SCR(D.active, uint8_t, nobs); SCR(D.cam_slot, int32_t, ncam); SCR(D.cam_used, uint8_t, ncam);
    SCR(D.Jr, double, 2 * (size_t)nobs); SCR(D.Ja, double, 12 * (size_t)nobs); SCR(D.Jo, double, 12 * (size_t)nobs);
    SCR(D.Jl, double, 2 * (size_t)nobs); SCR(D.chi2, double, nobs); SCR(D.dpos, uint8_t, nobs);
    SCR(D.sc_cam, double, MAX_N); SCR(D.sc_lm, double, npts); SCR(D.z, double, MAX_N);
    // accumulation buffer, zeroed once per LM iteration: [scal | rhs | gcam | cn_cam | S] (re-pointed per solve)
    SCR(D.scal, double, (size_t)SC_COUNT + 3 * (size_t)MAX_N + (size_t)MAX_N * MAX_N);
    SCR(D.ete, double, npts); SCR(D.ge, double, npts); SCR(D.flags, uint8_t, nobs);
#undef SCR
    cudaStream_t s = ctx->stream;
    if (all_host) {
        D.lm_ptr = (const int32_t*)(dpack + off_lp);
        OV2_CUDA(ctx, cudaMemcpyAsync(pose, dpack + off_pose, sizeof(double) * 7 * ncam, cudaMemcpyDeviceToDevice, s));
        OV2_CUDA(ctx, cudaMemcpyAsync(invd, dpack + off_invd, sizeof(double) * npts, cudaMemcpyDeviceToDevice, s));
    } else {
        D.lm_ptr = d_lmptr;
        OV2_CUDA(ctx, cudaMemcpyAsync(d_lmptr, lm_ptr.data(), sizeof(int32_t) * (npts + 1), cudaMemcpyHostToDevice, s));
        const cudaMemcpyKind kp = ov2_is_device_ptr(pb->pose) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        const cudaMemcpyKind ki = ov2_is_device_ptr(pb->lm_invdepth) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        OV2_CUDA(ctx, cudaMemcpyAsync(pose, pb->pose, sizeof(double) * 7 * ncam, kp, s));
        OV2_CUDA(ctx, cudaMemcpyAsync(invd, pb->lm_invdepth, sizeof(double) * npts, ki, s));
    }
    OV2_CUDA(ctx, cudaMemsetAsync(D.active, 1, nobs, s));
    OV2_CUDA(ctx, cudaMemsetAsync(D.flags, 0, nobs, s));
    OV2_CUDA(ctx, cudaMemsetAsync(D.sc_lm, 0, sizeof(double) * npts, s));
    OV2_CUDA(ctx, cudaMemsetAsync(D.sc_cam, 0, sizeof(double) * MAX_N, s));

    Summary s1, s2;
    if ((st = ba_ceres_solve(ctx, D, pose, cand_pose, invd, cand_invd, opts->max_iters_robust, opts->function_tolerance, &s1, sh)) != OV2_OK)
        return st;
    // outlier scan on the values the LAST Evaluate() left behind (optimizer.cpp:500-530)
    if (!ctx->ba_hscal) OV2_CUDA(ctx, cudaHostAlloc((void**)&ctx->ba_hscal, sizeof(double) * 2 * SC_COUNT, cudaHostAllocDefault));
    double* h = ctx->ba_hscal + SC_COUNT;   // second half: the LM controller uses the first
    OV2_CUDA(ctx, cudaMemsetAsync(D.scal + SC_NBAD, 0, 3 * sizeof(double), s));
    const int deact = opts->apply_l2_after_robust ? 1 : 0;
    OV2_LAUNCH(ctx, "ba_flag_kernel", ba_flag_kernel<<<grid1(nobs, 256), 256, 0, s>>>(D, (double)th_f, 1, deact));
    OV2_CUDA(ctx, cudaMemcpyAsync(h, D.scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, s));
    OV2_CUDA(ctx, cudaStreamSynchronize(s));
    res->iters_robust = s1.iterations;
    res->initial_cost = s1.initial_cost;
    res->final_cost = s1.final_cost;
    res->termination = s1.termination;
    res->n_outliers_first = (int)h[SC_NBAD];
    double nbad_global = h[SC_NBAD];
    if (sh && sh->fn) {
        // every rank must take the same branch: the refinement runs if ANY rank removed an observation
        if (sh->fn(sh->user, D.scal + SC_NBAD, 1, (void*)s) != 0) return ov2_fail(ctx, OV2_ERR_CUDA, "allreduce callback failed");
        OV2_CUDA(ctx, cudaMemcpyAsync(&nbad_global, D.scal + SC_NBAD, sizeof(double), cudaMemcpyDeviceToHost, s));
        OV2_CUDA(ctx, cudaStreamSynchronize(s));
    }
    if (opts->apply_l2_after_robust && opts->use_robust && nbad_global > 0) {
        // mono windows keep the Huber loss in the refinement; the wrapper is reset to the trivial loss only
        // when left-camera and other-frame right-camera residual lists are both non-empty (optimizer.cpp:606-608)
        // (sharded windows: the counts are per shard, so pass refine_loss explicitly there)
        int trivial = opts->refine_loss;
        if (trivial < 0) trivial = (h[SC_NLEFT] > 0.0 && h[SC_NRIGHT] > 0.0) ? 1 : 0;
        D.use_huber = trivial ? 0 : 1;
        if ((st = ba_ceres_solve(ctx, D, pose, cand_pose, invd, cand_invd, opts->max_iters_refine, opts->function_tolerance, &s2, sh)) != OV2_OK)
            return st;
        OV2_CUDA(ctx, cudaMemsetAsync(D.scal + SC_NBAD, 0, 3 * sizeof(double), s));
        OV2_LAUNCH(ctx, "ba_flag_kernel", ba_flag_kernel<<<grid1(nobs, 256), 256, 0, s>>>(D, (double)th_f, 2, 0));
        OV2_CUDA(ctx, cudaMemcpyAsync(h, D.scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, s));
        OV2_CUDA(ctx, cudaStreamSynchronize(s));
        res->iters_refine = s2.iterations;
        res->initial_cost = s2.initial_cost;
        res->final_cost = s2.final_cost;
        res->termination = s2.termination;
        res->n_outliers_second = (int)h[SC_NBAD];
    }
    // write-back of the states (the map update itself, optimizer.cpp:741-897, stays on the host)
    if (all_host) {
        // one D2H: [pose | invd | flags] gathered on the device into the pack block's head
        char* hp = (char*)ctx->ba_ws;
        const size_t o_pose = 0, o_invd = sizeof(double) * 7 * ncam, o_fl = o_invd + sizeof(double) * npts;
        const size_t out_bytes = o_fl + (size_t)nobs;
        OV2_CUDA(ctx, cudaMemcpyAsync(dpack + o_pose, pose, sizeof(double) * 7 * ncam, cudaMemcpyDeviceToDevice, s));
        OV2_CUDA(ctx, cudaMemcpyAsync(dpack + o_invd, invd, sizeof(double) * npts, cudaMemcpyDeviceToDevice, s));
        OV2_CUDA(ctx, cudaMemcpyAsync(dpack + o_fl, D.flags, nobs, cudaMemcpyDeviceToDevice, s));
        OV2_CUDA(ctx, cudaMemcpyAsync(hp, dpack, out_bytes, cudaMemcpyDeviceToHost, s));
        OV2_CUDA(ctx, cudaStreamSynchronize(s));
        memcpy(pb->pose, hp + o_pose, sizeof(double) * 7 * ncam);
        memcpy(pb->lm_invdepth, hp + o_invd, sizeof(double) * npts);
        if (outlier_out) memcpy(outlier_out, hp + o_fl, nobs);
    } else {
        OV2_CUDA(ctx, cudaMemcpyAsync(pb->pose, pose, sizeof(double) * 7 * ncam,
                                      ov2_is_device_ptr(pb->pose) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
        OV2_CUDA(ctx, cudaMemcpyAsync(pb->lm_invdepth, invd, sizeof(double) * npts,
                                      ov2_is_device_ptr(pb->lm_invdepth) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
        if (outlier_out)
            OV2_CUDA(ctx, cudaMemcpyAsync(outlier_out, D.flags, nobs,
                                          ov2_is_device_ptr(outlier_out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
        OV2_CUDA(ctx, cudaStreamSynchronize(s));
    }
    if (res->termination == 2) return ov2_fail(ctx, OV2_ERR_NUMERIC, "ov2_localba_solve: 5 consecutive invalid steps");
    return OV2_OK;
}

// Round-1 path (one launch per phase, LM controller on the host): kept selectable with OV2_BA_LEGACY=1 as a cross-check
// of the persistent kernel (ba_lm.cu), and as the carrier of the callback-based sharded solve below.
ov2_status ov2_localba_solve_legacy(ov2_ctx* ctx, const ov2_ba_problem* pb, const ov2_ba_opts* opts,
                                    ov2_ba_result* res, uint8_t* outlier_out) {
    return localba_impl(ctx, pb, opts, res, outlier_out, nullptr);
}

extern "C" ov2_status ov2_localba_solve_sharded(ov2_ctx* ctx, const ov2_ba_problem* pb, const ov2_ba_opts* opts,
                                                ov2_ba_result* res, uint8_t* outlier_out, ov2_allreduce_fn allreduce,
                                                void* user, int rank) {
    if (!allreduce) return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_localba_solve_sharded: allreduce callback is NULL");
    Shard sh{allreduce, user, rank};
    return localba_impl(ctx, pb, opts, res, outlier_out, &sh);
}

/* CPU ORACLE (C) for Optimizer::localBA's solve path.  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Same algorithm as oracle/ba_ref.py (see its header for the reference file:line map), written in
 * plain single-threaded C so the CPU baseline in bench.py is not a numpy artefact: the reference
 * runs Ceres with options.num_threads = 1 (/root/reference/src/optimizer.cpp:460).  It is a
 * RESTATEMENT, NOT CERES (Ceres / Eigen cannot be built in this container); tests pin it against
 * oracle/ba_ref.py to 1e-9.  Never linked into or called from the product path.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int ncam, npts, nobs;
    const double* K;
    double* pose;
    const uint8_t* pose_const;
    const int32_t* lm_anchor_cam;
    const double* lm_anchor_px;
    double* lm_invdepth;
    const int32_t* obs_cam;
    const int32_t* obs_lm;
    const double* obs_px;
} BaPb;

typedef struct { int iterations; double initial_cost, final_cost; int termination; } Summ;

#define SOPHUS_EPS 1e-10

static void quat_to_rot(const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

static void load_pose(const double* p, double* t, double* q) {
    t[0] = p[0]; t[1] = p[1]; t[2] = p[2];
    const double n = sqrt(p[3] * p[3] + p[4] * p[4] + p[5] * p[5] + p[6] * p[6]);
    q[0] = p[3] / n; q[1] = p[4] / n; q[2] = p[5] / n; q[3] = p[6] / n;
}

static void pose_plus(const double* pose, const double* d, double* out) {
    double t[3], q[4];
    load_pose(pose, t, q);
    const double ox = d[3], oy = d[4], oz = d[5], th2 = ox * ox + oy * oy + oz * oz;
    double imag, real, theta;
    if (th2 < SOPHUS_EPS * SOPHUS_EPS) {
        theta = 0;
        imag = 0.5 - th2 / 48.0 + th2 * th2 / 3840.0;
        real = 1.0 - th2 / 8.0 + th2 * th2 / 384.0;
    } else {
        theta = sqrt(th2);
        imag = sin(0.5 * theta) / theta;
        real = cos(0.5 * theta);
    }
    const double e[4] = {imag * ox, imag * oy, imag * oz, real};
    double Re[9], V[9];
    quat_to_rot(e, Re);
    if (theta < SOPHUS_EPS) memcpy(V, Re, sizeof(V));
    else {
        const double a = (1 - cos(theta)) / th2, b = (theta - sin(theta)) / (th2 * theta);
        const double O[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double o2 = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
                V[3 * i + j] = a * O[3 * i + j] + b * o2 + (i == j ? 1.0 : 0.0);
            }
    }
    double r[4];
    r[3] = e[3] * q[3] - e[0] * q[0] - e[1] * q[1] - e[2] * q[2];
    r[0] = e[3] * q[0] + e[0] * q[3] + e[1] * q[2] - e[2] * q[1];
    r[1] = e[3] * q[1] + e[1] * q[3] + e[2] * q[0] - e[0] * q[2];
    r[2] = e[3] * q[2] + e[2] * q[3] + e[0] * q[1] - e[1] * q[0];
    const double n = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
    for (int i = 0; i < 3; ++i)
        out[i] = V[3 * i] * d[0] + V[3 * i + 1] * d[1] + V[3 * i + 2] * d[2] + Re[3 * i] * t[0] + Re[3 * i + 1] * t[1] + Re[3 * i + 2] * t[2];
    for (int i = 0; i < 4; ++i) out[3 + i] = r[i] / n;
}

/* residual block i; if J != NULL writes robustified rows: r(2) Ja(12) Jo(12) Jl(2) */
static double eval_obs(const BaPb* pb, const double* pose, const double* invd, int i, int use_huber, double ha, double hb,
                       double* chi2, uint8_t* dpos, double* Jr, double* Ja, double* Jo, double* Jl) {
    const double fx = pb->K[0], fy = pb->K[1], cx = pb->K[2], cy = pb->K[3];
    const int lm = pb->obs_lm[i], ca = pb->lm_anchor_cam[lm], co = pb->obs_cam[i];
    double ta[3], qa[4], to[3], qo[4], Rwa[9], Rwc[9];
    load_pose(pose + 7 * ca, ta, qa);
    load_pose(pose + 7 * co, to, qo);
    quat_to_rot(qa, Rwa);
    quat_to_rot(qo, Rwc);
    const double zanch = 1.0 / invd[lm];
    const double ap[3] = {zanch * (pb->lm_anchor_px[2 * lm] - cx) / fx, zanch * (pb->lm_anchor_px[2 * lm + 1] - cy) / fy, zanch};
    double rp[3], wp[3], dv[3], lc[3];
    for (int k = 0; k < 3; ++k) {
        rp[k] = Rwa[3 * k] * ap[0] + Rwa[3 * k + 1] * ap[1] + Rwa[3 * k + 2] * ap[2];
        wp[k] = rp[k] + ta[k];
        dv[k] = wp[k] - to[k];
    }
    for (int k = 0; k < 3; ++k) lc[k] = Rwc[k] * dv[0] + Rwc[3 + k] * dv[1] + Rwc[6 + k] * dv[2];
    const double linvz = 1.0 / lc[2];
    const double r0 = fx * lc[0] * linvz + cx - pb->obs_px[2 * i], r1 = fy * lc[1] * linvz + cy - pb->obs_px[2 * i + 1];
    const double s = r0 * r0 + r1 * r1;
    *chi2 = s;
    *dpos = lc[2] > 0.0;
    double w = 1.0, cost = 0.5 * s;
    if (use_huber && s > hb) {
        const double rs = sqrt(s);
        cost = 0.5 * (2 * ha * rs - hb);
        w = sqrt(fmax(DBL_MIN, ha / rs));
    }
    if (Jr) {
        const double l2 = linvz * linvz;
        const double jc[6] = {linvz * fx, 0, -lc[0] * l2 * fx, 0, linvz * fy, -lc[1] * l2 * fy};
        double JR[6], JS[6];
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 3; ++b) JR[3 * a + b] = jc[3 * a] * Rwc[3 * b] + jc[3 * a + 1] * Rwc[3 * b + 1] + jc[3 * a + 2] * Rwc[3 * b + 2];
        for (int a = 0; a < 2; ++a) {
            const double j0 = JR[3 * a], j1 = JR[3 * a + 1], j2 = JR[3 * a + 2];
            JS[3 * a] = j1 * wp[2] - j2 * wp[1];
            JS[3 * a + 1] = j2 * wp[0] - j0 * wp[2];
            JS[3 * a + 2] = j0 * wp[1] - j1 * wp[0];
        }
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 3; ++b) {
                Ja[6 * a + b] = w * JR[3 * a + b]; Ja[6 * a + 3 + b] = -w * JS[3 * a + b];
                Jo[6 * a + b] = -w * JR[3 * a + b]; Jo[6 * a + 3 + b] = w * JS[3 * a + b];
            }
        Jl[0] = w * -zanch * (JR[0] * rp[0] + JR[1] * rp[1] + JR[2] * rp[2]);
        Jl[1] = w * -zanch * (JR[3] * rp[0] + JR[4] * rp[1] + JR[5] * rp[2]);
        Jr[0] = w * r0; Jr[1] = w * r1;
    }
    return cost;
}

static int cholesky_solve(double* A, double* b, int n) { /* in place, upper U'U, then solves */
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        if (!(d > 0.0) || !isfinite(d)) return 0;
        d = sqrt(d);
        A[(size_t)j * n + j] = d;
        for (int c = j + 1; c < n; ++c) A[(size_t)j * n + c] /= d;
        for (int r = j + 1; r < n; ++r) {
            const double u = A[(size_t)j * n + r];
            if (u == 0.0) continue;
            for (int c = r; c < n; ++c) A[(size_t)r * n + c] -= u * A[(size_t)j * n + c];
        }
    }
    for (int j = 0; j < n; ++j) {
        b[j] /= A[(size_t)j * n + j];
        for (int c = j + 1; c < n; ++c) b[c] -= A[(size_t)j * n + c] * b[j];
    }
    for (int j = n - 1; j >= 0; --j) {
        b[j] /= A[(size_t)j * n + j];
        for (int r = 0; r < j; ++r) b[r] -= A[(size_t)r * n + j] * b[j];
    }
    return 1;
}

static void ceres_solve(const BaPb* pb, double* pose, double* invd, const uint8_t* active, const int32_t* lm_ptr, int max_iters,
                        int use_huber, double ha, double ftol, double* chi2, uint8_t* dpos, Summ* out) {
    const int ncam = pb->ncam, npts = pb->npts, nobs = pb->nobs;
    const double hb = ha * ha;
    int* slot = (int*)malloc(sizeof(int) * ncam);
    uint8_t* used = (uint8_t*)calloc(ncam, 1);
    int any = 0;
    for (int i = 0; i < nobs; ++i)
        if (active[i]) { used[pb->obs_cam[i]] = 1; used[pb->lm_anchor_cam[pb->obs_lm[i]]] = 1; any = 1; }
    int ncv = 0;
    for (int c = 0; c < ncam; ++c) slot[c] = (used[c] && !pb->pose_const[c]) ? ncv++ : -1;
    memset(out, 0, sizeof(*out));
    if (!any) { free(slot); free(used); return; }
    const int n = 6 * ncv;
    double* Jr = (double*)malloc(sizeof(double) * 2 * nobs), *Ja = (double*)malloc(sizeof(double) * 12 * nobs);
    double* Jo = (double*)malloc(sizeof(double) * 12 * nobs), *Jl = (double*)malloc(sizeof(double) * 2 * nobs);
    double* S = (double*)malloc(sizeof(double) * (size_t)(n ? n : 1) * (n ? n : 1)), *rhs = (double*)malloc(sizeof(double) * (n + 1));
    double* gcam = (double*)malloc(sizeof(double) * (n + 1)), *cn = (double*)malloc(sizeof(double) * (n + 1)), *sc_cam = (double*)malloc(sizeof(double) * (n + 1));
    double* sc_lm = (double*)calloc(npts, sizeof(double)), *ete = (double*)malloc(sizeof(double) * npts), *ge = (double*)malloc(sizeof(double) * npts);
    double* cpose = (double*)malloc(sizeof(double) * 7 * ncam), *cinvd = (double*)malloc(sizeof(double) * npts);
    double* etf = (double*)malloc(sizeof(double) * 6 * (ncam + 1));
    int* eslot = (int*)malloc(sizeof(int) * (ncam + 1));
    memcpy(cpose, pose, sizeof(double) * 7 * ncam);
    memcpy(cinvd, invd, sizeof(double) * npts);
    double x_cost = 0;
    for (int i = 0; i < nobs; ++i)
        if (active[i]) x_cost += eval_obs(pb, pose, invd, i, use_huber, ha, hb, chi2 + i, dpos + i, Jr + 2 * i, Ja + 12 * i, Jo + 12 * i, Jl + 2 * i);
    out->initial_cost = x_cost;
    double minimum = DBL_MAX, xnorm = -1, radius = 1e4, dec = 2.0;
    int ok = 1, iteration = 0, ninvalid = 0, first = 1;
    for (;;) {
        if (ok && x_cost < minimum) minimum = x_cost;
        if (iteration >= max_iters) { out->termination = 1; break; }
        if (radius <= 1e-32) break;
        iteration++;
        memset(S, 0, sizeof(double) * (size_t)n * n);
        memset(rhs, 0, sizeof(double) * n); memset(gcam, 0, sizeof(double) * n); memset(cn, 0, sizeof(double) * n);
        double gmax = 0;
        /* pass 1: camera column norms (needed for nothing before the reduced solve), per-landmark blocks */
        for (int l = 0; l < npts; ++l) {
            double cnl = 0, gel = 0; int nact = 0;
            for (int p = lm_ptr[l]; p < lm_ptr[l + 1]; ++p)
                if (active[p]) { cnl += Jl[2 * p] * Jl[2 * p] + Jl[2 * p + 1] * Jl[2 * p + 1]; gel += Jl[2 * p] * Jr[2 * p] + Jl[2 * p + 1] * Jr[2 * p + 1]; nact++; }
            if (!nact) { ete[l] = 0; ge[l] = 0; continue; }
            if (first) sc_lm[l] = 1.0 / (1.0 + sqrt(cnl));
            const double sc = sc_lm[l], diag = fmin(fmax(cnl * sc * sc, 1e-6), 1e32);
            ete[l] = cnl + diag / (radius * sc * sc);
            ge[l] = gel;
            if (fabs(gel) > gmax) gmax = fabs(gel);
            const double inv = 1.0 / ete[l];
            const int sa = slot[pb->lm_anchor_cam[l]];
            int m = 0;
            if (sa >= 0) { eslot[0] = sa; memset(etf, 0, sizeof(double) * 6); m = 1; }
            for (int p = lm_ptr[l]; p < lm_ptr[l + 1]; ++p) {
                if (!active[p]) continue;
                const int so = slot[pb->obs_cam[p]];
                const double* A = Ja + 12 * p, *O = Jo + 12 * p;
                const double j0 = Jl[2 * p], j1 = Jl[2 * p + 1], r0 = Jr[2 * p], r1 = Jr[2 * p + 1];
                if (sa >= 0) {
                    for (int a = 0; a < 6; ++a) {
                        for (int b = a; b < 6; ++b) S[(size_t)(6 * sa + a) * n + 6 * sa + b] += A[a] * A[b] + A[6 + a] * A[6 + b];
                        gcam[6 * sa + a] += A[a] * r0 + A[6 + a] * r1;
                        cn[6 * sa + a] += A[a] * A[a] + A[6 + a] * A[6 + a];
                        etf[a] += j0 * A[a] + j1 * A[6 + a];
                    }
                }
                if (so >= 0) {
                    for (int a = 0; a < 6; ++a) {
                        for (int b = a; b < 6; ++b) S[(size_t)(6 * so + a) * n + 6 * so + b] += O[a] * O[b] + O[6 + a] * O[6 + b];
                        gcam[6 * so + a] += O[a] * r0 + O[6 + a] * r1;
                        cn[6 * so + a] += O[a] * O[a] + O[6 + a] * O[6 + a];
                        etf[6 * m + a] = j0 * O[a] + j1 * O[6 + a];
                    }
                    eslot[m] = so;
                    if (sa >= 0)
                        for (int a = 0; a < 6; ++a)
                            for (int b = 0; b < 6; ++b) {
                                const double v = A[a] * O[b] + A[6 + a] * O[6 + b];
                                if (sa < so) S[(size_t)(6 * sa + a) * n + 6 * so + b] += v; else S[(size_t)(6 * so + b) * n + 6 * sa + a] += v;
                            }
                    m++;
                }
            }
            for (int i = 0; i < m; ++i) {
                for (int j = 0; j < m; ++j) {
                    const int si = eslot[i], sj = eslot[j];
                    if (si > sj) continue;
                    for (int a = 0; a < 6; ++a)
                        for (int b = (si == sj ? a : 0); b < 6; ++b) S[(size_t)(6 * si + a) * n + 6 * sj + b] -= etf[6 * i + a] * etf[6 * j + b] * inv;
                }
                for (int a = 0; a < 6; ++a) rhs[6 * eslot[i] + a] -= etf[6 * i + a] * gel * inv;
            }
        }
        for (int i = 0; i < n; ++i) {
            if (first) sc_cam[i] = 1.0 / (1.0 + sqrt(cn[i]));
            const double diag = fmin(fmax(cn[i] * sc_cam[i] * sc_cam[i], 1e-6), 1e32);
            S[(size_t)i * n + i] += diag / (radius * sc_cam[i] * sc_cam[i]);
            rhs[i] += gcam[i];
        }
        first = 0;
        if (iteration == 1 || ok) {   /* GradientToleranceReached for the point the system was built at */
            for (int c = 0; c < ncam; ++c) {
                if (slot[c] < 0) continue;
                double g[6], pg[7];
                for (int k = 0; k < 6; ++k) g[k] = -gcam[6 * slot[c] + k];
                pose_plus(pose + 7 * c, g, pg);
                for (int k = 0; k < 7; ++k) if (fabs(pose[7 * c + k] - pg[k]) > gmax) gmax = fabs(pose[7 * c + k] - pg[k]);
            }
            if (gmax <= 1e-10) { iteration--; break; }
        }
        int valid = n ? cholesky_solve(S, rhs, n) : 1;   /* rhs now holds z */
        double mcc = 0, step2 = 0, candx2 = 0;
        if (valid) {
            for (int c = 0; c < ncam; ++c) {
                if (slot[c] < 0) continue;
                double d[6];
                for (int k = 0; k < 6; ++k) d[k] = -rhs[6 * slot[c] + k];
                pose_plus(pose + 7 * c, d, cpose + 7 * c);
                for (int k = 0; k < 7; ++k) { const double df = pose[7 * c + k] - cpose[7 * c + k]; step2 += df * df; candx2 += cpose[7 * c + k] * cpose[7 * c + k]; }
            }
            for (int l = 0; l < npts; ++l) {
                if (ete[l] == 0.0) continue;
                const int sa = slot[pb->lm_anchor_cam[l]];
                double acc = 0;
                for (int p = lm_ptr[l]; p < lm_ptr[l + 1]; ++p) {
                    if (!active[p]) continue;
                    const int so = slot[pb->obs_cam[p]];
                    double f0 = 0, f1 = 0;
                    if (sa >= 0) for (int k = 0; k < 6; ++k) { f0 += Ja[12 * p + k] * rhs[6 * sa + k]; f1 += Ja[12 * p + 6 + k] * rhs[6 * sa + k]; }
                    if (so >= 0) for (int k = 0; k < 6; ++k) { f0 += Jo[12 * p + k] * rhs[6 * so + k]; f1 += Jo[12 * p + 6 + k] * rhs[6 * so + k]; }
                    acc += Jl[2 * p] * f0 + Jl[2 * p + 1] * f1;
                }
                const double dl = -(ge[l] - acc) / ete[l];
                cinvd[l] = invd[l] + dl;
                step2 += dl * dl; candx2 += cinvd[l] * cinvd[l];
                for (int p = lm_ptr[l]; p < lm_ptr[l + 1]; ++p) {
                    if (!active[p]) continue;
                    const int so = slot[pb->obs_cam[p]];
                    double f0 = Jl[2 * p] * dl, f1 = Jl[2 * p + 1] * dl;
                    if (sa >= 0) for (int k = 0; k < 6; ++k) { f0 -= Ja[12 * p + k] * rhs[6 * sa + k]; f1 -= Ja[12 * p + 6 + k] * rhs[6 * sa + k]; }
                    if (so >= 0) for (int k = 0; k < 6; ++k) { f0 -= Jo[12 * p + k] * rhs[6 * so + k]; f1 -= Jo[12 * p + 6 + k] * rhs[6 * so + k]; }
                    mcc -= f0 * (Jr[2 * p] + 0.5 * f0) + f1 * (Jr[2 * p + 1] + 0.5 * f1);
                }
            }
            valid = isfinite(mcc) && mcc > 0.0;
        }
        if (!valid) {
            if (++ninvalid >= 5) { out->termination = 2; break; }
            radius /= dec; dec *= 2; ok = 0;
            continue;
        }
        ninvalid = 0;
        double cand = 0;
        for (int i = 0; i < nobs; ++i)
            if (active[i]) cand += eval_obs(pb, cpose, cinvd, i, use_huber, ha, hb, chi2 + i, dpos + i, NULL, NULL, NULL, NULL);
        if (!isfinite(cand)) cand = DBL_MAX;
        if (sqrt(step2) <= 1e-8 * (xnorm + 1e-8)) break;
        const double change = x_cost - cand;
        if (fabs(change) <= ftol * x_cost) break;
        const double rel = cand >= DBL_MAX ? -DBL_MAX : change / mcc;
        if (rel > 1e-3) {
            memcpy(pose, cpose, sizeof(double) * 7 * ncam);
            memcpy(invd, cinvd, sizeof(double) * npts);
            xnorm = sqrt(candx2);
            x_cost = 0;
            for (int i = 0; i < nobs; ++i)
                if (active[i]) x_cost += eval_obs(pb, pose, invd, i, use_huber, ha, hb, chi2 + i, dpos + i, Jr + 2 * i, Ja + 12 * i, Jo + 12 * i, Jl + 2 * i);
            ok = 1;
            radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rel - 1.0, 3)));
            dec = 2.0;
        } else {
            memcpy(cpose, pose, sizeof(double) * 7 * ncam);
            memcpy(cinvd, invd, sizeof(double) * npts);
            ok = 0; radius /= dec; dec *= 2;
        }
    }
    out->iterations = iteration;
    out->final_cost = minimum < DBL_MAX ? minimum : x_cost;
    free(slot); free(used); free(Jr); free(Ja); free(Jo); free(Jl); free(S); free(rhs); free(gcam); free(cn); free(sc_cam);
    free(sc_lm); free(ete); free(ge); free(cpose); free(cinvd); free(etf); free(eslot);
}

/* result[8]: iters_robust, iters_refine, initial_cost, final_cost, n_out1, n_out2, termination */
int ov2_oracle_local_ba(const BaPb* pb, int max_it1, int max_it2, double huber_th, double ftol, int use_robust, int l2_after,
                        double* result, uint8_t* flags) {
    const int nobs = pb->nobs, npts = pb->npts;
    const float thf = (float)huber_th;
    const double ha = (double)sqrtf(thf), th = (double)thf;
    int32_t* lm_ptr = (int32_t*)calloc(npts + 1, sizeof(int32_t));
    for (int i = 0; i < nobs; ++i) lm_ptr[pb->obs_lm[i] + 1]++;
    for (int l = 0; l < npts; ++l) lm_ptr[l + 1] += lm_ptr[l];
    uint8_t* active = (uint8_t*)malloc(nobs), *dpos = (uint8_t*)malloc(nobs);
    double* chi2 = (double*)malloc(sizeof(double) * nobs);
    memset(active, 1, nobs);
    memset(flags, 0, nobs);
    Summ s1, s2;
    ceres_solve(pb, pb->pose, pb->lm_invdepth, active, lm_ptr, max_it1, use_robust, ha, ftol, chi2, dpos, &s1);
    int nb1 = 0, nb2 = 0;
    for (int i = 0; i < nobs; ++i)
        if (active[i] && (chi2[i] > th || !dpos[i])) { flags[i] |= 1; nb1++; if (l2_after) active[i] = 0; }
    result[0] = s1.iterations; result[1] = 0; result[2] = s1.initial_cost; result[3] = s1.final_cost; result[4] = nb1; result[5] = 0; result[6] = s1.termination;
    if (l2_after && use_robust && nb1 > 0) {
        ceres_solve(pb, pb->pose, pb->lm_invdepth, active, lm_ptr, max_it2, 1, ha, ftol, chi2, dpos, &s2);
        for (int i = 0; i < nobs; ++i)
            if (active[i] && (chi2[i] > th || !dpos[i])) { flags[i] |= 2; nb2++; }
        result[1] = s2.iterations; result[2] = s2.initial_cost; result[3] = s2.final_cost; result[5] = nb2; result[6] = s2.termination;
    }
    free(lm_ptr); free(active); free(dpos); free(chi2);
    return 0;
}

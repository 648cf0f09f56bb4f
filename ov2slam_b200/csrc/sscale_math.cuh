// D (per-cell response of FeatureExtractor::detectSingleScale, /root/reference/src/feature_extractor.cpp:347-351):
// GaussianBlur 3x3 of the cell sub-matrix + cornerMinEigenVal(3, 3), as phases that a thread block runs
// with (tid, nthreads) - and that tests/test_host_logic.py compiles for the HOST and checks bit for bit
// against oracle/image_ref.py (min_eigen_ref(blur3_cell_ref(...))).  Every float operation is an explicit
// round-to-nearest primitive (device: __f*_rn intrinsics, host: plain ops compiled with
// -ffp-contract=off and fmaf), in OpenCV's order (see the header of frontend_sscale.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define SS_HD __host__ __device__ __forceinline__
#define SS_MUL(a, b) __fmul_rn((a), (b))
#define SS_ADD(a, b) __fadd_rn((a), (b))
#define SS_SUB(a, b) __fsub_rn((a), (b))
#define SS_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define SS_SQRT(a) __fsqrt_rn(a)
#define SS_D2F(a) __double2float_rn(a)
#else
#include <math.h>
#if defined(__CUDACC__)
#define SS_HD __host__ __device__ inline
#else
#define SS_HD inline
#endif
#define SS_MUL(a, b) ((a) * (b))
#define SS_ADD(a, b) ((a) + (b))
#define SS_SUB(a, b) ((a) - (b))
#define SS_FMA(a, b, c) fmaf((a), (b), (c))
#define SS_SQRT(a) sqrtf(a)
#define SS_D2F(a) ((float)(a))
#endif

namespace sscale {

SS_HD int refl(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }   // REFLECT_101, overshoot <= 1

// shared-memory carve-up (bytes) for a cs x cs cell: raw tile, blurred cell, Sobel products (double, 3 per pixel interleaved)
SS_HD size_t raw_bytes(int cs) { return (size_t)(((cs + 2) * (cs + 2) + 15) & ~15); }
SS_HD size_t blur_bytes(int cs) { return (size_t)((cs * cs + 15) & ~15); }
SS_HD size_t smem_bytes(int cs) { return raw_bytes(cs) + blur_bytes(cs) + 3 * sizeof(double) * (size_t)cs * cs; }

// Work decomposition of all three phases: one item = one column x and a run of SEG rows.  Whatever a row contributes to
// the three output rows that read it (horizontal 3-tap sums, the Sobel row quantities, the horizontal box sums) is formed
// ONCE per item and slid down the column in registers: 7 row visits per 5 outputs instead of 15, and no per-pixel
// index division (the first version spent 440 thread-instructions per pixel, mostly integer address arithmetic).
constexpr int SEG = 5;    // 35 x 7 = 245 items for the 256 threads of a 35-px cell

// phase 1: raw = (cs+2)^2 tile (cell + 1-px halo from the parent image) -> bl = blurred cell.
// S / 16 rounded half-to-even in the first 16*floor(cs/16) columns, half-up in the tail.
SS_HD void phase_blur(int tid, int nt, const uint8_t* raw, uint8_t* bl, int cs) {
    const int rw = cs + 2, nvec16 = (cs >> 4) << 4;
    const int nseg = (cs + SEG - 1) / SEG;
    for (int item = tid; item < cs * nseg; item += nt) {
        const int seg = item / cs, xx = item - seg * cs;
        const int y0 = seg * SEG;
        int y1 = y0 + SEG;
        if (y1 > cs) y1 = cs;
        const uint8_t* p = raw + y0 * rw + xx;                   // raw row y0 = the row above output row y0
        int h0 = (int)p[0] + 2 * (int)p[1] + (int)p[2];
        p += rw;
        int h1 = (int)p[0] + 2 * (int)p[1] + (int)p[2];
        const bool even = xx < nvec16;
        for (int yy = y0; yy < y1; ++yy) {
            p += rw;
            const int h2 = (int)p[0] + 2 * (int)p[1] + (int)p[2];
            const int S = h0 + 2 * h1 + h2;
            int v;
            if (even) {
                v = S >> 4;
                const int rem = S & 15;
                if (rem > 8 || (rem == 8 && (v & 1))) v++;
            } else {
                v = (S + 8) >> 4;
            }
            bl[yy * cs + xx] = (uint8_t)v;
            h0 = h1; h1 = h2;
        }
    }
}

// Sobel row quantities of blurred row `row` at column x (xm / xp = reflected neighbours): d = right - left (exact), q = the
// [k1 k2 k1] row filter - vector form (FMAs) in the first 32*floor(cs/32) columns, scalar form in the tail.
SS_HD void cov_row(const uint8_t* row, int xm, int xx, int xp, bool vec, float& d, float& q) {
    const float k1 = 1.0f / 3060.0f, k2 = 2.0f / 3060.0f;       // float32(s), float32(2 s) (= 2 k1 exactly)
    const int a = row[xm], c = row[xp];
    const float A = (float)a, B = (float)row[xx], C = (float)c;
    d = (float)(c - a);
    q = vec ? SS_FMA(k1, C, SS_FMA(k2, B, SS_MUL(k1, A))) : SS_ADD(SS_ADD(SS_MUL(k1, A), SS_MUL(k2, B)), SS_MUL(k1, C));
}

// phase 2: Sobel derivatives (scale 1/(4*3*255)) and their products, stored as double [pixel][xx, xy, yy]
// (a float32 -> double conversion is exact; phase 3 then adds without converting nine values per pixel)
SS_HD void phase_cov(int tid, int nt, const uint8_t* bl, double* cov, int cs) {
    const float k1 = 1.0f / 3060.0f, k2 = 2.0f / 3060.0f;
    const int nvec32 = (cs >> 5) << 5;
    const int nseg = (cs + SEG - 1) / SEG;
    for (int item = tid; item < cs * nseg; item += nt) {
        const int seg = item / cs, xx = item - seg * cs;
        const int y0 = seg * SEG;
        int y1 = y0 + SEG;
        if (y1 > cs) y1 = cs;
        const int xm = refl(xx - 1, cs), xp = refl(xx + 1, cs);
        const bool vec = xx < nvec32;
        float dm, qm, d0, q0, dp, qp;
        cov_row(bl + refl(y0 - 1, cs) * cs, xm, xx, xp, vec, dm, qm);
        cov_row(bl + y0 * cs, xm, xx, xp, vec, d0, q0);
        for (int yy = y0; yy < y1; ++yy) {
            cov_row(bl + refl(yy + 1, cs) * cs, xm, xx, xp, vec, dp, qp);
            // Dx: row [-1 0 1] (exact), column [k1 k2 k1] evaluated as fma(top + bottom, k1, mid * k2); Dy: column [-1 0 1]
            const float dx = SS_FMA(SS_ADD(dm, dp), k1, SS_MUL(d0, k2));
            const float dy = SS_SUB(qp, qm);
            double* o = cov + 3 * (size_t)(yy * cs + xx);
            o[0] = (double)SS_MUL(dx, dx);
            o[1] = (double)SS_MUL(dx, dy);
            o[2] = (double)SS_MUL(dy, dy);
            dm = d0; qm = q0; d0 = dp; q0 = qp;
        }
    }
}

// phase 3: 3x3 sums (exact in double whatever the order: nine float32 products spanning < 2^46; one rounding) and the
// minimal eigenvalue.  Also returns this thread's first maximum (row-major order: larger value, then smaller index) of the
// responses it wrote, so the cell's first maximum needs no second pass over the response map.
SS_HD void phase_response(int tid, int nt, const double* cov, float* out, int cs, float& best_v, int& best_i) {
    const int nseg = (cs + SEG - 1) / SEG;
    for (int item = tid; item < cs * nseg; item += nt) {
        const int seg = item / cs, xx = item - seg * cs;
        const int y0 = seg * SEG;
        int y1 = y0 + SEG;
        if (y1 > cs) y1 = cs;
        const int om = 3 * (refl(xx - 1, cs) - xx), op = 3 * (refl(xx + 1, cs) - xx);   // neighbour offsets (doubles)
        double a0[3], a1[3], a2[3];                       // horizontal sums of rows y-1, y, y+1 for the three planes
        {
            const double* ra = cov + 3 * (size_t)(refl(y0 - 1, cs) * cs + xx);
            const double* rb = cov + 3 * (size_t)(y0 * cs + xx);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < 3; ++k) {
                a0[k] = ra[om + k] + ra[k] + ra[op + k];
                a1[k] = rb[om + k] + rb[k] + rb[op + k];
            }
        }
        for (int yy = y0; yy < y1; ++yy) {
            const double* rc = cov + 3 * (size_t)(refl(yy + 1, cs) * cs + xx);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < 3; ++k) a2[k] = rc[om + k] + rc[k] + rc[op + k];
            const double sxx = a0[0] + a1[0] + a2[0], sxy = a0[1] + a1[1] + a2[1], syy = a0[2] + a1[2] + a2[2];
            const float fa = SS_MUL(SS_D2F(sxx), 0.5f);
            const float fb = SS_D2F(sxy);
            const float fc = SS_MUL(SS_D2F(syy), 0.5f);
            const float t = SS_SUB(fa, fc);
            const float q = SS_ADD(SS_MUL(t, t), SS_MUL(fb, fb));
            const float v = SS_SUB(SS_ADD(fa, fc), SS_SQRT(q));
            const int i = yy * cs + xx;
            out[i] = v;
            if (v > best_v || (v == best_v && i < best_i)) { best_v = v; best_i = i; }
            a0[0] = a1[0]; a0[1] = a1[1]; a0[2] = a1[2];
            a1[0] = a2[0]; a1[1] = a2[1]; a1[2] = a2[2];
        }
    }
}

}  // namespace sscale

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

#if defined(__CUDA_ARCH__)
#define SS_UNROLL _Pragma("unroll")
#else
#define SS_UNROLL
#endif

// phase 1: raw = (cs+2)^2 tile (cell + 1-px halo from the parent image) -> bl = blurred cell.
// S / 16 rounded half-to-even in the first 16*floor(cs/16) columns, half-up in the tail.
// The row loops of all phases run a FIXED SEG iterations (stores predicated on yy < cs): fully unrolled, the sliding
// window is register renaming and the row pointers are increments.
SS_HD void phase_blur(int tid, int nt, const uint8_t* raw, uint8_t* bl, int cs) {
    const int rw = cs + 2, nvec16 = (cs >> 4) << 4;
    const int nseg = (cs + SEG - 1) / SEG;
    for (int item = tid; item < cs * nseg; item += nt) {
        const int seg = item / cs, xx = item - seg * cs;
        const int y0 = seg * SEG;
        const uint8_t* p = raw + y0 * rw + xx;                   // raw row y0 = the row above output row y0
        uint8_t* o = bl + y0 * cs + xx;
        int h0 = (int)p[0] + 2 * (int)p[1] + (int)p[2];
        int h1 = (int)p[rw] + 2 * (int)p[rw + 1] + (int)p[rw + 2];
        const bool even = xx < nvec16;
        SS_UNROLL
        for (int k = 0; k < SEG; ++k) {
            if (y0 + k < cs) {
                const uint8_t* q = p + (k + 2) * rw;
                const int h2 = (int)q[0] + 2 * (int)q[1] + (int)q[2];
                const int S = h0 + 2 * h1 + h2;
                int v;
                if (even) {
                    v = S >> 4;
                    const int rem = S & 15;
                    if (rem > 8 || (rem == 8 && (v & 1))) v++;
                } else {
                    v = (S + 8) >> 4;
                }
                o[k * cs] = (uint8_t)v;
                h0 = h1; h1 = h2;
            }
        }
    }
}

// Sobel row quantities of blurred row `row` at column x (xm / xp = reflected neighbours): d = right - left (exact), q = the
// [k1 k2 k1] row filter - vector form (FMAs) in the first 32*floor(cs/32) columns, scalar form in the tail.
SS_HD void cov_row(const uint8_t* row, int om, int op, bool vec, float& d, float& q) {
    const float k1 = 1.0f / 3060.0f, k2 = 2.0f / 3060.0f;       // float32(s), float32(2 s) (= 2 k1 exactly)
    const int a = row[om], c = row[op];
    const float A = (float)a, B = (float)row[0], C = (float)c;
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
        const int om = refl(xx - 1, cs) - xx, op = refl(xx + 1, cs) - xx;
        const bool vec = xx < nvec32;
        const uint8_t* r0 = bl + y0 * cs + xx;                   // current row, at column xx
        float dm, qm, d0, q0, dp, qp;
        cov_row(y0 == 0 ? r0 + cs : r0 - cs, om, op, vec, dm, qm);     // REFLECT_101 inside the cell: row -1 = row 1
        cov_row(r0, om, op, vec, d0, q0);
        double* o = cov + 3 * (size_t)(y0 * cs + xx);
        SS_UNROLL
        for (int k = 0; k < SEG; ++k) {
            if (y0 + k < cs) {
                const uint8_t* rc = r0 + k * cs;
                cov_row(y0 + k + 1 < cs ? rc + cs : rc - cs, om, op, vec, dp, qp);     // row cs = row cs - 2
                // Dx: row [-1 0 1] (exact), column [k1 k2 k1] evaluated as fma(top + bottom, k1, mid * k2); Dy: column [-1 0 1]
                const float dx = SS_FMA(SS_ADD(dm, dp), k1, SS_MUL(d0, k2));
                const float dy = SS_SUB(qp, qm);
                o[3 * k * cs + 0] = (double)SS_MUL(dx, dx);
                o[3 * k * cs + 1] = (double)SS_MUL(dx, dy);
                o[3 * k * cs + 2] = (double)SS_MUL(dy, dy);
                dm = d0; qm = q0; d0 = dp; q0 = qp;
            }
        }
    }
}

SS_HD void box_row(const double* r, int om, int op, double* a) {
    a[0] = r[om] + r[0] + r[op];
    a[1] = r[om + 1] + r[1] + r[op + 1];
    a[2] = r[om + 2] + r[2] + r[op + 2];
}

// phase 3: 3x3 sums (exact in double whatever the order: nine float32 products spanning < 2^46; one rounding) and the
// minimal eigenvalue.  Also returns this thread's first maximum (row-major order: larger value, then smaller index) of the
// responses it wrote, so the cell's first maximum needs no second pass over the response map.
SS_HD void phase_response(int tid, int nt, const double* cov, float* out, int cs, float& best_v, int& best_i) {
    const int nseg = (cs + SEG - 1) / SEG;
    for (int item = tid; item < cs * nseg; item += nt) {
        const int seg = item / cs, xx = item - seg * cs;
        const int y0 = seg * SEG;
        const int om = 3 * (refl(xx - 1, cs) - xx), op = 3 * (refl(xx + 1, cs) - xx);   // neighbour offsets (doubles)
        const int rs = 3 * cs;                                                           // row stride (doubles)
        const double* r0 = cov + 3 * (size_t)(y0 * cs + xx);
        double a0[3], a1[3], a2[3];                       // horizontal sums of rows y-1, y, y+1 for the three planes
        box_row(y0 == 0 ? r0 + rs : r0 - rs, om, op, a0);
        box_row(r0, om, op, a1);
        SS_UNROLL
        for (int k = 0; k < SEG; ++k) {
            if (y0 + k < cs) {
                const double* rc = r0 + k * rs;
                box_row(y0 + k + 1 < cs ? rc + rs : rc - rs, om, op, a2);
                const double sxx = a0[0] + a1[0] + a2[0], sxy = a0[1] + a1[1] + a2[1], syy = a0[2] + a1[2] + a2[2];
                const float fa = SS_MUL(SS_D2F(sxx), 0.5f);
                const float fb = SS_D2F(sxy);
                const float fc = SS_MUL(SS_D2F(syy), 0.5f);
                const float t = SS_SUB(fa, fc);
                const float q = SS_ADD(SS_MUL(t, t), SS_MUL(fb, fb));
                const float v = SS_SUB(SS_ADD(fa, fc), SS_SQRT(q));
                const int i = (y0 + k) * cs + xx;
                out[i] = v;
                if (v > best_v || (v == best_v && i < best_i)) { best_v = v; best_i = i; }
                a0[0] = a1[0]; a0[1] = a1[1]; a0[2] = a1[2];
                a1[0] = a2[0]; a1[1] = a2[1]; a1[2] = a2[2];
            }
        }
    }
}

}  // namespace sscale

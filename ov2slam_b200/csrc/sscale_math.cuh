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

// shared-memory carve-up (bytes) for a cs x cs cell
SS_HD size_t raw_bytes(int cs) { return (size_t)(((cs + 2) * (cs + 2) + 15) & ~15); }
SS_HD size_t blur_bytes(int cs) { return (size_t)((cs * cs + 15) & ~15); }
SS_HD size_t smem_bytes(int cs) { return raw_bytes(cs) + blur_bytes(cs) + 3 * sizeof(float) * (size_t)cs * cs; }

// phase 1: raw = (cs+2)^2 tile (cell + 1-px halo from the parent image) -> bl = blurred cell.
// S / 16 rounded half-to-even in the first 16*floor(cs/16) columns, half-up in the tail.
SS_HD void phase_blur(int tid, int nt, const uint8_t* raw, uint8_t* bl, int cs) {
    const int rw = cs + 2, nvec16 = (cs >> 4) << 4;
    for (int i = tid; i < cs * cs; i += nt) {
        const int yy = i / cs, xx = i - yy * cs;
        const uint8_t* p = raw + yy * rw + xx;                    // top-left of the 3x3 window
        const int S = (int)p[0] + 2 * (int)p[1] + (int)p[2] + 2 * ((int)p[rw] + 2 * (int)p[rw + 1] + (int)p[rw + 2]) +
                      (int)p[2 * rw] + 2 * (int)p[2 * rw + 1] + (int)p[2 * rw + 2];
        int v;
        if (xx < nvec16) {
            v = S >> 4;
            const int rem = S & 15;
            if (rem > 8 || (rem == 8 && (v & 1))) v++;
        } else {
            v = (S + 8) >> 4;
        }
        bl[i] = (uint8_t)v;
    }
}

// phase 2: Sobel derivatives (scale 1/(4*3*255)) and their products
SS_HD void phase_cov(int tid, int nt, const uint8_t* bl, float* cxx, float* cxy, float* cyy, int cs) {
    const float k1 = 1.0f / 3060.0f, k2 = 2.0f / 3060.0f;       // float32(s), float32(2 s) (= 2 k1 exactly)
    const int nvec32 = (cs >> 5) << 5;
    for (int i = tid; i < cs * cs; i += nt) {
        const int yy = i / cs, xx = i - yy * cs;
        const int xm = refl(xx - 1, cs), xp = refl(xx + 1, cs);
        const uint8_t* rm = bl + refl(yy - 1, cs) * cs;
        const uint8_t* r0 = bl + yy * cs;
        const uint8_t* rp = bl + refl(yy + 1, cs) * cs;
        // Dx: row [-1 0 1] (exact), column [k1 k2 k1] evaluated as fma(top + bottom, k1, mid * k2)
        const float dm = (float)((int)rm[xp] - (int)rm[xm]);
        const float d0 = (float)((int)r0[xp] - (int)r0[xm]);
        const float dp = (float)((int)rp[xp] - (int)rp[xm]);
        const float dx = SS_FMA(SS_ADD(dm, dp), k1, SS_MUL(d0, k2));
        // Dy: row [k1 k2 k1] (vector form in the first 32*floor(cs/32) columns, scalar tail after), column [-1 0 1]
        const float Am = (float)rm[xm], Bm = (float)rm[xx], Cm = (float)rm[xp];
        const float Ap = (float)rp[xm], Bp = (float)rp[xx], Cp = (float)rp[xp];
        float qm, qp;
        if (xx < nvec32) {
            qm = SS_FMA(k1, Cm, SS_FMA(k2, Bm, SS_MUL(k1, Am)));
            qp = SS_FMA(k1, Cp, SS_FMA(k2, Bp, SS_MUL(k1, Ap)));
        } else {
            qm = SS_ADD(SS_ADD(SS_MUL(k1, Am), SS_MUL(k2, Bm)), SS_MUL(k1, Cm));
            qp = SS_ADD(SS_ADD(SS_MUL(k1, Ap), SS_MUL(k2, Bp)), SS_MUL(k1, Cp));
        }
        const float dy = SS_SUB(qp, qm);
        cxx[i] = SS_MUL(dx, dx);
        cxy[i] = SS_MUL(dx, dy);
        cyy[i] = SS_MUL(dy, dy);
    }
}

// phase 3: 3x3 sums (exact in double, one rounding) and the minimal eigenvalue.
// A work item is one column x and a run of SEG rows: the horizontal 3-sums of a row are formed once and reused by the
// three output rows that need them (sliding window), 2.4x fewer shared-memory loads / conversions / double adds than
// nine taps per pixel.  The sums are EXACT in double whatever the order (nine float32 products spanning < 2^46), so the
// result is bit-identical to the nine-tap form the oracle uses.
SS_HD void phase_response(int tid, int nt, const float* cxx, const float* cxy, const float* cyy, float* out, int cs) {
    constexpr int SEG = 5;    // 35 x 7 = 245 work items for the 256 threads of a 35-px cell
    const int nseg = (cs + SEG - 1) / SEG;
    for (int item = tid; item < cs * nseg; item += nt) {
        const int seg = item / cs, xx = item - seg * cs;
        const int y0 = seg * SEG;
        int y1 = y0 + SEG;
        if (y1 > cs) y1 = cs;
        const int xm = refl(xx - 1, cs), xp = refl(xx + 1, cs);
        double a0[3], a1[3], a2[3];                       // horizontal sums of rows y-1, y, y+1 for the three planes
        {
            const int ra = refl(y0 - 1, cs) * cs, rb = y0 * cs;
            a0[0] = (double)cxx[ra + xm] + (double)cxx[ra + xx] + (double)cxx[ra + xp];
            a0[1] = (double)cxy[ra + xm] + (double)cxy[ra + xx] + (double)cxy[ra + xp];
            a0[2] = (double)cyy[ra + xm] + (double)cyy[ra + xx] + (double)cyy[ra + xp];
            a1[0] = (double)cxx[rb + xm] + (double)cxx[rb + xx] + (double)cxx[rb + xp];
            a1[1] = (double)cxy[rb + xm] + (double)cxy[rb + xx] + (double)cxy[rb + xp];
            a1[2] = (double)cyy[rb + xm] + (double)cyy[rb + xx] + (double)cyy[rb + xp];
        }
        for (int yy = y0; yy < y1; ++yy) {
            const int rc = refl(yy + 1, cs) * cs;
            a2[0] = (double)cxx[rc + xm] + (double)cxx[rc + xx] + (double)cxx[rc + xp];
            a2[1] = (double)cxy[rc + xm] + (double)cxy[rc + xx] + (double)cxy[rc + xp];
            a2[2] = (double)cyy[rc + xm] + (double)cyy[rc + xx] + (double)cyy[rc + xp];
            const double sxx = a0[0] + a1[0] + a2[0], sxy = a0[1] + a1[1] + a2[1], syy = a0[2] + a1[2] + a2[2];
            const float fa = SS_MUL(SS_D2F(sxx), 0.5f);
            const float fb = SS_D2F(sxy);
            const float fc = SS_MUL(SS_D2F(syy), 0.5f);
            const float t = SS_SUB(fa, fc);
            const float q = SS_ADD(SS_MUL(t, t), SS_MUL(fb, fb));
            out[yy * cs + xx] = SS_SUB(SS_ADD(fa, fc), SS_SQRT(q));
            a0[0] = a1[0]; a0[1] = a1[1]; a0[2] = a1[2];
            a1[0] = a2[0]; a1[1] = a2[1]; a1[2] = a2[2];
        }
    }
}

}  // namespace sscale

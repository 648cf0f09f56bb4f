// K (level set-up): template patch staging, Scharr derivatives and the int16 template / derivative
// window of one Lucas-Kanade level, for the 9 x 9 window (all reference configurations,
// /root/reference/src/feature_tracker.cpp:35-137 via cv::calcOpticalFlowPyrLK).
//
// Written as plain per-lane functions (no warp intrinsics; the phases are separated by __syncwarp in
// the kernel) so that tests/test_host_logic.py can compile the SAME code for the host and check it
// against the definition, lane by lane (test_klt_level_setup_lane_code_matches_definition).
//
// Shared-memory layout per warp:
//   sP  : 12 rows x 16 bytes.  Row r holds image row iy-1+r; byte (r, c) of the 12 x 12 neighbourhood
//         (c = 0 is image column ix-1) lives at sP[r * 16 + off + c], off = (ix-1) & 3 on the word-aligned
//         interior path (rows are staged as 4 ALIGNED 32-bit words each), off = 0 on the border path.
//   sD  : 10 rows x 12 ints, (dx & 0xFFFF) | (dy << 16) of the Scharr derivative at image position
//         (iy + r, ix + c); 0 outside the image (cv::buildOpticalFlowPyramid's BORDER_CONSTANT planes).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define KLT_HD __host__ __device__ __forceinline__
#else
#define KLT_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define KLT_LDG8(p) __ldg(p)
#define KLT_LDG32(p) __ldg(p)
#else
#define KLT_LDG8(p) (*(p))
#define KLT_LDG32(p) (*(p))
#endif

namespace kltsetup {

constexpr int WIN = 9;
constexpr int PW = WIN + 3;      // 12: u8 neighbourhood
constexpr int DW = WIN + 1;      // 10: derivative patch
constexpr int SP_PITCH = 16;     // bytes
constexpr int SD_PITCH = 12;     // ints
constexpr int SP_BYTES = PW * SP_PITCH;   // 192
constexpr int SD_INTS = DW * SD_PITCH;    // 120

KLT_HD int reflect101_any(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}

// true when the whole 12 x 12 neighbourhood (and therefore every Scharr tap) lies inside the image
KLT_HD bool patch_interior(int ix, int iy, int lw, int lh) { return ix >= 1 && iy >= 1 && ix + PW - 1 <= lw && iy + PW - 1 <= lh; }

// Phase 1 (all lanes).  Returns off.  `aligned`: image base and pitch are multiples of 4.
KLT_HD int stage_patch(int lane, const uint8_t* img, int pitch, int lw, int lh, int ix, int iy, bool aligned, uint8_t* sP) {
    if (aligned && patch_interior(ix, iy, lw, lh)) {
        const int x0 = (ix - 1) & ~3;                 // first staged column (>= 0), word aligned
        uint32_t* sPw = reinterpret_cast<uint32_t*>(sP);
        for (int i = lane; i < PW * 4; i += 32) {
            const int r = i >> 2, q = i & 3;
            const int xw = x0 + 4 * q;
            // a word starting at or beyond lw holds no byte the patch needs (and may lie outside the buffer)
            sPw[i] = xw < lw ? KLT_LDG32(reinterpret_cast<const uint32_t*>(img + (size_t)(iy - 1 + r) * pitch + xw)) : 0u;
        }
        return (ix - 1) & 3;
    }
    for (int i = lane; i < PW * PW; i += 32) {
        const int r = i / PW, c = i - r * PW;
        const int y = reflect101_any(iy - 1 + r, lh), x = reflect101_any(ix - 1 + c, lw);
        sP[r * SP_PITCH + c] = KLT_LDG8(img + (size_t)y * pitch + x);
    }
    return 0;
}

// Phase 2 (all lanes): Scharr.  Lane l < 30 owns derivative row l / 3, columns 4 (l % 3) .. +3 (the
// last group has 2 columns): 3 x 6 byte loads, the vertical [3 10 3] / [-1 0 1] passes once per column.
KLT_HD void scharr_rows(int lane, const uint8_t* sP, int off, int ix, int iy, int lw, int lh, int* sD) {
    if (lane >= 30) return;
    const int r = lane / 3, g = lane - 3 * r;
    const int c0 = 4 * g, nout = g == 2 ? 2 : 4;
    const uint8_t* p0 = sP + r * SP_PITCH + off + c0;   // neighbourhood row r (image row iy-1+r), column c0
    int vs[6], vd[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        if (j < nout + 2) {
            const int a = p0[j], b = p0[SP_PITCH + j], c = p0[2 * SP_PITCH + j];
            vs[j] = (a + c) * 3 + b * 10;
            vd[j] = c - a;
        } else {
            vs[j] = 0; vd[j] = 0;
        }
    }
    const int y = iy + r;
    const bool yin = y >= 0 && y < lh;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < nout) {
            const int x = ix + c0 + j;
            int dx = 0, dy = 0;
            if (yin && x >= 0 && x < lw) {
                dx = vs[j + 2] - vs[j];
                dy = (vd[j] + vd[j + 2]) * 3 + vd[j + 1] * 10;
            }
            sD[r * SD_PITCH + c0 + j] = (int)(((unsigned)dx & 0xFFFFu) | ((unsigned)dy << 16));
        }
    }
}

// Phase 3 (lanes 0..26; others get zeros): window pixels 3l .. 3l+2 = row l / 3, columns 3 (l % 3) .. +2.
// OpenCV's fixed point: I = (bilinear(u8) * 2^14 + 2^8) >> 9, Ix / Iy = (bilinear(s16) + 2^13) >> 14.
KLT_HD void template_rows(int lane, const uint8_t* sP, int off, const int* sD, int iw00, int iw01, int iw10, int iw11,
                          short* Iv, short* Ixv, short* Iyv, int& sA11, int& sA12, int& sA22) {
    sA11 = 0; sA12 = 0; sA22 = 0;
    Iv[0] = Iv[1] = Iv[2] = 0; Ixv[0] = Ixv[1] = Ixv[2] = 0; Iyv[0] = Iyv[1] = Iyv[2] = 0;
    if (lane >= 27) return;
    const int y = lane / 3, x0 = 3 * (lane - 3 * y);
    const uint8_t* q = sP + (y + 1) * SP_PITCH + off + x0 + 1;
    int t[4], b[4], dxt[4], dyt[4], dxb[4], dyb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        t[j] = q[j];
        b[j] = q[SP_PITCH + j];
        const int d0 = sD[y * SD_PITCH + x0 + j], d1 = sD[(y + 1) * SD_PITCH + x0 + j];
        dxt[j] = (int)(short)(d0 & 0xFFFF); dyt[j] = d0 >> 16;
        dxb[j] = (int)(short)(d1 & 0xFFFF); dyb[j] = d1 >> 16;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int ival = (t[k] * iw00 + t[k + 1] * iw01 + b[k] * iw10 + b[k + 1] * iw11 + (1 << 8)) >> 9;
        const int ixval = (dxt[k] * iw00 + dxt[k + 1] * iw01 + dxb[k] * iw10 + dxb[k + 1] * iw11 + (1 << 13)) >> 14;
        const int iyval = (dyt[k] * iw00 + dyt[k + 1] * iw01 + dyb[k] * iw10 + dyb[k + 1] * iw11 + (1 << 13)) >> 14;
        Iv[k] = (short)ival; Ixv[k] = (short)ixval; Iyv[k] = (short)iyval;
        sA11 += ixval * ixval;
        sA12 += ixval * iyval;
        sA22 += iyval * iyval;
    }
}

// dp2a: signed 16-bit halves of a times UNSIGNED bytes of b (lo: bytes 0, 1; hi: bytes 2, 3), plus c
KLT_HD int dp2a_lo_su16(unsigned a, unsigned b, int c) {
#if defined(__CUDA_ARCH__)
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return c + (int)(short)(a & 0xFFFFu) * (int)(b & 255u) + (int)(short)(a >> 16) * (int)((b >> 8) & 255u);
#endif
}
KLT_HD int dp2a_hi_su16(unsigned a, unsigned b, int c) {
#if defined(__CUDA_ARCH__)
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return c + (int)(short)(a & 0xFFFFu) * (int)((b >> 16) & 255u) + (int)(short)(a >> 16) * (int)(b >> 24);
#endif
}
KLT_HD unsigned funnel_r(unsigned lo, unsigned hi, unsigned sh) {   // bytes of hi:lo shifted right by sh bits (sh = 0, 8, 16, 24)
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}

// Phases 2 + 3 in one for the INTERIOR case (patch_interior(): every Scharr tap is a real pixel, no reflected border and
// no zero plane).  The bilinear interpolation and the Scharr filter are both exact integer linear maps, so they commute:
//   bilinear(Scharr(P))  ==  Scharr(W),   W[r][c] = P[r][c] iw00 + P[r][c+1] iw01 + P[r+1][c] iw10 + P[r+1][c+1] iw11
// with W <= 255 * 16385 (22 bits) and the Scharr sums of W below 2^27: same integers as template_rows() produces from the
// int16 derivative planes, then the same roundings (I: (W + 2^8) >> 9, Ix / Iy: (. + 2^13) >> 14).  A lane forms the 3 x 5 W
// values around its three pixels straight from the staged bytes with dp2a; nothing is written to sD and no second
// __syncwarp is needed.  Lane l < 27 owns window pixels (l / 3, 3 (l % 3) ..+2) as in template_rows().
KLT_HD void template_direct(int lane, const uint8_t* sP, int off, int iw00, int iw01, int iw10, int iw11,
                            short* Iv, short* Ixv, short* Iyv, int& sA11, int& sA12, int& sA22) {
    sA11 = 0; sA12 = 0; sA22 = 0;
    Iv[0] = Iv[1] = Iv[2] = 0; Ixv[0] = Ixv[1] = Ixv[2] = 0; Iyv[0] = Iyv[1] = Iyv[2] = 0;
    if (lane >= 27) return;
    const int y = lane / 3, x0 = 3 * (lane - 3 * y);
    const unsigned W01 = ((unsigned)iw00 & 0xFFFFu) | ((unsigned)iw01 << 16);
    const unsigned W23 = ((unsigned)iw10 & 0xFFFFu) | ((unsigned)iw11 << 16);
    // neighbourhood rows y .. y+3, bytes bo .. bo+5 of each 16-byte row (bo = off + x0 <= 9)
    const int bo = off + x0, wi = bo >> 2;
    const unsigned sh = (unsigned)(bo & 3) * 8u;
    unsigned lo[4], hi[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(sP + (y + r) * SP_PITCH);
        const unsigned w0 = rw[wi], w1 = rw[wi + 1], w2 = wi < 2 ? rw[wi + 2] : 0u;
        lo[r] = funnel_r(w0, w1, sh);      // bytes bo .. bo+3
        hi[r] = funnel_r(w1, w2, sh);      // bytes bo+4 .. bo+7 (only bo+4, bo+5 are used)
    }
    int W[3][5];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const unsigned t = lo[r], b = lo[r + 1];
        const unsigned t3 = funnel_r(lo[r], hi[r], 24u), b3 = funnel_r(lo[r + 1], hi[r + 1], 24u);
        W[r][0] = dp2a_lo_su16(W01, t, dp2a_lo_su16(W23, b, 0));
        W[r][1] = dp2a_lo_su16(W01, t >> 8, dp2a_lo_su16(W23, b >> 8, 0));
        W[r][2] = dp2a_hi_su16(W01, t, dp2a_hi_su16(W23, b, 0));
        W[r][3] = dp2a_lo_su16(W01, t3, dp2a_lo_su16(W23, b3, 0));
        W[r][4] = dp2a_lo_su16(W01, hi[r], dp2a_lo_su16(W23, hi[r + 1], 0));
    }
    int vs[5], vd[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        vs[c] = (W[0][c] + W[2][c]) * 3 + W[1][c] * 10;
        vd[c] = W[2][c] - W[0][c];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int ival = (W[1][k + 1] + (1 << 8)) >> 9;
        const int ixval = (vs[k + 2] - vs[k] + (1 << 13)) >> 14;
        const int iyval = ((vd[k] + vd[k + 2]) * 3 + vd[k + 1] * 10 + (1 << 13)) >> 14;
        Iv[k] = (short)ival; Ixv[k] = (short)ixval; Iyv[k] = (short)iyval;
        sA11 += ixval * ixval;
        sA12 += ixval * iyval;
        sA22 += iyval * iyval;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Row-per-lane forms (three keypoints per warp, fb_klt3_kernel): a lane owns window ROW r (9 pixels) of its keypoint.
// Interior windows only.  Same integers as template_rows() / the packed iteration path, checked on the host
// (tests/test_host_logic.py).

// 12 bytes of an image row starting at column x (x + 11 < lw guaranteed by patch_interior; the aligned word that holds
// bytes beyond column lw - 1 is not loaded): w[0..2] = bytes x .. x+11
KLT_HD void load_row12(const uint8_t* row, int x, int lw, unsigned w[3]) {
    const int x0 = x & ~3;
    const unsigned sh = (unsigned)(x & 3) * 8u;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(row + x0);
    const unsigned a0 = KLT_LDG32(p), a1 = KLT_LDG32(p + 1), a2 = KLT_LDG32(p + 2);
    const unsigned a3 = (sh != 0u && x0 + 12 < lw) ? KLT_LDG32(p + 3) : 0u;
    w[0] = funnel_r(a0, a1, sh);
    w[1] = funnel_r(a1, a2, sh);
    w[2] = funnel_r(a2, a3, sh);
}

// the eleven adjacent byte pairs (c, c+1), c = 0..10, of a 12-byte row as dp2a operands:
// W[c] = dp2a(W01, top pair c) + dp2a(W23, bottom pair c)
KLT_HD void interp_row11(const unsigned t[3], const unsigned b[3], unsigned W01, unsigned W23, int W[11]) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const unsigned tw = t[q], bw = b[q];
        W[4 * q + 0] = dp2a_lo_su16(W01, tw, dp2a_lo_su16(W23, bw, 0));
        W[4 * q + 1] = dp2a_lo_su16(W01, tw >> 8, dp2a_lo_su16(W23, bw >> 8, 0));
        W[4 * q + 2] = dp2a_hi_su16(W01, tw, dp2a_hi_su16(W23, bw, 0));
        if (q < 2) {
            const unsigned t3 = funnel_r(tw, t[q + 1], 24u), b3 = funnel_r(bw, b[q + 1], 24u);
            W[4 * q + 3] = dp2a_lo_su16(W01, t3, dp2a_lo_su16(W23, b3, 0));
        }
    }
}

// template row r: nb[k] = the 12-byte neighbourhood rows r + k (k = 0..3; neighbourhood row 0 is image row iy - 1, its
// byte 0 image column ix - 1).  Outputs the nine I / Ix / Iy of window row r and the row's share of the A sums.
KLT_HD void template_row9(const unsigned nb[4][3], int iw00, int iw01, int iw10, int iw11, short* Iv, short* Ixv, short* Iyv,
                          int& sA11, int& sA12, int& sA22, int& sabs) {
    const unsigned W01 = ((unsigned)iw00 & 0xFFFFu) | ((unsigned)iw01 << 16);
    const unsigned W23 = ((unsigned)iw10 & 0xFFFFu) | ((unsigned)iw11 << 16);
    int W0[11], W1[11], W2[11];
    interp_row11(nb[0], nb[1], W01, W23, W0);
    interp_row11(nb[1], nb[2], W01, W23, W1);
    interp_row11(nb[2], nb[3], W01, W23, W2);
    int vs[11], vd[11];
#pragma unroll
    for (int c = 0; c < 11; ++c) {
        vs[c] = (W0[c] + W2[c]) * 3 + W1[c] * 10;
        vd[c] = W2[c] - W0[c];
    }
    sA11 = 0; sA12 = 0; sA22 = 0; sabs = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int ival = (W1[k + 1] + (1 << 8)) >> 9;
        const int ixval = (vs[k + 2] - vs[k] + (1 << 13)) >> 14;
        const int iyval = ((vd[k] + vd[k + 2]) * 3 + vd[k + 1] * 10 + (1 << 13)) >> 14;
        Iv[k] = (short)ival; Ixv[k] = (short)ixval; Iyv[k] = (short)iyval;
        sA11 += ixval * ixval;
        sA12 += ixval * iyval;
        sA22 += iyval * iyval;
        sabs += (ixval < 0 ? -ixval : ixval) + (iyval < 0 ? -iyval : iyval);
    }
}

// one LK iteration's share of window row r: search rows (jy + r, jy + r + 1) as 12-byte rows starting at column jx
// (only bytes 0..9 are used)
KLT_HD void mismatch_row9(const unsigned top[3], const unsigned bot[3], int iw00, int iw01, int iw10, int iw11,
                          const short* Iv, const short* Ixv, const short* Iyv, int& sb1, int& sb2) {
    const unsigned W01 = ((unsigned)iw00 & 0xFFFFu) | ((unsigned)iw01 << 16);
    const unsigned W23 = ((unsigned)iw10 & 0xFFFFu) | ((unsigned)iw11 << 16);
    int J[11];
    interp_row11(top, bot, W01, W23, J);
    sb1 = 0; sb2 = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int jval = (J[k] + (1 << 8)) >> 9;
        const int diff = jval - (int)Iv[k];
        sb1 += diff * (int)Ixv[k];
        sb2 += diff * (int)Iyv[k];
    }
}

}  // namespace kltsetup

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

}  // namespace kltsetup

// B: binary descriptors, one warp per keypoint.
//
// Reference behaviour replaced: FeatureExtractor::describeBRIEF, non-contrib branch
// cv::ORB::create(500, 1., 0).compute(im, kps, descs) (/root/reference/src/feature_extractor.cpp:
// 224-285, :245).  For the keypoints the reference passes (size 1, angle -1, octave 0) that is:
//   drop keypoints whose rounded centre is closer than 31 px to the border,
//   smooth the image with a float32 separable 7-tap Gaussian (sigma 2) and round to u8,
//   256 intensity comparisons on the ORB pattern around c = cvRound(pt)   (SURVEY.md A.5).
// Instead of smoothing the whole image (2 x W x H bytes of traffic per frame) each warp smooths
// only the 26x26 window the pattern touches, from a 32x32 raw window, in shared memory, with
// exactly OpenCV's float operation order (rows: sequential FMA chain; columns: symmetric pairs +
// FMA; pinned bit-exact by oracle/image_ref.py::smooth7_ref against cv2).  Keypoints are >= 31 px
// from the border, so the 32x32 window never needs the REFLECT_101 border.
//
// The descriptor is DATA-driven (ov2_describe_config): mode OV2_DESC_ORB_FALLBACK is the above; mode
// OV2_DESC_BRIEF32 is the reference's DEFAULT build (CMakeLists.txt:12 WITH_OPENCV_CONTRIB,
// feature_extractor.cpp:242-243 cv::xfeatures2d::BriefDescriptorExtractor::create()): 256 comparisons of 9 x 9
// box sums (an integral-image lookup in opencv_contrib's brief.cpp, `smoothedSum`) at the test pairs of
// generated_32.i around (int)(pt + 0.5), first test in bit 7 of its byte, keypoints whose rounded centre is
// closer than PATCH_SIZE/2 + KERNEL_SIZE/2 = 28 px to the border dropped.  The 256 pairs ship only with
// opencv_contrib (not in this image): the caller loads them (scripts/brief_table_from_contrib.py turns
// generated_32.i into the table), the kernel (`describe_box_kernel`) is table-agnostic.
#include "ov2_common.cuh"
#include "../../include/ov2_orb_pattern.h"

namespace {

constexpr int WARPS = 4;

__device__ __align__(16) signed char g_pat[256][4];   // global (not __constant__): every lane reads its own 8 pairs
__constant__ float c_gk[7];

struct DescArgs {
    const uint8_t* img; int w, h, pitch; long long fstride;
    int n;
    const int32_t* frame_idx; int first_frame, per_frame;
    const float2* pts;
    uint8_t* desc; uint8_t* valid;
};

__global__ void __launch_bounds__(WARPS * 32) describe_kernel(DescArgs A) {
    __shared__ float sraw[WARPS][32 * 33];   // 32x32 raw window as float (+1 pad: conflict-free rows); row-filtered IN PLACE
    __shared__ uint8_t ssm[WARPS][26 * 28];  // 26 x 26 smoothed (+2 pad)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * WARPS + warp;
    if (i >= A.n) return;
    const float2 pt = A.pts[i];
    const int cx = __float2int_rn(pt.x), cy = __float2int_rn(pt.y);
    uint8_t* dout = A.desc + (size_t)i * 32;
    // KeyPointsFilter::runByImageBorder(kps, size, 31) tests the rounded point; x < 0 marks an empty slot
    if (pt.x < 0.f || cx < 31 || cy < 31 || cx >= A.w - 31 || cy >= A.h - 31) {
        dout[lane] = 0;
        if (lane == 0) A.valid[i] = 0;
        return;
    }
    const int frame = A.frame_idx ? A.frame_idx[i] : A.first_frame + i / A.per_frame;
    const uint8_t* img = A.img + A.fstride * frame + (size_t)(cy - 16) * A.pitch + (cx - 16);
    float* raw = sraw[warp];
    float* row = raw;   // the row filter overwrites column c only after its taps c..c+6 sit in registers
    uint8_t* sm = ssm[warp];
    // raw window rows cy-16..cy+15, cols cx-16..cx+15 (lane = column): 32 coalesced 32-byte reads
#pragma unroll 8
    for (int r = 0; r < 32; ++r) raw[r * 33 + lane] = (float)__ldg(img + (size_t)r * A.pitch + lane);
    __syncwarp();
    const float k0 = c_gk[0], k1 = c_gk[1], k2 = c_gk[2], k3 = c_gk[3];
    // row filter, lane = row: output col c (0..25) <-> image col cx-13+c, taps raw cols c..c+6,
    // sequential FMA chain exactly as OpenCV's float row filter evaluates it
    {
        const float* p = raw + lane * 33;
        float x0 = p[0], x1 = p[1], x2 = p[2], x3 = p[3], x4 = p[4], x5 = p[5];
#pragma unroll
        for (int c = 0; c < 26; ++c) {
            const float x6 = p[c + 6];
            float acc = x0 * k0;
            acc = __fmaf_rn(x1, k1, acc);
            acc = __fmaf_rn(x2, k2, acc);
            acc = __fmaf_rn(x3, k3, acc);
            acc = __fmaf_rn(x4, k2, acc);
            acc = __fmaf_rn(x5, k1, acc);
            acc = __fmaf_rn(x6, k0, acc);
            row[lane * 33 + c] = acc;
            x0 = x1; x1 = x2; x2 = x3; x3 = x4; x4 = x5; x5 = x6;
        }
    }
    __syncwarp();
    // column filter, lane = column: output row q (0..25) <-> image row cy-13+q, taps rows q..q+6
    // (centre q+3), symmetric pairs + FMA as OpenCV's float column filter, rint -> u8
    if (lane < 26) {
        const float* p = row + lane;
        float y0 = p[0], y1 = p[33], y2 = p[2 * 33], y3 = p[3 * 33], y4 = p[4 * 33], y5 = p[5 * 33];
#pragma unroll
        for (int q = 0; q < 26; ++q) {
            const float y6 = p[(q + 6) * 33];
            float acc = y3 * k3;
            acc = __fmaf_rn(y2 + y4, k2, acc);
            acc = __fmaf_rn(y1 + y5, k1, acc);
            acc = __fmaf_rn(y0 + y6, k0, acc);
            int v = __float2int_rn(acc);
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            sm[q * 28 + lane] = (uint8_t)v;
            y0 = y1; y1 = y2; y2 = y3; y3 = y4; y4 = y5; y5 = y6;
        }
    }
    __syncwarp();
    // 8 tests per lane -> one descriptor byte per lane
    unsigned byte = 0;
    const int4 pa = __ldg(reinterpret_cast<const int4*>(&g_pat[lane * 8][0]));
    const int4 pb = __ldg(reinterpret_cast<const int4*>(&g_pat[lane * 8 + 4][0]));
    const int pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int t0 = (int)(signed char)(pw[k] & 255), t1 = (int)(signed char)((pw[k] >> 8) & 255);
        const int t2 = (int)(signed char)((pw[k] >> 16) & 255), t3 = (int)(signed char)((pw[k] >> 24) & 255);
        int a = sm[(t1 + 13) * 28 + (t0 + 13)];
        int b = sm[(t3 + 13) * 28 + (t2 + 13)];
        byte |= (unsigned)(a < b) << k;
    }
    dout[lane] = (uint8_t)byte;
    if (lane == 0) A.valid[i] = 1;
}


// ---- BRIEF-32 (box-smoothed tests): one warp per keypoint.
// raw  : 57 x 57 window (rows cy-28 .. cy+28), pitch 60 bytes
// hs   : horizontal 9-sums, hs[r][c] = sum raw[r][c .. c+8], c = 0..48 (centre column c+4), pitch 50 u16
// A test point (y, x), |y|, |x| <= 24, is the 9 x 9 box centred on raw (y+28, x+28):
//   S(y, x) = sum_{r = y+24 .. y+32} hs[r][x+24]      (exact integers, <= 81 * 255)
constexpr int BW = 57, BRAW_PITCH = 60, BHS_W = 49, BHS_PITCH = 50;

struct BoxArgs {
    DescArgs d;
    const signed char* table;   // [256][4] = (y0, x0, y1, x1) of SMOOTHED(y0, x0) < SMOOTHED(y1, x1), generated_32.i order
};

__global__ void __launch_bounds__(WARPS * 32) describe_box_kernel(BoxArgs B) {
    __shared__ __align__(4) uint8_t sraw[WARPS][BW * BRAW_PITCH];
    __shared__ __align__(4) unsigned short shs[WARPS][BW * BHS_PITCH];
    const DescArgs& A = B.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * WARPS + warp;
    if (i >= A.n) return;
    const float2 pt = A.pts[i];
    uint8_t* dout = A.desc + (size_t)i * 32;
    // KeyPointsFilter::runByImageBorder(kps, size, 28) tests the cvRound'ed point
    const int rx = __float2int_rn(pt.x), ry = __float2int_rn(pt.y);
    if (pt.x < 0.f || rx < 28 || ry < 28 || rx >= A.w - 28 || ry >= A.h - 28) {
        dout[lane] = 0;
        if (lane == 0) A.valid[i] = 0;
        return;
    }
    // smoothedSum: img_x = (int)(pt.x + 0.5) + x  (truncation of a double sum; pt >= 27.5 here so floor == trunc)
    const int cx = (int)((double)pt.x + 0.5), cy = (int)((double)pt.y + 0.5);
    const int frame = A.frame_idx ? A.frame_idx[i] : A.first_frame + i / A.per_frame;
    const uint8_t* img = A.img + A.fstride * frame;
    uint8_t* raw = sraw[warp];
    unsigned short* hs = shs[warp];
    // cx can exceed the rounded centre by one at a .5 tie (window one column / row past the image when W - 29 is
    // even): those taps are clamped to the last column / row (OpenCV reads the integral image out of its row there)
    for (int r = 0; r < BW; ++r) {
        int y = cy - 28 + r;
        y = y > A.h - 1 ? A.h - 1 : y;
        const uint8_t* rowp = img + (size_t)y * A.pitch;
        int x0 = cx - 28 + lane, x1 = x0 + 32;
        x0 = x0 > A.w - 1 ? A.w - 1 : x0;
        x1 = x1 > A.w - 1 ? A.w - 1 : x1;
        raw[r * BRAW_PITCH + lane] = __ldg(rowp + x0);
        if (lane + 32 < BW) raw[r * BRAW_PITCH + lane + 32] = __ldg(rowp + x1);
    }
    __syncwarp();
    // horizontal sliding 9-sums, lane = row
    for (int r = lane; r < BW; r += 32) {
        const uint8_t* p = raw + r * BRAW_PITCH;
        unsigned short* o = hs + r * BHS_PITCH;
        int s = 0;
#pragma unroll
        for (int c = 0; c < 9; ++c) s += p[c];
        o[0] = (unsigned short)s;
#pragma unroll 8
        for (int c = 1; c < BHS_W; ++c) {
            s += (int)p[c + 8] - (int)p[c - 1];
            o[c] = (unsigned short)s;
        }
    }
    __syncwarp();
    // 8 tests per lane -> descriptor byte `lane`, first test in bit 7 (generated_32.i: (t0 << 7) + (t1 << 6) + ...)
    const int4 pa = __ldg(reinterpret_cast<const int4*>(B.table + (size_t)lane * 32));
    const int4 pb = __ldg(reinterpret_cast<const int4*>(B.table + (size_t)lane * 32 + 16));
    const int pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
    unsigned byte = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int y0 = (int)(signed char)(pw[k] & 255), x0 = (int)(signed char)((pw[k] >> 8) & 255);
        const int y1 = (int)(signed char)((pw[k] >> 16) & 255), x1 = (int)(signed char)((pw[k] >> 24) & 255);
        const unsigned short* pa0 = hs + (y0 + 24) * BHS_PITCH + (x0 + 24);
        const unsigned short* pb0 = hs + (y1 + 24) * BHS_PITCH + (x1 + 24);
        int a = 0, b = 0;
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            a += pa0[r * BHS_PITCH];
            b += pb0[r * BHS_PITCH];
        }
        byte |= (unsigned)(a < b) << (7 - k);
    }
    dout[lane] = (uint8_t)byte;
    if (lane == 0) A.valid[i] = 1;
}

bool g_tables_loaded[64] = {false};

}  // namespace


extern "C" ov2_status ov2_describe_config(ov2_ctx* ctx, int mode, const int8_t* pairs) {
    if (!ctx) return OV2_ERR_INVALID;
    if (mode == OV2_DESC_ORB_FALLBACK) {
        if (pairs) return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_describe_config: the ORB-fallback pattern is built in (pass NULL)");
        ctx->desc_mode = OV2_DESC_ORB_FALLBACK;
        return OV2_OK;
    }
    if (mode != OV2_DESC_BRIEF32) return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_describe_config: unknown descriptor mode");
    if (!pairs)
        return ov2_fail(ctx, OV2_ERR_INVALID,
                        "ov2_describe_config: BRIEF-32 needs the 256 test pairs of opencv_contrib's generated_32.i (not redistributed here)");
    int8_t host[1024];
    if (ov2_is_device_ptr(pairs)) OV2_CUDA(ctx, cudaMemcpy(host, pairs, 1024, cudaMemcpyDeviceToHost));
    else memcpy(host, pairs, 1024);
    for (int k = 0; k < 1024; ++k)
        if (host[k] < -24 || host[k] > 24)
            return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_describe_config: test offsets must lie in the 48 x 48 patch (|offset| <= 24)");
    if (!ctx->desc_table) OV2_CUDA(ctx, cudaMalloc(&ctx->desc_table, 1024));
    OV2_CUDA(ctx, cudaMemcpyAsync(ctx->desc_table, host, 1024, cudaMemcpyHostToDevice, ctx->stream));
    OV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `host` is a stack buffer
    ctx->desc_mode = OV2_DESC_BRIEF32;
    return OV2_OK;
}

extern "C" ov2_status ov2_describe(ov2_ctx* ctx, const ov2_pyr* pyr, int n, const int32_t* frame_idx, int first_frame, int per_frame,
                                   const float* pts, uint8_t* desc32_out, uint8_t* valid_out) {
    if (!ctx || !pyr || !pyr->l0 || n < 0) return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_describe: bad arguments");
    if (n == 0) return OV2_OK;  // feature_extractor.cpp:226-229
    if (!pts || !desc32_out || !valid_out || (!frame_idx && per_frame <= 0))
        return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_describe: null array");
    if (!frame_idx && (first_frame < 0 || first_frame + (n + per_frame - 1) / per_frame > pyr->batch))
        return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_describe: more frames than pyramid slots");
    ov2_status st = ov2_begin(ctx);
    if (st != OV2_OK) return st;
    if (ctx->device < 64 && !g_tables_loaded[ctx->device]) {
        // cv::getGaussianKernel(7, 2, CV_32F) as exact float32 values (pinned in
        // tests/test_oracle_image.py against cv2.getGaussianKernel)
        const float kf[7] = {0x1.1f5f62p-4f, 0x1.0c70fcp-3f, 0x1.869472p-3f, 0x1.ba95c0p-3f,
                             0x1.869472p-3f, 0x1.0c70fcp-3f, 0x1.1f5f62p-4f};
        OV2_CUDA(ctx, cudaMemcpyToSymbol(c_gk, kf, sizeof(kf)));
        OV2_CUDA(ctx, cudaMemcpyToSymbol(g_pat, OV2_ORB_PATTERN, sizeof(OV2_ORB_PATTERN)));
        g_tables_loaded[ctx->device] = true;
    }
    DescArgs A;
    A.img = pyr->l0; A.w = pyr->w[0]; A.h = pyr->h[0]; A.pitch = (int)pyr->l0_pitch; A.fstride = (long long)pyr->l0_fstride;
    A.n = n; A.first_frame = first_frame; A.per_frame = per_frame;
    const void* d = nullptr;
    void* o = nullptr;
    if ((st = ov2_stage_in(ctx, frame_idx, sizeof(int32_t) * (size_t)n, &d)) != OV2_OK) return st;
    A.frame_idx = (const int32_t*)d;
    if ((st = ov2_stage_in(ctx, pts, sizeof(float) * 2 * (size_t)n, &d)) != OV2_OK) return st;
    A.pts = (const float2*)d;
    if ((st = ov2_stage_out(ctx, desc32_out, (size_t)32 * n, &o)) != OV2_OK) return st;
    A.desc = (uint8_t*)o;
    if ((st = ov2_stage_out(ctx, valid_out, (size_t)n, &o)) != OV2_OK) return st;
    A.valid = (uint8_t*)o;
    if (ctx->desc_mode == OV2_DESC_BRIEF32) {
        BoxArgs B;
        B.d = A;
        B.table = (const signed char*)ctx->desc_table;
        OV2_LAUNCH(ctx, "describe_box_kernel", describe_box_kernel<<<div_up(n, WARPS), WARPS * 32, 0, ctx->stream>>>(B));
    } else {
        OV2_LAUNCH(ctx, "describe_kernel", describe_kernel<<<div_up(n, WARPS), WARPS * 32, 0, ctx->stream>>>(A));
    }
    return ov2_end(ctx);
}

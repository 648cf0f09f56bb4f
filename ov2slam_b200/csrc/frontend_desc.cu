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

bool g_tables_loaded[64] = {false};

}  // namespace

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
    OV2_LAUNCH(ctx, "describe_kernel", describe_kernel<<<div_up(n, WARPS), WARPS * 32, 0, ctx->stream>>>(A));
    return ov2_end(ctx);
}

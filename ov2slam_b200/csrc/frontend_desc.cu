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

__constant__ signed char c_pat[256][4];
__constant__ float c_gk[7];

struct DescArgs {
    const uint8_t* img; int w, h, pitch; long long fstride;
    int n;
    const int32_t* frame_idx; int first_frame, per_frame;
    const float2* pts;
    uint8_t* desc; uint8_t* valid;
};

__global__ void __launch_bounds__(WARPS * 32) describe_kernel(DescArgs A) {
    __shared__ uint8_t sraw[WARPS][32 * 32];
    __shared__ float srow[WARPS][32 * 27];   // 32 rows x 26 cols (+1 pad)
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
    uint8_t* raw = sraw[warp];
    float* row = srow[warp];
    uint8_t* sm = ssm[warp];
    // raw window rows cy-16..cy+15, cols cx-16..cx+15: lane = column
    for (int r = 0; r < 32; ++r) raw[r * 32 + lane] = __ldg(img + (size_t)r * A.pitch + lane);
    __syncwarp();
    // row filter: output col c (0..25) <-> image col cx-13+c, taps raw cols c..c+6
    const float k0 = c_gk[0], k1 = c_gk[1], k2 = c_gk[2], k3 = c_gk[3];
    for (int e = lane; e < 32 * 26; e += 32) {
        int r = e / 26, c = e - r * 26;
        const uint8_t* p = raw + r * 32 + c;
        float acc = (float)p[0] * k0;
        acc = __fmaf_rn((float)p[1], k1, acc);
        acc = __fmaf_rn((float)p[2], k2, acc);
        acc = __fmaf_rn((float)p[3], k3, acc);
        acc = __fmaf_rn((float)p[4], k2, acc);
        acc = __fmaf_rn((float)p[5], k1, acc);
        acc = __fmaf_rn((float)p[6], k0, acc);
        row[r * 27 + c] = acc;
    }
    __syncwarp();
    // column filter: output row q (0..25) <-> image row cy-13+q, taps rows q..q+6 (centre q+3)
    for (int e = lane; e < 26 * 26; e += 32) {
        int q = e / 26, c = e - q * 26;
        const float* p = row + q * 27 + c;
        float acc = p[3 * 27] * k3;
        acc = __fmaf_rn(p[2 * 27] + p[4 * 27], k2, acc);
        acc = __fmaf_rn(p[1 * 27] + p[5 * 27], k1, acc);
        acc = __fmaf_rn(p[0] + p[6 * 27], k0, acc);
        int v = __float2int_rn(acc);
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        sm[q * 28 + c] = (uint8_t)v;
    }
    __syncwarp();
    // 8 tests per lane -> one descriptor byte per lane
    unsigned byte = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const signed char* t = c_pat[lane * 8 + k];
        int a = sm[(t[1] + 13) * 28 + (t[0] + 13)];
        int b = sm[(t[3] + 13) * 28 + (t[2] + 13)];
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
        OV2_CUDA(ctx, cudaMemcpyToSymbol(c_pat, OV2_ORB_PATTERN, sizeof(OV2_ORB_PATTERN)));
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
    describe_kernel<<<div_up(n, WARPS), WARPS * 32, 0, ctx->stream>>>(A);
    OV2_CHECK_LAUNCH(ctx, "describe_kernel");
    return ov2_end(ctx);
}

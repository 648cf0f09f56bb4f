// P: image pyramid (cv::pyrDown semantics), one launch per level over the whole frame batch.
//
// Reference behaviour replaced: cv::buildOpticalFlowPyramid(img, pyr, Size(9,9), 3)
// (/root/reference/src/visual_front_end.cpp:1172).  out(y,x) = (sum_{i,j} k_i k_j in(2y+i-2,
// 2x+j-2) + 128) >> 8, k = [1 4 6 4 1], BORDER_REFLECT_101, size ((W+1)/2, (H+1)/2): pure
// integer arithmetic, bit-exact against cv2.pyrDown.  HBM-bound: reads the level once (through
// a shared-memory tile with a 2-px halo), writes a quarter of it.
#include "ov2_common.cuh"

namespace {

constexpr int TX = 32, TY = 8;          // threads
constexpr int OUT_W = TX * 4;           // 128 output pixels per tile row
constexpr int OUT_H = TY;               // 8 output rows
constexpr int IN_ROWS = 2 * OUT_H + 3;  // 19
constexpr int IN_WORDS = 66;            // 264 bytes: input cols [2*x0-4, 2*x0+260)

__global__ void __launch_bounds__(TX* TY)
pyr_down_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch, long long sfstride,
                uint8_t* __restrict__ dst, int dw, int dh, int dpitch, long long dfstride, int first) {
    __shared__ uint32_t tile[IN_ROWS][IN_WORDS];
    const int frame = first + blockIdx.z;
    src += sfstride * frame;
    dst += dfstride * frame;
    const int ox0 = blockIdx.x * OUT_W, oy0 = blockIdx.y * OUT_H;
    const int ix0 = 2 * ox0 - 4;  // input column of tile byte 0 (multiple of 4 minus 4)
    const int iy0 = 2 * oy0 - 2;
    const int tid = threadIdx.y * TX + threadIdx.x;

    const bool interior = ix0 >= 0 && ix0 + IN_WORDS * 4 <= sw && iy0 >= 0 && iy0 + IN_ROWS <= sh &&
                          ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)spitch) & 3) == 0;
    if (interior) {
        for (int i = tid; i < IN_ROWS * IN_WORDS; i += TX * TY) {
            int r = i / IN_WORDS, c = i - r * IN_WORDS;
            tile[r][c] = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)(iy0 + r) * spitch + ix0) + c);
        }
    } else {
        uint8_t* tb = reinterpret_cast<uint8_t*>(&tile[0][0]);
        for (int i = tid; i < IN_ROWS * IN_WORDS * 4; i += TX * TY) {
            int r = i / (IN_WORDS * 4), c = i - r * (IN_WORDS * 4);
            int y = reflect101(iy0 + r, sh), x = reflect101(ix0 + c, sw);
            // far outside (only beyond what any in-range output needs): clamp for safety
            y = clampi(y, 0, sh - 1);
            x = clampi(x, 0, sw - 1);
            tb[i] = __ldg(src + (size_t)y * spitch + x);
        }
    }
    __syncthreads();

    const int oy = oy0 + threadIdx.y;
    const int ox = ox0 + 4 * threadIdx.x;
    if (oy >= dh || ox >= dw) return;
    // vertical 5-tap on the 11 input columns this thread needs: tile bytes 8*tx+2 .. 8*tx+12
    int v[16];
    const int r0 = 2 * threadIdx.y;
#pragma unroll
    for (int wq = 0; wq < 4; ++wq) {
        uint32_t a = tile[r0][2 * threadIdx.x + wq], b = tile[r0 + 1][2 * threadIdx.x + wq],
                 c = tile[r0 + 2][2 * threadIdx.x + wq], d = tile[r0 + 3][2 * threadIdx.x + wq],
                 e = tile[r0 + 4][2 * threadIdx.x + wq];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int sh8 = 8 * k;
            v[4 * wq + k] = (int)((a >> sh8) & 255) + 4 * (int)((b >> sh8) & 255) + 6 * (int)((c >> sh8) & 255) +
                            4 * (int)((d >> sh8) & 255) + (int)((e >> sh8) & 255);
        }
    }
    uint32_t packed = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int b = 2 + 2 * j;  // v index of input col 2*(ox+j)-2
        int s = v[b] + 4 * v[b + 1] + 6 * v[b + 2] + 4 * v[b + 3] + v[b + 4];
        packed |= (uint32_t)((s + 128) >> 8) << (8 * j);
    }
    uint8_t* drow = dst + (size_t)oy * dpitch + ox;
    if (ox + 3 < dw && ((reinterpret_cast<uintptr_t>(drow)) & 3) == 0) {
        *reinterpret_cast<uint32_t*>(drow) = packed;
    } else {
        for (int j = 0; j < 4 && ox + j < dw; ++j) drow[j] = (uint8_t)(packed >> (8 * j));
    }
}

}  // namespace

extern "C" ov2_status ov2_pyr_build(ov2_ctx* ctx, ov2_pyr* p, const uint8_t* images, size_t row_stride,
                                    size_t frame_stride, int first, int count) {
    if (!ctx || !p || !images || first < 0 || count <= 0 || first + count > p->batch ||
        row_stride < (size_t)p->w[0])
        return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_pyr_build: bad arguments");
    ov2_status st = ov2_begin(ctx);
    if (st != OV2_OK) return st;
    if (ov2_is_device_ptr(images)) {
        const uint8_t* base = images - (ptrdiff_t)frame_stride * first;
        if (p->l0_mode == 1 || (p->l0_mode == 2 && (p->l0 != base || p->l0_pitch != row_stride ||
                                                      p->l0_fstride != frame_stride)))
            if (p->l0_mode == 1)
                return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_pyr_build: pyramid already owns level 0 (host images)");
        p->l0 = base;
        p->l0_pitch = row_stride;
        p->l0_fstride = frame_stride;
        p->l0_mode = 2;
    } else {
        if (p->l0_mode == 2)
            return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_pyr_build: pyramid aliases device images; cannot mix host images");
        if (!p->own[0]) {
            OV2_CUDA(ctx, cudaMalloc(&p->own[0], p->fstride[0] * (size_t)p->batch));
            p->l0 = p->own[0];
            p->l0_pitch = p->pitch[0];
            p->l0_fstride = p->fstride[0];
            p->l0_mode = 1;
        }
        if (frame_stride == row_stride * (size_t)p->h[0]) {
            OV2_CUDA(ctx, cudaMemcpy2DAsync(p->own[0] + p->fstride[0] * (size_t)first, p->pitch[0], images, row_stride,
                                            p->w[0], (size_t)p->h[0] * count, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            for (int k = 0; k < count; ++k)
                OV2_CUDA(ctx, cudaMemcpy2DAsync(p->own[0] + p->fstride[0] * (size_t)(first + k), p->pitch[0],
                                                images + frame_stride * (size_t)k, row_stride, p->w[0], p->h[0],
                                                cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    for (int l = 1; l < p->nlev; ++l) {
        const uint8_t* s = l == 1 ? p->l0 : p->own[l - 1];
        int spitch = (int)(l == 1 ? p->l0_pitch : p->pitch[l - 1]);
        long long sfs = (long long)(l == 1 ? p->l0_fstride : p->fstride[l - 1]);
        dim3 grid(div_up(p->w[l], OUT_W), div_up(p->h[l], OUT_H), count);
        OV2_LAUNCH(ctx, "pyr_down_kernel", pyr_down_kernel<<<grid, dim3(TX, TY), 0, ctx->stream>>>(s, p->w[l - 1], p->h[l - 1], spitch, sfs, p->own[l],
                                                               p->w[l], p->h[l], (int)p->pitch[l],
                                                               (long long)p->fstride[l], first));
    }
    return ov2_end(ctx);
}

// P: image pyramid (cv::pyrDown semantics), one launch per level over the whole frame batch.
//
// Reference behaviour replaced: cv::buildOpticalFlowPyramid(img, pyr, Size(9,9), 3)
// (/root/reference/src/visual_front_end.cpp:1172).  out(y,x) = (sum_{i,j} k_i k_j in(2y+i-2,
// 2x+j-2) + 128) >> 8, k = [1 4 6 4 1], BORDER_REFLECT_101, size ((W+1)/2, (H+1)/2): pure
// integer arithmetic, bit-exact against cv2.pyrDown.  HBM-bound: reads the level once (through
// a shared-memory tile with a 2-px halo), writes a quarter of it.
#include "ov2_common.cuh"

namespace {

constexpr int TX = 32, TY = 4;          // threads: each computes 4 pixels of TWO output rows
constexpr int OUT_W = TX * 4;           // 128 output pixels per tile row
constexpr int OUT_H = 2 * TY;           // 8 output rows
constexpr int IN_ROWS = 2 * OUT_H + 3;  // 19
constexpr int IN_WORDS = 66;            // 264 bytes: input cols [2*x0-4, 2*x0+260)

// 5-tap [1 4 6 4 1] on packed bytes with dp4a: the four outputs of a thread read input bytes
// 2*x0-2 .. 2*x0+8, i.e. bytes 2.. of word W0 up to byte 0 of word W3 (W0 = tile word 2*tx)
__device__ __forceinline__ void hrow4(const uint32_t* w, int& h0, int& h1, int& h2, int& h3) {
    const uint32_t W0 = w[0], W1 = w[1], W2 = w[2], W3 = w[3];
    h0 = (int)__dp4a(W0, 0x04010000u, __dp4a(W1, 0x00010406u, 0u));   // bytes: W0[2]*1 + W0[3]*4 + W1[0]*6 + W1[1]*4 + W1[2]*1
    h1 = (int)__dp4a(W1, 0x04060401u, __dp4a(W2, 0x00000001u, 0u));   // W1[0..3]*(1,4,6,4) + W2[0]*1
    h2 = (int)__dp4a(W1, 0x04010000u, __dp4a(W2, 0x00010406u, 0u));
    h3 = (int)__dp4a(W2, 0x04060401u, __dp4a(W3, 0x00000001u, 0u));
}

__global__ void __launch_bounds__(TX* TY)
pyr_down_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch, long long sfstride,
                uint8_t* __restrict__ dst, int dw, int dh, int dpitch, long long dfstride, int first) {
    __shared__ uint32_t tile[IN_ROWS][IN_WORDS];
    const int frame = first + blockIdx.z;
    src += sfstride * frame;
    dst += dfstride * frame;
    const int ox0 = blockIdx.x * OUT_W, oy0 = blockIdx.y * OUT_H;
    const int ix0 = 2 * ox0 - 4;  // input column of tile byte 0 (multiple of 4)
    const int iy0 = 2 * oy0 - 2;

    // Fixed 2-D mapping (no divisions): rows ty, ty+8, ty+16; words tx, tx+32 and the 2 tail words.
    // BORDER_REFLECT_101 in y only changes which row is read, so whole words are still loaded;
    // only words that stick out of [0, sw) in x are assembled byte by byte.
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)spitch) & 3) == 0;
    for (int r = threadIdx.y; r < IN_ROWS; r += TY) {
        int y = reflect101(iy0 + r, sh);
        y = clampi(y, 0, sh - 1);
        const uint8_t* grow = src + (size_t)y * spitch;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int c = threadIdx.x + 32 * k;
            if (c >= IN_WORDS) break;
            const int x0 = ix0 + 4 * c;
            uint32_t wv;
            if (x0 > sw + 1) {
                wv = 0;   // beyond the last column any existing output reads (2*(dw-1)+2 <= sw+1)
            } else if (aligned && x0 >= 0 && x0 + 4 <= sw) {
                wv = __ldg(reinterpret_cast<const uint32_t*>(grow + x0));
            } else {
                wv = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    int x = reflect101(x0 + b, sw);
                    x = clampi(x, 0, sw - 1);   // far outside: only beyond what any in-range output needs
                    wv |= (uint32_t)__ldg(grow + x) << (8 * b);
                }
            }
            tile[r][c] = wv;
        }
    }
    __syncthreads();

    // two output rows per thread (oy, oy+1) share 3 of their 5 input rows: 7 horizontal passes (dp4a)
    // instead of 10, then the vertical [1 4 6 4 1] on the running sums
    const int oy = oy0 + 2 * threadIdx.y;
    const int ox = ox0 + 4 * threadIdx.x;
    if (oy >= dh || ox >= dw) return;
    const int r0 = 4 * threadIdx.y;
    int h[7][4];
#pragma unroll
    for (int r = 0; r < 7; ++r) hrow4(&tile[r0 + r][2 * threadIdx.x], h[r][0], h[r][1], h[r][2], h[r][3]);
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        if (oy + o >= dh) break;
        int v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            v[j] = h[2 * o][j] + 4 * h[2 * o + 1][j] + 6 * h[2 * o + 2][j] + 4 * h[2 * o + 3][j] + h[2 * o + 4][j];
        const uint32_t packed = (uint32_t)((v[0] + 128) >> 8) | ((uint32_t)((v[1] + 128) >> 8) << 8) |
                                ((uint32_t)((v[2] + 128) >> 8) << 16) | ((uint32_t)((v[3] + 128) >> 8) << 24);
        uint8_t* drow = dst + (size_t)(oy + o) * dpitch + ox;
        if (ox + 3 < dw && ((reinterpret_cast<uintptr_t>(drow)) & 3) == 0) {
            *reinterpret_cast<uint32_t*>(drow) = packed;
        } else {
            for (int j = 0; j < 4 && ox + j < dw; ++j) drow[j] = (uint8_t)(packed >> (8 * j));
        }
    }
}

}  // namespace

// level 0: alias device images / upload host images (H2D on the context's stream)
ov2_status ov2_pyr_load_level0(ov2_ctx* ctx, ov2_pyr* p, const uint8_t* images, size_t row_stride, size_t frame_stride,
                               int first, int count) {
    if (!ctx || !p || !images || first < 0 || count <= 0 || first + count > p->batch ||
        row_stride < (size_t)p->w[0])
        return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_pyr_build: bad arguments");
    if (ov2_is_device_ptr(images)) {
        const uint8_t* base = images - (ptrdiff_t)frame_stride * first;
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
        if (frame_stride == row_stride * (size_t)p->h[0] && row_stride == p->pitch[0]) {
            // fully contiguous on both sides: ONE linear copy (a pitched copy is issued as W-byte rows and
            // reaches a fraction of the PCIe bandwidth)
            OV2_CUDA(ctx, cudaMemcpyAsync(p->own[0] + p->fstride[0] * (size_t)first, images, frame_stride * (size_t)count,
                                          cudaMemcpyHostToDevice, ctx->stream));
        } else if (frame_stride == row_stride * (size_t)p->h[0]) {
            OV2_CUDA(ctx, cudaMemcpy2DAsync(p->own[0] + p->fstride[0] * (size_t)first, p->pitch[0], images, row_stride,
                                            p->w[0], (size_t)p->h[0] * count, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            for (int k = 0; k < count; ++k)
                OV2_CUDA(ctx, cudaMemcpy2DAsync(p->own[0] + p->fstride[0] * (size_t)(first + k), p->pitch[0],
                                                images + frame_stride * (size_t)k, row_stride, p->w[0], p->h[0],
                                                cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    return OV2_OK;
}

// levels 1.. from level 0 (kernels only)
ov2_status ov2_pyr_make_levels(ov2_ctx* ctx, ov2_pyr* p, int first, int count) {
    for (int l = 1; l < p->nlev; ++l) {
        const uint8_t* s = l == 1 ? p->l0 : p->own[l - 1];
        int spitch = (int)(l == 1 ? p->l0_pitch : p->pitch[l - 1]);
        long long sfs = (long long)(l == 1 ? p->l0_fstride : p->fstride[l - 1]);
        dim3 grid(div_up(p->w[l], OUT_W), div_up(p->h[l], OUT_H), count);
        OV2_LAUNCH(ctx, "pyr_down_kernel", pyr_down_kernel<<<grid, dim3(TX, TY), 0, ctx->stream>>>(s, p->w[l - 1], p->h[l - 1], spitch, sfs, p->own[l],
                                                               p->w[l], p->h[l], (int)p->pitch[l],
                                                               (long long)p->fstride[l], first));
    }
    return OV2_OK;
}

extern "C" ov2_status ov2_pyr_build(ov2_ctx* ctx, ov2_pyr* p, const uint8_t* images, size_t row_stride,
                                    size_t frame_stride, int first, int count) {
    if (!ctx) return OV2_ERR_INVALID;
    ov2_status st = ov2_begin(ctx);
    if (st != OV2_OK) return st;
    if ((st = ov2_pyr_load_level0(ctx, p, images, row_stride, frame_stride, first, count)) != OV2_OK) return st;
    if ((st = ov2_pyr_make_levels(ctx, p, first, count)) != OV2_OK) return st;
    return ov2_end(ctx);
}

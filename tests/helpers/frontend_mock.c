/* TEST HELPER (not part of the product): an LD_PRELOAD interposer that replaces the device calls the FeatureTracker drop-in makes
 * (context, pyramid slots, ov2_fb_klt, ov2_line_min_sad) by a RECORDER with canned, rule-based answers, so the host-side flow of the
 * MapManager::stereoMatching drop-in (ov2slam_b200/host/map_manager_stereo_gpu.cpp: which keypoints go to which tracker call with
 * which prior, what happens to the answers) can be checked on a box without a GPU.  Every tracker / row-search call is appended to
 * $OV2_MOCK_LOG; tests/test_host_shim.py replays the same rules in Python.
 *   CLAHE        dst = 255 - src (row strides honoured), arguments logged
 *   row search   xprior = floor(x) - 2 when floor(x) >= 3 and floor(y) % 4 != 0, else -1
 *   tracker      prior += (-0.5, dy), dy = 4 when floor(pt.x) % 7 == 0 else 0.25;
 *                status = floor(3 pt.x + pt.y) % 4 != 0 on a 2-level call (nbpyrlvl 1), floor(pt.x + 2 pt.y) % 6 != 0 otherwise */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/ov2b200.h"

static int dummy_ctx;
struct mock_pyr { int w, h, nlev; };

static FILE* log_file(void) {
    const char* path = getenv("OV2_MOCK_LOG");
    return path ? fopen(path, "ab") : NULL;
}

ov2_status ov2_create(int device, ov2_ctx** out) { (void)device; *out = (ov2_ctx*)&dummy_ctx; return OV2_OK; }
void ov2_destroy(ov2_ctx* ctx) { (void)ctx; }
const char* ov2_last_error(const ov2_ctx* ctx) { (void)ctx; return "mock"; }
ov2_status ov2_pyr_create(ov2_ctx* ctx, int batch, int width, int height, int nlevels_extra, ov2_pyr** out) {
    (void)ctx; (void)batch;
    struct mock_pyr* p = (struct mock_pyr*)malloc(sizeof(struct mock_pyr));
    p->w = width; p->h = height; p->nlev = nlevels_extra;
    *out = (ov2_pyr*)p;
    return OV2_OK;
}
void ov2_pyr_destroy(ov2_pyr* pyr) { free(pyr); }
ov2_status ov2_pyr_build(ov2_ctx* ctx, ov2_pyr* pyr, const uint8_t* images, size_t row_stride, size_t frame_stride, int first, int count) {
    (void)ctx; (void)pyr; (void)images; (void)row_stride; (void)frame_stride; (void)first; (void)count;
    FILE* f = log_file();
    if (f) { int32_t tag = 3; fwrite(&tag, 4, 1, f); fclose(f); }
    return OV2_OK;
}

ov2_status ov2_line_min_sad(ov2_ctx* ctx, const ov2_pyr* left, const ov2_pyr* right, int level, int n, const int32_t* frame_idx,
                            int first_frame, int per_frame, const float* pts, int nwinsize, int goleft, float* xprior_out, float* l1err_out) {
    (void)ctx; (void)left; (void)right; (void)frame_idx; (void)first_frame; (void)per_frame;
    FILE* f = log_file();
    if (f) {
        int32_t hd[5] = {1, level, n, nwinsize, goleft};
        fwrite(hd, 4, 5, f); fwrite(pts, 4, 2 * (size_t)n, f);
        fclose(f);
    }
    for (int i = 0; i < n; ++i) {
        const int fx = (int)floorf(pts[2 * i]), fy = (int)floorf(pts[2 * i + 1]);
        xprior_out[i] = (fx >= 3 && fy % 4 != 0) ? (float)(fx - 2) : -1.f;
        l1err_out[i] = xprior_out[i] < 0 ? 255.f : 1.f;
    }
    return OV2_OK;
}

ov2_status ov2_fb_klt(ov2_ctx* ctx, const ov2_pyr* prev, const ov2_pyr* cur, const ov2_klt_params* prm, int n, const int32_t* frame_idx,
                      int first_frame, int per_frame, const uint8_t* nbpyrlvl, int nbpyrlvl_all, const float* kps, float* priors_inout,
                      uint8_t* status_out) {
    (void)ctx; (void)prev; (void)cur; (void)frame_idx; (void)first_frame; (void)per_frame; (void)nbpyrlvl;
    FILE* f = log_file();
    if (f) {
        int32_t hd[4] = {2, n, nbpyrlvl_all, prm->win};
        fwrite(hd, 4, 4, f); fwrite(kps, 4, 2 * (size_t)n, f); fwrite(priors_inout, 4, 2 * (size_t)n, f);
        fclose(f);
    }
    for (int i = 0; i < n; ++i) {
        const float x = kps[2 * i], y = kps[2 * i + 1];
        priors_inout[2 * i] += -0.5f;
        priors_inout[2 * i + 1] += ((int)floorf(x) % 7 == 0) ? 4.f : 0.25f;
        status_out[i] = nbpyrlvl_all == 1 ? ((int)floorf(3.f * x + y) % 4 != 0) : ((int)floorf(x + 2.f * y) % 6 != 0);
    }
    return OV2_OK;
}

ov2_status ov2_clahe(ov2_ctx* ctx, const uint8_t* src, uint8_t* dst, int width, int height, size_t row_stride, size_t frame_stride, int count,
                     double clip_limit, int tiles_x, int tiles_y) {
    (void)ctx; (void)frame_stride;
    FILE* f = log_file();
    if (f) {
        int32_t hd[8] = {4, width, height, (int32_t)row_stride, count, (int32_t)(clip_limit * 1000), tiles_x, tiles_y};
        hd[0] = 4;
        fwrite(hd, 4, 8, f);
        int32_t inplace = src == dst;
        fwrite(&inplace, 4, 1, f);
        fclose(f);
    }
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) dst[(size_t)y * row_stride + x] = (uint8_t)(255 - src[(size_t)y * row_stride + x]);
    return OV2_OK;
}

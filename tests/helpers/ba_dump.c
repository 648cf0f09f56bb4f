/* TEST HELPER (not part of the product): an LD_PRELOAD interposer that replaces the device calls of the Optimizer drop-in
 * (ov2slam_b200/host/optimizer_localba_gpu.cpp) by a recorder: the flat window the drop-in WOULD solve on the GPU is written to
 * $OV2_BA_DUMP and the solve reports "nothing changed".  tests/test_oracle_vs_reference_map.py compares that window (which keyframes
 * are constant, which landmarks and observations are in) with what the reference's own localBA does with the same map. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ov2b200.h"

static int dummy_ctx;
ov2_status ov2_create(int device, ov2_ctx** out) { (void)device; *out = (ov2_ctx*)&dummy_ctx; return OV2_OK; }
const char* ov2_last_error(const ov2_ctx* ctx) { (void)ctx; return "recorder"; }
ov2_status ov2_localba_request_stop(ov2_ctx* ctx, int stop) { (void)ctx; (void)stop; return OV2_OK; }

ov2_status ov2_localba_solve(ov2_ctx* ctx, const ov2_ba_problem* p, const ov2_ba_opts* op, ov2_ba_result* res, uint8_t* outlier_out) {
    (void)ctx; (void)op;
    const char* path = getenv("OV2_BA_DUMP");
    FILE* f = path ? fopen(path, "wb") : NULL;
    if (!f) return OV2_ERR_INVALID;
    int32_t hd[3] = {p->ncam, p->npts, p->nobs};
    fwrite(hd, 4, 3, f);
    fwrite(p->pose, 8, 7 * (size_t)p->ncam, f);
    fwrite(p->pose_const, 1, (size_t)p->ncam, f);
    fwrite(p->lm_anchor_cam, 4, (size_t)p->npts, f);
    fwrite(p->lm_anchor_px, 8, 2 * (size_t)p->npts, f);
    fwrite(p->lm_invdepth, 8, (size_t)p->npts, f);
    fwrite(p->obs_cam, 4, (size_t)p->nobs, f);
    fwrite(p->obs_lm, 4, (size_t)p->nobs, f);
    fwrite(p->obs_px, 8, 2 * (size_t)p->nobs, f);
    fclose(f);
    memset(res, 0, sizeof(*res));
    if (outlier_out) memset(outlier_out, 0, (size_t)p->nobs);
    return OV2_OK;
}

/* TEST HELPER (not part of the product): an LD_PRELOAD interposer that replaces ov2_create / ov2_match_to_map by a recorder, so
 * the host-side flattening of the Mapper::matchToMap drop-in (ov2slam_b200/host/mapper_match_gpu.cpp) can be checked on a box
 * without a GPU: the problem the shim WOULD hand to the GPU is written to $OV2_MATCH_DUMP and every keypoint is reported
 * unmatched.  tests/test_host_shim.py runs the oracle on the recorded problem and on the original scene and compares. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/ov2b200.h"

static int dummy_ctx;

ov2_status ov2_create(int device, ov2_ctx** out) { (void)device; *out = (ov2_ctx*)&dummy_ctx; return OV2_OK; }
const char* ov2_last_error(const ov2_ctx* ctx) { (void)ctx; return "recorder"; }

#define W(ptr, n, sz) do { if ((n) > 0 && fwrite((ptr), (sz), (size_t)(n), f) != (size_t)(n)) { fclose(f); return OV2_ERR_INVALID; } } while (0)

ov2_status ov2_match_to_map(ov2_ctx* ctx, const ov2_match_problem* p, int32_t* best_kp_out, float* best_dist_out,
                            int32_t* kp_match_out, float* kp_dist_out) {
    (void)ctx;
    const char* path = getenv("OV2_MATCH_DUMP");
    FILE* f = path ? fopen(path, "wb") : NULL;
    if (!f) return OV2_ERR_INVALID;
    int32_t hd[12] = {p->nkps, p->nmps, p->ndesc, p->nobs, p->nkfs, p->ncand, p->ncells, p->nbwcells, p->ncellsize, p->img_w, p->img_h,
                      p->dist != NULL};
    float fl[3] = {p->dmaxpxdist, p->fdistratio, p->view_th};
    double zero[5] = {0, 0, 0, 0, 0};
    W(hd, 12, 4); W(fl, 3, 4);
    W(p->K, 4, 8); W(p->dist ? p->dist : zero, 5, 8); W(p->Tcw, 12, 8); W(p->kf_Tcw, 12 * p->nkfs, 8); W(p->mp_xyz, 3 * p->nmps, 8);
    W(p->kp_px, 2 * p->nkps, 4); W(p->obs_px, 2 * p->nobs, 4);
    W(p->cell_ptr, p->ncells + 1, 4); W(p->cell_kp, p->cell_ptr[p->ncells], 4); W(p->kp_lm, p->nkps, 4); W(p->mp_desc_ptr, p->nmps + 1, 4);
    W(p->mp_obs_ptr, p->nmps + 1, 4); W(p->obs_kf, p->nobs, 4); W(p->cand_mp, p->ncand, 4);
    W(p->mp_kfmask, 4 * p->nmps, 8);
    W(p->desc, 32 * p->ndesc, 1);
    fclose(f);
    for (int i = 0; i < p->ncand; ++i) { best_kp_out[i] = -1; best_dist_out[i] = 0.f; }
    for (int i = 0; i < p->nkps; ++i) { kp_match_out[i] = -1; kp_dist_out[i] = 1024.f; }
    return OV2_OK;
}

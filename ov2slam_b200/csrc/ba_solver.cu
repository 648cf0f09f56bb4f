// L: local bundle adjustment (placeholder until the solver lands; fails loudly, never falls back).
#include "ov2_common.cuh"
extern "C" ov2_status ov2_localba_solve(ov2_ctx* ctx, const ov2_ba_problem*, const ov2_ba_opts*, ov2_ba_result*, uint8_t*) {
    return ov2_fail(ctx, OV2_ERR_INVALID, "ov2_localba_solve: not built in this revision");
}

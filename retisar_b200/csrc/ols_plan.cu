// Uniformly partitioned overlap-save FIR convolver (reference: OverlapSaveConvolver,
// ReTiSAR/_convolver.py:436-669).  Implementation pending: entry points report RTB_ERR_UNSUPPORTED.
#include "rtb_common.cuh"

using namespace rtb;

extern "C" {

int rtb_ols_plan_create(rtb_ols_plan **plan, int, int, int, int, int, int) {
    if (plan) *plan = nullptr;
    return fail(RTB_ERR_UNSUPPORTED, "partitioned overlap-save convolver not built yet");
}
int rtb_ols_plan_set_filter(rtb_ols_plan *, const float *) { return fail(RTB_ERR_UNSUPPORTED, "not built yet"); }
int rtb_ols_plan_set_passthrough(rtb_ols_plan *, int) { return fail(RTB_ERR_UNSUPPORTED, "not built yet"); }
int rtb_ols_plan_reset(rtb_ols_plan *) { return fail(RTB_ERR_UNSUPPORTED, "not built yet"); }
int rtb_ols_process_block(rtb_ols_plan *, const float *, float *, void *) { return fail(RTB_ERR_UNSUPPORTED, "not built yet"); }
int rtb_ols_process_block_host(rtb_ols_plan *, const float *, float *) { return fail(RTB_ERR_UNSUPPORTED, "not built yet"); }
int rtb_ols_plan_query(rtb_ols_plan *, int, int64_t *) { return fail(RTB_ERR_UNSUPPORTED, "not built yet"); }
int rtb_ols_plan_destroy(rtb_ols_plan *) { return RTB_OK; }

}  // extern "C"

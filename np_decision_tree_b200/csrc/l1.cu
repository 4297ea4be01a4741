// K3: L1 (MAE) split search -- placeholder until the order-statistic kernels land.
#include "tree_kernels.cuh"

int l1_level_search(b200dt_ctx *ctx, Session *, b200dt_candidate *, bool) {
    return ctx->fail_msg("L1 search: not implemented yet");
}
int l1_cummedian_device(b200dt_ctx *ctx, const double *, int64_t, double *) {
    return ctx->fail_msg("cummedian: not implemented yet");
}
int l1_leaf_medians(b200dt_ctx *ctx, Session *) { return ctx->fail_msg("L1 leaves: not implemented yet"); }

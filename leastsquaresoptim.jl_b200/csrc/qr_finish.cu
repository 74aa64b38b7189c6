// qr_finish.cu — rank-revealing finish on the small n x n R (Q-d): placeholder until the pivoted path lands.
#include "qr.cuh"
int small_qr_finish(lso_ctx* ctx, int64_t n, double* d_R, int64_t ld, double* d_c, double* d_x, int* rank_out) {
    LSO_TRY(tri_solve(ctx, n, d_R, ld, d_c, d_x, 0));
    *rank_out = (int)n;
    return LSO_OK;
}

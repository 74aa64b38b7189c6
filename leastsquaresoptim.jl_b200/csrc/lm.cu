// lm.cu — (f1) the scalar bookkeeping of one trust-region iteration with ONE host synchronisation.
// The reference's loop (levenberg_marquardt.jl:102-124, dogleg.jl:160-178) reads four reductions per iteration, each
// of which is a host round trip when bound one by one (sum(abs2, ftrial), sum(abs2, fpredict), maximum(abs, δx),
// maxabs_projected_gradient).  Here they are reduced into adjacent device scalars by kernels enqueued back to back,
// all-reduced over the ranks when J is row-sharded (ONE ncclAllReduce of the two m-dimension sums), and read back
// together.  The decisions themselves (ρ, Δ update, convergence) stay in the host loop, as in the reference.
#include "csc.cuh"

int lso_dev_sumabs2(lso_ctx* ctx, int64_t n, const double* x, double* d_out);
int lso_dev_maxabs(lso_ctx* ctx, int64_t n, const double* x, double* d_out);
int lso_dev_maxabs_projected(lso_ctx* ctx, int64_t n, const double* g, const double* x, const double* lo, const double* hi,
                             double* d_out);
int dense_predicted_ssr_dev(lso_ctx* ctx, int64_t m, int64_t n, const double* d_J, int64_t ld, const double* d_delta,
                            const double* d_f, double* d_fpredict, double* d_out);
int csc_predicted_ssr_dev(lso_csc* A, const double* d_delta, const double* d_f, double* d_fpredict, double* d_out);
extern "C" int lso_comm_allreduce_sum(lso_ctx* ctx, double* d_buf, int64_t count);

#define LM_SLOT 48      /* ctx->d_scalars[48..51] = trial ssr, predicted ssr, maxabs(δx), maxabs projected gradient */

extern "C" {

// maxabs_projected_gradient(g, x, lower, upper) (utils.jl:38-55) into the step's scalar block; no synchronisation.
// Call it where the reference computes it (LM:104, dogleg:101: before x is updated).
int lso_lm_gradient_norm_async(lso_ctx* ctx, int64_t n, const double* d_g, const double* d_x, const double* d_lower,
                               const double* d_upper) {
    LSO_REQUIRE(ctx, ctx && d_g && (d_x || (!d_lower && !d_upper)), "NULL pointer");
    LSO_ENTER(ctx);
    return lso_dev_maxabs_projected(ctx, n, d_g, d_x, d_lower, d_upper, ctx->d_scalars + LM_SLOT + 3);
}

// out[0] = sum(abs2, ftrial)                       LM:110 / dogleg:168     (summed over ranks when allreduce != 0)
// out[1] = sum(abs2, J δx - fcur) (fpredict)       LM:114-117 / dogleg:171-174   (summed over ranks when allreduce != 0)
// out[2] = maximum(abs, δx)                        utils.jl:21
// out[3] = the value left by lso_lm_gradient_norm_async
// Exactly one of (d_J, ld) or A_csc describes J (this rank's rows).  d_fpredict may be NULL.
int lso_lm_step_tail(lso_ctx* ctx, int64_t m, int64_t n, const double* d_J, int64_t ld, lso_csc* A_csc,
                     const double* d_dx, const double* d_fcur, const double* d_ftrial, double* d_fpredict, int allreduce,
                     double* out4) {
    LSO_REQUIRE(ctx, ctx && out4 && d_dx && d_fcur && d_ftrial, "NULL pointer");
    LSO_REQUIRE(ctx, (A_csc != nullptr) != (d_J != nullptr), "give exactly one of A_csc or d_J");
    LSO_ENTER(ctx);
    double* sc = ctx->d_scalars + LM_SLOT;
    LSO_TRY(lso_dev_sumabs2(ctx, m, d_ftrial, sc + 0));
    if (A_csc) {
        LSO_REQUIRE(ctx, A_csc->m == m && A_csc->n == n, "operator dimension mismatch");
        LSO_TRY(csc_predicted_ssr_dev(A_csc, d_dx, d_fcur, d_fpredict, sc + 1));
    } else {
        LSO_REQUIRE(ctx, ld >= m, "leading dimension < m");
        LSO_TRY(dense_predicted_ssr_dev(ctx, m, n, d_J, ld, d_dx, d_fcur, d_fpredict, sc + 1));
    }
    LSO_TRY(lso_dev_maxabs(ctx, n, d_dx, sc + 2));
    if (allreduce && ctx->nranks > 1) LSO_TRY(lso_comm_allreduce_sum(ctx, sc, 2));
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars + LM_SLOT, sc, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 4; ++i) out4[i] = ctx->h_scalars[LM_SLOT + i];
    return LSO_OK;
}

}  // extern "C"

// =====================================================================================================================
// (f2) device Jacobian producer: the `autodiff = :central` default of LeastSquaresProblem (types.jl:54-58,
// FiniteDiff.finite_difference_jacobian! with its default step: eps_j = cbrt(eps) * max(1, |x_j|)) for an f! that is a
// device callback — J never leaves the GPU.  Column j = (f(x + eps_j e_j) - f(x - eps_j e_j)) / (2 eps_j).
// =====================================================================================================================
__global__ void fd_perturb_kernel(double* __restrict__ x, long long j, double xj, double h) { x[j] = xj + h; }
__global__ void fd_column_kernel(long long m, const double* __restrict__ fp, const double* __restrict__ fm, double inv2h,
                                 double* __restrict__ col) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x)
        col[i] = (fp[i] - fm[i]) * inv2h;
}

extern "C" int lso_fd_jacobian_central(lso_ctx* ctx, int64_t m, int64_t n, lso_residual_fn f, void* user, double* d_x,
                                       double* d_J, int64_t ld, double* d_work /* 2 m */) {
    LSO_REQUIRE(ctx, ctx && f && d_x && d_J && d_work, "NULL pointer");
    LSO_REQUIRE(ctx, m >= 1 && n >= 1 && ld >= m, "bad dimensions");
    LSO_ENTER(ctx);
    std::vector<double> hx((size_t)n);
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(hx.data(), d_x, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double relstep = cbrt(2.220446049250313e-16);
    double* fp = d_work;
    double* fm = d_work + m;
    const unsigned grid = (unsigned)std::min<int64_t>(cdiv64(m, 256), (int64_t)ctx->num_sms * 8);
    for (int64_t j = 0; j < n; ++j) {
        const double xj = hx[(size_t)j];
        const double h = relstep * fmax(1.0, fabs(xj));
        fd_perturb_kernel<<<1, 1, 0, ctx->stream>>>(d_x, j, xj, h);
        int st = f(user, d_x, fp);
        if (st != 0) return lso_set_error(ctx, LSO_ERR_ARG, "residual callback failed with status %d", st);
        fd_perturb_kernel<<<1, 1, 0, ctx->stream>>>(d_x, j, xj, -h);
        st = f(user, d_x, fm);
        if (st != 0) return lso_set_error(ctx, LSO_ERR_ARG, "residual callback failed with status %d", st);
        fd_perturb_kernel<<<1, 1, 0, ctx->stream>>>(d_x, j, xj, 0.0);
        fd_column_kernel<<<grid, 256, 0, ctx->stream>>>(m, fp, fm, 1.0 / (2.0 * h), d_J + (size_t)j * (size_t)ld);
        ctx->launches += 4;
    }
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

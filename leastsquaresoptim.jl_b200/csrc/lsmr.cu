// lsmr.cu — LSMR (Fong & Saunders) Golub-Kahan bidiagonalisation on the device, as driven by
//   ldiv!(x, J, y, A::LSMRAllocatedSolver)               src/solver/iterative_lsmr.jl:179-198
//   ldiv!(x, J, y, damp, A::LSMRDampenedAllocatedSolver) src/solver/iterative_lsmr.jl:238-259
// through the operator wrappers PreconditionedMatrix (:12-51), DampenedMatrix/DampenedVector (:61-109),
// InverseDiagonal (:117-122) and the default preconditioner (:129-141); iteration: src/utils/lsmr.jl:53-238.
// Vectors stay in HBM; the ~25 scalar recurrences per iteration run on the host from three scalar
// read-backs (β, α, ‖x‖), exactly the quantities the reference's `norm` calls produce.
#include "csc.cuh"
#include <math.h>

int lso_dev_sumabs2(lso_ctx* ctx, int64_t n, const double* x, double* d_out);

struct lso_lsmr_ws {
    lso_ctx* ctx = nullptr;
    int64_t m = 0, n = 0;
    int damped = 0;
    double *P = nullptr, *tmp = nullptr, *tmp2 = nullptr, *v = nullptr, *h = nullptr, *hbar = nullptr,
           *zerosvector = nullptr, *u = nullptr;
};

// z[i] = z[i] + (alpha * x[i]) * y[i]      (map!((z,x,y) -> z + α*x*y, ...) iterative_lsmr.jl:92,107)
__global__ void addmul_kernel(int64_t n, double* __restrict__ z, double alpha, const double* __restrict__ x,
                              const double* __restrict__ y) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        z[i] = __dadd_rn(z[i], __dmul_rn(__dmul_rn(alpha, x[i]), y[i]));
}
// P = s > 0 ? 1/sqrt(s) : 0 with s = cs (+ damp)     (iterative_lsmr.jl:131-137)
__global__ void precond_kernel(int64_t n, double* __restrict__ P, const double* __restrict__ damp) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = P[i];
        if (damp) s = s + damp[i];
        P[i] = (s > 0.0) ? 1.0 / sqrt(s) : 0.0;
    }
}
// lsmr.jl:152-156:  hbar = c1*hbar + h ; x += c2*hbar ; h = c3*h + v
__global__ void lsmr_update_kernel(int64_t n, double c1, double c2, double c3, double* __restrict__ h,
                                   double* __restrict__ hbar, double* __restrict__ x, const double* __restrict__ v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double hi = h[i];
        const double hb = __dadd_rn(__dmul_rn(hbar[i], c1), hi);
        hbar[i] = hb;
        x[i] = fma(c2, hb, x[i]);
        h[i] = __dadd_rn(__dmul_rn(hi, c3), v[i]);
    }
}

static inline int grid_for(lso_ctx* ctx, int64_t n) {
    int64_t g = cdiv64(n, 256), cap = (int64_t)ctx->num_sms * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

struct Op {
    lso_ctx* ctx;
    lso_csc* csc;
    const double* J;
    int64_t ld, m, n;
    const double* diag;   // sqrt(damp) (DampenedMatrix.diagonal) or NULL
    const double* P;      // InverseDiagonal._
    double *tmp, *tmp2;
};

static int inner_mul_n(Op& A, double alpha, const double* x, double beta, double* y) {
    if (A.csc) return lso_csc_mul_n(A.csc, alpha, x, beta, y);
    return lso_dense_gemv_n(A.ctx, A.m, A.n, alpha, A.J, A.ld, x, beta, y);
}
static int inner_mul_t(Op& A, double alpha, const double* y, double beta, double* x) {
    if (A.csc) return lso_csc_mul_t(A.csc, alpha, y, beta, x);
    return lso_dense_gemv_t(A.ctx, A.m, A.n, alpha, A.J, A.ld, y, beta, x);
}

// b <- α A a + β b   (PreconditionedMatrix mul!, iterative_lsmr.jl:30-34, over DampenedMatrix :87-94 or J)
static int op_mul(Op& A, double alpha, const double* a, double beta, double* by, double* bx) {
    lso_ctx* ctx = A.ctx;
    LSO_TRY(lso_vec_mul(ctx, A.n, A.tmp, a, A.P));
    if (A.diag) {
        if (beta != 1.0) {
            LSO_TRY(lso_vec_scal(ctx, A.m, by, beta));
            LSO_TRY(lso_vec_scal(ctx, A.n, bx, beta));
        }
        LSO_TRY(inner_mul_n(A, alpha, A.tmp, 1.0, by));
        addmul_kernel<<<grid_for(ctx, A.n), 256, 0, ctx->stream>>>(A.n, bx, alpha, A.tmp, A.diag);
        LSO_CHECK_LAUNCH(ctx);
        return LSO_OK;
    }
    return inner_mul_n(A, alpha, A.tmp, beta, by);
}

// b <- α A' a + β b   (adjoint mul!, iterative_lsmr.jl:36-51 over :95-109)
static int op_mul_t(Op& A, double alpha, const double* ay, const double* ax, double beta, double* b) {
    lso_ctx* ctx = A.ctx;
    LSO_TRY(inner_mul_t(A, 1.0, ay, 0.0, A.tmp));
    if (A.diag) {
        addmul_kernel<<<grid_for(ctx, A.n), 256, 0, ctx->stream>>>(A.n, A.tmp, 1.0, ax, A.diag);
        LSO_CHECK_LAUNCH(ctx);
    }
    LSO_TRY(lso_vec_mul(ctx, A.n, A.tmp2, A.tmp, A.P));
    if (beta != 1.0) {
        if (beta == 0.0) LSO_TRY(lso_vec_fill(ctx, A.n, b, 0.0));
        else LSO_TRY(lso_vec_scal(ctx, A.n, b, beta));
    }
    return lso_vec_axpy(ctx, A.n, alpha, A.tmp2, b);
}

// norm(u) for a plain vector, sqrt(norm(y)^2 + norm(x)^2) for a DampenedVector (iterative_lsmr.jl:72)
static int split_norm(lso_ctx* ctx, int64_t m, const double* by, int64_t n, const double* bx, double* out) {
    LSO_TRY(lso_dev_sumabs2(ctx, m, by, ctx->d_scalars + 16));
    if (bx) LSO_TRY(lso_dev_sumabs2(ctx, n, bx, ctx->d_scalars + 17));
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars + 16, ctx->d_scalars + 16, 2 * sizeof(double), cudaMemcpyDeviceToHost,
                                        ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double ny = sqrt(ctx->h_scalars[16]);
    if (!bx) { *out = ny; return LSO_OK; }
    const double nx = sqrt(ctx->h_scalars[17]);
    *out = sqrt(ny * ny + nx * nx);
    return LSO_OK;
}

static inline double sq(double a) { return a * a; }

static int lsmr_run(Op& A, double* x, double* by, double* bx, double* v, double* h, double* hbar, double atol,
                    double btol, double conlim, int64_t maxiter, int64_t* iters_out, int* istop_out) {
    lso_ctx* ctx = A.ctx;
    const int64_t m = A.m, n = A.n;
    const double lambda = 0.0;                                 // damping lives in the augmented operator
    const double ctol = conlim > 0 ? 1.0 / conlim : 0.0;
    // u = b - A x ; β = ‖u‖ ; v = A'u ; α = ‖v‖            (lsmr.jl:73-78)
    LSO_TRY(op_mul(A, -1.0, x, 1.0, by, bx));
    double beta = 0, alpha = 0;
    LSO_TRY(split_norm(ctx, m, by, n, bx, &beta));
    if (beta > 0) {
        LSO_TRY(lso_vec_scal(ctx, m, by, 1.0 / beta));
        if (bx) LSO_TRY(lso_vec_scal(ctx, n, bx, 1.0 / beta));
    }
    LSO_TRY(op_mul_t(A, 1.0, by, bx, 0.0, v));
    LSO_TRY(lso_vec_nrm2(ctx, n, v, &alpha));
    if (alpha > 0) LSO_TRY(lso_vec_scal(ctx, n, v, 1.0 / alpha));

    double zetabar = alpha * beta, alphabar = alpha, rho = 1, rhobar = 1, cbar = 1, sbar = 0;
    LSO_TRY(lso_vec_copy(ctx, n, h, v));
    LSO_TRY(lso_vec_fill(ctx, n, hbar, 0.0));
    double betadd = beta, betad = 0, rhodold = 1, tautildeold = 0, thetatilde = 0, zeta = 0, d = 0;
    double normA2 = sq(alpha), maxrbar = 0, minrbar = 1e100;
    const double normb = beta;
    int istop = 0;
    double normr = beta, normAr = alpha * beta;
    int64_t iter = 0;
    if (normAr != 0) {
        while (iter < maxiter) {
            ++iter;
            LSO_TRY(op_mul(A, 1.0, v, -alpha, by, bx));                     // lsmr.jl:118
            LSO_TRY(split_norm(ctx, m, by, n, bx, &beta));
            if (beta > 0) {
                LSO_TRY(lso_vec_scal(ctx, m, by, 1.0 / beta));
                if (bx) LSO_TRY(lso_vec_scal(ctx, n, bx, 1.0 / beta));
                LSO_TRY(op_mul_t(A, 1.0, by, bx, -beta, v));                // lsmr.jl:122
                LSO_TRY(lso_vec_nrm2(ctx, n, v, &alpha));
                if (alpha > 0) LSO_TRY(lso_vec_scal(ctx, n, v, 1.0 / alpha));
            }
            // rotation Qhat_{k,2k+1}
            const double alphahat = sqrt(sq(alphabar) + sq(lambda));
            const double chat = alphabar / alphahat, shat = lambda / alphahat;
            // rotation Q_i turning B_i into R_i
            const double rhoold = rho;
            rho = sqrt(sq(alphahat) + sq(beta));
            const double c = alphahat / rho, s = beta / rho;
            const double thetanew = s * alpha;
            alphabar = c * alpha;
            // rotation Qbar_i turning R_i' into R_i^bar
            const double rhobarold = rhobar, zetaold = zeta;
            const double thetabar = sbar * rho, rhotemp = cbar * rho;
            rhobar = sqrt(sq(cbar * rho) + sq(thetanew));
            cbar = cbar * rho / rhobar;
            sbar = thetanew / rhobar;
            zeta = cbar * zetabar;
            zetabar = -sbar * zetabar;
            // h, hbar, x
            lsmr_update_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(n, -thetabar * rho / (rhoold * rhobarold),
                                                                         zeta / (rho * rhobar), -thetanew / rho, h, hbar, x, v);
            LSO_CHECK_LAUNCH(ctx);
            // estimate of ‖r‖
            const double betaacute = chat * betadd, betacheck = -shat * betadd;
            const double betahat = c * betaacute;
            betadd = -s * betaacute;
            const double thetatildeold = thetatilde;
            const double rhotildeold = sqrt(sq(rhodold) + sq(thetabar));
            const double ctildeold = rhodold / rhotildeold, stildeold = thetabar / rhotildeold;
            thetatilde = stildeold * rhobar;
            rhodold = ctildeold * rhobar;
            betad = -stildeold * betad + ctildeold * betahat;
            tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold;
            const double taud = (zeta - thetatilde * tautildeold) / rhodold;
            d = d + sq(betacheck);
            normr = sqrt(d + sq(betad - taud) + sq(betadd));
            // estimate of ‖A‖ and cond(A)
            normA2 = normA2 + sq(beta);
            const double normA = sqrt(normA2);
            normA2 = normA2 + sq(alpha);
            maxrbar = fmax(maxrbar, rhobarold);
            if (iter > 1) minrbar = fmin(minrbar, rhobarold);
            const double condA = fmax(maxrbar, rhotemp) / fmin(minrbar, rhotemp);
            // stopping tests (lsmr.jl:202-231)
            normAr = fabs(zetabar);
            double normx = 0;
            LSO_TRY(lso_vec_nrm2(ctx, n, x, &normx));
            const double test1 = normr / normb;
            const double test2 = normAr / (normA * normr);
            const double test3 = 1.0 / condA;
            const double t1 = test1 / (1.0 + normA * normx / normb);
            const double rtol = btol + atol * normA * normx / normb;
            if (iter >= maxiter) { istop = 7; break; }
            if (1.0 + test3 <= 1.0) { istop = 6; break; }
            if (1.0 + test2 <= 1.0) { istop = 5; break; }
            if (1.0 + t1 <= 1.0) { istop = 4; break; }
            if (test3 <= ctol) { istop = 3; break; }
            if (test2 <= atol) { istop = 2; break; }
            if (test1 <= rtol) { istop = 1; break; }
        }
    }
    *iters_out = iter;
    *istop_out = istop;
    return LSO_OK;
}

extern "C" {

int lso_lsmr_ws_create(lso_ctx* ctx, int64_t m, int64_t n, int damped, lso_lsmr_ws** out) {
    LSO_REQUIRE(ctx, ctx && out, "ctx/out is NULL");
    *out = nullptr;
    LSO_REQUIRE(ctx, m >= 1 && n >= 1, "m and n must be positive");
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    lso_lsmr_ws* ws = new (std::nothrow) lso_lsmr_ws();
    if (!ws) return lso_set_error(ctx, LSO_ERR_ALLOC, "host allocation failed");
    ws->ctx = ctx; ws->m = m; ws->n = n; ws->damped = damped;
    double** nv[] = {&ws->P, &ws->tmp, &ws->tmp2, &ws->v, &ws->h, &ws->hbar, &ws->zerosvector};
    for (auto p : nv) {
        cudaError_t e = cudaMalloc(p, n * sizeof(double));
        if (e != cudaSuccess) { cudaGetLastError(); lso_lsmr_ws_destroy(ws); return lso_set_error(ctx, LSO_ERR_ALLOC, "LSMR workspace: %s", cudaGetErrorString(e)); }
        cudaMemsetAsync(*p, 0, n * sizeof(double), ctx->stream);
    }
    cudaError_t e = cudaMalloc(&ws->u, m * sizeof(double));
    if (e != cudaSuccess) { cudaGetLastError(); lso_lsmr_ws_destroy(ws); return lso_set_error(ctx, LSO_ERR_ALLOC, "LSMR workspace: %s", cudaGetErrorString(e)); }
    *out = ws;
    return LSO_OK;
}

int lso_lsmr_ws_destroy(lso_lsmr_ws* ws) {
    if (!ws) return LSO_OK;
    cudaSetDevice(ws->ctx->device);
    cudaStreamSynchronize(ws->ctx->stream);
    cudaFree(ws->P); cudaFree(ws->tmp); cudaFree(ws->tmp2); cudaFree(ws->v); cudaFree(ws->h); cudaFree(ws->hbar);
    cudaFree(ws->zerosvector); cudaFree(ws->u);
    delete ws;
    return LSO_OK;
}

int lso_lsmr_solve(lso_lsmr_ws* ws, lso_csc* A_csc, const double* d_J, int64_t ld, const double* d_y, double* d_damp,
                   double* d_x, double atol, double btol, double conlim, int64_t maxiter, int64_t* iters_out,
                   int* istop_out) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    const int64_t m = ws->m, n = ws->n;
    LSO_REQUIRE(ctx, (A_csc != nullptr) != (d_J != nullptr), "give exactly one of A_csc or d_J");
    LSO_REQUIRE(ctx, d_y && d_x && iters_out && istop_out, "NULL pointer");
    if (A_csc) LSO_REQUIRE(ctx, A_csc->m == m && A_csc->n == n, "operator / workspace dimension mismatch");
    else LSO_REQUIRE(ctx, ld >= m, "leading dimension < m");
    LSO_TRY(lso_vec_fill(ctx, n, d_x, 0.0));                    // fill!(x, 0)
    LSO_TRY(lso_vec_copy(ctx, m, ws->u, d_y));                  // copyto!(u, y)
    if (d_damp) LSO_TRY(lso_vec_fill(ctx, n, ws->zerosvector, 0.0));
    LSO_TRY(lso_vec_fill(ctx, n, ws->tmp, 0.0));
    // preconditioner!(P, x, J, damp)
    if (A_csc) LSO_TRY(lso_csc_colsumabs2(A_csc, ws->P));
    else LSO_TRY(lso_dense_colsumabs2(ctx, m, n, d_J, ld, ws->P));
    precond_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(n, ws->P, d_damp);
    LSO_CHECK_LAUNCH(ctx);
    if (d_damp) LSO_TRY(lso_vec_sqrt(ctx, n, d_damp));          // map!(sqrt, damp, damp)
    Op A{ctx, A_csc, d_J, ld, m, n, d_damp, ws->P, ws->tmp, ws->tmp2};
    if (maxiter <= 0) maxiter = d_damp ? std::max<int64_t>(m + n, n) : std::max<int64_t>(m, n);
    LSO_TRY(lsmr_run(A, d_x, ws->u, d_damp ? ws->zerosvector : nullptr, ws->v, ws->h, ws->hbar, atol, btol, conlim,
                     maxiter, iters_out, istop_out));
    LSO_TRY(lso_vec_mul(ctx, n, ws->tmp, d_x, ws->P));           // ldiv!(tmp, P, x)
    return lso_vec_copy(ctx, n, d_x, ws->tmp);                   // copyto!(x, tmp)
}

}  // extern "C"

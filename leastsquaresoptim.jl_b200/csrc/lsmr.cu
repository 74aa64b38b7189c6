// lsmr.cu — LSMR (Fong & Saunders) Golub-Kahan bidiagonalisation on the device, as driven by
//   ldiv!(x, J, y, A::LSMRAllocatedSolver)               src/solver/iterative_lsmr.jl:179-198
//   ldiv!(x, J, y, damp, A::LSMRDampenedAllocatedSolver) src/solver/iterative_lsmr.jl:238-259
// through the operator wrappers PreconditionedMatrix (:12-51), DampenedMatrix/DampenedVector (:61-109),
// InverseDiagonal (:117-122) and the default preconditioner (:129-141); iteration: src/utils/lsmr.jl:53-238.
//
// Two drivers:
//  * FUSED (sparse CSC operator with a diagonal preconditioner — BASELINE configs[2]): three launches per iteration,
//    no host synchronisation inside an iteration.
//      K_fwd  u <- A v - α u      stream SpMV on the CSR mirror; the wrapper algebra (P∘v precomputed as `tmp`, the damping
//                                 rows diag∘tmp, the -α u term, the lazily applied 1/β of the previous normalisation) lives in
//                                 its epilogue, ‖u‖² is reduced in the same launch and the last CTA stores β, 1/β
//      K_adj  v <- A'u - β v      stream SpMᵀV on the CSC image gathering u/β on the fly; epilogue P∘(J'u + diag∘u.x) - β v,
//                                 ‖v‖²; the last CTA stores α and runs ALL the scalar recurrences of lsmr.jl:128-196
//      K_upd  v/α, h̄, x, h, tmp = P∘v, ‖x‖²; the last CTA evaluates the seven stopping tests (lsmr.jl:202-231)
//    The iteration state is a device struct; every kernel is a no-op once `done` is set, so the host enqueues
//    iterations in batches and reads {iter, istop, done} back once per batch (≤ 1 sync per iteration).
//  * GENERIC (dense J, or a user preconditioner given as a callback): the reference's wrappers op for op, scalars on
//    the host from three read-backs per iteration.
#include "csc.cuh"
#include <math.h>

int lso_dev_sumabs2(lso_ctx* ctx, int64_t n, const double* x, double* d_out);
int csc_colsumabs2_cached(lso_csc* A, double* d_out);

struct LsmrState {
    double alpha, beta, inv_alpha, inv_beta;          // inv_* : the scale rmul!(·, inv(norm)) would apply (1 when the norm is 0)
    double zetabar, alphabar, rho, rhobar, cbar, sbar;
    double betadd, betad, rhodold, tautildeold, thetatilde, zeta, d;
    double normA2, maxrbar, minrbar, normb, normr, normAr, normA, condA;
    double c1, c2, c3;                                // coefficients of this iteration's h̄ / x / h update
    double atol, btol, ctol;
    long long iter, maxiter;
    int istop, done, v_fresh, pad;                    // v_fresh: this iteration produced a new (unnormalised) v
    double ux_sumsq;                                  // row-sharded form: ‖b.x‖² of the (replicated) damping part of u
};
struct LsmrTail { long long iter; int istop, done; };

struct lso_lsmr_ws {
    lso_ctx* ctx = nullptr;
    int64_t m = 0, n = 0;
    int damped = 0;
    double *P = nullptr, *tmp = nullptr, *tmp2 = nullptr, *v = nullptr, *h = nullptr, *hbar = nullptr,
           *zerosvector = nullptr, *u = nullptr;
    LsmrState* d_state = nullptr;
    LsmrTail* h_tail = nullptr;      // pinned
    double* packed = nullptr;        // row-sharded form: [ J_k'u_k (n) | ‖u_k‖² ], all-reduced once per iteration
    int64_t last_syncs = 0, last_launches = 0;
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

// =====================================================================================================================
// fused driver
// =====================================================================================================================
__device__ __forceinline__ double sqd(double a) { return a * a; }

// β = ‖b‖ for x = 0: u = b - A·0 = b exactly for a finite operator (lsmr.jl:73-75), so the first product is not formed
__global__ void lsmr_init_kernel(LsmrState* st, const double* __restrict__ sumsq_y, double atol, double btol, double ctol,
                                 long long maxiter) {
    const double ny = sqrt(*sumsq_y);
    const double beta = sqrt(ny * ny + 0.0);            // DampenedVector norm with x-part 0 (iterative_lsmr.jl:72)
    st->beta = beta;
    st->inv_beta = beta > 0.0 ? 1.0 / beta : 1.0;
    st->alpha = 0.0; st->inv_alpha = 1.0;
    st->atol = atol; st->btol = btol; st->ctol = ctol;
    st->iter = 0; st->maxiter = maxiter; st->istop = 0; st->done = 0; st->v_fresh = 0;
}

// u <- A v - α u  (lsmr.jl:118 through iterative_lsmr.jl:30-34 and :87-94), segments = rows of the CSR mirror
struct LsmrFwd {
    LsmrState* st;
    const double* tmp;          // P∘v  (ldiv!(tmp, P, v), written by the update kernel)
    double* uy;                 // b.y  (unnormalised; 1/β applied on read)
    double* ux;                 // b.x  or NULL (undamped operator)
    const double* diag;         // sqrt(damp)
    long long n_extra;          // n when damped, else 0
    double alpha_l, inv_beta_prev;
    static constexpr bool DUAL = false;
    __device__ __forceinline__ bool begin() {
        if (st->done) return false;
        alpha_l = st->alpha;
        inv_beta_prev = st->inv_beta;
        return true;
    }
    __device__ __forceinline__ bool idle() const { return false; }
    __device__ __forceinline__ double gather(int c) const { return __ldg(tmp + c); }
    __device__ __forceinline__ double epilogue(long long i, double a) const {
        const double uold = uy[i] * inv_beta_prev;              // rmul!(u, inv(β)) of the previous iteration
        const double unew = fma(-alpha_l, uold, a);             // rmul!(b, -α); b.y += J tmp
        uy[i] = unew;
        return unew * unew;
    }
    __device__ __forceinline__ double epilogue2(long long, double, double) const { return 0.0; }
    __device__ __forceinline__ double extra(long long j) const {
        const double xold = ux[j] * inv_beta_prev;
        const double xnew = __dadd_rn(__dmul_rn(-alpha_l, xold), __dmul_rn(tmp[j], diag[j]));   // z + 1*x*y
        ux[j] = xnew;
        return xnew * xnew;
    }
    __device__ __forceinline__ void finish(double sy, double sx) const {
        const double ny = sqrt(sy), nx = sqrt(sx);
        const double beta = n_extra ? sqrt(ny * ny + nx * nx) : ny;    // iterative_lsmr.jl:72 / plain norm
        st->beta = beta;
        st->inv_beta = beta > 0.0 ? 1.0 / beta : 1.0;
    }
};

// α = ‖v‖ (unless the step was skipped), then lsmr.jl:80-113 (init) or one iteration of the rotations / norm estimates
// (lsmr.jl:128-200); runs in ONE thread (the last CTA to retire of the kernel that produced ‖v‖²)
__device__ void lsmr_recurrences(LsmrState& s, double sv, bool skip, int init) {
    double alpha = s.alpha;
    if (!skip) {
        alpha = sqrt(sv);
        s.alpha = alpha;
        s.inv_alpha = alpha > 0.0 ? 1.0 / alpha : 1.0;
        s.v_fresh = 1;
    } else {
        s.v_fresh = 0;
    }
    const double beta = s.beta;
    if (init) {                                      // lsmr.jl:80-113
        s.zetabar = alpha * beta; s.alphabar = alpha; s.rho = 1.0; s.rhobar = 1.0; s.cbar = 1.0; s.sbar = 0.0;
        s.betadd = beta; s.betad = 0.0; s.rhodold = 1.0; s.tautildeold = 0.0; s.thetatilde = 0.0; s.zeta = 0.0; s.d = 0.0;
        s.normA2 = alpha * alpha; s.maxrbar = 0.0; s.minrbar = 1e100;
        s.normb = beta; s.normr = beta; s.normAr = alpha * beta;
        s.normA = -1.0; s.condA = -1.0;
        s.c1 = 0.0; s.c2 = 0.0; s.c3 = 0.0;
        if (!(s.normAr != 0.0)) s.done = 1;          // lsmr.jl:115: exit if b = 0 or A'b = 0
        return;
    }
    s.iter += 1;
    const double lambda = 0.0;                       // damping lives in the augmented operator
    // rotation Qhat_{k,2k+1}
    const double alphahat = sqrt(sqd(s.alphabar) + sqd(lambda));
    const double chat = s.alphabar / alphahat, shat = lambda / alphahat;
    // rotation Q_i turning B_i into R_i
    const double rhoold = s.rho;
    const double rho = sqrt(sqd(alphahat) + sqd(beta));
    const double c = alphahat / rho, sn = beta / rho;
    const double thetanew = sn * alpha;
    s.alphabar = c * alpha;
    // rotation Qbar_i turning R_i' into R_i^bar
    const double rhobarold = s.rhobar, zetaold = s.zeta;
    const double thetabar = s.sbar * rho, rhotemp = s.cbar * rho;
    const double rhobar = sqrt(sqd(s.cbar * rho) + sqd(thetanew));
    s.cbar = s.cbar * rho / rhobar;
    s.sbar = thetanew / rhobar;
    const double zeta = s.cbar * s.zetabar;
    s.zetabar = -s.sbar * s.zetabar;
    s.rho = rho; s.rhobar = rhobar; s.zeta = zeta;
    s.c1 = -thetabar * rho / (rhoold * rhobarold);
    s.c2 = zeta / (rho * rhobar);
    s.c3 = -thetanew / rho;
    // estimate of ‖r‖
    const double betaacute = chat * s.betadd, betacheck = -shat * s.betadd;
    const double betahat = c * betaacute;
    s.betadd = -sn * betaacute;
    const double thetatildeold = s.thetatilde;
    const double rhotildeold = sqrt(sqd(s.rhodold) + sqd(thetabar));
    const double ctildeold = s.rhodold / rhotildeold, stildeold = thetabar / rhotildeold;
    s.thetatilde = stildeold * rhobar;
    s.rhodold = ctildeold * rhobar;
    s.betad = -stildeold * s.betad + ctildeold * betahat;
    s.tautildeold = (zetaold - thetatildeold * s.tautildeold) / rhotildeold;
    const double taud = (zeta - s.thetatilde * s.tautildeold) / s.rhodold;
    s.d = s.d + sqd(betacheck);
    s.normr = sqrt(s.d + sqd(s.betad - taud) + sqd(s.betadd));
    // estimate of ‖A‖ and cond(A)
    s.normA2 = s.normA2 + sqd(beta);
    s.normA = sqrt(s.normA2);
    s.normA2 = s.normA2 + sqd(alpha);
    s.maxrbar = fmax(s.maxrbar, rhobarold);
    if (s.iter > 1) s.minrbar = fmin(s.minrbar, rhobarold);
    s.condA = fmax(s.maxrbar, rhotemp) / fmin(s.minrbar, rhotemp);
    s.normAr = fabs(s.zetabar);
}


// v <- A'u - β v  (lsmr.jl:76,122 through iterative_lsmr.jl:36-51 and :95-109), segments = columns of the CSC image
struct LsmrAdj {
    LsmrState* st;
    const double* uy;
    const double* ux;           // NULL: undamped, or the initial product (b.x = 0)
    const double* diag;
    const double* P;
    double* v;
    int init;                   // 1: v = A'u (β' = 0), then initialise the recurrences (lsmr.jl:76-113)
    long long n_extra;
    double beta_l, inv_beta;
    bool skip;
    static constexpr bool DUAL = false;
    __device__ __forceinline__ bool begin() {
        if (st->done) return false;
        beta_l = st->beta;
        inv_beta = st->inv_beta;
        skip = !init && !(beta_l > 0.0);                           // lsmr.jl:120: v and α are kept when β == 0
        return true;
    }
    __device__ __forceinline__ bool idle() const { return skip; }
    __device__ __forceinline__ double gather(int r) const { return __ldg(uy + r) * inv_beta; }
    __device__ __forceinline__ double epilogue(long long j, double a) const {
        double t = a;
        if (ux) t = __dadd_rn(t, __dmul_rn(ux[j] * inv_beta, diag[j]));      // tmp += 1 * a.x * diag
        const double t2 = t * P[j];                                           // ldiv!(tmp2, P, tmp)
        const double vnew = init ? t2 : fma(-beta_l, v[j], t2);               // rmul!(v, -β); axpy!(1, tmp2, v)
        v[j] = vnew;
        return vnew * vnew;
    }
    __device__ __forceinline__ double epilogue2(long long, double, double) const { return 0.0; }
    __device__ __forceinline__ double extra(long long) const { return 0.0; }
    __device__ void finish(double sv, double) const { lsmr_recurrences(*st, sv, skip, init); }
};

// INIT: v /= α ; h = v ; hbar = 0 ; tmp = P∘v                                   (lsmr.jl:78, 89-90)
// else: v /= α (when a new v was formed) ; hbar, x, h updates (lsmr.jl:152-156) ; tmp = P∘v for the next product ;
//       ‖x‖² and, in the last CTA, the stopping tests (lsmr.jl:202-231)
template <bool INIT>
__global__ void __launch_bounds__(256)
lsmr_upd_kernel(LsmrState* st, long long n, double* __restrict__ v, double* __restrict__ h, double* __restrict__ hbar,
                double* __restrict__ x, const double* __restrict__ P, double* __restrict__ tmp,
                double* __restrict__ partials, unsigned int* __restrict__ counter) {
    __shared__ double red[32];
    __shared__ bool is_last;
    if (st->done) return;
    const bool scale = st->v_fresh && st->alpha > 0.0;
    const double inv_alpha = st->inv_alpha, c1 = st->c1, c2 = st->c2, c3 = st->c3;
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double vi = v[i];
        if (scale) { vi *= inv_alpha; v[i] = vi; }
        if (INIT) {
            h[i] = vi;
            hbar[i] = 0.0;
        } else {
            const double hi = h[i];
            const double hb = __dadd_rn(__dmul_rn(hbar[i], c1), hi);
            hbar[i] = hb;
            const double xi = fma(c2, hb, x[i]);
            x[i] = xi;
            h[i] = __dadd_rn(__dmul_rn(hi, c3), vi);
            acc = fma(xi, xi, acc);
        }
        tmp[i] = vi * P[i];
    }
    if (INIT) return;
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = acc;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double sx = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) sx += ((volatile double*)partials)[i];
    sx = block_sum(sx, red);
    if (threadIdx.x != 0) return;
    *counter = 0;
    LsmrState& s = *st;
    const double normx = sqrt(sx);
    const double test1 = s.normr / s.normb;
    const double test2 = s.normAr / (s.normA * s.normr);
    const double test3 = 1.0 / s.condA;
    const double t1 = test1 / (1.0 + s.normA * normx / s.normb);
    const double rtol = s.btol + s.atol * s.normA * normx / s.normb;
    int istop = 0;
    if (s.iter >= s.maxiter) istop = 7;
    else if (1.0 + test3 <= 1.0) istop = 6;
    else if (1.0 + test2 <= 1.0) istop = 5;
    else if (1.0 + t1 <= 1.0) istop = 4;
    else if (test3 <= s.ctol) istop = 3;
    else if (test2 <= s.atol) istop = 2;
    else if (test1 <= rtol) istop = 1;
    if (istop) { s.istop = istop; s.done = 1; }
}

__global__ void lsmr_tail_kernel(const LsmrState* st, LsmrTail* out) {
    out->iter = st->iter;
    out->istop = st->istop;
    out->done = st->done;
}

static int lsmr_fused_csc(lso_lsmr_ws* ws, lso_csc* A, const double* d_y, double* d_damp, double* d_x, const double* d_P_user,
                          double atol, double btol, double conlim, int64_t maxiter, int64_t* iters_out, int* istop_out) {
    lso_ctx* ctx = ws->ctx;
    const int64_t m = ws->m, n = ws->n;
    const int64_t launches0 = ctx->launches;
    ws->last_syncs = 0;
    LSO_CHECK_CUDA(ctx, cudaMemsetAsync(d_x, 0, n * sizeof(double), ctx->stream));                          // fill!(x, 0)
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(ws->u, d_y, m * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));   // copyto!(u, y)
    if (d_damp) LSO_CHECK_CUDA(ctx, cudaMemsetAsync(ws->zerosvector, 0, n * sizeof(double), ctx->stream));
    // preconditioner!(P, x, J, damp)  (default: iterative_lsmr.jl:130-138; user: README.md:47, P._ given by the caller)
    if (d_P_user) {
        LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(ws->P, d_P_user, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        LSO_TRY(csc_colsumabs2_cached(A, ws->P));
        precond_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(n, ws->P, d_damp);
        LSO_CHECK_LAUNCH(ctx);
    }
    if (d_damp) LSO_TRY(lso_vec_sqrt(ctx, n, d_damp));                                                     // map!(sqrt, damp, damp)
    LSO_TRY(csc_refresh_csr(A));
    if (maxiter <= 0) maxiter = d_damp ? std::max<int64_t>(m + n, n) : std::max<int64_t>(m, n);
    const double ctol = conlim > 0 ? 1.0 / conlim : 0.0;
    LsmrState* st = ws->d_state;
    LSO_TRY(lso_dev_sumabs2(ctx, m, ws->u, ctx->d_scalars + 16));
    lsmr_init_kernel<<<1, 1, 0, ctx->stream>>>(st, ctx->d_scalars + 16, atol, btol, ctol, (long long)maxiter);
    LSO_CHECK_LAUNCH(ctx);
    double* ux = d_damp ? ws->zerosvector : nullptr;
    unsigned int* cnt = ctx->d_counters + SP_COUNTER_SLOT;
    const int ugrid = (int)std::min<int64_t>(cdiv64(n, 256), (int64_t)ctx->num_sms * 8);
    {   // v = A'u ; α = ‖v‖ ; recurrences initialised ; v /= α, h = v, hbar = 0, tmp = P∘v
        LsmrAdj fa{st, ws->u, nullptr, d_damp, ws->P, ws->v, 1, 0, 0.0, 1.0, false};
        lso_prof_mark(ctx);
        LSO_TRY(spmv_stream_launch(ctx, A->Gc, fa, A->d_colptr, A->d_rowidx, A->d_val, A->d_cblk, A->ncblk, A->n));
        lso_prof_mark(ctx);
        lsmr_upd_kernel<true><<<ugrid, 256, 0, ctx->stream>>>(st, n, ws->v, ws->h, ws->hbar, d_x, ws->P, ws->tmp, ctx->d_partials, cnt);
        LSO_CHECK_LAUNCH(ctx);
    }
    LsmrFwd ff{st, ws->tmp, ws->u, ux, d_damp, d_damp ? n : 0, 0.0, 1.0};
    LsmrAdj fa{st, ws->u, ux, d_damp, ws->P, ws->v, 0, 0, 0.0, 1.0, false};
    int64_t enq = 0;
    int batch = 1;
    for (;;) {
        for (int b = 0; b < batch && enq < maxiter; ++b, ++enq) {
            lso_prof_mark(ctx);
            LSO_TRY(spmv_stream_launch(ctx, A->Gr, ff, A->d_rowptr, A->d_colidx, A->d_valr, A->d_rblk, A->nrblk, A->m));
            lso_prof_mark(ctx);
            lso_prof_mark(ctx);
            LSO_TRY(spmv_stream_launch(ctx, A->Gc, fa, A->d_colptr, A->d_rowidx, A->d_val, A->d_cblk, A->ncblk, A->n));
            lso_prof_mark(ctx);
            lsmr_upd_kernel<false><<<ugrid, 256, 0, ctx->stream>>>(st, n, ws->v, ws->h, ws->hbar, d_x, ws->P, ws->tmp, ctx->d_partials, cnt);
            LSO_CHECK_LAUNCH(ctx);
        }
        lsmr_tail_kernel<<<1, 1, 0, ctx->stream>>>(st, (LsmrTail*)(ctx->d_scalars + 32));
        LSO_CHECK_LAUNCH(ctx);
        LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(ws->h_tail, ctx->d_scalars + 32, sizeof(LsmrTail), cudaMemcpyDeviceToHost, ctx->stream));
        LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ws->last_syncs++;
        if (ws->h_tail->done || enq >= maxiter) break;
        if (enq >= 2 && batch < 8) batch *= 2;          // 1, 1, 2, 4, 8, 8, ... iterations per read-back
    }
    *iters_out = ws->h_tail->iter;
    *istop_out = ws->h_tail->istop;
    ctx->stat_spmv_bytes += (double)(2 * ws->h_tail->iter + 1) * (12.0 * (double)A->nnz + 8.0 * (double)(m + n));
    LSO_TRY(lso_vec_mul(ctx, n, d_x, d_x, ws->P));        // ldiv!(tmp, P, x); copyto!(x, tmp)
    ws->last_launches = ctx->launches - launches0;
    return LSO_OK;
}

// =====================================================================================================================
// row-sharded fused driver: every rank holds a block of rows of J (its own CSC + CSR images), u is
// sharded like the rows, the n-vectors v, h, h̄, x, P and the damping part of u are replicated.  Per iteration
//     u_k <- J_k (P∘v) - α u_k                 local rows; partial ‖u_k‖² into the packed buffer            (1 launch)
//     w_k  = J_k' u_k   (u NOT yet normalised)  partial A'u into the packed buffer                            (1 launch)
//     ONE ncclAllReduce of [ w (n) | ‖u‖² (1) ]
//     β = sqrt(Σ‖u_k‖² + ‖u.x‖²) ; v <- P∘(w/β + √D∘u.x/β) - β v ; α = ‖v‖ ; rotations (replicated)         (1 launch)
//     h̄, x, h updates, ‖x‖², stopping tests (replicated)                                                   (1 launch)
// so the exchange per iteration is n + 1 doubles and every rank takes the same decisions from the same numbers.
// =====================================================================================================================
struct LsmrFwdShard {
    LsmrState* st;
    const double* tmp;
    double* uy;
    double* ux;
    const double* diag;
    long long n_extra;
    double* packed_tail;        // packed[n]: this rank's ‖u_k‖²
    double alpha_l, inv_beta_prev;
    static constexpr bool DUAL = false;
    __device__ __forceinline__ bool begin() {
        if (st->done) return false;
        alpha_l = st->alpha;
        inv_beta_prev = st->inv_beta;
        return true;
    }
    __device__ __forceinline__ bool idle() const { return false; }
    __device__ __forceinline__ double gather(int c) const { return __ldg(tmp + c); }
    __device__ __forceinline__ double epilogue(long long i, double a) const {
        const double uold = uy[i] * inv_beta_prev;
        const double unew = fma(-alpha_l, uold, a);
        uy[i] = unew;
        return unew * unew;
    }
    __device__ __forceinline__ double epilogue2(long long, double, double) const { return 0.0; }
    __device__ __forceinline__ double extra(long long j) const {
        const double xold = ux[j] * inv_beta_prev;
        const double xnew = __dadd_rn(__dmul_rn(-alpha_l, xold), __dmul_rn(tmp[j], diag[j]));
        ux[j] = xnew;
        return xnew * xnew;
    }
    __device__ __forceinline__ void finish(double sy, double sx) const {
        *packed_tail = sy;              // summed over the ranks by the all-reduce
        st->ux_sumsq = sx;              // replicated: counted once
    }
};

struct LsmrAdjPart {
    LsmrState* st;
    const double* uy;
    double* packed;             // packed[j] = (J_k' u_k)[j]
    long long n_extra;
    static constexpr bool DUAL = false;
    __device__ __forceinline__ bool begin() { return !st->done; }
    __device__ __forceinline__ bool idle() const { return false; }
    __device__ __forceinline__ double gather(int r) const { return __ldg(uy + r); }
    __device__ __forceinline__ double epilogue(long long j, double a) const { packed[j] = a; return 0.0; }
    __device__ __forceinline__ double epilogue2(long long, double, double) const { return 0.0; }
    __device__ __forceinline__ double extra(long long) const { return 0.0; }
    __device__ __forceinline__ void finish(double, double) const {}
};

// after the all-reduce: β, the new v and α, then the recurrences in the last CTA to retire
__global__ void __launch_bounds__(256)
lsmr_vupd_shard_kernel(LsmrState* st, long long n, const double* __restrict__ packed, const double* __restrict__ ux,
                       const double* __restrict__ diag, const double* __restrict__ P, double* __restrict__ v, int init,
                       double* __restrict__ partials, unsigned int* __restrict__ counter) {
    __shared__ double red[32];
    __shared__ bool is_last;
    if (st->done) return;
    const double ny = sqrt(packed[n]);
    const double nx = (ux && !init) ? sqrt(st->ux_sumsq) : 0.0;
    const double beta = ux ? sqrt(ny * ny + nx * nx) : ny;             // iterative_lsmr.jl:72 / plain norm
    const double inv_beta = beta > 0.0 ? 1.0 / beta : 1.0;
    const bool skip = !init && !(beta > 0.0);                          // lsmr.jl:120
    double acc = 0.0;
    if (!skip) {
        for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
            double t = packed[j] * inv_beta;
            if (ux && !init) t = __dadd_rn(t, __dmul_rn(ux[j] * inv_beta, diag[j]));
            const double t2 = t * P[j];
            const double vnew = init ? t2 : fma(-beta, v[j], t2);
            v[j] = vnew;
            acc = fma(vnew, vnew, acc);
        }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = acc;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double sv = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) sv += ((volatile double*)partials)[i];
    sv = block_sum(sv, red);
    if (threadIdx.x != 0) return;
    *counter = 0;
    st->beta = beta;
    st->inv_beta = inv_beta;
    lsmr_recurrences(*st, sv, skip, init);
}

// st->beta etc. are only written by the v-update kernel; the init kernel of the sharded form just sets the controls
__global__ void lsmr_init_shard_kernel(LsmrState* st, double atol, double btol, double ctol, long long maxiter) {
    st->beta = 0.0; st->inv_beta = 1.0;              // u = y is read unscaled by the first forward product
    st->alpha = 0.0; st->inv_alpha = 1.0;
    st->atol = atol; st->btol = btol; st->ctol = ctol;
    st->iter = 0; st->maxiter = maxiter; st->istop = 0; st->done = 0; st->v_fresh = 0; st->ux_sumsq = 0.0;
}

extern "C" int lso_comm_allreduce_sum(lso_ctx* ctx, double* d_buf, int64_t count);

static int lsmr_fused_csc_sharded(lso_lsmr_ws* ws, lso_csc* A, const double* d_y, double* d_damp, double* d_x,
                                  const double* d_P_user, double atol, double btol, double conlim, int64_t maxiter,
                                  int64_t m_total, int64_t* iters_out, int* istop_out) {
    lso_ctx* ctx = ws->ctx;
    const int64_t m = ws->m, n = ws->n;
    const int64_t launches0 = ctx->launches;
    ws->last_syncs = 0;
    if (!ws->packed) LSO_CHECK_CUDA(ctx, cudaMalloc(&ws->packed, (size_t)(n + 1) * sizeof(double)));
    double* packed = ws->packed;
    LSO_CHECK_CUDA(ctx, cudaMemsetAsync(d_x, 0, n * sizeof(double), ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(ws->u, d_y, m * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    if (d_damp) LSO_CHECK_CUDA(ctx, cudaMemsetAsync(ws->zerosvector, 0, n * sizeof(double), ctx->stream));
    if (d_P_user) {
        LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(ws->P, d_P_user, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        LSO_TRY(csc_colsumabs2_cached(A, ws->P));                      // this rank's rows ...
        LSO_TRY(lso_comm_allreduce_sum(ctx, ws->P, n));                // ... summed: colsumabs2 of the whole J
        precond_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(n, ws->P, d_damp);
        LSO_CHECK_LAUNCH(ctx);
    }
    if (d_damp) LSO_TRY(lso_vec_sqrt(ctx, n, d_damp));
    LSO_TRY(csc_refresh_csr(A));
    if (m_total < m) m_total = m;
    if (maxiter <= 0) maxiter = d_damp ? std::max<int64_t>(m_total + n, n) : std::max<int64_t>(m_total, n);
    const double ctol = conlim > 0 ? 1.0 / conlim : 0.0;
    LsmrState* st = ws->d_state;
    lsmr_init_shard_kernel<<<1, 1, 0, ctx->stream>>>(st, atol, btol, ctol, (long long)maxiter);
    LSO_CHECK_LAUNCH(ctx);
    double* ux = d_damp ? ws->zerosvector : nullptr;
    unsigned int* cnt = ctx->d_counters + SP_COUNTER_SLOT;
    const int ugrid = (int)std::min<int64_t>(cdiv64(n, 256), (int64_t)ctx->num_sms * 8);
    LsmrAdjPart fp{st, ws->u, packed, 0};
    {   // β = ‖y‖ over all ranks, v = A'u / β scaled by P, α, recurrences; then v /= α, h = v, h̄ = 0, tmp = P∘v
        LSO_TRY(lso_dev_sumabs2(ctx, m, ws->u, packed + n));
        LSO_TRY(spmv_stream_launch(ctx, A->Gc, fp, A->d_colptr, A->d_rowidx, A->d_val, A->d_cblk, A->ncblk, A->n));
        LSO_TRY(lso_comm_allreduce_sum(ctx, packed, n + 1));
        lsmr_vupd_shard_kernel<<<ugrid, 256, 0, ctx->stream>>>(st, n, packed, ux, d_damp, ws->P, ws->v, 1, ctx->d_partials, cnt);
        LSO_CHECK_LAUNCH(ctx);
        lsmr_upd_kernel<true><<<ugrid, 256, 0, ctx->stream>>>(st, n, ws->v, ws->h, ws->hbar, d_x, ws->P, ws->tmp, ctx->d_partials, cnt);
        LSO_CHECK_LAUNCH(ctx);
    }
    LsmrFwdShard ff{st, ws->tmp, ws->u, ux, d_damp, d_damp ? n : 0, packed + n, 0.0, 1.0};
    int64_t enq = 0;
    int batch = 1;
    for (;;) {
        for (int b = 0; b < batch && enq < maxiter; ++b, ++enq) {
            lso_prof_mark(ctx);
            LSO_TRY(spmv_stream_launch(ctx, A->Gr, ff, A->d_rowptr, A->d_colidx, A->d_valr, A->d_rblk, A->nrblk, A->m));
            LSO_TRY(spmv_stream_launch(ctx, A->Gc, fp, A->d_colptr, A->d_rowidx, A->d_val, A->d_cblk, A->ncblk, A->n));
            lso_prof_mark(ctx);
            LSO_TRY(lso_comm_allreduce_sum(ctx, packed, n + 1));
            lsmr_vupd_shard_kernel<<<ugrid, 256, 0, ctx->stream>>>(st, n, packed, ux, d_damp, ws->P, ws->v, 0, ctx->d_partials, cnt);
            LSO_CHECK_LAUNCH(ctx);
            lsmr_upd_kernel<false><<<ugrid, 256, 0, ctx->stream>>>(st, n, ws->v, ws->h, ws->hbar, d_x, ws->P, ws->tmp, ctx->d_partials, cnt);
            LSO_CHECK_LAUNCH(ctx);
        }
        lsmr_tail_kernel<<<1, 1, 0, ctx->stream>>>(st, (LsmrTail*)(ctx->d_scalars + 32));
        LSO_CHECK_LAUNCH(ctx);
        LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(ws->h_tail, ctx->d_scalars + 32, sizeof(LsmrTail), cudaMemcpyDeviceToHost, ctx->stream));
        LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ws->last_syncs++;
        if (ws->h_tail->done || enq >= maxiter) break;
        if (enq >= 2 && batch < 8) batch *= 2;
    }
    *iters_out = ws->h_tail->iter;
    *istop_out = ws->h_tail->istop;
    ctx->stat_spmv_bytes += (double)(2 * ws->h_tail->iter + 1) * (12.0 * (double)A->nnz + 8.0 * (double)(m + n));
    LSO_TRY(lso_vec_mul(ctx, n, d_x, d_x, ws->P));
    ws->last_launches = ctx->launches - launches0;
    return LSO_OK;
}

// =====================================================================================================================
// generic driver (dense J, callback preconditioner)
// =====================================================================================================================
struct Op {
    lso_ctx* ctx;
    lso_csc* csc;
    const double* J;
    int64_t ld, m, n;
    const double* diag;   // sqrt(damp) (DampenedMatrix.diagonal) or NULL
    const double* P;      // InverseDiagonal._  (NULL when the preconditioner is a callback)
    lso_precond_fn pfn;   // ldiv!(out, P, in) as a callback on device pointers
    void* puser;
    double *tmp, *tmp2;
};

static int apply_P(Op& A, double* out, const double* in) {
    if (A.pfn) {
        const int st = A.pfn(A.puser, A.n, in, out);
        if (st != 0) return lso_set_error(A.ctx, LSO_ERR_ARG, "preconditioner callback failed with status %d", st);
        return LSO_OK;
    }
    return lso_vec_mul(A.ctx, A.n, out, in, A.P);
}
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
    LSO_TRY(apply_P(A, A.tmp, a));
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
    LSO_TRY(apply_P(A, A.tmp2, A.tmp));
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
    if (e == cudaSuccess) e = cudaMalloc(&ws->d_state, sizeof(LsmrState));
    if (e == cudaSuccess) e = cudaMallocHost(&ws->h_tail, sizeof(LsmrTail));
    if (e != cudaSuccess) { cudaGetLastError(); lso_lsmr_ws_destroy(ws); return lso_set_error(ctx, LSO_ERR_ALLOC, "LSMR workspace: %s", cudaGetErrorString(e)); }
    cudaMemsetAsync(ws->d_state, 0, sizeof(LsmrState), ctx->stream);
    *out = ws;
    return LSO_OK;
}

int lso_lsmr_ws_destroy(lso_lsmr_ws* ws) {
    if (!ws) return LSO_OK;
    cudaSetDevice(ws->ctx->device);
    cudaStreamSynchronize(ws->ctx->stream);
    cudaFree(ws->P); cudaFree(ws->tmp); cudaFree(ws->tmp2); cudaFree(ws->v); cudaFree(ws->h); cudaFree(ws->hbar);
    cudaFree(ws->zerosvector); cudaFree(ws->u); cudaFree(ws->d_state); cudaFree(ws->packed);
    if (ws->h_tail) cudaFreeHost(ws->h_tail);
    delete ws;
    return LSO_OK;
}

int lso_lsmr_solve_ex(lso_lsmr_ws* ws, lso_csc* A_csc, const double* d_J, int64_t ld, const double* d_y, double* d_damp,
                      double* d_x, double atol, double btol, double conlim, int64_t maxiter, const double* d_P_diag,
                      lso_precond_fn precond_apply, void* precond_user, int64_t* iters_out, int* istop_out) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    const int64_t m = ws->m, n = ws->n;
    LSO_REQUIRE(ctx, (A_csc != nullptr) != (d_J != nullptr), "give exactly one of A_csc or d_J");
    LSO_REQUIRE(ctx, d_y && d_x && iters_out && istop_out, "NULL pointer");
    LSO_REQUIRE(ctx, !(d_P_diag && precond_apply), "give the preconditioner as a diagonal OR as a callback");
    if (A_csc) LSO_REQUIRE(ctx, A_csc->m == m && A_csc->n == n, "operator / workspace dimension mismatch");
    else LSO_REQUIRE(ctx, ld >= m, "leading dimension < m");
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (A_csc && !precond_apply && ctx->opt_lsmr_fused)
        return lsmr_fused_csc(ws, A_csc, d_y, d_damp, d_x, d_P_diag, atol, btol, conlim, maxiter, iters_out, istop_out);
    const int64_t launches0 = ctx->launches;
    LSO_TRY(lso_vec_fill(ctx, n, d_x, 0.0));                    // fill!(x, 0)
    LSO_TRY(lso_vec_copy(ctx, m, ws->u, d_y));                  // copyto!(u, y)
    if (d_damp) LSO_TRY(lso_vec_fill(ctx, n, ws->zerosvector, 0.0));
    LSO_TRY(lso_vec_fill(ctx, n, ws->tmp, 0.0));
    // preconditioner!(P, x, J, damp): the default (iterative_lsmr.jl:130-138) unless the caller supplies P
    if (d_P_diag) {
        LSO_TRY(lso_vec_copy(ctx, n, ws->P, d_P_diag));
    } else if (!precond_apply) {
        if (A_csc) LSO_TRY(lso_csc_colsumabs2(A_csc, ws->P));
        else LSO_TRY(lso_dense_colsumabs2(ctx, m, n, d_J, ld, ws->P));
        precond_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(n, ws->P, d_damp);
        LSO_CHECK_LAUNCH(ctx);
    }
    if (d_damp) LSO_TRY(lso_vec_sqrt(ctx, n, d_damp));          // map!(sqrt, damp, damp)
    Op A{ctx, A_csc, d_J, ld, m, n, d_damp, precond_apply ? nullptr : ws->P, precond_apply, precond_user, ws->tmp, ws->tmp2};
    if (maxiter <= 0) maxiter = d_damp ? std::max<int64_t>(m + n, n) : std::max<int64_t>(m, n);
    LSO_TRY(lsmr_run(A, d_x, ws->u, d_damp ? ws->zerosvector : nullptr, ws->v, ws->h, ws->hbar, atol, btol, conlim,
                     maxiter, iters_out, istop_out));
    LSO_TRY(apply_P(A, ws->tmp, d_x));                           // ldiv!(tmp, P, x)
    LSO_TRY(lso_vec_copy(ctx, n, d_x, ws->tmp));                 // copyto!(x, tmp)
    ws->last_syncs = 3 * (*iters_out) + 2;
    ws->last_launches = ctx->launches - launches0;
    return LSO_OK;
}

int lso_lsmr_solve(lso_lsmr_ws* ws, lso_csc* A_csc, const double* d_J, int64_t ld, const double* d_y, double* d_damp,
                   double* d_x, double atol, double btol, double conlim, int64_t maxiter, int64_t* iters_out,
                   int* istop_out) {
    return lso_lsmr_solve_ex(ws, A_csc, d_J, ld, d_y, d_damp, d_x, atol, btol, conlim, maxiter, nullptr, nullptr, nullptr,
                             iters_out, istop_out);
}

// Row-sharded LSMR (one process per GPU, NCCL): A_csc / d_y are THIS rank's rows of J and y, d_damp / d_x / d_P_diag are
// replicated; m_total (rows of the whole J, for the default maxiter of lsmr.jl:57) may be 0 = unknown.  Falls back to
// lso_lsmr_solve_ex on a context without a communicator.
int lso_lsmr_solve_sharded(lso_lsmr_ws* ws, lso_csc* A_csc, const double* d_y, double* d_damp, double* d_x, double atol,
                           double btol, double conlim, int64_t maxiter, int64_t m_total, const double* d_P_diag,
                           int64_t* iters_out, int* istop_out) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    if (ctx->nranks <= 1)
        return lso_lsmr_solve_ex(ws, A_csc, nullptr, 0, d_y, d_damp, d_x, atol, btol, conlim, maxiter, d_P_diag, nullptr, nullptr,
                                 iters_out, istop_out);
    LSO_REQUIRE(ctx, A_csc && d_y && d_x && iters_out && istop_out, "NULL pointer (the sharded LSMR needs a CSC operator)");
    LSO_REQUIRE(ctx, A_csc->m == ws->m && A_csc->n == ws->n, "operator / workspace dimension mismatch");
    LSO_REQUIRE(ctx, ctx->opt_spmv >= 1, "the sharded LSMR runs on the fused sparse products (ctx option spmv >= 1)");
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    return lsmr_fused_csc_sharded(ws, A_csc, d_y, d_damp, d_x, d_P_diag, atol, btol, conlim, maxiter, m_total, iters_out, istop_out);
}

int lso_lsmr_ws_stats(lso_lsmr_ws* ws, int64_t* launches_out, int64_t* syncs_out) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    if (launches_out) *launches_out = ws->last_launches;
    if (syncs_out) *syncs_out = ws->last_syncs;
    return LSO_OK;
}

}  // extern "C"

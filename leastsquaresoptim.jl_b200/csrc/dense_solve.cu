// dense_solve.cu — the two dense `AbstractAllocatedSolver`s of the reference behind the C ABI:
//   DenseQRAllocatedSolver        src/solver/dense_qr.jl:6-88
//   DenseCholeskyAllocatedSolver  src/solver/dense_cholesky.jl:7-59
#include "qr.cuh"
#include "chol.cuh"

struct lso_dense_ws {
    lso_ctx* ctx = nullptr;
    int64_t m = 0, n = 0;
    int kind = 0;
    int damped = 0;
    // QR
    QRPlan plan;          // local factorisation of [J; sqrt(D) | y; 0]
    QRPlan plan_stack;    // multi-GPU: QR of the stacked R factors (created on first sharded solve)
    bool have_stack = false;
    int stack_P = 0;
    double* d_gather = nullptr;   // multi-GPU: nranks * (n x (n+1)) gathered [R | Q'y]
    // Cholesky
    CholPlan chol;
    // host-buffer staging
    double* d_J = nullptr;
    double* d_y = nullptr;
    double* d_damp = nullptr;
    double* d_x = nullptr;
    int last_rank = 0;
    // (f3) factor kept across the re-solves of a rejected trust-region step: [R_J | Q'y] of the last lso_qr_factor_keep /
    // sharded solve lives in d_gather; valid until J or y change
    bool kept = false;
    // (f3) re-damping: the triangular factor [R | c] of the last DAMPED solve (in last_plan->A) and its damping vector;
    // a re-solve with a larger damping on the same J, y is the QR of the 2n x n stack [R; sqrt(damp_new - damp_last)]
    QRPlan* last_plan = nullptr;
    // host-fed chunked factorisation with >= 3 chunks: a second workspace and stream, so that the (latency-bound) panel
    // trees of one chunk run under the trailing updates of the other
    QRPlan plan_twin[3];
    int n_twin = 0;
    cudaStream_t twin_stream[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t twin_done[3] = {nullptr, nullptr, nullptr};
    // panel-pipelined sharded solve: row blocks of the local R travel (all-gather) and enter the replicated stack QR while
    // the local factorisation is still running
    double* d_rowsend = nullptr;      // npanels x [ (n+1) columns x 32 rows ]
    double* d_rowgath = nullptr;      // npanels x P x [ (n+1) x 32 ]
    cudaStream_t comm_stream = nullptr, stack_stream = nullptr;
    std::vector<cudaEvent_t> ev_local, ev_gath;
    cudaEvent_t ev_stack_done = nullptr, ev_pipe_start = nullptr;
    int pipe_P = 0;
    // host-fed chunks, pipelined: chunk c's panel k is followed by its row block and the event ev_chunk[c * np + k]; the
    // stack QR of lso_qr_solve_kept follows panel by panel.  chunk_done[w]: the last work queued on chunk stream w.
    std::vector<cudaEvent_t> ev_chunk;
    cudaEvent_t chunk_done[4] = {nullptr, nullptr, nullptr, nullptr};
    int chunks_pending = 0;           // chunk streams (0 = none) the context stream has not been joined with yet
    bool kept_pipe = false;           // d_rowgath holds the row blocks of all chunks (the kept factor of the pipelined form)
    QRPlan plan_redamp;
    bool have_redamp = false;
    double* d_lastdamp = nullptr;
    double* d_redamp_slot = nullptr;   // n x (n+1) packed [R | c]  +  n doubles for the damping increment
};

// The pipelined host-fed factorisation leaves work on the chunk streams that the context stream has not waited for (that is
// its point: colsumabs2! and the damping run meanwhile).  Every other entry that touches the plans joins first.
static int ws_join_chunks(lso_dense_ws* ws) {
    lso_ctx* ctx = ws->ctx;
    for (int w = 0; w < ws->chunks_pending; ++w)
        LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ws->chunk_done[w], 0));
    ws->chunks_pending = 0;
    return LSO_OK;
}

// ---- Q-a: build [J; diag(sqrt(damp)) | y; 0] in the padded workspace (dense_qr.jl:32-36, 64-80) ----
__global__ void qr_assemble_kernel(long long m, long long n, long long M, const double* __restrict__ J, long long ldJ,
                                   const double* __restrict__ y, const double* __restrict__ damp,
                                   double* __restrict__ A, long long ld, long long Npad) {
    const long long col = blockIdx.y;
    double* __restrict__ dst = A + col * ld;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < ld; r += (long long)gridDim.x * blockDim.x) {
        double v = 0.0;
        if (col < n) {
            if (r < m) v = J[col * ldJ + r];
            else if (damp != nullptr && r == m + col) v = sqrt(damp[col]);
        } else if (col == Npad) {
            if (r < m) v = y[r];
        }
        dst[r] = v;
    }
}

__global__ void extract_rhs_kernel(long long n, const double* __restrict__ A, long long ld, long long Npad,
                                   double* __restrict__ c) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) c[i] = A[Npad * ld + i];
}

// min/max of |R_ii| : cheap screen for numerical rank deficiency of the unpivoted factor
__global__ void diag_minmax_kernel(int n, const double* __restrict__ R, long long ld, double* __restrict__ out) {
    __shared__ double smn[32], smx[32];
    double mn = INFINITY, mx = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v = fabs(R[(long long)i * ld + i]);
        if (!(v >= 0.0)) v = 0.0;   // NaN -> 0 => flagged
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = fmin(mn, smn[w]); mx = fmax(mx, smx[w]); }
        out[0] = mn;
        out[1] = mx;
    }
}

static int qr_assemble(lso_ctx* ctx, QRPlan* p, int64_t m, int64_t n, const double* d_J, int64_t ldJ, const double* d_y,
                       const double* d_damp) {
    dim3 grid((unsigned)std::min<int64_t>(cdiv64(p->ld, 256), 64), (unsigned)p->Nc);
    qr_assemble_kernel<<<grid, 256, 0, ctx->stream>>>(m, n, p->M, d_J, ldJ, d_y, d_damp, p->A, p->ld, p->Npad);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

int small_qr_finish(lso_ctx* ctx, int64_t n, double* d_R, int64_t ld, double* d_c, double* d_x, int* rank_out);
int qr_rank_screen(lso_ctx* ctx, int64_t n, const double* d_R, int64_t ld, double rcond, int* full_rank_out);

// (f3) after a damped solve: plan->A holds [R | c] with R'R = J'J + D; keep D so that a later re-solve with MORE damping
// on the same J, y (a rejected LM step, levenberg_marquardt.jl:77-87) only has to factor [R; sqrt(D_new - D)].
static int remember_damped_factor(lso_dense_ws* ws, QRPlan* plan, const double* d_damp) {
    lso_ctx* ctx = ws->ctx;
    if (!d_damp) { ws->last_plan = nullptr; return LSO_OK; }
    if (!ws->d_lastdamp) LSO_CHECK_CUDA(ctx, cudaMalloc(&ws->d_lastdamp, (size_t)ws->n * sizeof(double)));
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(ws->d_lastdamp, d_damp, (size_t)ws->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    ws->last_plan = plan;
    return LSO_OK;
}

// R (n x n upper, in plan->A) and c = Q'y (column Npad) -> x.  Full-rank fast path: back substitution.
static int qr_finish(lso_dense_ws* ws, QRPlan* p, double* d_x, int* rank_out, bool undamped) {
    lso_ctx* ctx = ws->ctx;
    const int64_t n = ws->n;
    double* c = p->A + p->Npad * p->ld;
    // screen: the reference's pivoted QR + incremental condition estimation declares rank deficiency when
    // cond(R(1:k,1:k)) > 1/(min(rows,cols)*eps).  max|r_ii|/min|r_ii| is a lower bound of cond(R); when it is
    // comfortably below that threshold the matrix is treated as full rank; otherwise take the pivoted,
    // rank-revealing small-R finish (Q-d), which reproduces the reference's minimum-norm solution.
    diag_minmax_kernel<<<1, 256, 0, ctx->stream>>>((int)n, p->A, p->ld, ctx->d_scalars + 8);
    LSO_CHECK_LAUNCH(ctx);
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars + 8, ctx->d_scalars + 8, 2 * sizeof(double), cudaMemcpyDeviceToHost,
                                        ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double mn = ctx->h_scalars[8], mx = ctx->h_scalars[9];
    // rcond = min(rows, cols) * eps with the rows of the system the REFERENCE factors (zero padding rows do not count)
    const int64_t sys_rows = (p == &ws->plan) ? (ws->damped ? ws->m + ws->n : ws->m) : p->M;
    const double rcond = (double)std::min<int64_t>(sys_rows, n) * 2.220446049250313e-16;
    bool suspicious = !(mn > 1e3 * rcond * mx);   // also true for NaN / zero matrix
    if (!suspicious && undamped && n > 1) {
        // an ill-conditioned J can hide behind a benign diagonal (Kahan-type matrices): for the undamped solves, where
        // nothing bounds cond(J), run the reference's incremental condition estimate on the unpivoted triangle as well
        int full = 0;
        LSO_TRY(qr_rank_screen(ctx, n, p->A, p->ld, rcond, &full));
        suspicious = !full;
    }
    if (!suspicious) {
        LSO_TRY(tri_solve(ctx, n, p->A, p->ld, c, d_x, 0));
        ws->last_rank = (int)n;
        if (rank_out) *rank_out = (int)n;
        return LSO_OK;
    }
    int rk = 0;
    LSO_TRY(small_qr_finish(ctx, n, p->A, p->ld, c, d_x, &rk));
    ws->last_rank = rk;
    if (rank_out) *rank_out = rk;
    return LSO_OK;
}

extern "C" {

int lso_dense_ws_create(lso_ctx* ctx, int64_t m, int64_t n, int solver_kind, int damped, lso_dense_ws** out) {
    LSO_REQUIRE(ctx, ctx && out, "ctx/out is NULL");
    *out = nullptr;
    LSO_REQUIRE(ctx, m >= 1 && n >= 1, "m and n must be positive");
    LSO_REQUIRE(ctx, solver_kind == LSO_SOLVER_QR || solver_kind == LSO_SOLVER_CHOLESKY, "unknown solver kind");
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    lso_dense_ws* ws = new (std::nothrow) lso_dense_ws();
    if (!ws) return lso_set_error(ctx, LSO_ERR_ALLOC, "host allocation failed");
    ws->ctx = ctx; ws->m = m; ws->n = n; ws->kind = solver_kind; ws->damped = damped;
    int st = LSO_OK;
    if (solver_kind == LSO_SOLVER_QR) {
        // LM: (m+n) x n augmented system (dense_qr.jl:50-54); Dogleg: m x n (dense_qr.jl:25-28)
        // (m < n: the reference sizes u = zeros(max(m, n)), dense_qr.jl:27; here J is padded with zero rows to n x n, which
        // has the same Gram matrix, hence the same pivots, R and minimum-norm solution)
        st = qr_plan_create(ctx, damped ? m + n : std::max(m, n), n, &ws->plan);
    } else {
        st = chol_plan_create(ctx, n, &ws->chol);
    }
    if (st != LSO_OK) { lso_dense_ws_destroy(ws); return st; }
    *out = ws;
    return LSO_OK;
}

int lso_dense_ws_destroy(lso_dense_ws* ws) {
    if (!ws) return LSO_OK;
    lso_ctx* ctx = ws->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ws->comm_stream) cudaStreamSynchronize(ws->comm_stream);
    if (ws->stack_stream) cudaStreamSynchronize(ws->stack_stream);
    for (int t = 0; t < ws->n_twin; ++t) if (ws->twin_stream[t]) cudaStreamSynchronize(ws->twin_stream[t]);
    qr_plan_destroy(&ws->plan);
    if (ws->have_stack) qr_plan_destroy(&ws->plan_stack);
    if (ws->have_redamp) qr_plan_destroy(&ws->plan_redamp);
    cudaFree(ws->d_rowsend); cudaFree(ws->d_rowgath);
    if (ws->comm_stream) cudaStreamDestroy(ws->comm_stream);
    if (ws->stack_stream) cudaStreamDestroy(ws->stack_stream);
    for (cudaEvent_t e : ws->ev_local) cudaEventDestroy(e);
    for (cudaEvent_t e : ws->ev_gath) cudaEventDestroy(e);
    for (cudaEvent_t e : ws->ev_chunk) cudaEventDestroy(e);
    for (cudaEvent_t e : ws->chunk_done) if (e) cudaEventDestroy(e);
    if (ws->ev_stack_done) cudaEventDestroy(ws->ev_stack_done);
    if (ws->ev_pipe_start) cudaEventDestroy(ws->ev_pipe_start);
    for (int t = 0; t < ws->n_twin; ++t) {
        qr_plan_destroy(&ws->plan_twin[t]);
        if (ws->twin_stream[t]) cudaStreamDestroy(ws->twin_stream[t]);
        if (ws->twin_done[t]) cudaEventDestroy(ws->twin_done[t]);
    }
    cudaFree(ws->d_lastdamp);
    cudaFree(ws->d_redamp_slot);
    chol_plan_destroy(&ws->chol);
    cudaFree(ws->d_gather);
    cudaFree(ws->d_J); cudaFree(ws->d_y); cudaFree(ws->d_damp); cudaFree(ws->d_x);
    delete ws;
    return LSO_OK;
}

int lso_qr_solve(lso_dense_ws* ws, const double* d_J, int64_t ld, const double* d_y, const double* d_damp,
                 double* d_x, int* rank_out) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, ws->kind == LSO_SOLVER_QR, "workspace was not created for QR");
    LSO_REQUIRE(ctx, d_J && d_y && d_x, "NULL pointer");
    LSO_REQUIRE(ctx, ld >= ws->m, "leading dimension < m");
    // dense_qr.jl:61 — the damped form needs the (m+n)-row workspace, the undamped form the m-row one
    LSO_REQUIRE(ctx, (d_damp != nullptr) == (ws->damped != 0), "length(u) should equal length(x) + length(y)");
    LSO_ENTER(ctx);
    LSO_TRY(ws_join_chunks(ws));
    ws->last_plan = nullptr;
    LSO_TRY(qr_assemble(ctx, &ws->plan, ws->m, ws->n, d_J, ld, d_y, d_damp));
    LSO_TRY(qr_factor(ctx, &ws->plan));
    LSO_TRY(qr_finish(ws, &ws->plan, d_x, rank_out, d_damp == nullptr));
    return remember_damped_factor(ws, &ws->plan, d_damp);
}

static int ensure_staging(lso_dense_ws* ws, int64_t ld) {
    lso_ctx* ctx = ws->ctx;
    if (!ws->d_J) {
        LSO_CHECK_CUDA(ctx, cudaMalloc(&ws->d_J, (size_t)ws->m * ws->n * sizeof(double)));
        LSO_CHECK_CUDA(ctx, cudaMalloc(&ws->d_y, (size_t)ws->m * sizeof(double)));
        LSO_CHECK_CUDA(ctx, cudaMalloc(&ws->d_damp, (size_t)ws->n * sizeof(double)));
        LSO_CHECK_CUDA(ctx, cudaMalloc(&ws->d_x, (size_t)ws->n * sizeof(double)));
    }
    (void)ld;
    return LSO_OK;
}

int lso_qr_solve_host(lso_dense_ws* ws, const double* h_J, int64_t ld, const double* h_y, const double* h_damp,
                      double* h_x, int* rank_out) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, h_J && h_y && h_x, "NULL pointer");
    LSO_TRY(ensure_staging(ws, ld));
    LSO_TRY(lso_upload_matrix(ctx, ws->d_J, ws->m, h_J, ld, ws->m, ws->n));
    LSO_TRY(lso_upload_async(ctx, ws->d_y, h_y, ws->m * sizeof(double)));
    if (h_damp) LSO_TRY(lso_upload_async(ctx, ws->d_damp, h_damp, ws->n * sizeof(double)));
    LSO_TRY(lso_qr_solve(ws, ws->d_J, ws->m, ws->d_y, h_damp ? ws->d_damp : nullptr, ws->d_x, rank_out));
    return lso_download(ctx, h_x, ws->d_x, ws->n * sizeof(double));
}

int lso_chol_solve(lso_dense_ws* ws, const double* d_J, int64_t ld, const double* d_y, const double* d_damp,
                   double* d_x) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, ws->kind == LSO_SOLVER_CHOLESKY, "workspace was not created for Cholesky");
    LSO_REQUIRE(ctx, d_J && d_y && d_x, "NULL pointer");
    LSO_REQUIRE(ctx, ld >= ws->m, "leading dimension < m");
    LSO_ENTER(ctx);
    return chol_solve(ctx, &ws->chol, ws->m, ws->n, d_J, ld, d_y, d_damp, d_x, 0);
}

// Row-sharded J (one rank per GPU): J, y are this rank's rows; [upper(J'J) | J'y] is summed over the ranks by ONE
// NCCL all-reduce, the factorisation and the solves are replicated (no reference counterpart).
int lso_chol_solve_sharded(lso_dense_ws* ws, const double* d_J, int64_t ld, const double* d_y, const double* d_damp,
                           double* d_x) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, ws->kind == LSO_SOLVER_CHOLESKY, "workspace was not created for Cholesky");
    LSO_REQUIRE(ctx, d_J && d_y && d_x, "NULL pointer");
    LSO_REQUIRE(ctx, ld >= ws->m, "leading dimension < m");
    LSO_ENTER(ctx);
    return chol_solve(ctx, &ws->chol, ws->m, ws->n, d_J, ld, d_y, d_damp, d_x, 1);
}

// (f3) re-solve with the J'J and J'f of the last lso_chol_solve / lso_chol_solve_sharded and a new damping
int lso_chol_solve_kept(lso_dense_ws* ws, const double* d_damp, double* d_x) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, ws->kind == LSO_SOLVER_CHOLESKY && d_x, "bad arguments");
    LSO_ENTER(ctx);
    return chol_solve_kept(ctx, &ws->chol, d_damp, d_x);
}

// Test hook: the sharded Cholesky algorithm with P shards of ws->m rows each emulated on one device (d_J is (P * m) x n).
int lso_debug_chol_solve_emulated_shards(lso_dense_ws* ws, int P, const double* d_J, int64_t ld, const double* d_y,
                                         const double* d_damp, double* d_x) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, ws->kind == LSO_SOLVER_CHOLESKY, "workspace was not created for Cholesky");
    LSO_REQUIRE(ctx, P >= 1 && P <= 64 && d_J && d_y && d_x, "bad arguments");
    LSO_REQUIRE(ctx, ld >= (int64_t)P * ws->m, "leading dimension < P * m");
    LSO_ENTER(ctx);
    return chol_solve_emulated(ctx, &ws->chol, P, ws->m, ws->n, d_J, ld, d_y, d_damp, d_x);
}

int lso_chol_solve_host(lso_dense_ws* ws, const double* h_J, int64_t ld, const double* h_y, const double* h_damp,
                        double* h_x) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, h_J && h_y && h_x, "NULL pointer");
    LSO_ENTER(ctx);
    LSO_TRY(ensure_staging(ws, ld));
    LSO_TRY(lso_upload_matrix(ctx, ws->d_J, ws->m, h_J, ld, ws->m, ws->n));
    LSO_TRY(lso_upload_async(ctx, ws->d_y, h_y, ws->m * sizeof(double)));
    if (h_damp) LSO_TRY(lso_upload_async(ctx, ws->d_damp, h_damp, ws->n * sizeof(double)));
    int st = lso_chol_solve(ws, ws->d_J, ws->m, ws->d_y, h_damp ? ws->d_damp : nullptr, ws->d_x);
    if (st != LSO_OK) return st;
    return lso_download(ctx, h_x, ws->d_x, ws->n * sizeof(double));
}

int lso_dense_ws_get_factor(lso_dense_ws* ws, double* h_R) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, h_R != nullptr, "NULL pointer");
    LSO_ENTER(ctx);
    const int64_t n = ws->n;
    const double* src = (ws->kind == LSO_SOLVER_QR) ? ws->plan.A : ws->chol.C;
    const int64_t ld = (ws->kind == LSO_SOLVER_QR) ? ws->plan.ld : ws->chol.ldc;
    LSO_CHECK_CUDA(ctx, cudaMemcpy2DAsync(h_R, n * sizeof(double), src, ld * sizeof(double), n * sizeof(double), n,
                                          cudaMemcpyDeviceToHost, ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = j + 1; i < n; ++i) h_R[j * n + i] = 0.0;
    return LSO_OK;
}

// ---- multi-GPU QR: TSQR over row shards ----------------------------------------------------------
int lso_comm_allgather(lso_ctx* ctx, const double* d_send, double* d_recv, int64_t count);

__global__ void pack_R_kernel(long long n, const double* __restrict__ A, long long ld, long long Npad,
                              double* __restrict__ out /* n x (n+1), ld = n */) {
    const long long col = blockIdx.y;   // 0..n  (n = rhs)
    const long long src_col = (col < n) ? col : Npad;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        double v = A[src_col * ld + r];
        if (col < n && r > col) v = 0.0;
        out[col * n + r] = v;
    }
}

// stack of the P upper-triangular R_k (and diag(sqrt(damp)) as one more triangle when damping is present), rows
// INTERLEAVED: stack row Q*r + i = row r of triangle i  (Q = P or P + 1).  Stack row rho then has no entry left of
// column floor(rho / Q), so panel j of the factorisation only involves the rows below Q*32*(j+1) (QRPlan::band):
// the replicated QR of the stack costs half of a dense one.  rhs column = the c_k interleaved the same way.
__global__ void stack_assemble_kernel(long long n, int P, int Q, const double* __restrict__ gathered,
                                      const double* __restrict__ damp, double* __restrict__ A, long long ld,
                                      long long Npad) {
    const long long col = blockIdx.y;
    double* __restrict__ dst = A + col * ld;
    const long long rows = (long long)Q * n;
    for (long long rho = blockIdx.x * (long long)blockDim.x + threadIdx.x; rho < ld; rho += (long long)gridDim.x * blockDim.x) {
        double v = 0.0;
        if ((col < n || col == Npad) && rho < rows) {
            const long long r = rho / Q;
            const int i = (int)(rho - r * Q);
            const long long scol = (col < n) ? col : n;
            if (i < P) v = gathered[(long long)i * n * (n + 1) + scol * n + r];
            else if (col < n && r == col) v = sqrt(damp[col]);
        }
        dst[rho] = v;
    }
}

// local QR of [J_k | y_k] and its n x (n+1) [R | Q'y] packed into `slot`
static int shard_local_R(lso_dense_ws* ws, const double* d_J, int64_t ld, const double* d_y, double* slot, int64_t rows = -1,
                         QRPlan* plan = nullptr) {
    lso_ctx* ctx = ws->ctx;
    const int64_t n = ws->n;
    if (rows < 0) rows = ws->m;
    if (!plan) plan = &ws->plan;
    LSO_TRY(qr_assemble(ctx, plan, rows, n, d_J, ld, d_y, nullptr));
    const int64_t M_full = plan->M;
    plan->M = std::max<int64_t>(rows, n);      // a damped workspace has m + n rows: only the rows of J are in use here
    const int st = qr_factor(ctx, plan);
    plan->M = M_full;
    LSO_TRY(st);
    dim3 grid((unsigned)std::min<int64_t>(cdiv64(n, 256), 64), (unsigned)(n + 1));
    pack_R_kernel<<<grid, 256, 0, ctx->stream>>>(n, plan->A, plan->ld, plan->Npad, slot);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}
static int shard_ensure_stack(lso_dense_ws* ws, int P) {
    lso_ctx* ctx = ws->ctx;
    const int64_t n = ws->n;
    if (ws->have_stack && ws->stack_P != P) {
        qr_plan_destroy(&ws->plan_stack);
        cudaFree(ws->d_gather);
        ws->d_gather = nullptr;
        ws->have_stack = false;
    }
    if (!ws->have_stack) {
        LSO_TRY(qr_plan_create(ctx, (int64_t)P * n + n, n, &ws->plan_stack));
        ws->have_stack = true;
        ws->stack_P = P;
        ws->plan_stack.band = P + 1;       // tune the launch schedule for the banded shape the stack solves have
        LSO_TRY(qr_plan_tune(ctx, &ws->plan_stack));
        LSO_CHECK_CUDA(ctx, cudaMalloc(&ws->d_gather, (size_t)(P + 1) * n * (n + 1) * sizeof(double)));
    }
    return LSO_OK;
}
// replicated QR of the interleaved stack of the P gathered triangles (+ the damping triangle), then the solve
static int shard_stack_solve(lso_dense_ws* ws, int P, const double* d_damp, double* d_x, int* rank_out) {
    lso_ctx* ctx = ws->ctx;
    const int64_t n = ws->n;
    QRPlan* ps = &ws->plan_stack;
    dim3 grid((unsigned)std::min<int64_t>(cdiv64(ps->ld, 256), 64), (unsigned)ps->Nc);
    const int Q = d_damp ? P + 1 : P;
    stack_assemble_kernel<<<grid, 256, 0, ctx->stream>>>(n, P, Q, ws->d_gather, d_damp, ps->A, ps->ld, ps->Npad);
    LSO_CHECK_LAUNCH(ctx);
    ps->M = (int64_t)Q * n;       // rows in use (the workspace was sized for P + 1 triangles)
    ps->band = Q;
    ws->last_plan = nullptr;
    LSO_TRY(qr_factor(ctx, ps));
    LSO_TRY(qr_finish(ws, ps, d_x, rank_out, d_damp == nullptr));
    return remember_damped_factor(ws, ps, d_damp);
}

// ---- panel-pipelined TSQR -----------------------------------------------------------------------------------------
// Row block k (rows 32k .. 32k+31) of a rank's R and of its Q'y is final as soon as the rank's panel k has been factorised
// and applied, and panel k of the stack QR only involves the row blocks 0..k of every rank (QRPlan::band).  So the stack
// factorisation runs panel by panel BEHIND the local one, on its own stream, fed by one small all-gather per panel: the
// two latency-bound chains (32 panel trees each) overlap instead of following each other.
// out: [(n+1) columns][32 rows] for row block k: columns 0..n-1 of R (zero below the diagonal / beyond row n), column n = Q'y
__global__ void pack_rowblock_kernel(long long n, const double* __restrict__ A, long long ld, long long Npad, long long k,
                                     double* __restrict__ out) {
    const long long total = (n + 1) * QB;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long col = e / QB, r = e - col * QB;
        const long long row = k * QB + r;
        double v = 0.0;
        if (row < n) {
            if (col == n) v = A[Npad * ld + row];
            else if (row <= col) v = A[col * ld + row];
        }
        out[e] = v;
    }
}
// rows Q*(32k + r) + i of the interleaved stack (all Nc columns) from the P gathered row blocks and the damping triangle
__global__ void stack_assemble_panel_kernel(long long n, int P, int Q, long long k, const double* __restrict__ gath,
                                            const double* __restrict__ damp, double* __restrict__ A, long long ld,
                                            long long Npad, long long Nc) {
    const long long col = blockIdx.y;                       // 0 .. Nc-1
    if (col >= Nc) return;
    const long long scol = (col < n) ? col : ((col == Npad) ? n : -1);
    for (int e = threadIdx.x; e < Q * QB; e += blockDim.x) {
        const int r = e / Q, i = e - r * Q;
        const long long row = k * QB + r;                   // row of the triangles
        if (row >= n) continue;                             // (the rows past Q n are zeroed by stack_zero_tail_kernel)
        double v = 0.0;
        if (scol >= 0) {
            if (i < P) v = gath[((long long)i * (n + 1) + scol) * QB + r];
            else if (scol == row) v = sqrt(damp[row]);
        }
        A[col * ld + (long long)Q * row + i] = v;
    }
}
// rows [row0, row1) of every column := 0 (what lies past the Q n rows in use: padding, and rows a solve with more triangles left)
__global__ void stack_zero_tail_kernel(double* __restrict__ A, long long ld, long long row0, long long row1) {
    const long long col = blockIdx.y;
    for (long long r = row0 + threadIdx.x; r < row1; r += blockDim.x) A[col * ld + r] = 0.0;
}

static int pipe_ensure(lso_dense_ws* ws, int P) {
    lso_ctx* ctx = ws->ctx;
    const int64_t n = ws->n;
    const int64_t np = ws->plan.Npad / QB;
    if (ws->pipe_P == P && ws->d_rowsend) return LSO_OK;
    cudaFree(ws->d_rowsend); cudaFree(ws->d_rowgath);
    ws->d_rowsend = ws->d_rowgath = nullptr;
    const size_t blk = (size_t)(n + 1) * QB;
    LSO_CHECK_CUDA(ctx, cudaMalloc(&ws->d_rowsend, (size_t)np * blk * sizeof(double)));
    LSO_CHECK_CUDA(ctx, cudaMalloc(&ws->d_rowgath, (size_t)np * (size_t)P * blk * sizeof(double)));
    if (!ws->comm_stream) {
        LSO_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ws->comm_stream, cudaStreamNonBlocking));
        LSO_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ws->stack_stream, cudaStreamNonBlocking));
        LSO_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ws->ev_stack_done, cudaEventDisableTiming));
        LSO_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ws->ev_pipe_start, cudaEventDisableTiming));
    }
    while ((int64_t)ws->ev_local.size() < np) {
        cudaEvent_t a, b;
        LSO_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        LSO_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        ws->ev_local.push_back(a);
        ws->ev_gath.push_back(b);
    }
    ws->pipe_P = P;
    return LSO_OK;
}

// The stack side of the pipelined forms: on the stack stream, panel k of the replicated stack QR follows the arrival of row
// block k of every triangle.  n_chunk_events == 0: the blocks of panel k are signalled by ev_gath[k] (all-gather / emulation);
// > 0: by ev_chunk[c * np + k] for each of that many host-fed chunks.  Ends with the finish on the context stream.
static int stack_solve_from_rowblocks(lso_dense_ws* ws, int P, int n_chunk_events, const double* d_damp, double* d_x, int* rank_out) {
    lso_ctx* ctx = ws->ctx;
    const int64_t n = ws->n;
    QRPlan* ps = &ws->plan_stack;
    const int Q = d_damp ? P + 1 : P;
    ps->M = (int64_t)Q * n;
    ps->band = Q;
    const int64_t np = ws->plan.Npad / QB;
    const int64_t nps = qr_num_panels(ps);
    const size_t blk = (size_t)(n + 1) * QB;
    cudaStream_t U = ctx->stream, S = ws->stack_stream;
    ws->last_plan = nullptr;
    if (n_chunk_events > 0) {                               // the damping was computed on the context stream after the chunks were queued
        LSO_CHECK_CUDA(ctx, cudaEventRecord(ws->ev_pipe_start, U));
        LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(S, ws->ev_pipe_start, 0));
    }
    {
        const long long row0 = (long long)Q * n, row1 = std::min<long long>(ps->ld, row0 + 2 * QH);
        dim3 grid(1, (unsigned)ps->Nc);
        stack_zero_tail_kernel<<<grid, 128, 0, S>>>(ps->A, ps->ld, row0, row1);
        ctx->launches++;
    }
    int st = LSO_OK;
    ctx->stream = S;
    for (int64_t k = 0; k < np && st == LSO_OK; ++k) {
        if (n_chunk_events > 0) for (int c = 0; c < n_chunk_events; ++c) cudaStreamWaitEvent(S, ws->ev_chunk[(size_t)c * np + k], 0);
        else cudaStreamWaitEvent(S, ws->ev_gath[k], 0);
        dim3 grid(1, (unsigned)ps->Nc);
        stack_assemble_panel_kernel<<<grid, 128, 0, S>>>(n, P, Q, k, ws->d_rowgath + (size_t)k * P * blk, d_damp, ps->A, ps->ld, ps->Npad, ps->Nc);
        ctx->launches++;
        if (k < nps) st = qr_factor_range(ctx, ps, k, k + 1);
    }
    ctx->stream = U;
    LSO_TRY(st);
    LSO_CHECK_LAUNCH(ctx);
    LSO_CHECK_CUDA(ctx, cudaEventRecord(ws->ev_stack_done, S));
    LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(U, ws->ev_stack_done, 0));
    if (n_chunk_events > 0) ws->chunks_pending = 0;         // the stack followed every chunk's last panel
    LSO_TRY(qr_finish(ws, ps, d_x, rank_out, d_damp == nullptr));
    return remember_damped_factor(ws, ps, d_damp);
}

int lso_comm_allgather_on(lso_ctx* ctx, const double* d_send, double* d_recv, int64_t count, cudaStream_t st);

// emulate != 0: the P shards are the row chunks of d_J on THIS device (test hook): they are factorised one after the other and
// their row blocks copied into place instead of all-gathered; the stack side is the code the ranks run.
static int qr_solve_sharded_pipelined(lso_dense_ws* ws, int P, bool emulate, const double* d_J, int64_t ld, const double* d_y,
                                      const double* d_damp, double* d_x, int* rank_out) {
    lso_ctx* ctx = ws->ctx;
    const int64_t n = ws->n;
    LSO_TRY(shard_ensure_stack(ws, P));
    LSO_TRY(pipe_ensure(ws, P));
    QRPlan* pl = &ws->plan;
    QRPlan* ps = &ws->plan_stack;
    const int Q = d_damp ? P + 1 : P;
    ps->M = (int64_t)Q * n;
    ps->band = Q;
    const int64_t np = pl->Npad / QB;                        // panels of the local plan (m_loc >= n)
    const int64_t nps = qr_num_panels(ps);
    const size_t blk = (size_t)(n + 1) * QB;
    const unsigned pgrid = (unsigned)std::min<int64_t>(cdiv64((int64_t)blk, 256), 64);
    cudaStream_t U = ctx->stream, Cs = ws->comm_stream, S = ws->stack_stream;
    ws->last_plan = nullptr;
    ws->kept = false; ws->kept_pipe = false;
    LSO_CHECK_CUDA(ctx, cudaEventRecord(ws->ev_pipe_start, U));
    LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(Cs, ws->ev_pipe_start, 0));
    LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(S, ws->ev_pipe_start, 0));
    int st = LSO_OK;
    if (emulate) {
        for (int i = 0; i < P && st == LSO_OK; ++i) {        // shard i: whole local QR, then its row blocks into slot i of every panel
            st = qr_assemble(ctx, pl, ws->m, n, d_J + (size_t)i * ws->m, ld, d_y + (size_t)i * ws->m, nullptr);
            const int64_t M_full = pl->M;
            pl->M = ws->m;
            if (st == LSO_OK) st = qr_factor(ctx, pl);
            pl->M = M_full;
            for (int64_t k = 0; k < np && st == LSO_OK; ++k) {
                pack_rowblock_kernel<<<pgrid, 256, 0, U>>>(n, pl->A, pl->ld, pl->Npad, k, ws->d_rowgath + ((size_t)k * P + i) * blk);
                ctx->launches++;
            }
        }
        LSO_TRY(st);
        LSO_CHECK_LAUNCH(ctx);
        for (int64_t k = 0; k < np; ++k) LSO_CHECK_CUDA(ctx, cudaEventRecord(ws->ev_gath[k], U));
    } else {
        LSO_TRY(qr_assemble(ctx, pl, ws->m, n, d_J, ld, d_y, nullptr));
        const int64_t M_full = pl->M;
        pl->M = ws->m;
        for (int64_t k = 0; k < np && st == LSO_OK; ++k) {
            st = qr_factor_range(ctx, pl, k, k + 1);
            if (st != LSO_OK) break;
            pack_rowblock_kernel<<<pgrid, 256, 0, U>>>(n, pl->A, pl->ld, pl->Npad, k, ws->d_rowsend + (size_t)k * blk);
            ctx->launches++;
            cudaEventRecord(ws->ev_local[k], U);
            cudaStreamWaitEvent(Cs, ws->ev_local[k], 0);
            st = lso_comm_allgather_on(ctx, ws->d_rowsend + (size_t)k * blk, ws->d_rowgath + (size_t)k * P * blk, (int64_t)blk, Cs);
            cudaEventRecord(ws->ev_gath[k], Cs);
        }
        pl->M = M_full;
        LSO_TRY(st);
        LSO_CHECK_LAUNCH(ctx);
    }
    return stack_solve_from_rowblocks(ws, P, 0, d_damp, d_x, rank_out);
}

int lso_qr_solve_sharded(lso_dense_ws* ws, const double* d_J, int64_t ld, const double* d_y, const double* d_damp,
                         double* d_x, int* rank_out) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    if (ctx->nranks <= 1) return lso_qr_solve(ws, d_J, ld, d_y, d_damp, d_x, rank_out);
    LSO_REQUIRE(ctx, ws->kind == LSO_SOLVER_QR, "workspace was not created for QR");
    LSO_REQUIRE(ctx, ws->damped == 0, "sharded QR: create the workspace with damped = 0 (damping rows join the stack)");
    LSO_REQUIRE(ctx, d_J && d_y && d_x, "NULL pointer");
    LSO_REQUIRE(ctx, ws->m >= ws->n, "sharded QR: each shard needs rows >= columns");
    LSO_ENTER(ctx);
    LSO_TRY(ws_join_chunks(ws));
    const int64_t n = ws->n;
    const int P = ctx->nranks;
    if (ctx->opt_qr_shard_pipeline && ws->plan.Npad / QB >= 2)
        return qr_solve_sharded_pipelined(ws, P, false, d_J, ld, d_y, d_damp, d_x, rank_out);
    LSO_TRY(shard_ensure_stack(ws, P));
    double* sendbuf = ws->d_gather + (size_t)P * n * (n + 1);
    LSO_TRY(shard_local_R(ws, d_J, ld, d_y, sendbuf));
    lso_prof_mark2(ctx);
    LSO_TRY(lso_comm_allgather(ctx, sendbuf, ws->d_gather, n * (n + 1)));
    lso_prof_mark2(ctx);
    ws->kept = true; ws->kept_pipe = false;
    return shard_stack_solve(ws, P, d_damp, d_x, rank_out);
}

// (f3) levenberg_marquardt.jl:77-87: after a REJECTED step the reference solves again with the same J and f and a new
// damping, and refactors everything.  Here the QR of [J | y] is done once per Jacobian:
//   lso_qr_factor_keep   J = Q R_J : keeps the n x (n+1) block [R_J | Q'y]  (cost of one undamped factorisation)
//   lso_qr_solve_kept    QR of the interleaved 2n x n stack [R_J ; sqrt(D) | Q'y ; 0]  (banded: n^3-scale, independent of m)
// which is the same least-squares problem as QR of [J ; sqrt(D) | y ; 0] (dense_qr.jl:64-88).  On a row-sharded
// workspace the gathered factors of the last lso_qr_solve_sharded are what is kept: a re-solve needs no local
// factorisation and NO collective.
int lso_qr_factor_keep(lso_dense_ws* ws, const double* d_J, int64_t ld, const double* d_y) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, ws->kind == LSO_SOLVER_QR, "workspace was not created for QR");
    LSO_REQUIRE(ctx, d_J && d_y, "NULL pointer");
    LSO_REQUIRE(ctx, ld >= ws->m, "leading dimension < m");
    LSO_REQUIRE(ctx, ws->m >= ws->n, "kept-factor path needs rows >= columns");
    LSO_ENTER(ctx);
    LSO_TRY(ws_join_chunks(ws));
    ws->kept = false; ws->kept_pipe = false;
    LSO_TRY(shard_ensure_stack(ws, 1));
    LSO_TRY(shard_local_R(ws, d_J, ld, d_y, ws->d_gather));
    ws->kept = true; ws->kept_pipe = false;
    return LSO_OK;
}

// The same, fed from HOST memory in P row chunks: chunk k (ws->m rows, the last one possibly shorter) is copied to the
// device on a copy stream while chunk k-1 is being factorised (TSQR over the chunks), so the H2D transfer of J — 14.5 ms
// of a 30 ms end-to-end LM step at 100 000 x 1 000 — runs under the factorisation instead of in front of it.  J and y
// also land in d_J (ld_d) / d_y for the passes of the iteration that need them whole.  Follow with lso_qr_solve_kept.
int lso_qr_factor_keep_host_chunks(lso_dense_ws* ws, int P, const int64_t* chunk_rows, const double* h_J, int64_t ld_h,
                                   const double* h_y, double* d_J, int64_t ld_d, double* d_y) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, ws->kind == LSO_SOLVER_QR && h_J && h_y && d_J && d_y && chunk_rows, "bad arguments");
    LSO_REQUIRE(ctx, P >= 1 && P <= 16, "between 1 and 16 row chunks");
    const int64_t n = ws->n;
    int64_t m_total = 0;
    for (int k = 0; k < P; ++k) {
        LSO_REQUIRE(ctx, chunk_rows[k] >= 1 && chunk_rows[k] <= ws->m, "a chunk has no rows or more rows than the workspace");
        m_total += chunk_rows[k];
    }
    LSO_REQUIRE(ctx, ld_h >= m_total && ld_d >= m_total, "leading dimension < rows");
    LSO_REQUIRE(ctx, ws->m >= n, "the workspace needs rows >= columns");
    LSO_ENTER(ctx);
    LSO_TRY(ws_join_chunks(ws));
    ws->kept = false; ws->kept_pipe = false;
    ws->last_plan = nullptr;
    if (!ctx->copy_stream) {
        LSO_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (auto& e : ctx->copy_ev) LSO_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    LSO_TRY(shard_ensure_stack(ws, P));
    // the copies may not overtake earlier readers of d_J / d_y on the compute stream
    LSO_CHECK_CUDA(ctx, cudaEventRecord(ctx->copy_ev[0], ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[0], 0));
    int64_t r0 = 0;
    for (int k = 0; k < P; ++k) {            // chunks cross PCIe in the order given
        const int64_t rows = chunk_rows[k];
        LSO_CHECK_CUDA(ctx, cudaMemcpy2DAsync(d_J + r0, ld_d * sizeof(double), h_J + r0, ld_h * sizeof(double), rows * sizeof(double),
                                              n, cudaMemcpyDefault, ctx->copy_stream));
        LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(d_y + r0, h_y + r0, rows * sizeof(double), cudaMemcpyDefault, ctx->copy_stream));
        LSO_CHECK_CUDA(ctx, cudaEventRecord(ctx->copy_ev[1 + k], ctx->copy_stream));
        r0 += rows;
    }
    // With three or more chunks they are factorised round-robin in up to `qr_twin` + 1 workspaces on as many streams: a
    // chunk's QR is mostly latency (32 panel trees) and the next chunk is usually there before it ends, so the panel trees
    // of one chunk run under the trailing updates of another.  TSQR does not care about the order of the triangles.
    int K = (P >= 3) ? 1 + (int)std::min<int64_t>(std::max<int64_t>(ctx->opt_qr_twin, 0), 3) : 1;
    if (K > P) K = P;
    while (ws->n_twin < K - 1) {
        const int t = ws->n_twin;
        LSO_TRY(qr_plan_create(ctx, ws->plan.M, n, &ws->plan_twin[t]));
        LSO_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ws->twin_stream[t], cudaStreamNonBlocking));
        LSO_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ws->twin_done[t], cudaEventDisableTiming));
        ws->n_twin = t + 1;
    }
    for (int t = 0; t < K - 1; ++t) LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(ws->twin_stream[t], ctx->copy_ev[0], 0));
    cudaStream_t main_stream = ctx->stream;
    int st = LSO_OK;
    r0 = 0;
    if (ctx->opt_qr_shard_pipeline && P >= 2 && ws->plan.Npad / QB >= 2) {
        // Pipelined form: every chunk is factorised panel by panel on a stream of its own (never the context stream), its row
        // block k packed and signalled after panel k; the context stream only waits for the LAST COPY, so the caller's
        // passes over the whole J (colsumabs2!, J'f) and the damping run while the last chunk is still being factorised,
        // and lso_qr_solve_kept's stack QR then follows that factorisation one panel behind (qr_solve_sharded_pipelined).
        LSO_TRY(pipe_ensure(ws, P));
        const int64_t np = ws->plan.Npad / QB;
        while ((int64_t)ws->ev_chunk.size() < (int64_t)P * np) {
            cudaEvent_t e;
            LSO_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ws->ev_chunk.push_back(e);
        }
        for (int w = 0; w < K; ++w)
            if (!ws->chunk_done[w]) LSO_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ws->chunk_done[w], cudaEventDisableTiming));
        LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(ws->comm_stream, ctx->copy_ev[0], 0));
        const size_t blk = (size_t)(n + 1) * QB;
        const unsigned pgrid = (unsigned)std::min<int64_t>(cdiv64((int64_t)blk, 256), 64);
        for (int c = 0; c < P && st == LSO_OK; ++c) {
            const int64_t rows = chunk_rows[c];
            const int w = c % K;
            QRPlan* pl = (w == 0) ? &ws->plan : &ws->plan_twin[w - 1];
            cudaStream_t cs = (w == 0) ? ws->comm_stream : ws->twin_stream[w - 1];
            ctx->stream = cs;
            if (cudaStreamWaitEvent(cs, ctx->copy_ev[1 + c], 0) != cudaSuccess) { st = lso_set_error(ctx, LSO_ERR_CUDA, "cudaStreamWaitEvent failed"); break; }
            st = qr_assemble(ctx, pl, rows, n, d_J + r0, ld_d, d_y + r0, nullptr);
            const int64_t M_full = pl->M;
            pl->M = std::max<int64_t>(rows, n);
            for (int64_t k = 0; k < np && st == LSO_OK; ++k) {
                st = qr_factor_range(ctx, pl, k, k + 1);
                if (st != LSO_OK) break;
                pack_rowblock_kernel<<<pgrid, 256, 0, cs>>>(n, pl->A, pl->ld, pl->Npad, k, ws->d_rowgath + ((size_t)k * P + c) * blk);
                ctx->launches++;
                cudaEventRecord(ws->ev_chunk[(size_t)c * np + k], cs);
            }
            pl->M = M_full;
            r0 += rows;
        }
        ctx->stream = main_stream;
        LSO_TRY(st);
        LSO_CHECK_LAUNCH(ctx);
        LSO_CHECK_CUDA(ctx, cudaEventRecord(ws->chunk_done[0], ws->comm_stream));
        for (int w = 1; w < K; ++w) LSO_CHECK_CUDA(ctx, cudaEventRecord(ws->chunk_done[w], ws->twin_stream[w - 1]));
        ws->chunks_pending = K;
        LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->copy_ev[P], 0));   // J and y are whole on the device
        ws->kept_pipe = true;
        return LSO_OK;
    }
    for (int k = 0; k < P && st == LSO_OK; ++k) {
        const int64_t rows = chunk_rows[k];
        const int w = k % K;                  // 0 = the workspace's own plan on the context stream
        ctx->stream = (w == 0) ? main_stream : ws->twin_stream[w - 1];
        if (cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[1 + k], 0) != cudaSuccess) { st = lso_set_error(ctx, LSO_ERR_CUDA, "cudaStreamWaitEvent failed"); break; }
        st = shard_local_R(ws, d_J + r0, ld_d, d_y + r0, ws->d_gather + (size_t)k * n * (n + 1), rows, (w == 0) ? &ws->plan : &ws->plan_twin[w - 1]);
        r0 += rows;
    }
    ctx->stream = main_stream;
    LSO_TRY(st);
    for (int t = 0; t < K - 1; ++t) {
        LSO_CHECK_CUDA(ctx, cudaEventRecord(ws->twin_done[t], ws->twin_stream[t]));
        LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ws->twin_done[t], 0));
    }
    ws->kept = true; ws->kept_pipe = false;
    return LSO_OK;
}

// uniform chunks of the workspace's m rows, the short remainder chunk first
int lso_qr_factor_keep_host(lso_dense_ws* ws, int64_t m_total, const double* h_J, int64_t ld_h, const double* h_y,
                            double* d_J, int64_t ld_d, double* d_y) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    const int64_t mc = ws->m;
    LSO_REQUIRE(ctx, mc >= 1 && m_total >= 1, "bad dimensions");
    const int P = (int)cdiv64(m_total, mc);
    LSO_REQUIRE(ctx, P >= 1 && P <= 16, "between 1 and 16 row chunks");
    int64_t rows[16];
    const int64_t rem = m_total - (int64_t)(P - 1) * mc;
    rows[0] = rem;
    for (int k = 1; k < P; ++k) rows[k] = mc;
    return lso_qr_factor_keep_host_chunks(ws, P, rows, h_J, ld_h, h_y, d_J, ld_d, d_y);
}

int lso_qr_solve_kept(lso_dense_ws* ws, const double* d_damp, double* d_x, int* rank_out) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, ws->kind == LSO_SOLVER_QR && d_x, "bad arguments");
    if (ws->kept_pipe && ws->have_stack) {
        LSO_ENTER(ctx);
        return stack_solve_from_rowblocks(ws, ws->stack_P, /*n_chunk_events=*/ws->stack_P, d_damp, d_x, rank_out);
    }
    if (!(ws->kept && ws->have_stack))
        return lso_set_error(ctx, LSO_ERR_UNSUPPORTED, "no factor kept: call lso_qr_factor_keep (or a non-pipelined lso_qr_solve_sharded) first");
    LSO_ENTER(ctx);
    return shard_stack_solve(ws, ws->stack_P, d_damp, d_x, rank_out);
}

int lso_qr_kept_invalidate(lso_dense_ws* ws) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    ws->kept = false; ws->kept_pipe = false;
    ws->last_plan = nullptr;
    return LSO_OK;
}

// e[i] = new[i] - last[i]; *flag = 1 when some e[i] is negative beyond rounding (the increment must be >= 0)
__global__ void redamp_increment_kernel(long long n, const double* __restrict__ dnew, const double* __restrict__ dlast,
                                        double* __restrict__ e, int* __restrict__ flag) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double d = dnew[i] - dlast[i];
        if (!(d >= -4.0e-16 * fabs(dlast[i]))) *flag = 1;      // also catches NaN
        e[i] = d > 0.0 ? d : 0.0;
    }
}

// (f3) Re-solve after a REJECTED trust-region step (levenberg_marquardt.jl:77-87: same J, same f, Δ shrunk, i.e. a
// LARGER damping D_new >= D_last elementwise).  The last damped solve on this workspace left R with R'R = J'J + D_last
// and c = Q'[y; 0]; since J'J + D_new = R'R + (D_new - D_last), the triangular factor of the new system is the R of
//     [ R ; sqrt(D_new - D_last) ]   (2n x n, banded: cost independent of m),   right-hand side [c ; 0],
// obtained by orthogonal transformations only — the same least-squares problem min ||[J; sqrt(D_new)] x - [y; 0]|| that the
// reference refactors from scratch (dense_qr.jl:64-88).  No pass over J, no collective on a sharded workspace.
// Returns LSO_ERR_UNSUPPORTED (and leaves x untouched) when there is no damped factor to start from or the damping did
// not grow; the caller then solves the ordinary way.
int lso_qr_solve_redamp(lso_dense_ws* ws, const double* d_damp_new, double* d_x, int* rank_out) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, ws->kind == LSO_SOLVER_QR && d_damp_new && d_x, "bad arguments");
    if (!ws->last_plan) return lso_set_error(ctx, LSO_ERR_UNSUPPORTED, "no damped factor to re-damp");
    LSO_ENTER(ctx);
    LSO_TRY(ws_join_chunks(ws));
    const int64_t n = ws->n;
    if (!ws->have_redamp) {
        LSO_TRY(qr_plan_create(ctx, 2 * n, n, &ws->plan_redamp));
        ws->have_redamp = true;
        ws->plan_redamp.band = 2;
        LSO_TRY(qr_plan_tune(ctx, &ws->plan_redamp));
        LSO_CHECK_CUDA(ctx, cudaMalloc(&ws->d_redamp_slot, ((size_t)n * (n + 1) + (size_t)n + 2) * sizeof(double)));
    }
    double* slot = ws->d_redamp_slot;
    double* inc = slot + (size_t)n * (n + 1);
    int* flag = (int*)(inc + n);
    LSO_CHECK_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    redamp_increment_kernel<<<(unsigned)std::min<int64_t>(cdiv64(n, 256), 1024), 256, 0, ctx->stream>>>(n, d_damp_new, ws->d_lastdamp, inc, flag);
    LSO_CHECK_LAUNCH(ctx);
    int h_flag = 0;
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_flag) return lso_set_error(ctx, LSO_ERR_UNSUPPORTED, "re-damping needs damp_new >= damp_last elementwise");
    QRPlan* src = ws->last_plan;
    {
        dim3 grid((unsigned)std::min<int64_t>(cdiv64(n, 256), 64), (unsigned)(n + 1));
        pack_R_kernel<<<grid, 256, 0, ctx->stream>>>(n, src->A, src->ld, src->Npad, slot);
        LSO_CHECK_LAUNCH(ctx);
    }
    QRPlan* ps = &ws->plan_redamp;
    {
        dim3 grid((unsigned)std::min<int64_t>(cdiv64(ps->ld, 256), 64), (unsigned)ps->Nc);
        stack_assemble_kernel<<<grid, 256, 0, ctx->stream>>>(n, 1, 2, slot, inc, ps->A, ps->ld, ps->Npad);
        LSO_CHECK_LAUNCH(ctx);
    }
    ps->M = 2 * n;
    ps->band = 2;
    ws->last_plan = nullptr;
    LSO_TRY(qr_factor(ctx, ps));
    LSO_TRY(qr_finish(ws, ps, d_x, rank_out, false));
    return remember_damped_factor(ws, ps, d_damp_new);
}

// Allocate everything a row-sharded solve will need (the stack workspace and the gather buffer) WITHOUT any collective, so
// that an allocation failure on one rank can be detected and agreed on by the host program before the first all-gather
// (a rank that fails inside lso_qr_solve_sharded would leave the others waiting in NCCL).
int lso_qr_prepare_sharded(lso_dense_ws* ws) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, ws->kind == LSO_SOLVER_QR && ws->damped == 0, "create the workspace for QR with damped = 0");
    LSO_ENTER(ctx);
    return shard_ensure_stack(ws, ctx->nranks > 1 ? ctx->nranks : 1);
}

// Test hook: the sharded algorithm with the P shards emulated on ONE device (the rows of J are cut into P equal
// chunks of ws->m rows that are factorised one after the other; no communicator needed).  d_J is (P * ws->m) x n.
int lso_debug_qr_solve_emulated_shards(lso_dense_ws* ws, int P, const double* d_J, int64_t ld, const double* d_y,
                                       const double* d_damp, double* d_x, int* rank_out) {
    if (!ws) return lso_set_error(nullptr, LSO_ERR_ARG, "ws is NULL");
    lso_ctx* ctx = ws->ctx;
    LSO_REQUIRE(ctx, ws->kind == LSO_SOLVER_QR && ws->damped == 0, "create the workspace for QR with damped = 0");
    LSO_REQUIRE(ctx, P >= 1 && P <= 64 && d_J && d_y && d_x, "bad arguments");
    LSO_REQUIRE(ctx, ws->m >= ws->n, "each shard needs rows >= columns");
    LSO_ENTER(ctx);
    LSO_TRY(ws_join_chunks(ws));
    const int64_t n = ws->n;
    if (ctx->opt_qr_shard_pipeline && ws->plan.Npad / QB >= 2)
        return qr_solve_sharded_pipelined(ws, P, true, d_J, ld, d_y, d_damp, d_x, rank_out);
    LSO_TRY(shard_ensure_stack(ws, P));
    for (int k = 0; k < P; ++k)
        LSO_TRY(shard_local_R(ws, d_J + (size_t)k * ws->m, ld, d_y + (size_t)k * ws->m, ws->d_gather + (size_t)k * n * (n + 1)));
    return shard_stack_solve(ws, P, d_damp, d_x, rank_out);
}

}  // extern "C"

/*
 * lsob200.h — C ABI of the B200-native inner solver for LeastSquaresOptim.jl's hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, returns an `int` status
 * (0 = ok, >0 = LAPACK-style `info`, <0 = argument / CUDA / NCCL error; text via
 * lso_last_error).  No C++ exception crosses this boundary.  All floating point is IEEE
 * binary64 ("Float64" on the Julia side).  Dense matrices are column-major with an explicit
 * leading dimension (Julia `Matrix{Float64}`: ld = size(J,1)).  Sparse matrices are CSC with
 * Julia's Int64 1-based `colptr` / `rowval` (converted once per pattern on import).
 *
 * Each group cites the reference interface it replaces (paths under the reference checkout,
 * matthieugomez/LeastSquaresOptim.jl v0.8.10).  INTEGRATION.md shows the Julia `ccall`
 * binding for each.
 *
 * Memory model: `double*` arguments named `d_*` are DEVICE pointers obtained from
 * lso_dev_alloc (resident mode: J, x, f, δ live in HBM across iterations).  The `*_host`
 * entry points take HOST pointers, borrowed for the duration of the call.
 * A context is single-caller (not thread-safe), one CUDA stream, one device.
 */
#ifndef LSOB200_H
#define LSOB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lso_ctx lso_ctx;           /* device + stream + scratch + (optional) NCCL comm */
typedef struct lso_dense_ws lso_dense_ws; /* workspace of DenseQRAllocatedSolver / DenseCholeskyAllocatedSolver */
typedef struct lso_csc lso_csc;           /* device image of a SparseMatrixCSC{Float64,Int64} (+ CSR mirror) */
typedef struct lso_lsmr_ws lso_lsmr_ws;   /* workspace of LSMRAllocatedSolver / LSMRDampenedAllocatedSolver */

/* ---- status codes ------------------------------------------------------------------- */
#define LSO_OK                 0
#define LSO_ERR_ARG           (-1)   /* DimensionMismatch / ArgumentError on the Julia side   */
#define LSO_ERR_CUDA          (-2)
#define LSO_ERR_NCCL          (-3)
#define LSO_ERR_ALLOC         (-4)
#define LSO_ERR_UNSUPPORTED   (-5)
#define LSO_ERR_NOT_FINITE    (-6)   /* IsFiniteException (src/utils/utils.jl:63-78)          */
/* positive return = LAPACK info: PosDefException(info) (dense_cholesky.jl:57),
 * RankDeficientException(info) (dense_cholesky.jl:33).                                        */

/* ---- library / context -------------------------------------------------------------- */
int         lso_version(void);
int         lso_device_count(int* count);
int         lso_ctx_create(int device, lso_ctx** out);
int         lso_ctx_destroy(lso_ctx* ctx);
const char* lso_last_error(lso_ctx* ctx);          /* ctx may be NULL: last global error */
int         lso_ctx_sync(lso_ctx* ctx);
void*       lso_ctx_stream(lso_ctx* ctx);          /* cudaStream_t, for CUDA-event timing by the caller */
/* Options (debug / cross-check switches; the defaults are the measured-fastest paths):
 *   "qr_apply"     0 plain-FMA trailing update, 1 first-generation DMMA kernel, 2 ping-pong DMMA kernel with one launch
 *                  per tree level (default), 3 the same kernel with all tree levels of a panel in one launch,
 *                  4 levels 0 and 1 one launch each and the levels above chained in one launch
 *   "qr_lookahead" 1 = panel trees on a second stream under the previous update (default 0)
 *   "qr_tune"      1 (default) = QR workspaces whose panel tree fits on a third of the SMs (row shards, the stacked R
 *                  factors) time four launch schedules once at creation and keep the fastest; setting "qr_apply" or
 *                  "qr_lookahead" explicitly turns this off
 *   "qr_shard_pipeline" 1 (default) = lso_qr_solve_sharded runs the replicated stack QR panel by panel behind the local
 *                  factorisation, fed by one small all-gather per panel; 0 = local QR, one all-gather, stack QR.
 *                  The same switch pipelines lso_qr_factor_keep_host[_chunks] + lso_qr_solve_kept: the chunks are
 *                  factorised panel by panel on streams of their own and the stack QR follows the last one.
 *   "qr_twin"      extra workspaces / streams (0..3, default 2) among which lso_qr_factor_keep_host[_chunks] rotates its
 *                  chunks when there are three or more (the panel trees of one chunk run under the updates of another)
 *   "syrk"         0 plain-FMA syrk, 1 DMMA syrk, 3 (default) tcgen05 syrk for m >= 8192 and n >= 512 and DMMA otherwise,
 *                  2 tcgen05 syrk always: int8 digit matrices (Ozaki scheme) multiplied by
 *                  tcgen05.mma.kind::i8 into TMEM, operands fed by TMA, fp64 reconstructed exactly
 *   "ozaki_slices" 7-bit digits per value for "syrk" = 2 (2..8, default 8: representation error 2^-57)
 *   "trisolve"     1 (default) = the triangular solves behind every dense solve run on n/32 CTAs (diagonal blocks chained
 *                  through an L2 mailbox), 0 = the single-CTA left-looking kernel
 *   "spmv"         0 first-generation sparse products, 1 stream kernels (shared-memory staging), 2 warp kernels (default)
 *   "lsmr_fused"   0 LSMR with host-side scalars (3 syncs per iteration), 1 fused device-resident LSMR (default)
 *   "profile"      see lso_ctx_profile_read */
int         lso_ctx_set_option(lso_ctx* ctx, const char* key, int64_t value);
int         lso_ctx_launch_count(lso_ctx* ctx, int64_t* out, int reset); /* kernels launched by this library */
/* With option "profile" = 1 every launch of the dominant kernel of a solve (QR trailing update, syrk, SpMV pair) is
 * bracketed by CUDA events on the context stream; this returns their summed duration and count, and resets. */
int         lso_ctx_profile_read(lso_ctx* ctx, double* total_ms, int64_t* launches);
/* Algorithmic work issued through this context since the last reset (the numerators of bench.py's rooflines):
 *   "qr_update_flops"  trailing-update flops of the QR factorisations (sum over panels of 4*32*active rows*trailing columns)
 *   "qr_flops"         2 M n^2 - 2/3 n^3 per factorisation          "syrk_flops"  m n (n+1) per J'J
 *   "syrk_i8_macs"     int8 multiply-accumulates issued to tcgen05 by the digit-matrix syrk (S(S+1)/2 digit products per tile)
 *   "spmv_bytes"       12 B per stored entry + 8 B per vector element per sparse product */
int         lso_ctx_stat(lso_ctx* ctx, const char* key, double* out, int reset);
/* second channel of the same option: the collective of a sharded solve (ncclAllReduce / ncclAllGather) */
int         lso_ctx_profile_read_collective(lso_ctx* ctx, double* total_ms, int64_t* launches);

/* ---- device memory (Julia GC owns nothing on the device; finalizers call *_free) ----- */
int lso_dev_alloc(lso_ctx* ctx, size_t nbytes, void** d_out);
int lso_dev_free(lso_ctx* ctx, void* d_ptr);
int lso_host_alloc_pinned(lso_ctx* ctx, size_t nbytes, void** h_out);
int lso_host_free_pinned(lso_ctx* ctx, void* h_ptr);
/* page-lock / release memory the caller owns (Julia Arrays are pageable; J, x, y live for the whole optimize! run,
 * types.jl:141-157, so the glue registers them once and every H2D copy of J then runs at pinned-memory speed) */
int lso_host_register(lso_ctx* ctx, void* h_ptr, size_t nbytes);
int lso_host_unregister(lso_ctx* ctx, void* h_ptr);
int lso_upload(lso_ctx* ctx, void* d_dst, const void* h_src, size_t nbytes);     /* synchronous */
int lso_download(lso_ctx* ctx, void* h_dst, const void* d_src, size_t nbytes);   /* synchronous */
int lso_upload_async(lso_ctx* ctx, void* d_dst, const void* h_src, size_t nbytes);
int lso_download_async(lso_ctx* ctx, void* h_dst, const void* d_src, size_t nbytes);
int lso_upload_matrix(lso_ctx* ctx, double* d_dst, int64_t ld_dst, const double* h_src, int64_t ld_src,
                      int64_t rows, int64_t cols);

/* ---- vector duck type: src/utils/lsmr.jl:30-44, and what the optimizers call on x / f
 *      (levenberg_marquardt.jl:60,84-86,89-98,106,111,116-117,127,135; dogleg.jl:68,90,93,105-106,
 *       114,117,122-145,148-160,168,173-174,190; utils.jl:24,39-55,165-176)               ---- */
int lso_vec_fill(lso_ctx* ctx, int64_t n, double* d_x, double value);                 /* fill!     */
int lso_vec_copy(lso_ctx* ctx, int64_t n, double* d_dst, const double* d_src);        /* copyto!   */
int lso_vec_scal(lso_ctx* ctx, int64_t n, double* d_x, double alpha);                 /* rmul!     */
int lso_vec_axpy(lso_ctx* ctx, int64_t n, double alpha, const double* d_x, double* d_y);          /* axpy! */
int lso_vec_axpby(lso_ctx* ctx, int64_t n, double alpha, const double* d_x, double beta, double* d_y);
int lso_vec_mul(lso_ctx* ctx, int64_t n, double* d_out, const double* d_x, const double* d_y);    /* map!(*,…) */
int lso_vec_div(lso_ctx* ctx, int64_t n, double* d_out, const double* d_x, const double* d_y);    /* map!(/,…) dogleg.jl:105 */
int lso_vec_sqrt(lso_ctx* ctx, int64_t n, double* d_x);                                           /* map!(sqrt,…) iterative_lsmr.jl:252 */
int lso_vec_clamp(lso_ctx* ctx, int64_t n, double* d_x, double lo, double hi);                    /* clamp!    */
int lso_vec_sum(lso_ctx* ctx, int64_t n, const double* d_x, double* out);                         /* sum       */
int lso_vec_sumabs2(lso_ctx* ctx, int64_t n, const double* d_x, double* out);                     /* sum(abs2,·) */
int lso_vec_nrm2(lso_ctx* ctx, int64_t n, const double* d_x, double* out);                        /* norm      */
int lso_vec_maxabs(lso_ctx* ctx, int64_t n, const double* d_x, double* out);                      /* maximum(abs,·) */
int lso_vec_dot(lso_ctx* ctx, int64_t n, const double* d_x, const double* d_y, double* out);
int lso_vec_wdot(lso_ctx* ctx, int64_t n, const double* d_x, const double* d_y, const double* d_w, double* out); /* utils.jl:165-173 */
int lso_vec_check_finite(lso_ctx* ctx, int64_t n, const double* d_x, int64_t* first_bad /* -1 if none */);      /* utils.jl:70-75 */
/* box projection of the step: δ[i] = min(δ[i], x[i]-lower[i]); δ[i] = max(δ[i], x[i]-upper[i])
 * (levenberg_marquardt.jl:89-98, dogleg.jl:148-157). d_lower / d_upper may be NULL. */
int lso_vec_box_project(lso_ctx* ctx, int64_t n, double* d_delta, const double* d_x,
                        const double* d_lower, const double* d_upper);
/* maxabs_projected_gradient (utils.jl:39-55). d_lower / d_upper may be NULL. */
int lso_vec_maxabs_projected(lso_ctx* ctx, int64_t n, const double* d_g, const double* d_x,
                             const double* d_lower, const double* d_upper, double* out);
/* LM damping build, levenberg_marquardt.jl:84-86:
 *   mean = sum(dtd)/n; clamp!(dtd, min_diag*mean, max_diag*mean); rmul!(dtd, inv_delta)   */
int lso_lm_damping(lso_ctx* ctx, int64_t n, double* d_dtd, double min_diag, double max_diag, double inv_delta);
/* Dogleg step blend, dogleg.jl:120-145; returns wnorm_δx. */
int lso_dogleg_blend(lso_ctx* ctx, int64_t n, double* d_dx, const double* d_gn, const double* d_gr,
                     const double* d_dtd, double delta, double alpha, double wnorm_gn, double wnorm_gr,
                     double* wnorm_dx_out);

/* ---- dense operator: mul!(y,J,x,α,β), mul!(x,J',y,α,β), colsumabs2!(x,J)
 *      (README.md:37-43; utils.jl:139-144; levenberg_marquardt.jl:82,102,114; dogleg.jl:85,99,109,171) */
int lso_dense_colsumabs2(lso_ctx* ctx, int64_t m, int64_t n, const double* d_J, int64_t ld, double* d_out);
int lso_dense_gemv_n(lso_ctx* ctx, int64_t m, int64_t n, double alpha, const double* d_J, int64_t ld,
                     const double* d_x, double beta, double* d_y);
int lso_dense_gemv_t(lso_ctx* ctx, int64_t m, int64_t n, double alpha, const double* d_J, int64_t ld,
                     const double* d_y, double beta, double* d_x);
/* fused: dtd = colsumabs2(J) and g = J'f in ONE pass over J (LM:82 + LM:102; dogleg:85 + :99) */
int lso_dense_colsumabs2_gemv_t(lso_ctx* ctx, int64_t m, int64_t n, const double* d_J, int64_t ld,
                                const double* d_f, double* d_dtd, double* d_g);
/* fused: fpredict = J*δ - f ; *ssr_out = sum(abs2, fpredict)  (LM:114-117; dogleg:171-174).
 * d_fpredict may be NULL (only the scalar is wanted). */
int lso_dense_predicted_ssr(lso_ctx* ctx, int64_t m, int64_t n, const double* d_J, int64_t ld,
                            const double* d_delta, const double* d_f, double* d_fpredict, double* ssr_out);

/* ---- dense solvers: DenseQRAllocatedSolver (src/solver/dense_qr.jl:25-28, 50-54) and
 *      DenseCholeskyAllocatedSolver (src/solver/dense_cholesky.jl:19-21)                     ---- */
#define LSO_SOLVER_QR        1
#define LSO_SOLVER_CHOLESKY  2
/* damped != 0: LevenbergMarquardt workspace ((m+n) x n augmented system), else Dogleg workspace. */
int lso_dense_ws_create(lso_ctx* ctx, int64_t m, int64_t n, int solver_kind, int damped, lso_dense_ws** out);
int lso_dense_ws_destroy(lso_dense_ws* ws);
/* ldiv!(x, J, y, damp, A::DenseQRAllocatedSolver)  — dense_qr.jl:56-88  (d_damp != NULL)
 * ldiv!(x, J, y, A::DenseQRAllocatedSolver)        — dense_qr.jl:30-42  (d_damp == NULL)
 * Solves min ||[J; diag(sqrt(damp))] x - [y; 0]||_2 by Householder QR.  J, y, damp untouched.
 * rank_out (may be NULL) receives the numerical rank found (dgelsy-style, rcond = min(rows,cols)*eps). */
int lso_qr_solve(lso_dense_ws* ws, const double* d_J, int64_t ld, const double* d_y,
                 const double* d_damp, double* d_x, int* rank_out);
/* ldiv!(x, J, y, damp, A::DenseCholeskyAllocatedSolver) — dense_cholesky.jl:43-59 (d_damp != NULL)
 * ldiv!(x, J, y, A::DenseCholeskyAllocatedSolver)       — dense_cholesky.jl:29-35 (d_damp == NULL)
 * Always a LOCAL solve of the J it is given (also on a context that has a communicator); the row-sharded form is
 * lso_chol_solve_sharded.  Returns info > 0 when the matrix is not positive definite / rank deficient. */
int lso_chol_solve(lso_dense_ws* ws, const double* d_J, int64_t ld, const double* d_y,
                   const double* d_damp, double* d_x);
/* Host-buffer forms (what `ldiv!` on plain Julia Arrays binds): copy J, y, damp H2D, solve, copy x D2H. */
int lso_qr_solve_host(lso_dense_ws* ws, const double* h_J, int64_t ld, const double* h_y,
                      const double* h_damp, double* h_x, int* rank_out);
int lso_chol_solve_host(lso_dense_ws* ws, const double* h_J, int64_t ld, const double* h_y,
                        const double* h_damp, double* h_x);
/* introspection for tests: copy the n x n upper-triangular factor (column-major, ld = n) to the host */
int lso_dense_ws_get_factor(lso_dense_ws* ws, double* h_R);

/* ---- multi-GPU (row-sharded dense J): one process per GPU, NCCL over NVLink.
 *      No reference counterpart (the reference is single-process); see DESIGN.md §multi-GPU. ---- */
int lso_comm_unique_id(void* id128 /* 128 bytes out */);
int lso_comm_init_rank(lso_ctx* ctx, int nranks, int rank, const void* id128);
int lso_comm_destroy(lso_ctx* ctx);
int lso_comm_allreduce_sum(lso_ctx* ctx, double* d_buf, int64_t count);
/* Row-sharded Cholesky path (BASELINE.json configs[3]): every rank passes its row shard of J and y; the packed
 * [upper(J'J) by columns | J'y] (n(n+1)/2 + n doubles) is summed by ONE ncclAllReduce per solve, the factorisation and
 * the solves are replicated, all ranks get x.  Math of dense_cholesky.jl:43-59 / :29-35 on the whole J. */
int lso_chol_solve_sharded(lso_dense_ws* ws, const double* d_J, int64_t ld, const double* d_y,
                           const double* d_damp, double* d_x);
/* (f3) re-solve of a rejected step: J'J and J'f of the last lso_chol_solve / lso_chol_solve_sharded (kept packed,
 * before damping) with a new damping — no pass over J, no collective (levenberg_marquardt.jl:77-87). */
int lso_chol_solve_kept(lso_dense_ws* ws, const double* d_damp, double* d_x);
/* Test hook: the same algorithm (partial products per shard, packed, summed in rank order) with the P shards emulated on
 * one device: d_J is (P * m) x n where m is the workspace's shard row count. */
int lso_debug_chol_solve_emulated_shards(lso_dense_ws* ws, int P, const double* d_J, int64_t ld, const double* d_y,
                                         const double* d_damp, double* d_x);
/* (f3) Re-solves of a rejected trust-region step (levenberg_marquardt.jl:77-87 solves again with the same J and f and a
 * new damping; the reference refactors everything, dense_qr.jl:64-88).  lso_qr_factor_keep does the QR of [J | y] once
 * and keeps the n x (n+1) block [R_J | Q'y]; lso_qr_solve_kept solves min ||[J; sqrt(D)] x - [y; 0]|| from it by a QR
 * of the banded 2n x n stack [R_J; sqrt(D)] (cost independent of m; d_damp == NULL solves the undamped problem).  On a
 * row-sharded workspace the factors gathered by the last lso_qr_solve_sharded are kept: the re-solve has no collective.
 * lso_qr_kept_invalidate: J or y changed. */
int lso_qr_factor_keep(lso_dense_ws* ws, const double* d_J, int64_t ld, const double* d_y);
int lso_qr_solve_kept(lso_dense_ws* ws, const double* d_damp, double* d_x, int* rank_out);
/* lso_qr_factor_keep fed from HOST memory (what `ldiv!` on plain Julia Arrays needs): J (m_total x n, ld_h) and y are
 * copied in row chunks of the workspace's m rows on a copy stream and each chunk is factorised as soon as it has landed
 * (TSQR over the chunks), so the PCIe transfer runs under the factorisation; J and y also end up in d_J (ld_d) / d_y.
 * Follow with lso_qr_solve_kept (the damping is only needed there, so colsumabs2!(dtd, J) can wait for the whole J). */
int lso_qr_factor_keep_host(lso_dense_ws* ws, int64_t m_total, const double* h_J, int64_t ld_h, const double* h_y,
                            double* d_J, int64_t ld_d, double* d_y);
/* the same with an explicit list of chunk sizes (each <= the workspace's m, sum = rows of J), sent in that order; with
 * three or more chunks they are factorised round-robin in up to "qr_twin" + 1 workspaces on as many streams.
 * With "qr_shard_pipeline" (default) the call returns with the context stream ordered after the COPIES only: d_J and
 * d_y are whole for the caller's passes (colsumabs2!, J'f), the factorisations are still running on their own streams,
 * and lso_qr_solve_kept joins them panel by panel.  Any other call on the workspace joins them first. */
int lso_qr_factor_keep_host_chunks(lso_dense_ws* ws, int P, const int64_t* chunk_rows, const double* h_J, int64_t ld_h,
                                   const double* h_y, double* d_J, int64_t ld_d, double* d_y);
int lso_qr_kept_invalidate(lso_dense_ws* ws);
/* Cheaper still, and from the FIRST rejection on: the last damped solve on the workspace left R with R'R = J'J + D_last.
 * A rejected LM step re-solves with the same J, f and a LARGER damping (Delta shrinks, levenberg_marquardt.jl:77-87,135),
 * and J'J + D_new = R'R + (D_new - D_last): lso_qr_solve_redamp factors the banded 2n x n stack
 * [R; sqrt(D_new - D_last)] (right-hand side [Q'y; 0]) — orthogonal transformations only, the same least-squares problem
 * the reference refactors from scratch, at a cost independent of m.  Returns LSO_ERR_UNSUPPORTED (x untouched) when no
 * damped factor is available (J or y changed: call lso_qr_kept_invalidate) or the damping did not grow elementwise. */
int lso_qr_solve_redamp(lso_dense_ws* ws, const double* d_damp_new, double* d_x, int* rank_out);
/* TSQR over row shards for the QR path: every rank passes its shard of J and y; all ranks get x. */
int lso_qr_solve_sharded(lso_dense_ws* ws, const double* d_J, int64_t ld, const double* d_y,
                         const double* d_damp, double* d_x, int* rank_out);
/* allocates the stack workspace / gather buffer of lso_qr_solve_sharded up front, with no collective, so that the host
 * program can agree on an allocation failure before any rank enters NCCL */
int lso_qr_prepare_sharded(lso_dense_ws* ws);
/* Test hook: the sharded algorithm (local QR per shard, interleaved stack of the R factors, banded QR of the stack)
 * with the P shards emulated on one device: d_J is (P * m) x n where m is the workspace's shard row count. */
int lso_debug_qr_solve_emulated_shards(lso_dense_ws* ws, int P, const double* d_J, int64_t ld, const double* d_y,
                                       const double* d_damp, double* d_x, int* rank_out);

/* ---- sparse operator: SparseMatrixCSC mul! both ways + colsumabs2! (utils.jl:146-151;
 *      lsmr.jl:73,76,118,122; levenberg_marquardt.jl:82,102,114)                             ---- */
int lso_csc_create(lso_ctx* ctx, int64_t m, int64_t n, int64_t nnz,
                   const int64_t* h_colptr_1based, const int64_t* h_rowval_1based, lso_csc** out);
int lso_csc_destroy(lso_csc* A);
/* g! changed the sparsity PATTERN (setindex! into a sparse J, test/nonlinearsolvers.jl:526-530): re-import colptr /
 * rowval (same m, n; nnz may differ), rebuild the CSR mirror and the CTA partitions; values are zeroed. */
int lso_csc_update_pattern(lso_csc* A, int64_t nnz, const int64_t* h_colptr_1based, const int64_t* h_rowval_1based);
int lso_csc_set_values_host(lso_csc* A, const double* h_nzval);     /* after g!(J,x) wrote nonzeros(J) */
int lso_csc_set_values_dev(lso_csc* A, const double* d_nzval);
double* lso_csc_values(lso_csc* A);                                  /* device nzval (CSC order), writable by device g! */
int lso_csc_values_changed(lso_csc* A);                              /* refresh the CSR mirror after writing lso_csc_values */
/* A device g! may write BOTH images and skip the mirror gather (12 + 16 B per entry of random traffic):
 * lso_csc_values_csr is nzval in CSR order (entry k of it is entry lso_csc_csr_perm[k] of lso_csc_values; int32, 0-based
 * rowptr / colidx), lso_csc_values_changed_both declares both images current. lso_csc_gather_csr permutes any
 * nnz-vector from CSC into CSR order (e.g. constant factors of the Jacobian, once per pattern). */
double* lso_csc_values_csr(lso_csc* A);
const int* lso_csc_csr_rowptr(lso_csc* A);
const int* lso_csc_csr_colidx(lso_csc* A);
const int* lso_csc_csr_perm(lso_csc* A);
int lso_csc_values_changed_both(lso_csc* A);
int lso_csc_gather_csr(lso_csc* A, const double* d_in_csc_order, double* d_out_csr_order);
int lso_csc_mul_n(lso_csc* A, double alpha, const double* d_x, double beta, double* d_y);   /* y = α J x + β y  */
int lso_csc_mul_t(lso_csc* A, double alpha, const double* d_y, double beta, double* d_x);   /* x = α J'y + β x  */
int lso_csc_colsumabs2(lso_csc* A, double* d_out);                   /* cached per J (utils.jl:146-151) */
/* fused: dtd = colsumabs2(J) and g = J'f in ONE pass over the CSC image (LM:82 + LM:102) */
int lso_csc_colsumabs2_gemv_t(lso_csc* A, const double* d_f, double* d_dtd, double* d_g);
/* fused: fpredict = J*δ - f ; *ssr_out = sum(abs2, fpredict)  (LM:114-117; dogleg:171-174). d_fpredict may be NULL. */
int lso_csc_predicted_ssr(lso_csc* A, const double* d_delta, const double* d_f, double* d_fpredict, double* ssr_out);

/* ---- LSMR solvers: src/solver/iterative_lsmr.jl:161-198 (undamped) and :221-259 (damped),
 *      running src/utils/lsmr.jl:53-238 on the device.  Exactly one of (A_csc) or (d_J, ld) is
 *      given: the operator is the CSC image or a dense column-major J.                        ---- */
int lso_lsmr_ws_create(lso_ctx* ctx, int64_t m, int64_t n, int damped, lso_lsmr_ws** out);
int lso_lsmr_ws_destroy(lso_lsmr_ws* ws);
/* d_damp == NULL: ldiv!(x,J,y,A::LSMRAllocatedSolver) (atol=btol=1e-6 by default);
 * d_damp != NULL: ldiv!(x,J,y,damp,A::LSMRDampenedAllocatedSolver) (btol=0.5 at iterative_lsmr.jl:255);
 *                 damp is overwritten by sqrt(damp) exactly as at iterative_lsmr.jl:252.
 * maxiter <= 0 selects the reference default max(rows, cols) (lsmr.jl:55).
 * iters_out: LSMR iterations (the reference returns mvps = 2*iters); istop_out: lsmr.jl:224-231. */
int lso_lsmr_solve(lso_lsmr_ws* ws, lso_csc* A_csc, const double* d_J, int64_t ld,
                   const double* d_y, double* d_damp, double* d_x,
                   double atol, double btol, double conlim, int64_t maxiter,
                   int64_t* iters_out, int* istop_out);
/* User preconditioner (README.md:47; iterative_lsmr.jl:143-145, 216-218: `preconditioner!(P, x, J, λ)` has been run
 * by the caller): either d_P_diag, the vector of an InverseDiagonal (iterative_lsmr.jl:117-122; stays on the fused
 * sparse path), or precond_apply, `ldiv!(out, P, in)` as a callback on device pointers that enqueues its work on
 * lso_ctx_stream (generic path: the reference's wrappers op for op).  Both NULL = the default preconditioner. */
typedef int (*lso_precond_fn)(void* user, int64_t n, const double* d_in, double* d_out);
int lso_lsmr_solve_ex(lso_lsmr_ws* ws, lso_csc* A_csc, const double* d_J, int64_t ld,
                      const double* d_y, double* d_damp, double* d_x,
                      double atol, double btol, double conlim, int64_t maxiter,
                      const double* d_P_diag, lso_precond_fn precond_apply, void* precond_user,
                      int64_t* iters_out, int* istop_out);
/* (f4) LSMR on a row-sharded J, one process per GPU (needs lso_comm_init_rank): A_csc / d_y are THIS rank's rows of J
 * and y (the workspace's m = the rank's row count), d_damp / d_x / d_P_diag are replicated.  u is sharded like the rows;
 * v, h, hbar, x and the damping part of u are replicated.  Per iteration (iterative_lsmr.jl:30-51, lsmr.jl:116-156):
 * local J_k (P.v), local J_k'u_k with u not yet normalised, ONE ncclAllReduce of [J'u (n) | ||u||^2 (1)], then beta,
 * the new v, alpha, the rotations and the stopping tests replicated on every rank: 4 launches + 1 collective per
 * iteration, the same <= 1 host synchronisation as the single-GPU form.  m_total = rows of the whole J (default maxiter,
 * lsmr.jl:55; 0 = unknown).  Same answer as lso_lsmr_solve on the whole J up to the order of the row sums.  On a context
 * without a communicator this IS lso_lsmr_solve_ex. */
int lso_lsmr_solve_sharded(lso_lsmr_ws* ws, lso_csc* A_csc, const double* d_y, double* d_damp, double* d_x,
                           double atol, double btol, double conlim, int64_t maxiter, int64_t m_total,
                           const double* d_P_diag, int64_t* iters_out, int* istop_out);
/* kernel launches and host synchronisations of the last solve on this workspace (lsmr.jl:116-231 runs on the device:
 * 3 launches per iteration, at most one synchronisation per iteration) */
int lso_lsmr_ws_stats(lso_lsmr_ws* ws, int64_t* launches_out, int64_t* syncs_out);

/* ---- (f1) the four reductions a trust-region iteration reads, with ONE host synchronisation
 *      (levenberg_marquardt.jl:104,110,114-117, dogleg.jl:101,168,171-174, utils.jl:21).  lso_lm_gradient_norm_async
 *      enqueues maxabs_projected_gradient(g, x, lower, upper) (utils.jl:38-55) where the reference computes it (before x
 *      moves); lso_lm_step_tail enqueues sum(abs2, ftrial), fpredict = J*dx - fcur with sum(abs2, fpredict) and
 *      maximum(abs, dx), all-reduces the two m-dimension sums over the ranks when `allreduce` != 0 (row-sharded J) and
 *      returns out4 = {trial ssr, predicted ssr, maxabs dx, maxabs projected gradient}. ---- */
int lso_lm_gradient_norm_async(lso_ctx* ctx, int64_t n, const double* d_g, const double* d_x, const double* d_lower,
                               const double* d_upper);
int lso_lm_step_tail(lso_ctx* ctx, int64_t m, int64_t n, const double* d_J, int64_t ld, lso_csc* A_csc,
                     const double* d_dx, const double* d_fcur, const double* d_ftrial, double* d_fpredict, int allreduce,
                     double* out4);

/* ---- (f2) device Jacobian producer: `autodiff = :central` of LeastSquaresProblem (src/types.jl:54-58: FiniteDiff's
 *      finite_difference_jacobian! with step cbrt(eps) * max(1, |x_j|)) for a residual f!(out, x) given as a callback on
 *      DEVICE pointers that enqueues its work on lso_ctx_stream.  d_x is perturbed in place and restored; d_work holds
 *      2 m doubles.  2 n residual evaluations, J is written on the device. ---- */
typedef int (*lso_residual_fn)(void* user, const double* d_x, double* d_out);
int lso_fd_jacobian_central(lso_ctx* ctx, int64_t m, int64_t n, lso_residual_fn f, void* user, double* d_x,
                            double* d_J, int64_t ld, double* d_work);

/* ---- synthetic workload generators and residual models used by bench.py / tests
 *      (counter-based hash, bit-identical on CPU and GPU).  Harness, not boundary. ---- */
int lso_synth_dense_matrix(lso_ctx* ctx, int64_t m, int64_t n, int64_t row_offset, uint64_t seed,
                           double* d_A, int64_t ld);
int lso_synth_vector(lso_ctx* ctx, int64_t n, int64_t offset, uint64_t seed, double scale, double* d_x);
/* r = t + c t^2 - b with t = A x ;  J = diag(1 + 2 c t) A */
int lso_synth_residual(lso_ctx* ctx, int64_t m, int64_t n, const double* d_A, int64_t ld, const double* d_x,
                       const double* d_b, double c, double* d_t, double* d_r);
int lso_synth_residual_from_t(lso_ctx* ctx, int64_t m, const double* d_t, const double* d_b, double c, double* d_r);
int lso_synth_jacobian(lso_ctx* ctx, int64_t m, int64_t n, const double* d_A, int64_t ld, const double* d_t,
                       double c, double* d_J, int64_t ldJ);
int lso_synth_csc_pattern(int64_t m, int64_t n, int64_t nnz_per_col, uint64_t seed,
                          int64_t* h_colptr_1based, int64_t* h_rowval_1based); /* host, deterministic */
/* the same with the rows of column j confined to a window of `window` rows around j*m/n (a Jacobian with locality) */
int lso_synth_csc_pattern_banded(int64_t m, int64_t n, int64_t nnz_per_col, int64_t window, uint64_t seed,
                                 int64_t* h_colptr_1based, int64_t* h_rowval_1based);
int lso_synth_csc_jacobian(lso_csc* A, const double* d_Aval, const double* d_t, double c); /* nzval = (1+2c t[row]) * Aval */
/* the same device g! writing both the CSC and the CSR image (d_Aval_csr = lso_csc_gather_csr of d_Aval) */
int lso_synth_csc_jacobian_both(lso_csc* A, const double* d_Aval, const double* d_Aval_csr, const double* d_t, double c);

/* ---- micro-benchmarks that give the roofline denominators this library reports against ---- */
int lso_bench_fp64_mma_peak(lso_ctx* ctx, int iters, double* tflops_out);   /* DMMA m8n8k4 issue-bound loop */
int lso_bench_fp64_fma_peak(lso_ctx* ctx, int iters, double* tflops_out);   /* DFMA issue-bound loop */
/* DMMA with the trailing-update kernel's register pattern (operands change every k-step) at 8/16/32 warps per SM */
int lso_bench_fp64_mma_pattern(lso_ctx* ctx, int iters, int mode, double* tflops_out);
int lso_bench_hbm_copy(lso_ctx* ctx, size_t nbytes, int iters, double* gbs_out);

#ifdef __cplusplus
}
#endif
#endif /* LSOB200_H */

// synth.cu — bench/test harness, NOT part of the reference boundary: counter-based synthetic problem
// generators (bit-identical on CPU — see oracle/synth_ref.py — and GPU), the polynomial residual model
// r(x) = t + c t^2 - b, t = A x, with Jacobian diag(1 + 2 c t) A  (SURVEY.md §8d), and the micro-benchmarks
// that measure this device's fp64 tensor / fp64 FMA / HBM-copy peaks (roofline denominators).
#include "csc.cuh"

__host__ __device__ __forceinline__ uint64_t lso_mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
__host__ __device__ __forceinline__ uint64_t lso_hash(uint64_t seed, uint64_t i, uint64_t j) {
    return lso_mix64(seed ^ lso_mix64(i * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL) ^
                     lso_mix64(j * 0xC2B2AE3D27D4EB4FULL + 0x165667B19E3779F9ULL));
}
// uniform in [-1, 1): exact 52-bit fraction
__host__ __device__ __forceinline__ double lso_unif(uint64_t h) {
    return (double)(h >> 12) * (1.0 / 2251799813685248.0) - 1.0;   // 2^-51
}
__host__ __device__ __forceinline__ double lso_colscale(uint64_t seed, uint64_t j) {
    const int e = (int)(lso_hash(seed + 1, 0x5CA1EULL, j) % 13ULL) - 6;   // 2^-6 .. 2^6
    return (e >= 0) ? (double)(1ULL << e) : 1.0 / (double)(1ULL << (-e));
}

__global__ void synth_matrix_kernel(long long m, long long n, long long row_offset, uint64_t seed,
                                    double* __restrict__ A, long long ld) {
    const long long j = blockIdx.y;
    const double s = lso_colscale(seed, (uint64_t)j);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x)
        A[j * ld + i] = lso_unif(lso_hash(seed, (uint64_t)(row_offset + i), (uint64_t)j)) * s;
}
__global__ void synth_vector_kernel(long long n, long long offset, uint64_t seed, double scale, double* __restrict__ x) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = scale * lso_unif(lso_hash(seed, (uint64_t)(offset + i), 0xFFFFFFFFULL));
}
__global__ void synth_residual_kernel(long long m, const double* __restrict__ t, const double* __restrict__ b, double c,
                                      double* __restrict__ r) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const double ti = t[i];
        r[i] = __dadd_rn(__dadd_rn(ti, __dmul_rn(__dmul_rn(c, ti), ti)), -b[i]);
    }
}
__global__ void synth_jacobian_kernel(long long m, long long n, const double* __restrict__ A, long long ld,
                                      const double* __restrict__ t, double c, double* __restrict__ J, long long ldJ) {
    const long long j = blockIdx.y;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const double w = __dadd_rn(1.0, __dmul_rn(2.0 * c, t[i]));
        J[j * ldJ + i] = __dmul_rn(w, A[j * ld + i]);
    }
}
__global__ void synth_csc_jacobian_kernel(long long n, const int* __restrict__ colptr, const int* __restrict__ rowidx,
                                          const double* __restrict__ aval, const double* __restrict__ t, double c,
                                          double* __restrict__ val) {
    const int lane = threadIdx.x & 31;
    const long long j = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (j >= n) return;
    for (int k = colptr[j] + lane; k < colptr[j + 1]; k += 32) {
        const double w = __dadd_rn(1.0, __dmul_rn(2.0 * c, t[rowidx[k]]));
        val[k] = __dmul_rn(w, aval[k]);
    }
}

// CSR image of the same Jacobian: valr[k] = (1 + 2 c t[row]) * aval_csr[k], G lanes per row (rows are contiguous)
template <int G>
__global__ void synth_csr_jacobian_kernel(long long m, const int* __restrict__ rowptr, const double* __restrict__ aval_csr,
                                          const double* __restrict__ t, double c, double* __restrict__ valr) {
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long i = gid / G;
    if (i >= m) return;
    const double w = __dadd_rn(1.0, __dmul_rn(2.0 * c, t[i]));
    for (int k = rowptr[i] + (int)(gid % G); k < rowptr[i + 1]; k += G) valr[k] = __dmul_rn(w, aval_csr[k]);
}

// ---- peaks ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* __restrict__ out) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
// same accumulator pattern as the QR trailing update (2 x 4 accumulators, 2 + 4 distinct operand registers per k-step)
template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) dmma_pattern_kernel(int iters, double* __restrict__ out) {
    double c[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { c[i][j][0] = threadIdx.x * 1e-9; c[i][j][1] = i + j; }
    double a[2], b[4];
    a[0] = 1.0 + threadIdx.x * 1e-12; a[1] = 1.0 - threadIdx.x * 2e-12;
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = 1.0 + (threadIdx.x + j) * 3e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
        a[0] += 1e-13; a[1] -= 1e-13;          // operands change every k-step, as when they come from shared memory
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] += 1e-13;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
    if (s == 12345.678) out[0] = s;
}
// dependency-distance probe: NACC independent accumulators per warp, 8 warps per SM (the update kernel's occupancy)
template <int NACC>
__global__ void __launch_bounds__(256) dmma_dist_kernel(int iters, double* __restrict__ out) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8 / NACC; ++r)
#pragma unroll
            for (int i = 0; i < NACC; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        a += 1e-13; b -= 1e-13;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double* __restrict__ out) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9 + i;
    const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}
__global__ void copy_kernel(long long n2, const double2* __restrict__ src, double2* __restrict__ dst) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

static int time_ms(lso_ctx* ctx, cudaEvent_t e0, cudaEvent_t e1, float* ms) {
    LSO_CHECK_CUDA(ctx, cudaEventSynchronize(e1));
    LSO_CHECK_CUDA(ctx, cudaEventElapsedTime(ms, e0, e1));
    return LSO_OK;
}

extern "C" {

int lso_synth_dense_matrix(lso_ctx* ctx, int64_t m, int64_t n, int64_t row_offset, uint64_t seed, double* d_A, int64_t ld) {
    LSO_REQUIRE(ctx, ctx && d_A && ld >= m && m >= 1 && n >= 1, "bad arguments");
    LSO_ENTER(ctx);
    dim3 grid((unsigned)std::min<int64_t>(cdiv64(m, 256), 128), (unsigned)n);
    synth_matrix_kernel<<<grid, 256, 0, ctx->stream>>>(m, n, row_offset, seed, d_A, ld);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}
int lso_synth_vector(lso_ctx* ctx, int64_t n, int64_t offset, uint64_t seed, double scale, double* d_x) {
    LSO_REQUIRE(ctx, ctx && d_x && n >= 1, "bad arguments");
    LSO_ENTER(ctx);
    synth_vector_kernel<<<(unsigned)std::min<int64_t>(cdiv64(n, 256), 4096), 256, 0, ctx->stream>>>(n, offset, seed, scale, d_x);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}
int lso_synth_residual(lso_ctx* ctx, int64_t m, int64_t n, const double* d_A, int64_t ld, const double* d_x,
                       const double* d_b, double c, double* d_t, double* d_r) {
    LSO_REQUIRE(ctx, ctx && d_A && d_x && d_b && d_t && d_r, "NULL pointer");
    LSO_ENTER(ctx);
    LSO_TRY(lso_dense_gemv_n(ctx, m, n, 1.0, d_A, ld, d_x, 0.0, d_t));
    synth_residual_kernel<<<(unsigned)std::min<int64_t>(cdiv64(m, 256), 4096), 256, 0, ctx->stream>>>(m, d_t, d_b, c, d_r);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}
// r = t + c t^2 - b for a t = A x formed by the caller (sparse model: t comes from lso_csc_mul_n)
int lso_synth_residual_from_t(lso_ctx* ctx, int64_t m, const double* d_t, const double* d_b, double c, double* d_r) {
    LSO_REQUIRE(ctx, ctx && d_t && d_b && d_r && m >= 1, "bad arguments");
    LSO_ENTER(ctx);
    synth_residual_kernel<<<(unsigned)std::min<int64_t>(cdiv64(m, 256), 4096), 256, 0, ctx->stream>>>(m, d_t, d_b, c, d_r);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}
int lso_synth_jacobian(lso_ctx* ctx, int64_t m, int64_t n, const double* d_A, int64_t ld, const double* d_t, double c,
                       double* d_J, int64_t ldJ) {
    LSO_REQUIRE(ctx, ctx && d_A && d_t && d_J && ld >= m && ldJ >= m, "bad arguments");
    LSO_ENTER(ctx);
    dim3 grid((unsigned)std::min<int64_t>(cdiv64(m, 256), 128), (unsigned)n);
    synth_jacobian_kernel<<<grid, 256, 0, ctx->stream>>>(m, n, d_A, ld, d_t, c, d_J, ldJ);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}
int lso_synth_csc_pattern(int64_t m, int64_t n, int64_t nnz_per_col, uint64_t seed, int64_t* h_colptr, int64_t* h_rowval) {
    if (!h_colptr || !h_rowval || nnz_per_col < 1 || nnz_per_col > m) return lso_set_error(nullptr, LSO_ERR_ARG, "bad arguments");
    // stratified rows: entry k of column j falls in stratum k of [0, m) => sorted and distinct by construction
    for (int64_t j = 0; j < n; ++j) {
        h_colptr[j] = j * nnz_per_col + 1;
        for (int64_t k = 0; k < nnz_per_col; ++k) {
            const int64_t lo = (m * k) / nnz_per_col, hi = (m * (k + 1)) / nnz_per_col;
            h_rowval[j * nnz_per_col + k] = lo + (int64_t)(lso_hash(seed + 2, (uint64_t)j, (uint64_t)k) % (uint64_t)(hi - lo)) + 1;
        }
    }
    h_colptr[n] = n * nnz_per_col + 1;
    return LSO_OK;
}
// The same construction inside a BAND: the rows of column j are stratified over a window of `window` rows centred on
// j * m / n (clipped to [0, m)).  A Jacobian with this kind of locality (PDE stencils, bundle adjustment, time series) is
// what lets the gathers of the sparse products hit L1 / whole L2 sectors; the uniform pattern above has none by design.
int lso_synth_csc_pattern_banded(int64_t m, int64_t n, int64_t nnz_per_col, int64_t window, uint64_t seed, int64_t* h_colptr,
                                 int64_t* h_rowval) {
    if (!h_colptr || !h_rowval || nnz_per_col < 1 || window < nnz_per_col || window > m)
        return lso_set_error(nullptr, LSO_ERR_ARG, "bad arguments");
    for (int64_t j = 0; j < n; ++j) {
        h_colptr[j] = j * nnz_per_col + 1;
        int64_t w0 = (int64_t)(((__int128)j * m) / n) - window / 2;
        if (w0 < 0) w0 = 0;
        if (w0 + window > m) w0 = m - window;
        for (int64_t k = 0; k < nnz_per_col; ++k) {
            const int64_t lo = (window * k) / nnz_per_col, hi = (window * (k + 1)) / nnz_per_col;
            h_rowval[j * nnz_per_col + k] = w0 + lo + (int64_t)(lso_hash(seed + 2, (uint64_t)j, (uint64_t)k) % (uint64_t)(hi - lo)) + 1;
        }
    }
    h_colptr[n] = n * nnz_per_col + 1;
    return LSO_OK;
}
int lso_synth_csc_jacobian(lso_csc* A, const double* d_Aval, const double* d_t, double c) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    lso_ctx* ctx = A->ctx;
    LSO_REQUIRE(ctx, d_Aval && d_t, "NULL pointer");
    LSO_ENTER(ctx);
    synth_csc_jacobian_kernel<<<(unsigned)cdiv64(A->n * 32, 256), 256, 0, ctx->stream>>>(A->n, A->d_colptr, A->d_rowidx, d_Aval, d_t, c, A->d_val);
    LSO_CHECK_LAUNCH(ctx);
    A->csr_dirty = true;
    A->colsq_valid = false;
    return LSO_OK;
}
// the same g! writing BOTH images (no mirror gather): d_Aval_csr = lso_csc_gather_csr(A, d_Aval), once per pattern
int lso_synth_csc_jacobian_both(lso_csc* A, const double* d_Aval, const double* d_Aval_csr, const double* d_t, double c) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    lso_ctx* ctx = A->ctx;
    LSO_REQUIRE(ctx, d_Aval && d_Aval_csr && d_t, "NULL pointer");
    LSO_ENTER(ctx);
    LSO_TRY(lso_synth_csc_jacobian(A, d_Aval, d_t, c));
    if (A->Gr >= 8) synth_csr_jacobian_kernel<8><<<(unsigned)cdiv64(A->m * 8, 256), 256, 0, ctx->stream>>>(A->m, A->d_rowptr, d_Aval_csr, d_t, c, A->d_valr);
    else synth_csr_jacobian_kernel<2><<<(unsigned)cdiv64(A->m * 2, 256), 256, 0, ctx->stream>>>(A->m, A->d_rowptr, d_Aval_csr, d_t, c, A->d_valr);
    LSO_CHECK_LAUNCH(ctx);
    A->csr_dirty = false;
    return LSO_OK;
}

int lso_bench_fp64_mma_peak(lso_ctx* ctx, int iters, double* tflops_out) {
    LSO_REQUIRE(ctx, ctx && tflops_out && iters > 0, "bad arguments");
    LSO_ENTER(ctx);
    cudaEvent_t e0, e1;
    LSO_CHECK_CUDA(ctx, cudaEventCreate(&e0));
    LSO_CHECK_CUDA(ctx, cudaEventCreate(&e1));
    const int grid = ctx->num_sms * 8;
    dmma_peak_kernel<<<grid, 256, 0, ctx->stream>>>(iters / 8 + 1, ctx->d_partials);
    LSO_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    dmma_peak_kernel<<<grid, 256, 0, ctx->stream>>>(iters, ctx->d_partials);
    LSO_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    ctx->launches += 2;
    float ms = 0;
    LSO_TRY(time_ms(ctx, e0, e1, &ms));
    const double flops = (double)grid * 8.0 /*warps*/ * (double)iters * 8.0 * 512.0;
    *tflops_out = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return LSO_OK;
}
// mode 1: 8 warps / SM (the trailing-update kernel's occupancy), mode 2: 16 warps / SM, mode 3: 32 warps / SM
int lso_bench_fp64_mma_pattern(lso_ctx* ctx, int iters, int mode, double* tflops_out) {
    LSO_REQUIRE(ctx, ctx && tflops_out && iters > 0, "bad arguments");
    LSO_ENTER(ctx);
    cudaEvent_t e0, e1;
    LSO_CHECK_CUDA(ctx, cudaEventCreate(&e0));
    LSO_CHECK_CUDA(ctx, cudaEventCreate(&e1));
    const int ctas_per_sm = (mode == 1) ? 1 : (mode == 2 ? 2 : (mode == 3 ? 4 : 1));
    const int grid = ctx->num_sms * ctas_per_sm;
    auto launch = [&](int it) {
        if (mode <= 3) dmma_pattern_kernel<8><<<grid, 256, 0, ctx->stream>>>(it, ctx->d_partials);
        else if (mode == 12) dmma_dist_kernel<2><<<grid, 256, 0, ctx->stream>>>(it, ctx->d_partials);   // 8 warps/SM, distance 2
        else if (mode == 14) dmma_dist_kernel<4><<<grid, 256, 0, ctx->stream>>>(it, ctx->d_partials);
        else dmma_dist_kernel<8><<<grid, 256, 0, ctx->stream>>>(it, ctx->d_partials);
    };
    launch(iters / 8 + 1);
    LSO_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    launch(iters);
    LSO_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    ctx->launches += 2;
    float ms = 0;
    LSO_TRY(time_ms(ctx, e0, e1, &ms));
    const double flops = (double)grid * 8.0 * (double)iters * 8.0 * 512.0;
    *tflops_out = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return LSO_OK;
}
int lso_bench_fp64_fma_peak(lso_ctx* ctx, int iters, double* tflops_out) {
    LSO_REQUIRE(ctx, ctx && tflops_out && iters > 0, "bad arguments");
    LSO_ENTER(ctx);
    cudaEvent_t e0, e1;
    LSO_CHECK_CUDA(ctx, cudaEventCreate(&e0));
    LSO_CHECK_CUDA(ctx, cudaEventCreate(&e1));
    const int grid = ctx->num_sms * 8;
    dfma_peak_kernel<<<grid, 256, 0, ctx->stream>>>(iters / 8 + 1, ctx->d_partials);
    LSO_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    dfma_peak_kernel<<<grid, 256, 0, ctx->stream>>>(iters, ctx->d_partials);
    LSO_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    ctx->launches += 2;
    float ms = 0;
    LSO_TRY(time_ms(ctx, e0, e1, &ms));
    const double flops = (double)grid * 256.0 * (double)iters * 16.0 * 2.0;
    *tflops_out = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return LSO_OK;
}
int lso_bench_hbm_copy(lso_ctx* ctx, size_t nbytes, int iters, double* gbs_out) {
    LSO_REQUIRE(ctx, ctx && gbs_out && iters > 0 && nbytes >= 4096, "bad arguments");
    LSO_ENTER(ctx);
    nbytes &= ~(size_t)15;
    double2 *a = nullptr, *b = nullptr;
    LSO_CHECK_CUDA(ctx, cudaMalloc(&a, nbytes));
    LSO_CHECK_CUDA(ctx, cudaMalloc(&b, nbytes));
    LSO_CHECK_CUDA(ctx, cudaMemsetAsync(a, 1, nbytes, ctx->stream));
    cudaEvent_t e0, e1;
    LSO_CHECK_CUDA(ctx, cudaEventCreate(&e0));
    LSO_CHECK_CUDA(ctx, cudaEventCreate(&e1));
    const long long n2 = (long long)(nbytes / 16);
    const int grid = ctx->num_sms * 16;
    double best = 0;
    copy_kernel<<<grid, 256, 0, ctx->stream>>>(n2, a, b);
    for (int it = 0; it < iters; ++it) {
        LSO_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        copy_kernel<<<grid, 256, 0, ctx->stream>>>(n2, a, b);
        LSO_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        ctx->launches++;
        float ms = 0;
        LSO_TRY(time_ms(ctx, e0, e1, &ms));
        const double gbs = 2.0 * (double)nbytes / (ms * 1e-3) / 1e9;
        if (gbs > best) best = gbs;
    }
    *gbs_out = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(a); cudaFree(b);
    return LSO_OK;
}

}  // extern "C"

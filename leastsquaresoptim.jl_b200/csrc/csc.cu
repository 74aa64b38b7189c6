// csc.cu — sparse operator kernels (S-a..S-d in SURVEY.md Appendix B):
//   mul!(y, J, x, α, β) / mul!(x, J', y, α, β) for SparseMatrixCSC   [stdlib SparseArrays], called from
//   src/utils/lsmr.jl:73,76,118,122 and levenberg_marquardt.jl:102,114;
//   colsumabs2!(v, J::SparseMatrixCSC)   src/utils/utils.jl:146-151.
// HBM-bound: 12 bytes per stored entry (8 value + 4 index) plus vector traffic.
#include "csc.cuh"
#include <vector>

// ---- adjoint product: one warp per column (gather-dot), fixed summation order -----------------------
__global__ void __launch_bounds__(256)
csc_mul_t_kernel(long long n, const int* __restrict__ colptr, const int* __restrict__ rowidx,
                 const double* __restrict__ val, const double* __restrict__ y, double alpha, double beta,
                 double* __restrict__ x) {
    const int lane = threadIdx.x & 31;
    const long long j = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (j >= n) return;
    const int k0 = colptr[j], k1 = colptr[j + 1];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int k = k0 + lane;
    // 4 independent (value, index, gather) chains in flight per lane
    for (; k + 96 < k1; k += 128) {
        const double v0 = val[k], v1 = val[k + 32], v2 = val[k + 64], v3 = val[k + 96];
        const int r0 = rowidx[k], r1 = rowidx[k + 32], r2 = rowidx[k + 64], r3 = rowidx[k + 96];
        a0 = fma(v0, y[r0], a0);
        a1 = fma(v1, y[r1], a1);
        a2 = fma(v2, y[r2], a2);
        a3 = fma(v3, y[r3], a3);
    }
    {   // tail: up to 3 more entries, still issued together
        const bool p0 = k < k1, p1 = k + 32 < k1, p2 = k + 64 < k1;
        const double v0 = p0 ? val[k] : 0.0, v1 = p1 ? val[k + 32] : 0.0, v2 = p2 ? val[k + 64] : 0.0;
        const int r0 = p0 ? rowidx[k] : 0, r1 = p1 ? rowidx[k + 32] : 0, r2 = p2 ? rowidx[k + 64] : 0;
        if (p0) a0 = fma(v0, y[r0], a0);
        if (p1) a1 = fma(v1, y[r1], a1);
        if (p2) a2 = fma(v2, y[r2], a2);
    }
    a0 = (a0 + a1) + (a2 + a3);
    a1 = 0.0;
    double acc = warp_sum(a0 + a1);
    if (lane == 0) {
        acc *= alpha;
        x[j] = (beta == 0.0) ? acc : fma(beta, x[j], acc);
    }
}

// ---- column sums of squares ------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
csc_colsumabs2_kernel(long long n, const int* __restrict__ colptr, const double* __restrict__ val,
                      double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long j = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (j >= n) return;
    const int k0 = colptr[j], k1 = colptr[j + 1];
    double a0 = 0.0, a1 = 0.0;
    int k = k0 + lane;
    for (; k + 32 < k1; k += 64) {
        const double v0 = val[k], v1 = val[k + 32];
        a0 = fma(v0, v0, a0);
        a1 = fma(v1, v1, a1);
    }
    if (k < k1) { const double v = val[k]; a0 = fma(v, v, a0); }
    double acc = warp_sum(a0 + a1);
    if (lane == 0) out[j] = acc;
}

// ---- forward product on the CSR mirror: G lanes per row -----------------------------------------------
template <int G>
__global__ void __launch_bounds__(256)
csr_mul_n_kernel(long long m, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                 const double* __restrict__ val, const double* __restrict__ x, double alpha, double beta,
                 double* __restrict__ y) {
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long i = gid / G;
    const int sub = (int)(gid % G);
    double acc = 0.0;
    if (i < m) {
        const int k0 = rowptr[i], k1 = rowptr[i + 1];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int k = k0 + sub; k < k1; k += 4 * G) {      // 4 predicated chains in flight per lane
            const bool p1 = k + G < k1, p2 = k + 2 * G < k1, p3 = k + 3 * G < k1;
            const double v0 = val[k], v1 = p1 ? val[k + G] : 0.0, v2 = p2 ? val[k + 2 * G] : 0.0, v3 = p3 ? val[k + 3 * G] : 0.0;
            const int c0 = colidx[k], c1 = p1 ? colidx[k + G] : 0, c2 = p2 ? colidx[k + 2 * G] : 0, c3 = p3 ? colidx[k + 3 * G] : 0;
            const double x0 = x[c0], x1 = p1 ? x[c1] : 0.0, x2 = p2 ? x[c2] : 0.0, x3 = p3 ? x[c3] : 0.0;
            a0 = fma(v0, x0, a0);
            a1 = fma(v1, x1, a1);
            a2 = fma(v2, x2, a2);
            a3 = fma(v3, x3, a3);
        }
        acc = (a0 + a1) + (a2 + a3);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (i < m && sub == 0) {
        acc *= alpha;
        y[i] = (beta == 0.0) ? acc : fma(beta, y[i], acc);
    }
}

__global__ void csr_gather_values_kernel(long long nnz, const int* __restrict__ perm, const double* __restrict__ val,
                                         double* __restrict__ valr) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x)
        valr[k] = val[perm[k]];
}

int csc_refresh_csr(lso_csc* A) {
    lso_ctx* ctx = A->ctx;
    if (!A->csr_dirty || A->nnz == 0) { A->csr_dirty = false; return LSO_OK; }
    int64_t g = cdiv64(A->nnz, 256);
    if (g > (int64_t)ctx->num_sms * 16) g = (int64_t)ctx->num_sms * 16;
    csr_gather_values_kernel<<<(unsigned)g, 256, 0, ctx->stream>>>(A->nnz, A->d_perm, A->d_val, A->d_valr);
    LSO_CHECK_LAUNCH(ctx);
    A->csr_dirty = false;
    return LSO_OK;
}

extern "C" {

int lso_csc_create(lso_ctx* ctx, int64_t m, int64_t n, int64_t nnz, const int64_t* h_colptr, const int64_t* h_rowval,
                   lso_csc** out) {
    LSO_REQUIRE(ctx, ctx && out, "ctx/out is NULL");
    *out = nullptr;
    LSO_REQUIRE(ctx, m >= 1 && n >= 1 && nnz >= 0, "bad dimensions");
    LSO_REQUIRE(ctx, nnz < 2147483647LL && m < 2147483647LL && n < 2147483647LL, "dimensions exceed int32 indexing");
    LSO_REQUIRE(ctx, h_colptr && (nnz == 0 || h_rowval), "NULL pointer");
    LSO_REQUIRE(ctx, h_colptr[0] == 1 && h_colptr[n] == nnz + 1, "colptr must be 1-based with colptr[n+1] == nnz+1");
    std::vector<int> colptr(n + 1), rowidx(nnz), rowptr(m + 1, 0), colidx(nnz), perm(nnz);
    for (int64_t j = 0; j <= n; ++j) {
        if (j > 0 && h_colptr[j] < h_colptr[j - 1]) return lso_set_error(ctx, LSO_ERR_ARG, "colptr is not monotone");
        colptr[j] = (int)(h_colptr[j] - 1);
    }
    for (int64_t k = 0; k < nnz; ++k) {
        const int64_t r = h_rowval[k] - 1;
        if (r < 0 || r >= m) return lso_set_error(ctx, LSO_ERR_ARG, "rowval out of range");
        rowidx[k] = (int)r;
        rowptr[r + 1]++;
    }
    for (int64_t i = 0; i < m; ++i) rowptr[i + 1] += rowptr[i];
    {   // stable counting sort by row: within a row, entries keep ascending column order
        std::vector<int> next(rowptr.begin(), rowptr.end() - 1);
        for (int64_t j = 0; j < n; ++j)
            for (int k = colptr[j]; k < colptr[j + 1]; ++k) {
                const int pos = next[rowidx[k]]++;
                colidx[pos] = (int)j;
                perm[pos] = k;
            }
    }
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    lso_csc* A = new (std::nothrow) lso_csc();
    if (!A) return lso_set_error(ctx, LSO_ERR_ALLOC, "host allocation failed");
    A->ctx = ctx; A->m = m; A->n = n; A->nnz = nnz;
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&A->d_colptr, (n + 1) * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&A->d_rowidx, nz * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&A->d_val, nz * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&A->d_rowptr, (m + 1) * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&A->d_colidx, nz * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&A->d_perm, nz * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&A->d_valr, nz * sizeof(double));
    if (e != cudaSuccess) {
        cudaGetLastError();
        lso_csc_destroy(A);
        return lso_set_error(ctx, LSO_ERR_ALLOC, "CSC image: %s", cudaGetErrorString(e));
    }
    cudaMemcpyAsync(A->d_colptr, colptr.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(A->d_rowptr, rowptr.data(), (m + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    if (nnz) {
        cudaMemcpyAsync(A->d_rowidx, rowidx.data(), nnz * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(A->d_colidx, colidx.data(), nnz * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(A->d_perm, perm.data(), nnz * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    }
    cudaMemsetAsync(A->d_val, 0, nz * sizeof(double), ctx->stream);
    cudaMemsetAsync(A->d_valr, 0, nz * sizeof(double), ctx->stream);
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // host vectors die at return
    *out = A;
    return LSO_OK;
}

int lso_csc_destroy(lso_csc* A) {
    if (!A) return LSO_OK;
    cudaSetDevice(A->ctx->device);
    cudaStreamSynchronize(A->ctx->stream);
    cudaFree(A->d_colptr); cudaFree(A->d_rowidx); cudaFree(A->d_val);
    cudaFree(A->d_rowptr); cudaFree(A->d_colidx); cudaFree(A->d_perm); cudaFree(A->d_valr);
    delete A;
    return LSO_OK;
}

int lso_csc_set_values_host(lso_csc* A, const double* h_nzval) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    LSO_REQUIRE(A->ctx, A->nnz == 0 || h_nzval, "NULL pointer");
    LSO_CHECK_CUDA(A->ctx, cudaMemcpyAsync(A->d_val, h_nzval, A->nnz * sizeof(double), cudaMemcpyHostToDevice, A->ctx->stream));
    LSO_CHECK_CUDA(A->ctx, cudaStreamSynchronize(A->ctx->stream));
    A->csr_dirty = true;
    return LSO_OK;
}

int lso_csc_set_values_dev(lso_csc* A, const double* d_nzval) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    LSO_REQUIRE(A->ctx, A->nnz == 0 || d_nzval, "NULL pointer");
    LSO_CHECK_CUDA(A->ctx, cudaMemcpyAsync(A->d_val, d_nzval, A->nnz * sizeof(double), cudaMemcpyDeviceToDevice, A->ctx->stream));
    A->csr_dirty = true;
    return LSO_OK;
}

double* lso_csc_values(lso_csc* A) { return A ? A->d_val : nullptr; }

int lso_csc_values_changed(lso_csc* A) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    A->csr_dirty = true;
    return LSO_OK;
}

int lso_csc_mul_n(lso_csc* A, double alpha, const double* d_x, double beta, double* d_y) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    lso_ctx* ctx = A->ctx;
    LSO_REQUIRE(ctx, d_x && d_y, "NULL pointer");
    LSO_TRY(csc_refresh_csr(A));
    const double avg = (double)A->nnz / (double)A->m;
    if (avg <= 6.0) {
        csr_mul_n_kernel<4><<<(unsigned)cdiv64(A->m * 4, 256), 256, 0, ctx->stream>>>(A->m, A->d_rowptr, A->d_colidx, A->d_valr, d_x, alpha, beta, d_y);
    } else if (avg <= 48.0) {
        csr_mul_n_kernel<8><<<(unsigned)cdiv64(A->m * 8, 256), 256, 0, ctx->stream>>>(A->m, A->d_rowptr, A->d_colidx, A->d_valr, d_x, alpha, beta, d_y);
    } else {
        csr_mul_n_kernel<32><<<(unsigned)cdiv64(A->m * 32, 256), 256, 0, ctx->stream>>>(A->m, A->d_rowptr, A->d_colidx, A->d_valr, d_x, alpha, beta, d_y);
    }
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

int lso_csc_mul_t(lso_csc* A, double alpha, const double* d_y, double beta, double* d_x) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    lso_ctx* ctx = A->ctx;
    LSO_REQUIRE(ctx, d_x && d_y, "NULL pointer");
    csc_mul_t_kernel<<<(unsigned)cdiv64(A->n * 32, 256), 256, 0, ctx->stream>>>(A->n, A->d_colptr, A->d_rowidx, A->d_val, d_y, alpha, beta, d_x);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

int lso_csc_colsumabs2(lso_csc* A, double* d_out) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    lso_ctx* ctx = A->ctx;
    LSO_REQUIRE(ctx, d_out, "NULL pointer");
    csc_colsumabs2_kernel<<<(unsigned)cdiv64(A->n * 32, 256), 256, 0, ctx->stream>>>(A->n, A->d_colptr, A->d_val, d_out);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

}  // extern "C"

// csc.cu — sparse operator kernels (S-a..S-d in SURVEY.md Appendix B):
//   mul!(y, J, x, α, β) / mul!(x, J', y, α, β) for SparseMatrixCSC   [stdlib SparseArrays], called from
//   src/utils/lsmr.jl:73,76,118,122 and levenberg_marquardt.jl:102,114;
//   colsumabs2!(v, J::SparseMatrixCSC)   src/utils/utils.jl:146-151.
// HBM-bound: 12 bytes per stored entry (8 value + 4 index) plus vector traffic.  The products run through the
// stream kernel of csc.cuh (coalesced 128-bit loads, shared-memory segmented reduction); the first-generation
// warp-per-segment kernels are kept behind ctx option "spmv" = 0 as a cross-check.
#include "csc.cuh"
#include <algorithm>
#include <vector>

// ---- first-generation adjoint product: one warp per column (gather-dot), fixed summation order ----------------
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
    double acc = warp_sum(a0);
    if (lane == 0) {
        acc *= alpha;
        x[j] = (beta == 0.0) ? acc : fma(beta, x[j], acc);
    }
}

// ---- column sums of squares (values only: 8 B per stored entry) -----------------------------------------------
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

// ---- first-generation forward product on the CSR mirror: G lanes per row ---------------------------------------
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

// ---- functors of the stream kernel for the plain operator interface ---------------------------------------------
struct MulFunctor {               // out[s] = alpha * sum + beta * out[s]
    const double* x;
    double* out;
    double alpha, beta;
    long long n_extra;
    static constexpr bool DUAL = false;
    __device__ __forceinline__ bool begin() { return true; }
    __device__ __forceinline__ bool idle() const { return false; }
    __device__ __forceinline__ double gather(int i) const { return __ldg(x + i); }
    __device__ __forceinline__ double epilogue(long long s, double a) const {
        a *= alpha;
        out[s] = (beta == 0.0) ? a : fma(beta, out[s], a);
        return 0.0;
    }
    __device__ __forceinline__ double epilogue2(long long, double, double) const { return 0.0; }
    __device__ __forceinline__ double extra(long long) const { return 0.0; }
    __device__ __forceinline__ void finish(double, double) const {}
};
struct PredictFunctor {           // fpredict = J δ - f ; ssr = sum(abs2, fpredict)   (LM:114-117, dogleg:171-174)
    const double* x;
    const double* f;
    double* fpredict;             // may be NULL
    double* out;                  // device scalar
    long long n_extra;
    static constexpr bool DUAL = false;
    __device__ __forceinline__ bool begin() { return true; }
    __device__ __forceinline__ bool idle() const { return false; }
    __device__ __forceinline__ double gather(int i) const { return __ldg(x + i); }
    __device__ __forceinline__ double epilogue(long long s, double a) const {
        const double r = a - f[s];
        if (fpredict) fpredict[s] = r;
        return r * r;
    }
    __device__ __forceinline__ double epilogue2(long long, double, double) const { return 0.0; }
    __device__ __forceinline__ double extra(long long) const { return 0.0; }
    __device__ __forceinline__ void finish(double sm, double) const { *out = sm; }
};
struct ColsqGradFunctor {         // dtd = colsumabs2(J) and g = J'f in one pass (LM:82 + LM:102)
    const double* f;
    double* dtd;
    double* g;
    long long n_extra;
    static constexpr bool DUAL = true;
    __device__ __forceinline__ bool begin() { return true; }
    __device__ __forceinline__ bool idle() const { return false; }
    __device__ __forceinline__ double gather(int i) const { return __ldg(f + i); }
    __device__ __forceinline__ double epilogue(long long, double) const { return 0.0; }
    __device__ __forceinline__ double epilogue2(long long s, double a, double a2) const {
        g[s] = a;
        dtd[s] = a2;
        return 0.0;
    }
    __device__ __forceinline__ double extra(long long) const { return 0.0; }
    __device__ __forceinline__ void finish(double, double) const {}
};

// ---- pattern import: Int64 1-based CSC -> int32 0-based CSC + CSR mirror + CTA partitions --------------------------
static void build_blocks(const int* ptr, int64_t nseg, std::vector<int>& blk) {
    blk.clear();
    blk.push_back(0);
    int64_t s = 0;
    while (s < nseg) {
        int64_t e = s + 1;                                    // a CTA owns at least one segment
        const int64_t base = ptr[s] & ~1;                     // the slice starts on a 16-byte boundary
        while (e < nseg && e - s < SP_SEGMAX && (int64_t)ptr[e + 1] - base <= SP_CHUNK) ++e;
        blk.push_back((int)e);
        s = e;
    }
}
static int lanes_for(double avg) {
    if (avg > 96.0) return 32;
    if (avg > 48.0) return 16;
    if (avg > 12.0) return 8;
    if (avg > 6.0) return 4;
    if (avg > 3.0) return 2;
    return 1;
}

static void csc_free_buffers(lso_csc* A) {
    cudaFree(A->d_colptr); cudaFree(A->d_rowidx); cudaFree(A->d_val);
    cudaFree(A->d_rowptr); cudaFree(A->d_colidx); cudaFree(A->d_perm); cudaFree(A->d_valr);
    cudaFree(A->d_rblk); cudaFree(A->d_cblk); cudaFree(A->d_colsq);
    A->d_colptr = A->d_rowidx = A->d_rowptr = A->d_colidx = A->d_perm = A->d_rblk = A->d_cblk = nullptr;
    A->d_val = A->d_valr = A->d_colsq = nullptr;
    A->cap_nnz = 0;
}

// (re)build the device image for the pattern (h_colptr, h_rowval); values are zeroed
static int csc_import_pattern(lso_csc* A, int64_t nnz, const int64_t* h_colptr, const int64_t* h_rowval) {
    lso_ctx* ctx = A->ctx;
    const int64_t m = A->m, n = A->n;
    LSO_REQUIRE(ctx, nnz >= 0 && nnz < 2147483000LL, "nnz exceeds int32 indexing");
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
    std::vector<int> rblk, cblk;
    build_blocks(rowptr.data(), m, rblk);
    build_blocks(colptr.data(), n, cblk);
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (nnz > A->cap_nnz || A->d_colptr == nullptr) {
        csc_free_buffers(A);
        const size_t nz = (size_t)nnz + 8;                    // padding: the stream kernel reads entries in aligned pairs
        cudaError_t e = cudaSuccess;
        if (e == cudaSuccess) e = cudaMalloc(&A->d_colptr, (n + 1) * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&A->d_rowidx, nz * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&A->d_val, nz * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc(&A->d_rowptr, (m + 1) * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&A->d_colidx, nz * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&A->d_perm, nz * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&A->d_valr, nz * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc(&A->d_rblk, (m + 2) * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&A->d_cblk, (n + 2) * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&A->d_colsq, n * sizeof(double));
        if (e != cudaSuccess) {
            cudaGetLastError();
            csc_free_buffers(A);
            return lso_set_error(ctx, LSO_ERR_ALLOC, "CSC image: %s", cudaGetErrorString(e));
        }
        A->cap_nnz = nnz;
    }
    A->nnz = nnz;
    A->nrblk = (int)rblk.size() - 1;
    A->ncblk = (int)cblk.size() - 1;
    A->Gr = lanes_for((double)nnz / (double)m);
    A->Gc = lanes_for((double)nnz / (double)n);
    const size_t nz = (size_t)A->cap_nnz + 8;
    cudaMemsetAsync(A->d_val, 0, nz * sizeof(double), ctx->stream);
    cudaMemsetAsync(A->d_valr, 0, nz * sizeof(double), ctx->stream);
    cudaMemsetAsync(A->d_rowidx, 0, nz * sizeof(int), ctx->stream);
    cudaMemsetAsync(A->d_colidx, 0, nz * sizeof(int), ctx->stream);
    cudaMemcpyAsync(A->d_colptr, colptr.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(A->d_rowptr, rowptr.data(), (m + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(A->d_rblk, rblk.data(), rblk.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(A->d_cblk, cblk.data(), cblk.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    if (nnz) {
        cudaMemcpyAsync(A->d_rowidx, rowidx.data(), nnz * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(A->d_colidx, colidx.data(), nnz * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(A->d_perm, perm.data(), nnz * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    }
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // host vectors die at return
    A->csr_dirty = true;
    A->colsq_valid = false;
    return LSO_OK;
}

static inline void csc_values_touched(lso_csc* A) {
    A->csr_dirty = true;
    A->colsq_valid = false;
}

// dtd = colsumabs2(J), through the per-J cache
int csc_colsumabs2_cached(lso_csc* A, double* d_out) {
    lso_ctx* ctx = A->ctx;
    if (!A->colsq_valid) {
        csc_colsumabs2_kernel<<<(unsigned)cdiv64(A->n * 32, 256), 256, 0, ctx->stream>>>(A->n, A->d_colptr, A->d_val, A->d_colsq);
        LSO_CHECK_LAUNCH(ctx);
        A->colsq_valid = true;
    }
    if (d_out != A->d_colsq)
        LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(d_out, A->d_colsq, A->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return LSO_OK;
}

// fpredict = J δ - f, *d_out (device scalar) = sum(abs2, fpredict); no synchronisation
int csc_predicted_ssr_dev(lso_csc* A, const double* d_delta, const double* d_f, double* d_fpredict, double* d_out) {
    lso_ctx* ctx = A->ctx;
    LSO_TRY(csc_refresh_csr(A));
    PredictFunctor f{d_delta, d_f, d_fpredict, d_out, 0};
    return spmv_stream_launch(ctx, A->Gr, f, A->d_rowptr, A->d_colidx, A->d_valr, A->d_rblk, A->nrblk, A->m);
}

extern "C" {

int lso_csc_create(lso_ctx* ctx, int64_t m, int64_t n, int64_t nnz, const int64_t* h_colptr, const int64_t* h_rowval,
                   lso_csc** out) {
    LSO_REQUIRE(ctx, ctx && out, "ctx/out is NULL");
    *out = nullptr;
    LSO_REQUIRE(ctx, m >= 1 && n >= 1 && nnz >= 0, "bad dimensions");
    LSO_REQUIRE(ctx, m < 2147483647LL && n < 2147483647LL, "dimensions exceed int32 indexing");
    lso_csc* A = new (std::nothrow) lso_csc();
    if (!A) return lso_set_error(ctx, LSO_ERR_ALLOC, "host allocation failed");
    A->ctx = ctx; A->m = m; A->n = n;
    const int st = csc_import_pattern(A, nnz, h_colptr, h_rowval);
    if (st != LSO_OK) { lso_csc_destroy(A); return st; }
    *out = A;
    return LSO_OK;
}

int lso_csc_update_pattern(lso_csc* A, int64_t nnz, const int64_t* h_colptr, const int64_t* h_rowval) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    return csc_import_pattern(A, nnz, h_colptr, h_rowval);
}

int lso_csc_destroy(lso_csc* A) {
    if (!A) return LSO_OK;
    cudaSetDevice(A->ctx->device);
    cudaStreamSynchronize(A->ctx->stream);
    csc_free_buffers(A);
    delete A;
    return LSO_OK;
}

int lso_csc_set_values_host(lso_csc* A, const double* h_nzval) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    LSO_REQUIRE(A->ctx, A->nnz == 0 || h_nzval, "NULL pointer");
    LSO_CHECK_CUDA(A->ctx, cudaSetDevice(A->ctx->device));
    LSO_CHECK_CUDA(A->ctx, cudaMemcpyAsync(A->d_val, h_nzval, A->nnz * sizeof(double), cudaMemcpyHostToDevice, A->ctx->stream));
    LSO_CHECK_CUDA(A->ctx, cudaStreamSynchronize(A->ctx->stream));
    csc_values_touched(A);
    return LSO_OK;
}

int lso_csc_set_values_dev(lso_csc* A, const double* d_nzval) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    LSO_REQUIRE(A->ctx, A->nnz == 0 || d_nzval, "NULL pointer");
    LSO_CHECK_CUDA(A->ctx, cudaSetDevice(A->ctx->device));
    LSO_CHECK_CUDA(A->ctx, cudaMemcpyAsync(A->d_val, d_nzval, A->nnz * sizeof(double), cudaMemcpyDeviceToDevice, A->ctx->stream));
    csc_values_touched(A);
    return LSO_OK;
}

double* lso_csc_values(lso_csc* A) { return A ? A->d_val : nullptr; }
double* lso_csc_values_csr(lso_csc* A) { return A ? A->d_valr : nullptr; }
const int* lso_csc_csr_rowptr(lso_csc* A) { return A ? A->d_rowptr : nullptr; }
const int* lso_csc_csr_colidx(lso_csc* A) { return A ? A->d_colidx : nullptr; }
const int* lso_csc_csr_perm(lso_csc* A) { return A ? A->d_perm : nullptr; }

int lso_csc_values_changed(lso_csc* A) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    csc_values_touched(A);
    return LSO_OK;
}

int lso_csc_values_changed_both(lso_csc* A) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    A->csr_dirty = false;          // the device g! wrote lso_csc_values AND lso_csc_values_csr
    A->colsq_valid = false;
    return LSO_OK;
}

int lso_csc_gather_csr(lso_csc* A, const double* d_in_csc_order, double* d_out_csr_order) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    lso_ctx* ctx = A->ctx;
    LSO_REQUIRE(ctx, A->nnz == 0 || (d_in_csc_order && d_out_csr_order), "NULL pointer");
    if (A->nnz == 0) return LSO_OK;
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t g = cdiv64(A->nnz, 256);
    if (g > (int64_t)ctx->num_sms * 16) g = (int64_t)ctx->num_sms * 16;
    csr_gather_values_kernel<<<(unsigned)g, 256, 0, ctx->stream>>>(A->nnz, A->d_perm, d_in_csc_order, d_out_csr_order);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

int lso_csc_mul_n(lso_csc* A, double alpha, const double* d_x, double beta, double* d_y) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    lso_ctx* ctx = A->ctx;
    LSO_REQUIRE(ctx, d_x && d_y, "NULL pointer");
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    LSO_TRY(csc_refresh_csr(A));
    ctx->stat_spmv_bytes += 12.0 * (double)A->nnz + 8.0 * (double)(A->m + A->n);
    lso_prof_mark(ctx);
    if (ctx->opt_spmv) {
        MulFunctor f{d_x, d_y, alpha, beta, 0};
        LSO_TRY(spmv_stream_launch(ctx, A->Gr, f, A->d_rowptr, A->d_colidx, A->d_valr, A->d_rblk, A->nrblk, A->m));
    } else {
        const double avg = (double)A->nnz / (double)A->m;
        if (avg <= 6.0) {
            csr_mul_n_kernel<4><<<(unsigned)cdiv64(A->m * 4, 256), 256, 0, ctx->stream>>>(A->m, A->d_rowptr, A->d_colidx, A->d_valr, d_x, alpha, beta, d_y);
        } else if (avg <= 48.0) {
            csr_mul_n_kernel<8><<<(unsigned)cdiv64(A->m * 8, 256), 256, 0, ctx->stream>>>(A->m, A->d_rowptr, A->d_colidx, A->d_valr, d_x, alpha, beta, d_y);
        } else {
            csr_mul_n_kernel<32><<<(unsigned)cdiv64(A->m * 32, 256), 256, 0, ctx->stream>>>(A->m, A->d_rowptr, A->d_colidx, A->d_valr, d_x, alpha, beta, d_y);
        }
        LSO_CHECK_LAUNCH(ctx);
    }
    lso_prof_mark(ctx);
    return LSO_OK;
}

int lso_csc_mul_t(lso_csc* A, double alpha, const double* d_y, double beta, double* d_x) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    lso_ctx* ctx = A->ctx;
    LSO_REQUIRE(ctx, d_x && d_y, "NULL pointer");
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->stat_spmv_bytes += 12.0 * (double)A->nnz + 8.0 * (double)(A->m + A->n);
    lso_prof_mark(ctx);
    if (ctx->opt_spmv) {
        MulFunctor f{d_y, d_x, alpha, beta, 0};
        LSO_TRY(spmv_stream_launch(ctx, A->Gc, f, A->d_colptr, A->d_rowidx, A->d_val, A->d_cblk, A->ncblk, A->n));
    } else {
        csc_mul_t_kernel<<<(unsigned)cdiv64(A->n * 32, 256), 256, 0, ctx->stream>>>(A->n, A->d_colptr, A->d_rowidx, A->d_val, d_y, alpha, beta, d_x);
        LSO_CHECK_LAUNCH(ctx);
    }
    lso_prof_mark(ctx);
    return LSO_OK;
}

int lso_csc_colsumabs2(lso_csc* A, double* d_out) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    lso_ctx* ctx = A->ctx;
    LSO_REQUIRE(ctx, d_out, "NULL pointer");
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    return csc_colsumabs2_cached(A, d_out);
}

int lso_csc_colsumabs2_gemv_t(lso_csc* A, const double* d_f, double* d_dtd, double* d_g) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    lso_ctx* ctx = A->ctx;
    LSO_REQUIRE(ctx, d_f && d_dtd && d_g, "NULL pointer");
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (A->colsq_valid) {          // a rejected LM step: same J, only J'f is needed again (and f did not change either)
        LSO_TRY(csc_colsumabs2_cached(A, d_dtd));
        return lso_csc_mul_t(A, 1.0, d_f, 0.0, d_g);
    }
    ColsqGradFunctor f{d_f, A->d_colsq, d_g, 0};
    LSO_TRY(spmv_stream_launch(ctx, A->Gc, f, A->d_colptr, A->d_rowidx, A->d_val, A->d_cblk, A->ncblk, A->n));
    A->colsq_valid = true;
    return csc_colsumabs2_cached(A, d_dtd);
}

int lso_csc_predicted_ssr(lso_csc* A, const double* d_delta, const double* d_f, double* d_fpredict, double* ssr_out) {
    if (!A) return lso_set_error(nullptr, LSO_ERR_ARG, "A is NULL");
    lso_ctx* ctx = A->ctx;
    LSO_REQUIRE(ctx, d_delta && d_f && ssr_out, "NULL pointer");
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    LSO_TRY(csc_predicted_ssr_dev(A, d_delta, d_f, d_fpredict, ctx->d_scalars + 2));
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars + 2, ctx->d_scalars + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *ssr_out = ctx->h_scalars[2];
    return LSO_OK;
}

}  // extern "C"

// dense.cu — HBM-bound passes over a dense column-major Jacobian (G1, G3, D1 in SURVEY.md App. B):
//   colsumabs2!(dtd, J)          src/utils/utils.jl:139-144
//   mul!(x, J', y, α, β)         LM:102, dogleg:99, dense_cholesky.jl:32,56   (dgemv 'T')
//   mul!(y, J, x, α, β)          LM:114, dogleg:109,171                        (dgemv 'N')
// plus the fused forms  {colsumabs2 + J'f}  and  {J δ − f, ‖·‖²}  that read J once instead of twice.
// Algorithmic traffic: 8·m·n bytes per pass.  Reductions are two-stage in a fixed order (deterministic).
#include "common.cuh"

// ---- column-wise reductions: one CTA per (column, row-split) --------------------------------------
// MODE bit0: sum of squares ; bit1: dot with f
template <int MODE>
__global__ void __launch_bounds__(256)
col_reduce_kernel(int64_t m, int64_t n, const double* __restrict__ J, int64_t ld, const double* __restrict__ f,
                  int64_t rows_per_split, double* __restrict__ out_sq, double* __restrict__ out_dot, int nsplit) {
    __shared__ double sm[32];
    const int64_t j = blockIdx.x;
    const int s = blockIdx.y;
    const int64_t r0 = (int64_t)s * rows_per_split;
    int64_t r1 = r0 + rows_per_split;
    if (r1 > m) r1 = m;
    const double* __restrict__ col = J + j * ld;
    double a_sq = 0.0, a_dot = 0.0;
    int64_t i = r0 + threadIdx.x;
    // 8 independent loads in flight per thread
    for (; i + 7 * 256 < r1; i += 8 * 256) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = col[i + u * 256];
        if (MODE & 1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a_sq = fma(v[u], v[u], a_sq);
        }
        if (MODE & 2) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a_dot = fma(v[u], f[i + u * 256], a_dot);
        }
    }
    for (; i < r1; i += 256) {
        double v = col[i];
        if (MODE & 1) a_sq = fma(v, v, a_sq);
        if (MODE & 2) a_dot = fma(v, f[i], a_dot);
    }
    if (MODE & 1) {
        a_sq = block_sum(a_sq, sm);
        if (threadIdx.x == 0) out_sq[(int64_t)s * n + j] = a_sq;
    }
    if (MODE & 2) {
        a_dot = block_sum(a_dot, sm);
        if (threadIdx.x == 0) out_dot[(int64_t)s * n + j] = a_dot;
    }
}

// second stage: out[j] = alpha * sum_s part[s*n + j] + beta * out[j]
__global__ void col_reduce_final_kernel(int64_t n, int nsplit, const double* __restrict__ part, double alpha,
                                        double beta, double* __restrict__ out) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    double a = 0.0;
    for (int s = 0; s < nsplit; ++s) a += part[(int64_t)s * n + j];
    a *= alpha;
    out[j] = (beta == 0.0) ? a : fma(beta, out[j], a);
}

struct ColSplit { int nsplit; int64_t rows_per_split; };
static ColSplit choose_split(lso_ctx* ctx, int64_t m, int64_t n) {
    int64_t want = cdiv64((int64_t)ctx->num_sms * 8, n > 0 ? n : 1);
    int64_t maxs = cdiv64(m, 4096);
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    // partial buffer capacity: 2 arrays of nsplit*n doubles
    while (want > 1 && 2 * want * n > LSO_PARTIALS) --want;
    ColSplit cs;
    cs.rows_per_split = roundup64(cdiv64(m, want), 256);
    cs.nsplit = (int)cdiv64(m, cs.rows_per_split);
    if (cs.nsplit < 1) cs.nsplit = 1;
    return cs;
}

// scratch for big n (nsplit*n > LSO_PARTIALS/2 can't happen when nsplit==1 and we write directly)
static int col_reduce(lso_ctx* ctx, int mode, int64_t m, int64_t n, const double* J, int64_t ld, const double* f,
                      double* d_sq, double alpha, double beta, double* d_dot) {
    LSO_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    LSO_REQUIRE(ctx, m >= 0 && n >= 0 && ld >= m, "bad dimensions");
    if (n == 0) return LSO_OK;
    LSO_REQUIRE(ctx, J != nullptr, "J is NULL");
    LSO_REQUIRE(ctx, n <= 2147483647LL, "n too large");
    LSO_ENTER(ctx);
    ColSplit cs = choose_split(ctx, m, n);
    const bool direct = (cs.nsplit == 1 && alpha == 1.0 && beta == 0.0);
    if (!direct) LSO_REQUIRE(ctx, 2 * (int64_t)cs.nsplit * n <= LSO_PARTIALS, "n too large for split reduction");
    double* p_sq = direct ? d_sq : ctx->d_partials;
    double* p_dot = direct ? d_dot : ctx->d_partials + (int64_t)cs.nsplit * n;
    dim3 grid((unsigned)n, (unsigned)cs.nsplit);
    if (mode == 1) col_reduce_kernel<1><<<grid, 256, 0, ctx->stream>>>(m, n, J, ld, f, cs.rows_per_split, p_sq, p_dot, cs.nsplit);
    else if (mode == 2) col_reduce_kernel<2><<<grid, 256, 0, ctx->stream>>>(m, n, J, ld, f, cs.rows_per_split, p_sq, p_dot, cs.nsplit);
    else col_reduce_kernel<3><<<grid, 256, 0, ctx->stream>>>(m, n, J, ld, f, cs.rows_per_split, p_sq, p_dot, cs.nsplit);
    LSO_CHECK_LAUNCH(ctx);
    if (!direct) {
        int g = (int)cdiv64(n, 256);
        if (mode & 1) {
            col_reduce_final_kernel<<<g, 256, 0, ctx->stream>>>(n, cs.nsplit, p_sq, 1.0, 0.0, d_sq);
            LSO_CHECK_LAUNCH(ctx);
        }
        if (mode & 2) {
            col_reduce_final_kernel<<<g, 256, 0, ctx->stream>>>(n, cs.nsplit, p_dot, alpha, beta, d_dot);
            LSO_CHECK_LAUNCH(ctx);
        }
    }
    return LSO_OK;
}

// ---- y = α J x + β y : one row per thread, coalesced along rows, x broadcast from L1 ---------------
// FUSE: y_out = J x − f and block partial of Σ y_out²  (predicted residual, LM:114-117)
template <bool FUSE>
__global__ void __launch_bounds__(256)
gemv_n_kernel(int64_t m, int64_t n, double alpha, const double* __restrict__ J, int64_t ld,
              const double* __restrict__ x, double beta, double* __restrict__ y, const double* __restrict__ f,
              double* __restrict__ partials, unsigned int* __restrict__ counter, double* __restrict__ out) {
    __shared__ double sm[32];
    __shared__ bool is_last;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double acc = 0.0;
    if (i < m) {
        const double* __restrict__ row = J + i;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        int64_t j = 0;
        for (; j + 7 < n; j += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = row[(j + u) * ld];
            a0 = fma(v[0], x[j + 0], a0); a1 = fma(v[1], x[j + 1], a1);
            a2 = fma(v[2], x[j + 2], a2); a3 = fma(v[3], x[j + 3], a3);
            a0 = fma(v[4], x[j + 4], a0); a1 = fma(v[5], x[j + 5], a1);
            a2 = fma(v[6], x[j + 6], a2); a3 = fma(v[7], x[j + 7], a3);
        }
        for (; j < n; ++j) a0 = fma(row[j * ld], x[j], a0);
        acc = (a0 + a1) + (a2 + a3);
    }
    if (!FUSE) {
        if (i < m) {
            double r = alpha * acc;
            y[i] = (beta == 0.0) ? r : fma(beta, y[i], r);
        }
        return;
    }
    double r = 0.0;
    if (i < m) {
        r = acc - f[i];                 // mul!(fpredict, J, δx); axpy!(-1, fcur, fpredict)
        if (y) y[i] = r;
    }
    double s = block_sum(r * r, sm);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s;
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double a = 0.0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) a += ((volatile double*)partials)[k];
        a = block_sum(a, sm);
        if (threadIdx.x == 0) { *out = a; *counter = 0; }
    }
}

// column-split variant for tall-but-not-huge J: grid.y chunks of columns give enough CTAs (bytes in flight) to
// saturate HBM; partial row sums go to a small scratch and a combine kernel finishes (deterministic order).
__global__ void __launch_bounds__(256)
gemv_n_part_kernel(int64_t m, int64_t n, const double* __restrict__ J, int64_t ld, const double* __restrict__ x,
                   int64_t cols_per_chunk, double* __restrict__ part) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int64_t j0 = (int64_t)blockIdx.y * cols_per_chunk;
    int64_t j1 = j0 + cols_per_chunk;
    if (j1 > n) j1 = n;
    const double* __restrict__ row = J + i;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int64_t j = j0;
    for (; j + 7 < j1; j += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = row[(j + u) * ld];
        a0 = fma(v[0], x[j + 0], a0); a1 = fma(v[1], x[j + 1], a1);
        a2 = fma(v[2], x[j + 2], a2); a3 = fma(v[3], x[j + 3], a3);
        a0 = fma(v[4], x[j + 4], a0); a1 = fma(v[5], x[j + 5], a1);
        a2 = fma(v[6], x[j + 6], a2); a3 = fma(v[7], x[j + 7], a3);
    }
    for (; j < j1; ++j) a0 = fma(row[j * ld], x[j], a0);
    part[(int64_t)blockIdx.y * m + i] = (a0 + a1) + (a2 + a3);
}

template <bool FUSE>
__global__ void __launch_bounds__(256)
gemv_n_combine_kernel(int64_t m, int nchunk, const double* __restrict__ part, double alpha, double beta,
                      double* __restrict__ y, const double* __restrict__ f, double* __restrict__ partials,
                      unsigned int* __restrict__ counter, double* __restrict__ out) {
    __shared__ double sm[32];
    __shared__ bool is_last;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double acc = 0.0;
    if (i < m)
        for (int c = 0; c < nchunk; ++c) acc += part[(int64_t)c * m + i];
    if (!FUSE) {
        if (i < m) {
            const double r = alpha * acc;
            y[i] = (beta == 0.0) ? r : fma(beta, y[i], r);
        }
        return;
    }
    double r = 0.0;
    if (i < m) {
        r = acc - f[i];
        if (y) y[i] = r;
    }
    double s = block_sum(r * r, sm);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s;
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double a = 0.0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) a += ((volatile double*)partials)[k];
        a = block_sum(a, sm);
        if (threadIdx.x == 0) { *out = a; *counter = 0; }
    }
}

static int gemv_n_chunks(lso_ctx* ctx, int64_t m, int64_t n) {
    const int64_t rows_ctas = cdiv64(m, 256);
    int64_t cs = cdiv64((int64_t)ctx->num_sms * 8, rows_ctas);
    if (cs > 16) cs = 16;
    if (cs > n / 64) cs = n / 64;
    return (int)(cs < 1 ? 1 : cs);
}

static int gemv_scratch(lso_ctx* ctx, size_t doubles, double** out) {
    if (ctx->gemv_cap < doubles) {
        LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_gemv);
        ctx->d_gemv = nullptr;
        ctx->gemv_cap = 0;
        LSO_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_gemv, doubles * sizeof(double)));
        ctx->gemv_cap = doubles;
    }
    *out = ctx->d_gemv;
    return LSO_OK;
}

int lso_fetch_scalar(lso_ctx* ctx, int slot, double* out);

extern "C" {

int lso_dense_colsumabs2(lso_ctx* ctx, int64_t m, int64_t n, const double* d_J, int64_t ld, double* d_out) {
    LSO_REQUIRE(ctx, ctx && (n == 0 || d_out), "NULL pointer");
    return col_reduce(ctx, 1, m, n, d_J, ld, nullptr, d_out, 1.0, 0.0, nullptr);
}

int lso_dense_gemv_t(lso_ctx* ctx, int64_t m, int64_t n, double alpha, const double* d_J, int64_t ld,
                     const double* d_y, double beta, double* d_x) {
    LSO_REQUIRE(ctx, ctx && (n == 0 || d_x) && (m == 0 || d_y), "NULL pointer");
    return col_reduce(ctx, 2, m, n, d_J, ld, d_y, nullptr, alpha, beta, d_x);
}

int lso_dense_colsumabs2_gemv_t(lso_ctx* ctx, int64_t m, int64_t n, const double* d_J, int64_t ld,
                                const double* d_f, double* d_dtd, double* d_g) {
    LSO_REQUIRE(ctx, ctx && (n == 0 || (d_dtd && d_g)) && (m == 0 || d_f), "NULL pointer");
    return col_reduce(ctx, 3, m, n, d_J, ld, d_f, d_dtd, 1.0, 0.0, d_g);
}

int lso_dense_gemv_n(lso_ctx* ctx, int64_t m, int64_t n, double alpha, const double* d_J, int64_t ld,
                     const double* d_x, double beta, double* d_y) {
    LSO_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    LSO_REQUIRE(ctx, m >= 0 && n >= 0 && ld >= m, "bad dimensions");
    if (m == 0) return LSO_OK;
    LSO_REQUIRE(ctx, d_y && (n == 0 || (d_J && d_x)), "NULL pointer");
    LSO_ENTER(ctx);
    const int cs = gemv_n_chunks(ctx, m, n);
    if (cs > 1) {
        double* part = nullptr;
        LSO_TRY(gemv_scratch(ctx, (size_t)cs * m, &part));
        const int64_t cpc = roundup64(cdiv64(n, cs), 8);
        dim3 grid((unsigned)cdiv64(m, 256), (unsigned)cdiv64(n, cpc));
        gemv_n_part_kernel<<<grid, 256, 0, ctx->stream>>>(m, n, d_J, ld, d_x, cpc, part);
        LSO_CHECK_LAUNCH(ctx);
        gemv_n_combine_kernel<false><<<(unsigned)cdiv64(m, 256), 256, 0, ctx->stream>>>(m, (int)grid.y, part, alpha, beta, d_y,
                                                                                    nullptr, nullptr, nullptr, nullptr);
        LSO_CHECK_LAUNCH(ctx);
        return LSO_OK;
    }
    gemv_n_kernel<false><<<(unsigned)cdiv64(m, 256), 256, 0, ctx->stream>>>(m, n, alpha, d_J, ld, d_x, beta, d_y, nullptr,
                                                                          nullptr, nullptr, nullptr);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

}  // extern "C"

// fpredict = J δ - f, *d_out (device scalar) = sum(abs2, fpredict); no synchronisation
int dense_predicted_ssr_dev(lso_ctx* ctx, int64_t m, int64_t n, const double* d_J, int64_t ld, const double* d_delta,
                            const double* d_f, double* d_fpredict, double* d_out) {
    int64_t g = cdiv64(m, 256);
    LSO_REQUIRE(ctx, g <= LSO_PARTIALS, "m too large");
    const int cs = gemv_n_chunks(ctx, m, n);
    if (cs > 1) {
        double* part = nullptr;
        LSO_TRY(gemv_scratch(ctx, (size_t)cs * m, &part));
        const int64_t cpc = roundup64(cdiv64(n, cs), 8);
        dim3 grid((unsigned)g, (unsigned)cdiv64(n, cpc));
        gemv_n_part_kernel<<<grid, 256, 0, ctx->stream>>>(m, n, d_J, ld, d_delta, cpc, part);
        LSO_CHECK_LAUNCH(ctx);
        gemv_n_combine_kernel<true><<<(unsigned)g, 256, 0, ctx->stream>>>(m, (int)grid.y, part, 1.0, 0.0, d_fpredict, d_f,
                                                                         ctx->d_partials, ctx->d_counters + 1, d_out);
        LSO_CHECK_LAUNCH(ctx);
        return LSO_OK;
    }
    gemv_n_kernel<true><<<(unsigned)g, 256, 0, ctx->stream>>>(m, n, 1.0, d_J, ld, d_delta, 0.0, d_fpredict, d_f,
                                                             ctx->d_partials, ctx->d_counters + 1, d_out);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

extern "C" int lso_dense_predicted_ssr(lso_ctx* ctx, int64_t m, int64_t n, const double* d_J, int64_t ld,
                                       const double* d_delta, const double* d_f, double* d_fpredict, double* ssr_out) {
    LSO_REQUIRE(ctx, ctx && ssr_out, "NULL pointer");
    LSO_REQUIRE(ctx, m >= 0 && n >= 0 && ld >= m, "bad dimensions");
    if (m == 0) { *ssr_out = 0.0; return LSO_OK; }
    LSO_REQUIRE(ctx, d_f && (n == 0 || (d_J && d_delta)), "NULL pointer");
    LSO_ENTER(ctx);
    LSO_TRY(dense_predicted_ssr_dev(ctx, m, n, d_J, ld, d_delta, d_f, d_fpredict, ctx->d_scalars + 3));
    return lso_fetch_scalar(ctx, 3, ssr_out);
}

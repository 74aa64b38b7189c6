// core.cu — context lifecycle, device memory, and the vector duck-type kernels (G4 in SURVEY.md
// Appendix B): everything LeastSquaresOptim.jl's optimizers and lsmr! call on n- and m-vectors.
#include "common.cuh"
#include <stdarg.h>
#include <math.h>

std::string g_lso_last_error;

int lso_set_error(lso_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_lso_last_error = buf;
    if (ctx) ctx->last_error = buf;
    return code;
}

extern "C" {

int lso_version(void) { return 100; }

int lso_device_count(int* count) {
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return lso_set_error(nullptr, LSO_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return LSO_OK;
}

int lso_ctx_create(int device, lso_ctx** out) {
    if (!out) return lso_set_error(nullptr, LSO_ERR_ARG, "out is NULL");
    *out = nullptr;
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0)
        return lso_set_error(nullptr, LSO_ERR_CUDA,
                             "no CUDA device available (%s): the lsob200 product path has no CPU fallback",
                             cudaGetErrorString(e));
    if (device < 0 || device >= cnt) return lso_set_error(nullptr, LSO_ERR_ARG, "device %d out of range", device);
    lso_ctx* ctx = new (std::nothrow) lso_ctx();
    if (!ctx) return lso_set_error(nullptr, LSO_ERR_ALLOC, "host allocation failed");
    ctx->device = device;
    LSO_CHECK_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    LSO_CHECK_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    LSO_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    LSO_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_partials, LSO_PARTIALS * sizeof(double)));
    LSO_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_counters, 64 * sizeof(unsigned int)));
    LSO_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 64 * sizeof(unsigned int), ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_scalars, LSO_NSCALARS * sizeof(double)));
    LSO_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->d_scalars, 0, LSO_NSCALARS * sizeof(double), ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaMallocHost(&ctx->h_scalars, LSO_NSCALARS * sizeof(double)));
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = ctx;
    return LSO_OK;
}

int lso_comm_destroy(lso_ctx* ctx);

int lso_ctx_destroy(lso_ctx* ctx) {
    if (!ctx) return LSO_OK;
    cudaSetDevice(ctx->device);
    lso_comm_destroy(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_partials);
    cudaFree(ctx->d_counters);
    cudaFree(ctx->d_trimail);
    cudaFree(ctx->d_scalars);
    cudaFree(ctx->d_finish);
    cudaFree(ctx->d_gemv);
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->prof2_events) cudaEventDestroy(e);
    cudaFreeHost(ctx->h_scalars);
    if (ctx->copy_stream) {
        cudaStreamDestroy(ctx->copy_stream);
        for (auto& e : ctx->copy_ev) if (e) cudaEventDestroy(e);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return LSO_OK;
}

const char* lso_last_error(lso_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_lso_last_error.c_str(); }

int lso_ctx_sync(lso_ctx* ctx) {
    LSO_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    LSO_ENTER(ctx);
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LSO_OK;
}

void* lso_ctx_stream(lso_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int lso_ctx_set_option(lso_ctx* ctx, const char* key, int64_t value) {
    LSO_REQUIRE(ctx, ctx && key, "ctx/key is NULL");
    if (!strcmp(key, "qr_apply")) { ctx->opt_qr_apply = value; ctx->opt_qr_tune = 0; }
    else if (!strcmp(key, "syrk")) ctx->opt_syrk = value;
    else if (!strcmp(key, "qr_lookahead")) { ctx->opt_qr_lookahead = value; ctx->opt_qr_tune = 0; }
    else if (!strcmp(key, "qr_tune")) ctx->opt_qr_tune = value;
    else if (!strcmp(key, "qr_twin")) ctx->opt_qr_twin = value;
    else if (!strcmp(key, "qr_shard_pipeline")) ctx->opt_qr_shard_pipeline = value;
    else if (!strcmp(key, "spmv")) ctx->opt_spmv = value;
    else if (!strcmp(key, "trisolve")) ctx->opt_trisolve = value;
    else if (!strcmp(key, "ozaki_slices")) { if (value < 2 || value > 8) return lso_set_error(ctx, LSO_ERR_ARG, "ozaki_slices must be 2..8"); ctx->opt_ozaki_slices = value; }
    else if (!strcmp(key, "lsmr_fused")) ctx->opt_lsmr_fused = value;
    else if (!strcmp(key, "profile")) { ctx->opt_profile = value; ctx->prof_used = 0; ctx->prof2_used = 0; }
    else return lso_set_error(ctx, LSO_ERR_ARG, "unknown option '%s'", key);
    return LSO_OK;
}

int lso_ctx_launch_count(lso_ctx* ctx, int64_t* out, int reset) {
    LSO_REQUIRE(ctx, ctx && out, "ctx/out is NULL");
    *out = ctx->launches;
    if (reset) ctx->launches = 0;
    return LSO_OK;
}

int lso_ctx_profile_read(lso_ctx* ctx, double* total_ms, int64_t* launches) {
    LSO_REQUIRE(ctx, ctx && total_ms && launches, "NULL pointer");
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < ctx->prof_used; i += 2) {
        float ms = 0.f;
        LSO_CHECK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_events[i], ctx->prof_events[i + 1]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = (int64_t)(ctx->prof_used / 2);
    ctx->prof_used = 0;
    return LSO_OK;
}

int lso_ctx_stat(lso_ctx* ctx, const char* key, double* out, int reset) {
    LSO_REQUIRE(ctx, ctx && key && out, "NULL pointer");
    double* p = nullptr;
    if (!strcmp(key, "qr_update_flops")) p = &ctx->stat_qr_update_flops;
    else if (!strcmp(key, "qr_flops")) p = &ctx->stat_qr_flops;
    else if (!strcmp(key, "syrk_flops")) p = &ctx->stat_syrk_flops;
    else if (!strcmp(key, "syrk_i8_macs")) p = &ctx->stat_syrk_i8_macs;
    else if (!strcmp(key, "spmv_bytes")) p = &ctx->stat_spmv_bytes;
    else return lso_set_error(ctx, LSO_ERR_ARG, "unknown statistic '%s'", key);
    *out = *p;
    if (reset) *p = 0.0;
    return LSO_OK;
}

int lso_ctx_profile_read_collective(lso_ctx* ctx, double* total_ms, int64_t* launches) {
    LSO_REQUIRE(ctx, ctx && total_ms && launches, "NULL pointer");
    LSO_ENTER(ctx);
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < ctx->prof2_used; i += 2) {
        float ms = 0.f;
        LSO_CHECK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof2_events[i], ctx->prof2_events[i + 1]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = (int64_t)(ctx->prof2_used / 2);
    ctx->prof2_used = 0;
    return LSO_OK;
}

// ---- memory -------------------------------------------------------------------------------------
int lso_dev_alloc(lso_ctx* ctx, size_t nbytes, void** d_out) {
    LSO_REQUIRE(ctx, ctx && d_out, "ctx/d_out is NULL");
    *d_out = nullptr;
    if (nbytes == 0) nbytes = 16;
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(d_out, nbytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return lso_set_error(ctx, LSO_ERR_ALLOC, "cudaMalloc(%zu bytes): %s", nbytes, cudaGetErrorString(e));
    }
    return LSO_OK;
}
int lso_dev_free(lso_ctx* ctx, void* d_ptr) {
    LSO_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (!d_ptr) return LSO_OK;
    LSO_ENTER(ctx);
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaFree(d_ptr));
    return LSO_OK;
}
int lso_host_alloc_pinned(lso_ctx* ctx, size_t nbytes, void** h_out) {
    LSO_REQUIRE(ctx, ctx && h_out, "ctx/h_out is NULL");
    if (nbytes == 0) nbytes = 16;
    LSO_ENTER(ctx);
    cudaError_t e = cudaMallocHost(h_out, nbytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return lso_set_error(ctx, LSO_ERR_ALLOC, "cudaMallocHost(%zu bytes): %s", nbytes, cudaGetErrorString(e));
    }
    return LSO_OK;
}
// Page-lock memory the CALLER owns (a Julia Array, a numpy array): `optimize!` re-uses J, x and y for the whole run
// (types.jl:141-157), so the glue registers them once when the problem is allocated and every later H2D copy of J runs at
// pinned-memory speed (measured: 30 ms per LM step at 100 000 x 1 000 against 85 ms from pageable memory).
int lso_host_register(lso_ctx* ctx, void* h_ptr, size_t nbytes) {
    LSO_REQUIRE(ctx, ctx && h_ptr, "NULL pointer");
    LSO_ENTER(ctx);
    cudaError_t e = cudaHostRegister(h_ptr, nbytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return LSO_OK; }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return lso_set_error(ctx, LSO_ERR_ALLOC, "cudaHostRegister(%zu bytes): %s", nbytes, cudaGetErrorString(e));
    }
    return LSO_OK;
}
int lso_host_unregister(lso_ctx* ctx, void* h_ptr) {
    LSO_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (h_ptr) {
        cudaError_t e = cudaHostUnregister(h_ptr);
        if (e != cudaSuccess) cudaGetLastError();
    }
    return LSO_OK;
}
int lso_host_free_pinned(lso_ctx* ctx, void* h_ptr) {
    LSO_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (h_ptr) LSO_CHECK_CUDA(ctx, cudaFreeHost(h_ptr));
    return LSO_OK;
}
int lso_upload_async(lso_ctx* ctx, void* d_dst, const void* h_src, size_t nbytes) {
    LSO_REQUIRE(ctx, ctx && (nbytes == 0 || (d_dst && h_src)), "NULL pointer");
    LSO_ENTER(ctx);
    if (nbytes) LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    return LSO_OK;
}
int lso_download_async(lso_ctx* ctx, void* h_dst, const void* d_src, size_t nbytes) {
    LSO_REQUIRE(ctx, ctx && (nbytes == 0 || (h_dst && d_src)), "NULL pointer");
    LSO_ENTER(ctx);
    if (nbytes) LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    return LSO_OK;
}
int lso_upload(lso_ctx* ctx, void* d_dst, const void* h_src, size_t nbytes) {
    LSO_TRY(lso_upload_async(ctx, d_dst, h_src, nbytes));
    return lso_ctx_sync(ctx);
}
int lso_download(lso_ctx* ctx, void* h_dst, const void* d_src, size_t nbytes) {
    LSO_TRY(lso_download_async(ctx, h_dst, d_src, nbytes));
    return lso_ctx_sync(ctx);
}
int lso_upload_matrix(lso_ctx* ctx, double* d_dst, int64_t ld_dst, const double* h_src, int64_t ld_src,
                      int64_t rows, int64_t cols) {
    LSO_REQUIRE(ctx, ctx && d_dst && h_src, "NULL pointer");
    LSO_REQUIRE(ctx, ld_dst >= rows && ld_src >= rows, "leading dimension < rows");
    if (rows == 0 || cols == 0) return LSO_OK;
    LSO_ENTER(ctx);
    LSO_CHECK_CUDA(ctx, cudaMemcpy2DAsync(d_dst, ld_dst * sizeof(double), h_src, ld_src * sizeof(double),
                                          rows * sizeof(double), cols, cudaMemcpyHostToDevice, ctx->stream));
    return LSO_OK;
}

}  // extern "C"

// =================================================================================================
// elementwise kernels
// =================================================================================================
enum { EW_FILL, EW_COPY, EW_SCAL, EW_AXPY, EW_AXPBY, EW_MUL, EW_DIV, EW_SQRT, EW_CLAMP };

template <int OP>
__global__ void ew_kernel(int64_t n, double* __restrict__ out, const double* __restrict__ x,
                          const double* __restrict__ y, double a, double b) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double r;
        if (OP == EW_FILL) r = a;
        else if (OP == EW_COPY) r = x[i];
        else if (OP == EW_SCAL) r = out[i] * a;
        else if (OP == EW_AXPY) r = fma(a, x[i], out[i]);          // y += a x
        else if (OP == EW_AXPBY) r = (b == 0.0) ? a * x[i] : fma(a, x[i], b * out[i]);
        else if (OP == EW_MUL) r = x[i] * y[i];
        else if (OP == EW_DIV) r = x[i] / y[i];
        else if (OP == EW_SQRT) r = sqrt(out[i]);
        else /* EW_CLAMP */ { double v = out[i]; r = v < a ? a : (v > b ? b : v); }   // Julia clamp: NaN passes through
        out[i] = r;
    }
}

static inline int ew_grid(lso_ctx* ctx, int64_t n) {
    int64_t g = cdiv64(n, 256);
    int64_t cap = (int64_t)ctx->num_sms * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <int OP>
static int ew_launch(lso_ctx* ctx, int64_t n, double* out, const double* x, const double* y, double a, double b) {
    LSO_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    LSO_REQUIRE(ctx, n >= 0, "negative length");
    if (n == 0) return LSO_OK;
    LSO_REQUIRE(ctx, out != nullptr, "NULL vector");
    LSO_ENTER(ctx);
    ew_kernel<OP><<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(n, out, x, y, a, b);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

// =================================================================================================
// reductions: one launch, per-block partials + last-block final pass in fixed order (deterministic)
// =================================================================================================
enum { RD_SUM, RD_SUMABS2, RD_MAXABS, RD_DOT, RD_WDOT, RD_MAXABS_PROJ };

template <int OP>
__device__ __forceinline__ double rd_elem(int64_t i, const double* __restrict__ x, const double* __restrict__ y,
                                          const double* __restrict__ w, const double* __restrict__ lo,
                                          const double* __restrict__ hi) {
    if (OP == RD_SUM) return x[i];
    if (OP == RD_SUMABS2) { double v = x[i]; return v * v; }
    if (OP == RD_MAXABS) return fabs(x[i]);
    if (OP == RD_DOT) return x[i] * y[i];
    if (OP == RD_WDOT) return w[i] * x[i] * y[i];          // utils.jl:170: out += w[i] * x[i] * y[i]
    // RD_MAXABS_PROJ: utils.jl:44-53 — x is g, y is the iterate
    double gi = x[i];
    if (lo && y[i] <= lo[i] && gi > 0.0) gi = 0.0;
    else if (hi && y[i] >= hi[i] && gi < 0.0) gi = 0.0;
    return fabs(gi);
}

template <int OP>
__global__ void reduce_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                              const double* __restrict__ w, const double* __restrict__ lo,
                              const double* __restrict__ hi, double* __restrict__ partials,
                              unsigned int* __restrict__ counter, double* __restrict__ out) {
    __shared__ double sm[32];
    __shared__ bool is_last;
    constexpr bool ISMAX = (OP == RD_MAXABS || OP == RD_MAXABS_PROJ);
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = rd_elem<OP>(i, x, y, w, lo, hi);
        if (ISMAX) acc = (OP == RD_MAXABS) ? nanmax(acc, v) : ((v > acc) ? v : acc);   // utils.jl:52: a > m && (m = a)
        else acc += v;
    }
    acc = ISMAX ? block_nanmax(acc, sm) : block_sum(acc, sm);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = acc;
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double a2 = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
            double v = ((volatile double*)partials)[i];
            if (ISMAX) a2 = nanmax(a2, v); else a2 += v;
        }
        a2 = ISMAX ? block_nanmax(a2, sm) : block_sum(a2, sm);
        if (threadIdx.x == 0) {
            *out = a2;
            *counter = 0;
        }
    }
}

// Launch a reduction whose result lands in device memory d_out (no sync).
template <int OP>
int rd_launch_dev(lso_ctx* ctx, int64_t n, const double* x, const double* y, const double* w, const double* lo,
                  const double* hi, double* d_out) {
    LSO_REQUIRE(ctx, n >= 0, "negative length");
    LSO_REQUIRE(ctx, n == 0 || x != nullptr, "NULL vector");
    LSO_ENTER(ctx);
    int64_t g = cdiv64(n, 256 * 4);
    if (g < 1) g = 1;
    int64_t cap = (int64_t)ctx->num_sms * 8;
    if (g > cap) g = cap;
    reduce_kernel<OP><<<(int)g, 256, 0, ctx->stream>>>(n, x, y, w, lo, hi, ctx->d_partials, ctx->d_counters, d_out);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

int lso_fetch_scalar(lso_ctx* ctx, int slot, double* out) {
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars + slot, ctx->d_scalars + slot, sizeof(double),
                                        cudaMemcpyDeviceToHost, ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = ctx->h_scalars[slot];
    return LSO_OK;
}

template <int OP>
static int rd_launch(lso_ctx* ctx, int64_t n, const double* x, const double* y, const double* w, const double* lo,
                     const double* hi, double* out) {
    LSO_REQUIRE(ctx, ctx && out, "ctx/out is NULL");
    LSO_TRY((rd_launch_dev<OP>(ctx, n, x, y, w, lo, hi, ctx->d_scalars + 0)));
    return lso_fetch_scalar(ctx, 0, out);
}

// explicit instantiations used from other translation units
int lso_dev_sumabs2(lso_ctx* ctx, int64_t n, const double* x, double* d_out) {
    return rd_launch_dev<RD_SUMABS2>(ctx, n, x, nullptr, nullptr, nullptr, nullptr, d_out);
}
int lso_dev_sum(lso_ctx* ctx, int64_t n, const double* x, double* d_out) {
    return rd_launch_dev<RD_SUM>(ctx, n, x, nullptr, nullptr, nullptr, nullptr, d_out);
}
int lso_dev_maxabs(lso_ctx* ctx, int64_t n, const double* x, double* d_out) {
    return rd_launch_dev<RD_MAXABS>(ctx, n, x, nullptr, nullptr, nullptr, nullptr, d_out);
}
int lso_dev_maxabs_projected(lso_ctx* ctx, int64_t n, const double* g, const double* x, const double* lo, const double* hi,
                             double* d_out) {
    if (!lo && !hi) return rd_launch_dev<RD_MAXABS>(ctx, n, g, nullptr, nullptr, nullptr, nullptr, d_out);
    return rd_launch_dev<RD_MAXABS_PROJ>(ctx, n, g, x, nullptr, lo, hi, d_out);
}

// ---- LM damping (levenberg_marquardt.jl:84-86), mean read from a device scalar -------------------
__global__ void lm_damping_kernel(int64_t n, double* __restrict__ dtd, const double* __restrict__ d_sum,
                                  double min_diag, double max_diag, double inv_delta) {
    const double mean = *d_sum / (double)n;
    const double lo = min_diag * mean, hi = max_diag * mean;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = dtd[i];
        v = v < lo ? lo : (v > hi ? hi : v);
        dtd[i] = v * inv_delta;
    }
}

// ---- box projection (LM:89-98; dogleg:148-157) ---------------------------------------------------
__global__ void box_project_kernel(int64_t n, double* __restrict__ d, const double* __restrict__ x,
                                   const double* __restrict__ lo, const double* __restrict__ hi) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = d[i];
        // Julia's min / max propagate NaN (fmin / fmax would silently clamp a NaN step to the bound)
        if (lo) { const double b = x[i] - lo[i]; v = (v != v || b != b) ? NAN : (b < v ? b : v); }
        if (hi) { const double b = x[i] - hi[i]; v = (v != v || b != b) ? NAN : (b > v ? b : v); }
        d[i] = v;
    }
}

// ---- check_isfinite (utils.jl:70-75): index of the first non-finite entry, or n ------------------
__global__ void first_nonfinite_kernel(int64_t n, const double* __restrict__ x, unsigned long long* __restrict__ first) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (!isfinite(x[i])) atomicMin(first, (unsigned long long)i);
    }
}

// ---- dogleg blend (dogleg.jl:137-143): δx = β δgn + α(1-β) δgr ------------------------------------
__global__ void blend_kernel(int64_t n, double* __restrict__ dx, const double* __restrict__ gn,
                             const double* __restrict__ gr, double beta, double coef) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = gn[i] * beta;          // copyto!; rmul!(δx, β)
        dx[i] = fma(coef, gr[i], v);      // axpy!(α(1-β), δgr, δx)
    }
}

extern "C" {

int lso_vec_fill(lso_ctx* ctx, int64_t n, double* d_x, double value) {
    return ew_launch<EW_FILL>(ctx, n, d_x, nullptr, nullptr, value, 0);
}
int lso_vec_copy(lso_ctx* ctx, int64_t n, double* d_dst, const double* d_src) {
    return ew_launch<EW_COPY>(ctx, n, d_dst, d_src, nullptr, 0, 0);
}
int lso_vec_scal(lso_ctx* ctx, int64_t n, double* d_x, double alpha) {
    return ew_launch<EW_SCAL>(ctx, n, d_x, nullptr, nullptr, alpha, 0);
}
int lso_vec_axpy(lso_ctx* ctx, int64_t n, double alpha, const double* d_x, double* d_y) {
    return ew_launch<EW_AXPY>(ctx, n, d_y, d_x, nullptr, alpha, 0);
}
int lso_vec_axpby(lso_ctx* ctx, int64_t n, double alpha, const double* d_x, double beta, double* d_y) {
    return ew_launch<EW_AXPBY>(ctx, n, d_y, d_x, nullptr, alpha, beta);
}
int lso_vec_mul(lso_ctx* ctx, int64_t n, double* d_out, const double* d_x, const double* d_y) {
    return ew_launch<EW_MUL>(ctx, n, d_out, d_x, d_y, 0, 0);
}
int lso_vec_div(lso_ctx* ctx, int64_t n, double* d_out, const double* d_x, const double* d_y) {
    return ew_launch<EW_DIV>(ctx, n, d_out, d_x, d_y, 0, 0);
}
int lso_vec_sqrt(lso_ctx* ctx, int64_t n, double* d_x) {
    return ew_launch<EW_SQRT>(ctx, n, d_x, nullptr, nullptr, 0, 0);
}
int lso_vec_clamp(lso_ctx* ctx, int64_t n, double* d_x, double lo, double hi) {
    return ew_launch<EW_CLAMP>(ctx, n, d_x, nullptr, nullptr, lo, hi);
}
int lso_vec_sum(lso_ctx* ctx, int64_t n, const double* d_x, double* out) {
    return rd_launch<RD_SUM>(ctx, n, d_x, nullptr, nullptr, nullptr, nullptr, out);
}
int lso_vec_sumabs2(lso_ctx* ctx, int64_t n, const double* d_x, double* out) {
    return rd_launch<RD_SUMABS2>(ctx, n, d_x, nullptr, nullptr, nullptr, nullptr, out);
}
int lso_vec_nrm2(lso_ctx* ctx, int64_t n, const double* d_x, double* out) {
    LSO_TRY((rd_launch<RD_SUMABS2>(ctx, n, d_x, nullptr, nullptr, nullptr, nullptr, out)));
    *out = sqrt(*out);
    return LSO_OK;
}
int lso_vec_maxabs(lso_ctx* ctx, int64_t n, const double* d_x, double* out) {
    return rd_launch<RD_MAXABS>(ctx, n, d_x, nullptr, nullptr, nullptr, nullptr, out);
}
int lso_vec_dot(lso_ctx* ctx, int64_t n, const double* d_x, const double* d_y, double* out) {
    LSO_REQUIRE(ctx, n == 0 || d_y, "NULL vector");
    return rd_launch<RD_DOT>(ctx, n, d_x, d_y, nullptr, nullptr, nullptr, out);
}
int lso_vec_wdot(lso_ctx* ctx, int64_t n, const double* d_x, const double* d_y, const double* d_w, double* out) {
    LSO_REQUIRE(ctx, n == 0 || (d_y && d_w), "NULL vector");
    return rd_launch<RD_WDOT>(ctx, n, d_x, d_y, d_w, nullptr, nullptr, out);
}
int lso_vec_maxabs_projected(lso_ctx* ctx, int64_t n, const double* d_g, const double* d_x, const double* d_lower,
                             const double* d_upper, double* out) {
    if (!d_lower && !d_upper)   // utils.jl:42: (haslower || hasupper) || return maximum(abs, g)
        return rd_launch<RD_MAXABS>(ctx, n, d_g, nullptr, nullptr, nullptr, nullptr, out);
    LSO_REQUIRE(ctx, n == 0 || d_x, "NULL vector");
    return rd_launch<RD_MAXABS_PROJ>(ctx, n, d_g, d_x, nullptr, d_lower, d_upper, out);
}

int lso_vec_check_finite(lso_ctx* ctx, int64_t n, const double* d_x, int64_t* first_bad) {
    LSO_REQUIRE(ctx, ctx && first_bad, "ctx/first_bad is NULL");
    *first_bad = -1;
    if (n == 0) return LSO_OK;
    LSO_ENTER(ctx);
    unsigned long long* d_first = (unsigned long long*)(ctx->d_scalars + 1);
    unsigned long long init = (unsigned long long)n;
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(d_first, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    first_nonfinite_kernel<<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(n, d_x, d_first);
    LSO_CHECK_LAUNCH(ctx);
    unsigned long long res = 0;
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(&res, d_first, sizeof(res), cudaMemcpyDeviceToHost, ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if ((int64_t)res < n) {
        *first_bad = (int64_t)res;
        return LSO_ERR_NOT_FINITE;
    }
    return LSO_OK;
}

int lso_vec_box_project(lso_ctx* ctx, int64_t n, double* d_delta, const double* d_x, const double* d_lower,
                        const double* d_upper) {
    LSO_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (n == 0 || (!d_lower && !d_upper)) return LSO_OK;
    LSO_REQUIRE(ctx, d_delta && d_x, "NULL vector");
    LSO_ENTER(ctx);
    box_project_kernel<<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(n, d_delta, d_x, d_lower, d_upper);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

int lso_lm_damping(lso_ctx* ctx, int64_t n, double* d_dtd, double min_diag, double max_diag, double inv_delta) {
    LSO_REQUIRE(ctx, ctx && (n == 0 || d_dtd), "NULL pointer");
    if (n == 0) return LSO_OK;
    LSO_ENTER(ctx);
    LSO_TRY(lso_dev_sum(ctx, n, d_dtd, ctx->d_scalars + 2));
    lm_damping_kernel<<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(n, d_dtd, ctx->d_scalars + 2, min_diag, max_diag, inv_delta);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

int lso_dogleg_blend(lso_ctx* ctx, int64_t n, double* d_dx, const double* d_gn, const double* d_gr,
                     const double* d_dtd, double delta, double alpha, double wnorm_gn, double wnorm_gr,
                     double* wnorm_dx_out) {
    LSO_REQUIRE(ctx, ctx && wnorm_dx_out, "NULL pointer");
    if (wnorm_gn <= delta) {                       // dogleg.jl:120-123
        LSO_TRY(lso_vec_copy(ctx, n, d_dx, d_gn));
        *wnorm_dx_out = wnorm_gn;
    } else if (wnorm_gr * alpha >= delta) {        // dogleg.jl:124-130
        LSO_TRY(lso_vec_copy(ctx, n, d_dx, d_gr));
        LSO_TRY(lso_vec_scal(ctx, n, d_dx, delta / wnorm_gr));
        *wnorm_dx_out = delta;
    } else {                                       // dogleg.jl:131-145
        double wd = 0;
        LSO_TRY(lso_vec_wdot(ctx, n, d_gr, d_gn, d_dtd, &wd));
        double b_dot_a = alpha * wd;
        double a_sq = (alpha * wnorm_gr) * (alpha * wnorm_gr);
        double bma_sq = a_sq - 2 * b_dot_a + wnorm_gn * wnorm_gn;
        double c = b_dot_a - a_sq;
        double d = sqrt(c * c + bma_sq * (delta * delta - a_sq));
        double beta = (c <= 0) ? (d - c) / bma_sq : (delta * delta - a_sq) / (d + c);
        if (n > 0) {
            blend_kernel<<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(n, d_dx, d_gn, d_gr, beta, alpha * (1 - beta));
            LSO_CHECK_LAUNCH(ctx);
        }
        double w2 = 0;
        LSO_TRY(lso_vec_wdot(ctx, n, d_dx, d_dx, d_dtd, &w2));
        *wnorm_dx_out = sqrt(w2);
    }
    return LSO_OK;
}

}  // extern "C"

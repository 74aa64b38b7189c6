// chol.cu — DenseCholeskyAllocatedSolver (src/solver/dense_cholesky.jl:29-59) on the device:
//   C-a  syrk   cholm = J'J            (dense_cholesky.jl:31,48)  — fp64 tensor pipe (DMMA), split-K
//        gemv   x = J'y                (dense_cholesky.jl:32,56)
//   C-b  one NCCL all-reduce of [J'J | J'y] when J is row-sharded over ranks
//   C-c  cholm[i,i] += damp[i] (:51-53), upper Cholesky (:33,57), two triangular solves
#include "chol.cuh"
#include "qr.cuh"
#include <limits.h>

// ---------------------------------------------------------------------------------------------------
// syrk on the tensor pipe.  CTA tile 128 x 128 of C, 8 warps each 64 x 32 (8 x 4 DMMA accumulators),
// K (rows of J) streamed in chunks of 16 through a 4-stage cp.async ring.  Both operands are slices of
// J itself: column-major J makes them "K-major", so the 4x8 / 8x4 DMMA fragments are 4 consecutive
// rows of 8 columns; the padded column stride SY_SK = 20 doubles spreads them over 16 distinct banks.
// ---------------------------------------------------------------------------------------------------
#define SY_TS 128
#define SY_KC 16
#define SY_SK 20
#define SY_NST 4
#define SY_STAGE_DOUBLES (2 * SY_TS * SY_SK)
#define SY_SMEM_BYTES (SY_NST * SY_STAGE_DOUBLES * 8)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void dmma2(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void decode_pair(int p, int& bi, int& bj) {
    int j = 0;
    while ((j + 1) * (j + 2) / 2 <= p) ++j;
    bj = j;
    bi = p - j * (j + 1) / 2;
}

// SUB = false: out = J'J (per split-K slab).  SUB = true: out -= J'J on the upper triangle only — the rank-w update of the
// blocked Cholesky (J = the w x rest row block of the factor just computed, out = the trailing matrix).
template <bool ALIGN16, bool SUB = false>
__global__ void __launch_bounds__(256, 1)
syrk_mma_kernel(long long m, long long n, const double* __restrict__ J, long long ld, long long rows_per_split,
                double* __restrict__ out, long long ldc, long long slab_stride) {
    extern __shared__ __align__(16) double sysm[];
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wi = wrp >> 2, wj = wrp & 3;
    int bi, bj;
    decode_pair(blockIdx.x, bi, bj);
    const long long r_begin = (long long)blockIdx.y * rows_per_split;
    long long r_end = r_begin + rows_per_split;
    if (r_end > m) r_end = m;
    const int niter = (r_end > r_begin) ? (int)((r_end - r_begin + SY_KC - 1) / SY_KC) : 0;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(sysm);

    auto load_stage = [&](int it) {
        const int s = it % SY_NST;
        const long long rk = r_begin + (long long)it * SY_KC;
        const uint32_t sbase = smem_base + (uint32_t)(s * SY_STAGE_DOUBLES) * 8u;
        if (ALIGN16) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = tid + 256 * u;
                const int col = e >> 3, part = e & 7;
                const int which = col >> 7, cc = col & 127;
                const long long gcol = (long long)(which ? bj : bi) * SY_TS + cc;
                const long long row = rk + 2 * part;
                long long rem = (r_end - row) * 8;
                int bytes = (gcol < n && rem > 0) ? (int)(rem > 16 ? 16 : rem) : 0;
                const double* src = bytes ? (J + gcol * ld + row) : J;
                cp_async16(sbase + (uint32_t)(which * SY_TS * SY_SK + cc * SY_SK + 2 * part) * 8u, src, bytes);
            }
        } else {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int e = tid + 256 * u;
                const int col = e >> 4, part = e & 15;
                const int which = col >> 7, cc = col & 127;
                const long long gcol = (long long)(which ? bj : bi) * SY_TS + cc;
                const long long row = rk + part;
                int bytes = (gcol < n && row < r_end) ? 8 : 0;
                const double* src = bytes ? (J + gcol * ld + row) : J;
                cp_async8(sbase + (uint32_t)(which * SY_TS * SY_SK + cc * SY_SK + part) * 8u, src, bytes);
            }
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
    for (int s = 0; s < SY_NST - 1; ++s) {
        if (s < niter) load_stage(s);
        cp_async_commit();
    }
    for (int it = 0; it < niter; ++it) {
        cp_async_wait<SY_NST - 2>();
        __syncthreads();
        if (it + SY_NST - 1 < niter) load_stage(it + SY_NST - 1);
        cp_async_commit();
        const double* As = sysm + (it % SY_NST) * SY_STAGE_DOUBLES;
        const double* Bs = As + SY_TS * SY_SK;
#pragma unroll
        for (int ks = 0; ks < SY_KC / 4; ++ks) {
            double a[8], b[4];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) a[mi] = As[(64 * wi + 8 * mi + g) * SY_SK + 4 * ks + t];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[(32 * wj + 8 * ni + g) * SY_SK + 4 * ks + t];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma2(acc[mi][ni], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();
    double* __restrict__ o = out + (long long)blockIdx.y * slab_stride;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        const long long i = (long long)bi * SY_TS + 64 * wi + 8 * mi + g;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const long long j = (long long)bj * SY_TS + 32 * wj + 8 * ni + 2 * t;
            if (i < n) {
                if (SUB) {
                    if (j < n && i <= j) o[j * ldc + i] -= acc[mi][ni][0];
                    if (j + 1 < n && i <= j + 1) o[(j + 1) * ldc + i] -= acc[mi][ni][1];
                } else {
                    if (j < n) o[j * ldc + i] = acc[mi][ni][0];
                    if (j + 1 < n) o[(j + 1) * ldc + i] = acc[mi][ni][1];
                }
            }
        }
    }
}

// plain-FMA syrk (cross-check path; ctx option syrk = 0): 64 x 64 tile, 4 x 4 per thread
__global__ void __launch_bounds__(256)
syrk_fma_kernel(long long m, long long n, const double* __restrict__ J, long long ld, long long rows_per_split,
                double* __restrict__ out, long long ldc, long long slab_stride) {
    __shared__ double As[16][65], Bs[16][65];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    int bi, bj;
    decode_pair(blockIdx.x, bi, bj);
    const long long r_begin = (long long)blockIdx.y * rows_per_split;
    long long r_end = r_begin + rows_per_split;
    if (r_end > m) r_end = m;
    double acc[4][4] = {};
    for (long long rk = r_begin; rk < r_end; rk += 16) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + 256 * u;
            const int k = e & 15, c = e >> 4;
            const long long row = rk + k;
            const long long ca = (long long)bi * 64 + c, cb = (long long)bj * 64 + c;
            As[k][c] = (row < r_end && ca < n) ? J[ca * ld + row] : 0.0;
            Bs[k][c] = (row < r_end && cb < n) ? J[cb * ld + row] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = As[k][4 * ty + u]; b[u] = Bs[k][4 * tx + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
        }
    }
    double* __restrict__ o = out + (long long)blockIdx.y * slab_stride;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const long long i = (long long)bi * 64 + 4 * ty + u, j = (long long)bj * 64 + 4 * tx + v;
            if (i < n && j < n) o[j * ldc + i] = acc[u][v];
        }
}

__global__ void slab_reduce_kernel(long long count, int nslab, const double* __restrict__ part, long long slab_stride,
                                   double* __restrict__ C) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        double a = 0.0;
        for (int z = 0; z < nslab; ++z) a += part[z * slab_stride + i];
        C[i] = a;
    }
}

__global__ void add_diag_kernel(long long n, double* __restrict__ C, long long ldc, const double* __restrict__ damp) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) C[i * ldc + i] += damp[i];
}

// ---------------------------------------------------------------------------------------------------
// blocked right-looking upper Cholesky:  for each 32-column block k:
//   panel kernel : R11 = chol(C11) and R11^{-1} (every CTA redundantly, one warp, registers), R12 = R11^{-T} C12
//   update kernel: C22 -= R12' R12 on the upper 64 x 64 tiles
// ---------------------------------------------------------------------------------------------------
#define PP_THREADS 128
__global__ void __launch_bounds__(PP_THREADS)
potrf_panel_kernel(int n, double* __restrict__ C, long long ldc, int k0, int* __restrict__ info) {
    // R11 = chol(C11) in registers of warp 0 (lane = column, every CTA redundantly; row j is published to shared memory at
    // step j).  Then every thread takes one column of C12 through the forward substitution R11' y = c in AXPY form: the
    // updates of the later entries are independent, the dependent chain is one multiply + one FMA per step — the same FMA
    // count as a product with an explicit inverse, without the 32-step inversion in front of it (and 3 000 fewer
    // straight-line instructions: a third of the first version's stall samples were instruction-fetch misses).  A
    // shared-memory Cholesky with rolled loops was tried and is slower (8.1 vs 7.1 ms per potrf at n = 4 000).
    __shared__ __align__(16) double Rm[32][34];    // Rm[j][c] = R11[j][c]   (row j published at step j)
    __shared__ double invd[32];                    // 1 / R11[j][j]
    const int tid = threadIdx.x, lane = tid & 31;
    const int w = (n - k0 < 32) ? (n - k0) : 32;
    // this thread's column of C12 (independent of R11: in flight during the factorisation)
    const int c12 = k0 + w + blockIdx.x * PP_THREADS + tid;
    double cv[32];
    {
        const double* __restrict__ col = C + (long long)(c12 < n ? c12 : k0) * ldc + k0;
#pragma unroll
        for (int j = 0; j < 32; ++j) cv[j] = (c12 < n && j < w) ? col[j] : 0.0;
    }
    if (tid < 32) {
        double a[32];
        {
            const double* __restrict__ col = C + (long long)(k0 + (lane < w ? lane : 0)) * ldc + k0;
#pragma unroll
            for (int r = 0; r < 32; ++r) a[r] = (lane < w && r <= lane) ? col[r] : ((r == lane) ? 1.0 : 0.0);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            double d = __shfl_sync(0xffffffffu, a[j], j);
            if (j < w && !(d > 0.0)) {
                if (blockIdx.x == 0 && lane == 0) atomicMin(info, k0 + j + 1);
                d = 1.0;
            }
            const double rs = rsqrt(d);
            const double rjc = (lane > j) ? a[j] * rs : ((lane == j) ? d * rs : 0.0);     // R[j][lane]
            a[j] = rjc;
            Rm[j][lane] = rjc;
            if (lane == j) invd[j] = rs;
            __syncwarp();
#pragma unroll
            for (int i = j + 1; i < 32; ++i) {
                const double rji = Rm[j][i];
                if (i <= lane) a[i] = fma(-rji, rjc, a[i]);
            }
        }
        if (blockIdx.x == 0 && lane < w) {
            double* __restrict__ col = C + (long long)(k0 + lane) * ldc + k0;
#pragma unroll
            for (int r = 0; r < 32; ++r)
                if (r <= lane) col[r] = a[r];
        }
    }
    __syncthreads();
    if (c12 < n) {
        // y = R11^{-T} c:  y_i = c_i / R_ii, then c_j -= R[i][j] y_i for j > i
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const double yi = cv[i] * invd[i];
            cv[i] = yi;
#pragma unroll
            for (int j = (i + 1) & ~1; j < 32; j += 2) {
                const double2 r = *reinterpret_cast<const double2*>(&Rm[i][j]);
                if (j > i) cv[j] = fma(-r.x, yi, cv[j]);
                cv[j + 1] = fma(-r.y, yi, cv[j + 1]);
            }
        }
        double* __restrict__ col = C + (long long)c12 * ldc + k0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < w) col[j] = cv[j];
    }
}

__global__ void __launch_bounds__(256)
potrf_update_kernel(int n, double* __restrict__ C, long long ldc, int k0, int w) {
    __shared__ double Ri[32][65], Rj[32][65];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    int bi, bj;
    decode_pair(blockIdx.x, bi, bj);
    const int t0 = k0 + w;
    const int i0 = t0 + bi * 64, j0 = t0 + bj * 64;
    for (int e = tid; e < 32 * 64; e += 256) {
        const int k = e & 31, c = e >> 5;
        Ri[k][c] = (k < w && i0 + c < n) ? C[(long long)(i0 + c) * ldc + k0 + k] : 0.0;
        Rj[k][c] = (k < w && j0 + c < n) ? C[(long long)(j0 + c) * ldc + k0 + k] : 0.0;
    }
    __syncthreads();
    double acc[4][4] = {};
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = Ri[k][4 * ty + u]; b[u] = Rj[k][4 * tx + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int i = i0 + 4 * ty + u, j = j0 + 4 * tx + v;
            if (i < n && j < n && i <= j) C[(long long)j * ldc + i] -= acc[u][v];
        }
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

int chol_plan_create(lso_ctx* ctx, int64_t n, CholPlan* p) {
    *p = CholPlan();
    LSO_REQUIRE(ctx, n <= 24000, "Cholesky path: n too large (the single-CTA triangular solves hold the right-hand side in shared memory)");
    p->n = n;
    p->ldc = roundup64(n, 32);
    size_t cbytes = (size_t)(p->ldc * n + roundup64(n, 32)) * sizeof(double);
    cudaError_t e = cudaMalloc(&p->C, cbytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return lso_set_error(ctx, LSO_ERR_ALLOC, "Cholesky workspace: cudaMalloc(%zu): %s", cbytes, cudaGetErrorString(e));
    }
    LSO_CHECK_CUDA(ctx, cudaMemsetAsync(p->C, 0, cbytes, ctx->stream));
    p->rhs = p->C + p->ldc * n;
    int64_t ntb = cdiv64(n, SY_TS);
    int64_t pairs = ntb * (ntb + 1) / 2;
    // split K so that the launch fills whole waves of SMs (one 128 x 128 tile CTA per SM): the smallest split <= 8
    // whose last wave is at least 95 % full, else the fullest one  (n = 4000: 528 tile pairs -> 3 slabs, 10.7 / 11 waves)
    int64_t ksplit = 1;
    {
        double best = 0.0;
        for (int64_t sp = 1; sp <= 8; ++sp) {
            const double waves = (double)(pairs * sp) / ctx->num_sms;
            const double eff = waves / (double)cdiv64(pairs * sp, ctx->num_sms);
            if (eff > best + 1e-9) { best = eff; ksplit = sp; }
            if (eff >= 0.95) break;
        }
    }
    p->part_cap = ksplit > 1 ? ksplit : 0;
    if (p->part_cap) {
        size_t pb = (size_t)p->part_cap * p->ldc * n * sizeof(double);
        e = cudaMalloc(&p->part, pb);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return lso_set_error(ctx, LSO_ERR_ALLOC, "syrk split-K workspace: cudaMalloc(%zu): %s", pb, cudaGetErrorString(e));
        }
        LSO_CHECK_CUDA(ctx, cudaMemsetAsync(p->part, 0, pb, ctx->stream));
    }
    LSO_CHECK_CUDA(ctx, cudaMalloc(&p->d_info, sizeof(int)));
    {   // packed all-reduce buffer [upper(J'J) by columns | J'y]: n(n+1)/2 + n doubles, x2 for the
        // emulated-shards test hook's running sum
        const size_t pk = (size_t)(n * (n + 1) / 2 + n);
        e = cudaMalloc(&p->packed, 2 * pk * sizeof(double));
        if (e != cudaSuccess) {
            cudaGetLastError();
            return lso_set_error(ctx, LSO_ERR_ALLOC, "Cholesky all-reduce buffer: cudaMalloc(%zu): %s", 2 * pk * sizeof(double), cudaGetErrorString(e));
        }
        p->packed_len = (int64_t)pk;
    }
    static bool attr_done_dev[LSO_MAX_DEVICES] = {};      // function attributes are per device
    bool& attr_done = attr_done_dev[ctx->device % LSO_MAX_DEVICES];
    if (!attr_done) {
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(syrk_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SY_SMEM_BYTES));
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(syrk_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SY_SMEM_BYTES));
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute((syrk_mma_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, SY_SMEM_BYTES));
        attr_done = true;
    }
    return LSO_OK;
}

void chol_plan_destroy(CholPlan* p) {
    if (!p) return;
    cudaFree(p->C);
    cudaFree(p->part);
    cudaFree(p->d_info);
    cudaFree(p->packed);
    oz_plan_destroy(p->oz);
    *p = CholPlan();
}

int syrk_upper(lso_ctx* ctx, CholPlan* p, int64_t m, int64_t n, const double* d_J, int64_t ld) {
    const int64_t slab = p->ldc * n;
    ctx->stat_syrk_flops += (double)m * (double)n * (double)(n + 1);
    int64_t ksplit = p->part_cap > 1 ? p->part_cap : 1;
    // keep every split at least a few chunks long
    while (ksplit > 1 && cdiv64(m, ksplit) < 4 * SY_KC) --ksplit;
    int64_t rows_per_split = roundup64(cdiv64(m, ksplit), SY_KC);
    ksplit = cdiv64(m, rows_per_split);
    if (ksplit < 1) ksplit = 1;
    // tcgen05 int8 digit products (ozaki.cu): always with "syrk" = 2; with the default "syrk" = 3 for shapes where the
    // 128 x 128 tiles and the digit split pay off, falling back to the DMMA kernel if the digit matrices (as large as J
    // itself) cannot be allocated
    if (ctx->opt_syrk == 2 || (ctx->opt_syrk == 3 && m >= 8192 && n >= 512)) {
        const int st = oz_syrk_upper(ctx, &p->oz, (int)ctx->opt_ozaki_slices, m, n, d_J, ld, p->C, p->ldc, p->part, p->part_cap);
        if (st == LSO_OK) { ctx->stat_syrk_i8_macs += (double)ctx->opt_ozaki_slices * (double)(ctx->opt_ozaki_slices + 1) / 2.0 * (double)roundup64(m, 128) * 128.0 * 128.0 * (double)(cdiv64(n, 128) * (cdiv64(n, 128) + 1) / 2); return st; }
        if (ctx->opt_syrk == 2 || st != LSO_ERR_ALLOC) return st;
    }
    double* out = (ksplit > 1) ? p->part : p->C;
    if (ctx->opt_syrk) {
        int64_t ntb = cdiv64(n, SY_TS);
        dim3 grid((unsigned)(ntb * (ntb + 1) / 2), (unsigned)ksplit);
        const bool al = (((uintptr_t)d_J & 15) == 0) && (ld % 2 == 0);
        if (al) syrk_mma_kernel<true><<<grid, 256, SY_SMEM_BYTES, ctx->stream>>>(m, n, d_J, ld, rows_per_split, out, p->ldc, slab);
        else syrk_mma_kernel<false><<<grid, 256, SY_SMEM_BYTES, ctx->stream>>>(m, n, d_J, ld, rows_per_split, out, p->ldc, slab);
    } else {
        int64_t ntb = cdiv64(n, 64);
        dim3 grid((unsigned)(ntb * (ntb + 1) / 2), (unsigned)ksplit);
        syrk_fma_kernel<<<grid, 256, 0, ctx->stream>>>(m, n, d_J, ld, rows_per_split, out, p->ldc, slab);
    }
    LSO_CHECK_LAUNCH(ctx);
    if (ksplit > 1) {
        int64_t g = cdiv64(slab, 256);
        if (g > (int64_t)ctx->num_sms * 8) g = (int64_t)ctx->num_sms * 8;
        slab_reduce_kernel<<<(unsigned)g, 256, 0, ctx->stream>>>(slab, (int)ksplit, p->part, slab, p->C);
        LSO_CHECK_LAUNCH(ctx);
    }
    return LSO_OK;
}

// rank-w update of the row strip [t0, i_end) x [t0, n) of the trailing matrix only (upper triangle), t0 = k0 + w: inside a
// 128-column super-block only the rows that the NEXT sub-panels factor need the update right away
__global__ void __launch_bounds__(256)
potrf_strip_update_kernel(int n, double* __restrict__ C, long long ldc, int k0, int w, int i_end) {
    __shared__ double Ri[32][65], Rj[32][65];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int t0 = k0 + w;
    const int i0 = t0 + blockIdx.y * 64, j0 = t0 + blockIdx.x * 64;
    if (j0 + 63 < i0) return;                      // tile entirely below the diagonal
    for (int e = tid; e < 32 * 64; e += 256) {
        const int k = e & 31, c = e >> 5;
        Ri[k][c] = (k < w && i0 + c < i_end) ? C[(long long)(i0 + c) * ldc + k0 + k] : 0.0;
        Rj[k][c] = (k < w && j0 + c < n) ? C[(long long)(j0 + c) * ldc + k0 + k] : 0.0;
    }
    __syncthreads();
    double acc[4][4] = {};
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = Ri[k][4 * ty + u]; b[u] = Rj[k][4 * tx + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int i = i0 + 4 * ty + u, j = j0 + 4 * tx + v;
            if (i < i_end && j < n && i <= j) C[(long long)j * ldc + i] -= acc[u][v];
        }
}

int potrf_upper(lso_ctx* ctx, CholPlan* p, int* info_out) {
    const int n = (int)p->n;
    set_int_kernel<<<1, 1, 0, ctx->stream>>>(p->d_info, INT_MAX);
    LSO_CHECK_LAUNCH(ctx);
    // Two-level blocking: 128-column super-blocks of four 32-column panels.  Inside a super-block a panel's rank-32 update
    // only goes to the row strip the next panels of the block factor; the rest of the trailing matrix gets ONE rank-128
    // update per super-block on the fp64 tensor pipe (the DMMA syrk kernel in subtract mode: its operand is the
    // 128 x rest row block just computed, K-major like the rows of J) — a quarter of the trailing-matrix traffic.
    const int SB = 128;
    for (int b0 = 0; b0 < n; b0 += SB) {
        const int bend = (b0 + SB < n) ? b0 + SB : n;
        const bool big = ctx->opt_syrk && (bend - b0 == SB) && (n - bend >= 256);
        for (int k0 = b0; k0 < bend; k0 += 32) {
            const int w = (n - k0 < 32) ? (n - k0) : 32;
            const int rest = n - k0 - w;
            int g = rest > 0 ? (rest + PP_THREADS - 1) / PP_THREADS : 1;
            potrf_panel_kernel<<<g, PP_THREADS, 0, ctx->stream>>>(n, p->C, p->ldc, k0, p->d_info);
            LSO_CHECK_LAUNCH(ctx);
            if (rest <= 0) continue;
            if (big) {
                const int t0 = k0 + w;
                if (t0 < bend) {
                    dim3 grid((unsigned)((n - t0 + 63) / 64), (unsigned)((bend - t0 + 63) / 64));
                    potrf_strip_update_kernel<<<grid, 256, 0, ctx->stream>>>(n, p->C, p->ldc, k0, w, bend);
                    LSO_CHECK_LAUNCH(ctx);
                }
            } else {
                int nt = (rest + 63) / 64;
                potrf_update_kernel<<<nt * (nt + 1) / 2, 256, 0, ctx->stream>>>(n, p->C, p->ldc, k0, w);
                LSO_CHECK_LAUNCH(ctx);
            }
        }
        if (big) {
            const int rest = n - bend;
            const long long ntb = (rest + SY_TS - 1) / SY_TS;
            dim3 grid((unsigned)(ntb * (ntb + 1) / 2), 1);
            const double* P = p->C + (long long)bend * p->ldc + b0;          // SB x rest, ld = ldc
            double* T = p->C + (long long)bend * p->ldc + bend;
            syrk_mma_kernel<true, true><<<grid, 256, SY_SMEM_BYTES, ctx->stream>>>(SB, rest, P, p->ldc, SB, T, p->ldc, 0);
            LSO_CHECK_LAUNCH(ctx);
        }
    }
    int info = 0;
    LSO_CHECK_CUDA(ctx, cudaMemcpyAsync(&info, p->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *info_out = (info == INT_MAX) ? 0 : info;
    return LSO_OK;
}

extern "C" int lso_comm_allreduce_sum(lso_ctx* ctx, double* d_buf, int64_t count);
extern "C" int lso_dense_gemv_t(lso_ctx* ctx, int64_t m, int64_t n, double alpha, const double* d_J, int64_t ld,
                                const double* d_y, double beta, double* d_x);

// [upper(C) by columns | rhs]  <->  the ldc x n workspace.  MODE 0: pack, 1: unpack, 2: packed accumulate (dst += src)
template <int MODE>
__global__ void chol_pack_kernel(long long n, double* __restrict__ C, long long ldc, double* __restrict__ rhs,
                                 double* __restrict__ packed, const double* __restrict__ src) {
    const long long j = blockIdx.y;                       // column 0..n-1, n = the right-hand side
    if (MODE == 2) {
        const long long len = n * (n + 1) / 2 + n;
        for (long long i = (blockIdx.y * (long long)gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x; i < len;
             i += (long long)gridDim.y * gridDim.x * blockDim.x)
            packed[i] += src[i];
        return;
    }
    if (j < n) {
        double* __restrict__ pk = packed + j * (j + 1) / 2;
        double* __restrict__ col = C + j * ldc;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i <= j; i += (long long)gridDim.x * blockDim.x) {
            if (MODE == 0) pk[i] = col[i]; else col[i] = pk[i];
        }
    } else {
        double* __restrict__ pk = packed + n * (n + 1) / 2;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            if (MODE == 0) pk[i] = rhs[i]; else rhs[i] = pk[i];
        }
    }
}

// local part: C (upper tiles) = J'J, rhs = J'y for this rank's rows
static int chol_local(lso_ctx* ctx, CholPlan* p, int64_t m, int64_t n, const double* d_J, int64_t ld, const double* d_y) {
    lso_prof_mark(ctx);
    LSO_TRY(syrk_upper(ctx, p, m, n, d_J, ld));                               // mul!(cholm, J', J)
    lso_prof_mark(ctx);
    return lso_dense_gemv_t(ctx, m, n, 1.0, d_J, ld, d_y, 0.0, p->rhs);       // mul!(x, J', y)
}
static int chol_pack(lso_ctx* ctx, CholPlan* p, int mode, double* packed, const double* src) {
    const int64_t n = p->n;
    dim3 grid((unsigned)std::min<int64_t>(cdiv64(n, 256), 16), (unsigned)(n + 1));
    if (mode == 0) chol_pack_kernel<0><<<grid, 256, 0, ctx->stream>>>(n, p->C, p->ldc, p->rhs, packed, nullptr);
    else if (mode == 1) chol_pack_kernel<1><<<grid, 256, 0, ctx->stream>>>(n, p->C, p->ldc, p->rhs, packed, nullptr);
    else chol_pack_kernel<2><<<grid, 256, 0, ctx->stream>>>(n, p->C, p->ldc, p->rhs, packed, src);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}
// cholm[i,i] += damp[i]; factor; two triangular solves
static int chol_finish(lso_ctx* ctx, CholPlan* p, const double* d_damp, double* d_x) {
    const int64_t n = p->n;
    if (d_damp) {                                                             // dense_cholesky.jl:51-53
        add_diag_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, ctx->stream>>>(n, p->C, p->ldc, d_damp);
        LSO_CHECK_LAUNCH(ctx);
    }
    int info = 0;
    LSO_TRY(potrf_upper(ctx, p, &info));
    if (info > 0) {
        lso_set_error(ctx, info, d_damp ? "PosDefException: matrix is not positive definite; Cholesky factorization failed (info=%d)"
                                        : "RankDeficientException(%d)", info);
        return info;
    }
    LSO_TRY(tri_solve(ctx, n, p->C, p->ldc, p->rhs, d_x, 1));   // R' z = J'y
    LSO_TRY(tri_solve(ctx, n, p->C, p->ldc, d_x, d_x, 0));      // R x = z
    return LSO_OK;
}

// sharded != 0: J, y are this rank's row shard; ONE all-reduce of the packed [upper(J'J) | J'y].
// The packed [upper(J'J) | J'y] (before damping) stays in p->packed: chol_solve_kept re-solves from it.
int chol_solve(lso_ctx* ctx, CholPlan* p, int64_t m, int64_t n, const double* d_J, int64_t ld, const double* d_y,
               const double* d_damp, double* d_x, int sharded) {
    p->kept = false;
    LSO_TRY(chol_local(ctx, p, m, n, d_J, ld, d_y));
    LSO_TRY(chol_pack(ctx, p, 0, p->packed, nullptr));
    if (sharded && ctx->nranks > 1) {
        lso_prof_mark2(ctx);
        LSO_TRY(lso_comm_allreduce_sum(ctx, p->packed, p->packed_len));
        lso_prof_mark2(ctx);
        LSO_TRY(chol_pack(ctx, p, 1, p->packed, nullptr));
    }
    p->kept = true;
    return chol_finish(ctx, p, d_damp, d_x);
}

// (f3) re-solve of a rejected trust-region step (levenberg_marquardt.jl:77-87: same J and f, new damping): J'J and J'f
// of the last chol_solve are unpacked again; no pass over J, no collective.
int chol_solve_kept(lso_ctx* ctx, CholPlan* p, const double* d_damp, double* d_x) {
    LSO_REQUIRE(ctx, p->kept, "no J'J kept: call lso_chol_solve first");
    LSO_TRY(chol_pack(ctx, p, 1, p->packed, nullptr));
    return chol_finish(ctx, p, d_damp, d_x);
}

// Test hook: the sharded algorithm with the P row shards emulated on ONE device: the P partial [upper(J'J) | J'y] are
// formed one after the other, packed exactly as for the all-reduce, and summed in rank order.
int chol_solve_emulated(lso_ctx* ctx, CholPlan* p, int P, int64_t ms, int64_t n, const double* d_J, int64_t ld,
                        const double* d_y, const double* d_damp, double* d_x) {
    double* sum = p->packed + p->packed_len;
    p->kept = false;
    for (int k = 0; k < P; ++k) {
        LSO_TRY(chol_local(ctx, p, ms, n, d_J + (size_t)k * ms, ld, d_y + (size_t)k * ms));
        LSO_TRY(chol_pack(ctx, p, 0, k == 0 ? sum : p->packed, nullptr));
        if (k > 0) LSO_TRY(chol_pack(ctx, p, 2, sum, p->packed));
    }
    LSO_TRY(chol_pack(ctx, p, 1, sum, nullptr));
    return chol_finish(ctx, p, d_damp, d_x);
}

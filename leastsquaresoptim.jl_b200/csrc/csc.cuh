// csc.cuh — device image of a SparseMatrixCSC{Float64,Int64}: CSC arrays (int32, 0-based) for the
// adjoint product / column norms and a CSR mirror (pattern built once per sparsity pattern, values
// refreshed by a gather after every g!, or written directly by a device g!) for the forward product, so
// both products are deterministic gathers with no floating-point atomics.
//
// Both products run through ONE kernel template, `spmv_stream_kernel` below ("stream" SpMV): a CTA owns a run of
// consecutive segments (rows of the CSR mirror / columns of the CSC image) whose entries are one contiguous slice
// of at most SP_CHUNK stored entries.  The slice is streamed with fully coalesced 128-bit value loads and 64-bit
// index loads (every lane reads 2 consecutive entries per load instruction; the warp reads 512 + 256 contiguous
// bytes), the products val * gather(idx) are parked in shared memory, and G lanes per segment reduce them in a
// fixed order (lane-strided partial sums, then a shuffle tree): no atomics, bit-reproducible.  The functor F
// supplies the gather (what is read at the entry's index), the epilogue (what is done with a segment's sum; it
// returns the segment's contribution to a grid-wide sum of squares) and the finisher, which the LAST CTA to retire
// runs on the deterministic total — this is where LSMR's norms and scalar recurrences live (lsmr.cu).
#pragma once
#include "common.cuh"

struct lso_csc {
    lso_ctx* ctx = nullptr;
    int64_t m = 0, n = 0, nnz = 0;
    int64_t cap_nnz = 0, cap_m = 0, cap_n = 0;   // allocated capacities (lso_csc_update_pattern re-uses the buffers)
    int* d_colptr = nullptr;   // n+1
    int* d_rowidx = nullptr;   // nnz (+ padding)
    double* d_val = nullptr;   // nnz (CSC order)
    int* d_rowptr = nullptr;   // m+1
    int* d_colidx = nullptr;   // nnz (CSR order)
    int* d_perm = nullptr;     // nnz: CSR position -> CSC position
    double* d_valr = nullptr;  // nnz (CSR order)
    bool csr_dirty = true;
    // stream-SpMV partitions: segment ranges per CTA
    int* d_rblk = nullptr;     // nrblk+1 first row of each CTA (CSR mirror)
    int* d_cblk = nullptr;     // ncblk+1 first column of each CTA (CSC)
    int nrblk = 0, ncblk = 0;
    int Gr = 8, Gc = 32;       // lanes per row / per column in the segmented reduction
    // colsumabs2 cache (S-d: computed once per J, shared by LM:82 and the LSMR preconditioner iterative_lsmr.jl:131)
    double* d_colsq = nullptr;
    bool colsq_valid = false;
};

int csc_refresh_csr(lso_csc* A);

#define SP_THREADS 256
#define SP_CHUNK 2048            /* stored entries per CTA slice: 8 per thread */
#define SP_SEGMAX 2048           /* segments per CTA (bounds the epilogue loop when there are empty segments) */
#define SP_WARP_CTAS 6           /* resident CTAs per SM of the warp kernel = its persistent grid per SM */
#define SP_COUNTER_SLOT 1        /* ctx->d_counters slot of the retirement ticket */

#ifdef __CUDACC__
// F interface (the kernel works on a thread-local copy of the functor, so begin() may cache device scalars in members):
//   bool   begin()                               false => the whole launch is a no-op (device-side loop guard)
//   bool   idle() const                          true => skip the vector work but still run finish() (uniform per launch)
//   double gather(int idx) const                 value multiplied with the stored entry whose index is idx
//   double epilogue(long long seg, double sum)   consume a segment's sum; returns its contribution to the grid sum
//   static constexpr bool DUAL                   also accumulate sum(val^2) per segment -> epilogue2(seg, sum, sumsq)
//   long long n_extra; double extra(long long i) elementwise tail handled by CTAs >= nblk (contribution returned)
//   void finish(double sum_main, double sum_extra)   run by ONE thread after every CTA has retired
template <int G, class F>
__global__ void __launch_bounds__(SP_THREADS)
spmv_stream_kernel(const F f_in, const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                   const int* __restrict__ blk, int nblk, double* __restrict__ partials, unsigned int* __restrict__ counter) {
    __shared__ __align__(16) double prod[SP_CHUNK];
    __shared__ __align__(16) double prod2[F::DUAL ? SP_CHUNK : 2];
    __shared__ double red[32];
    __shared__ bool is_last;
    F f = f_in;
    if (!f.begin()) return;
    const int tid = threadIdx.x;
    double contrib = 0.0;
    if (f.idle()) {
        // nothing to do for the vectors (LSMR: beta == 0 keeps v and alpha), the finisher still runs
    } else if ((int)blockIdx.x < nblk) {
        const int s0 = blk[blockIdx.x], s1 = blk[blockIdx.x + 1];
        const int k0 = ptr[s0], k1 = ptr[s1];
        const int ka = k0 & ~1;
        if (k1 - ka <= SP_CHUNK) {
            // ---- stream the slice: 4 x (128-bit values, 64-bit indices) per thread, all loads issued before the gathers ----
            double2 v[4];
            int2 c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = ka + 2 * (tid + SP_THREADS * u);
                if (e < k1) {
                    v[u] = __ldg(reinterpret_cast<const double2*>(val + e));
                    c[u] = __ldg(reinterpret_cast<const int2*>(idx + e));
                } else {
                    v[u] = make_double2(0.0, 0.0);
                    c[u] = make_int2(0, 0);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = ka + 2 * (tid + SP_THREADS * u);
                const bool ok0 = e >= k0 && e < k1, ok1 = e + 1 < k1;      // e + 1 >= k0 always holds (ka >= k0 - 1)
                const double g0 = ok0 ? f.gather(c[u].x) : 0.0, g1 = ok1 ? f.gather(c[u].y) : 0.0;
                if (e < ka + SP_CHUNK) {
                    *reinterpret_cast<double2*>(prod + (e - ka)) = make_double2(ok0 ? v[u].x * g0 : 0.0, ok1 ? v[u].y * g1 : 0.0);
                    if (F::DUAL) *reinterpret_cast<double2*>(prod2 + (e - ka)) = make_double2(ok0 ? v[u].x * v[u].x : 0.0, ok1 ? v[u].y * v[u].y : 0.0);
                }
            }
            __syncthreads();
            // ---- segmented reduction: G lanes per segment, fixed order ----
            const int sub = tid % G;
            for (int s = s0 + tid / G; s - (tid / G) < s1; s += SP_THREADS / G) {     // uniform trip count within a group
                double a = 0.0, a2 = 0.0;
                const bool live = s < s1;
                if (live) {
                    const int b0 = ptr[s] - ka, b1 = ptr[s + 1] - ka;
                    for (int k = b0 + sub; k < b1; k += G) {
                        a += prod[k];
                        if (F::DUAL) a2 += prod2[k];
                    }
                }
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    if (F::DUAL) a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                }
                if (live && sub == 0) contrib += F::DUAL ? f.epilogue2(s, a, a2) : f.epilogue(s, a);
            }
        } else {
            // ---- one long segment (s1 == s0 + 1 by construction of the partition): the CTA walks it chunk by chunk ----
            double a = 0.0, a2 = 0.0;
            for (int e = ka + 2 * tid; e < k1; e += 2 * SP_THREADS) {
                const double2 vv = __ldg(reinterpret_cast<const double2*>(val + e));
                const int2 cc = __ldg(reinterpret_cast<const int2*>(idx + e));
                if (e >= k0) { a = fma(vv.x, f.gather(cc.x), a); if (F::DUAL) a2 = fma(vv.x, vv.x, a2); }
                if (e + 1 < k1) { a = fma(vv.y, f.gather(cc.y), a); if (F::DUAL) a2 = fma(vv.y, vv.y, a2); }
            }
            a = block_sum(a, red);
            if (F::DUAL) a2 = block_sum(a2, red);
            if (tid == 0) contrib = F::DUAL ? f.epilogue2(s0, a, a2) : f.epilogue(s0, a);
        }
    } else {
        // elementwise tail (e.g. the damping rows of LSMR's augmented operator)
        const long long i = (long long)(blockIdx.x - nblk) * SP_THREADS + tid;
        if (i < f.n_extra) contrib = f.extra(i);
    }
    // ---- deterministic grid-wide sum: per-CTA partial, the last CTA to retire adds them in index order ----
    contrib = block_sum(contrib, red);
    if (tid == 0) {
        partials[blockIdx.x] = contrib;
        __threadfence();
        const unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double sm = 0.0, sx = 0.0;
        for (int i = tid; i < nblk; i += SP_THREADS) sm += ((volatile double*)partials)[i];
        for (int i = nblk + tid; i < (int)gridDim.x; i += SP_THREADS) sx += ((volatile double*)partials)[i];
        sm = block_sum(sm, red);
        sx = block_sum(sx, red);
        if (tid == 0) {
            *counter = 0;
            f.finish(sm, sx);
        }
    }
}

// "warp" SpMV (ctx option "spmv" = 2, the default): G lanes per segment, everything in registers and shuffles, no shared
// memory staging and no block barrier on the data path, persistent grid.  A random-pattern product is bound by the
// gathers, not by the stream: every gathered double costs one L1TEX wavefront and one 32-byte L2 sector (measured:
// profiles/r2_ncu_spmv.txt), so the kernel is built to keep as many gathers in flight as possible — each lane loads
// entries in aligned PAIRS (one 128-bit value load + one 64-bit index load), two pairs per loop trip, i.e. up to four
// independent gathers per lane before the first use — and to spend as few wavefronts as possible on anything else.
// Same functor interface as the stream kernel; sums are formed in a fixed order (lane-sequential FMA chain, shuffle
// tree over the G lanes, per-CTA partial in segment order, last CTA adds the partials in index order): bit-reproducible.
template <int G, class F>
__global__ void __launch_bounds__(SP_THREADS, SP_WARP_CTAS)
spmv_warp_kernel(const F f_in, const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                 long long nseg, double* __restrict__ partials, unsigned int* __restrict__ counter) {
    __shared__ double red[32];
    __shared__ bool is_last;
    F f = f_in;
    if (!f.begin()) return;
    const int tid = threadIdx.x;
    double contrib = 0.0, contrib_x = 0.0;
    if (!f.idle()) {
        constexpr int SPB = SP_THREADS / G;                       // segments per CTA per pass
        const int sub = tid % G;
        for (long long base = (long long)blockIdx.x * SPB; base < nseg; base += (long long)gridDim.x * SPB) {
            const long long s = base + tid / G;
            const bool live = s < nseg;
            double a = 0.0, a2 = 0.0;
            if (live) {
                const int k0 = ptr[s], k1 = ptr[s + 1];
                const int ka = k0 & ~1;                          // pairs start on a 16-byte boundary
                for (int e = ka + 2 * sub; e < k1; e += 4 * G) {
                    const int e2 = e + 2 * G;
                    const bool h2 = e2 < k1;
                    const double2 v0 = __ldg(reinterpret_cast<const double2*>(val + e));
                    const int2 c0 = __ldg(reinterpret_cast<const int2*>(idx + e));
                    double2 v1 = make_double2(0.0, 0.0);
                    int2 c1 = make_int2(0, 0);
                    if (h2) {
                        v1 = __ldg(reinterpret_cast<const double2*>(val + e2));
                        c1 = __ldg(reinterpret_cast<const int2*>(idx + e2));
                    }
                    const bool o00 = e >= k0, o01 = e + 1 < k1, o10 = h2, o11 = e2 + 1 < k1;    // e2 >= k0 always
                    const double g00 = o00 ? f.gather(c0.x) : 0.0, g01 = o01 ? f.gather(c0.y) : 0.0;
                    const double g10 = o10 ? f.gather(c1.x) : 0.0, g11 = o11 ? f.gather(c1.y) : 0.0;
                    if (o00) { a = fma(v0.x, g00, a); if (F::DUAL) a2 = fma(v0.x, v0.x, a2); }
                    if (o01) { a = fma(v0.y, g01, a); if (F::DUAL) a2 = fma(v0.y, v0.y, a2); }
                    if (o10) { a = fma(v1.x, g10, a); if (F::DUAL) a2 = fma(v1.x, v1.x, a2); }
                    if (o11) { a = fma(v1.y, g11, a); if (F::DUAL) a2 = fma(v1.y, v1.y, a2); }
                }
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                if (F::DUAL) a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            }
            if (live && sub == 0) contrib += F::DUAL ? f.epilogue2(s, a, a2) : f.epilogue(s, a);
        }
        // elementwise tail (the damping rows of LSMR's augmented operator)
        for (long long i = (long long)blockIdx.x * SP_THREADS + tid; i < f.n_extra; i += (long long)gridDim.x * SP_THREADS)
            contrib_x += f.extra(i);
    }
    // ---- deterministic grid-wide sums: per-CTA partials, the last CTA to retire adds them in index order ----
    contrib = block_sum(contrib, red);
    contrib_x = block_sum(contrib_x, red);
    if (tid == 0) {
        partials[blockIdx.x] = contrib;
        partials[gridDim.x + blockIdx.x] = contrib_x;
        __threadfence();
        const unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double sm = 0.0, sx = 0.0;
        for (int i = tid; i < (int)gridDim.x; i += SP_THREADS) {
            sm += ((volatile double*)partials)[i];
            sx += ((volatile double*)partials)[gridDim.x + i];
        }
        sm = block_sum(sm, red);
        sx = block_sum(sx, red);
        if (tid == 0) {
            *counter = 0;
            f.finish(sm, sx);
        }
    }
}

// launch helper.  stream kernel: grid = nblk segment CTAs + the CTAs of the elementwise tail; warp kernel: persistent grid
template <class F>
static inline int spmv_stream_launch(lso_ctx* ctx, int G, const F& f, const int* ptr, const int* idx, const double* val,
                                     const int* blk, int nblk, long long nseg) {
    double* part = ctx->d_partials;
    unsigned int* cnt = ctx->d_counters + SP_COUNTER_SLOT;
    if (ctx->opt_spmv >= 2) {
        if (nseg <= 0 && f.n_extra <= 0) return LSO_OK;
        const long long want = std::max<long long>((nseg * G + SP_THREADS - 1) / SP_THREADS, (f.n_extra + SP_THREADS - 1) / SP_THREADS);
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)ctx->num_sms * SP_WARP_CTAS));
        switch (G) {
            case 1: spmv_warp_kernel<1, F><<<grid, SP_THREADS, 0, ctx->stream>>>(f, ptr, idx, val, nseg, part, cnt); break;
            case 2: spmv_warp_kernel<2, F><<<grid, SP_THREADS, 0, ctx->stream>>>(f, ptr, idx, val, nseg, part, cnt); break;
            case 4: spmv_warp_kernel<4, F><<<grid, SP_THREADS, 0, ctx->stream>>>(f, ptr, idx, val, nseg, part, cnt); break;
            case 8: spmv_warp_kernel<8, F><<<grid, SP_THREADS, 0, ctx->stream>>>(f, ptr, idx, val, nseg, part, cnt); break;
            case 16: spmv_warp_kernel<16, F><<<grid, SP_THREADS, 0, ctx->stream>>>(f, ptr, idx, val, nseg, part, cnt); break;
            default: spmv_warp_kernel<32, F><<<grid, SP_THREADS, 0, ctx->stream>>>(f, ptr, idx, val, nseg, part, cnt); break;
        }
        LSO_CHECK_LAUNCH(ctx);
        return LSO_OK;
    }
    const long long extra = (f.n_extra + SP_THREADS - 1) / SP_THREADS;
    const unsigned grid = (unsigned)(nblk + extra);
    if (grid == 0) return LSO_OK;
    if ((size_t)grid > (size_t)LSO_PARTIALS) return lso_set_error(ctx, LSO_ERR_UNSUPPORTED, "sparse product: too many CTAs for the partials buffer");
    switch (G) {
        case 1: spmv_stream_kernel<1, F><<<grid, SP_THREADS, 0, ctx->stream>>>(f, ptr, idx, val, blk, nblk, part, cnt); break;
        case 2: spmv_stream_kernel<2, F><<<grid, SP_THREADS, 0, ctx->stream>>>(f, ptr, idx, val, blk, nblk, part, cnt); break;
        case 4: spmv_stream_kernel<4, F><<<grid, SP_THREADS, 0, ctx->stream>>>(f, ptr, idx, val, blk, nblk, part, cnt); break;
        case 8: spmv_stream_kernel<8, F><<<grid, SP_THREADS, 0, ctx->stream>>>(f, ptr, idx, val, blk, nblk, part, cnt); break;
        case 16: spmv_stream_kernel<16, F><<<grid, SP_THREADS, 0, ctx->stream>>>(f, ptr, idx, val, blk, nblk, part, cnt); break;
        default: spmv_stream_kernel<32, F><<<grid, SP_THREADS, 0, ctx->stream>>>(f, ptr, idx, val, blk, nblk, part, cnt); break;
    }
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}
#endif

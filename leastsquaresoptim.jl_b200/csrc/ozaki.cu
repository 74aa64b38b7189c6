// ozaki.cu — J'J on the 5th-generation tensor cores: tcgen05.mma.kind::i8 with TMEM accumulators, operands staged by
// the TMA unit (cp.async.bulk.tensor), fp64 results reconstructed exactly from integer partial products (the Ozaki
// scheme).  Replaces `mul!(cholm, J', J)` (src/solver/dense_cholesky.jl:31,48) when ctx option "syrk" = 2.
//
// tcgen05.mma has no f64 kind; the DMMA pipe (mma.sync m8n8k4) tops out at ~37 TFLOP/s.  The int8 pipe is two orders of
// magnitude faster, and integer accumulation is exact, so an fp64 product can be split into exact int8 pieces:
//   1. every column j of J is scaled by 2^-e_j (|J[k,j]| 2^-e_j <= 1/2) and cut into S signed digits of 7 bits,
//         J[k,j] = 2^e_j * ( sum_{p=1..S} D_p[k,j] 2^(-7p) + r ),   D_p in [-64, 64],  |r| <= 2^(-7S-1)
//      (round-to-nearest digit extraction; every step is exact in fp64).  D_p are int8 matrices, K-major.
//   2. (J'J)[i,j] = 2^(e_i+e_j) * sum_{d=2..S+1} 2^(-7d) * sum_{p+q=d} (D_p' D_q)[i,j]    (terms with p+q > S+1 are below
//      the truncation error of step 1 and are dropped).  Each inner sum is an int8 GEMM accumulated EXACTLY in int32
//      (|D_p D_q| <= 2^12, at most 8 pairs per d, K chunks of <= 32768 rows: < 2^31).
//   3. the S integer accumulators are converted to fp64, scaled by powers of two (exact) and added, smallest terms first.
// With S = 8 the representation error is 2^-57 and the dropped pairs contribute (S-1) 2^(-7S-2) = 2^-55, both relative to
// 2^(e_i+e_j): below fp64 rounding for columns whose entries are of comparable size, K 2^(5-7S) max|J_i| max|J_j| at worst.
//
// Kernel: one CTA per 128 x 128 tile of the upper triangle of J'J and per K chunk.  Warp 0 = TMA producer (3-stage ring,
// one 32-row K step of all needed digit matrices per stage, 32-byte-swizzled K-major boxes), warp 1 = MMA issuer (one
// elected thread, UTCIMMA 128x128x32), warps 2-9 = epilogue (tcgen05.ld, fp64 reconstruction in registers).  TMEM holds
// four 128 x 128 int32 accumulators (512 columns), one per value of d, so the S values of d are covered in ceil(S/4)
// passes over the K chunk, the pass with the smallest terms first.
#include "chol.cuh"
#include <cuda.h>
#include <limits.h>
#include <math.h>

#define OZ_T 128                     /* tile edge */
#define OZ_BK 32                     /* K bytes per stage = K of one UTCIMMA */
#define OZ_MAXS 8
#define OZ_STAGES 3
#define OZ_SLICE_BYTES (OZ_T * OZ_BK)                       /* 4096 */
#define OZ_STAGE_BYTES (2 * OZ_MAXS * OZ_SLICE_BYTES)       /* 65536 */
#define OZ_SMEM_BYTES (OZ_STAGES * OZ_STAGE_BYTES + 1024 + 256)
#define OZ_THREADS 320
#define OZ_KCHUNK_MAX 32768          /* rows accumulated in int32 before the accumulators are drained: 8 pairs * 32768 * 2^12 = 2^30 < 2^31 */
#define OZ_CHUNK_STEPS (OZ_KCHUNK_MAX / OZ_BK)
#define OZ_W 7                       /* bits per digit */

struct __align__(64) OzMaps { CUtensorMap m[OZ_MAXS]; };

struct OzPlan {
    int64_t m_cap = 0, n = 0, kpad = 0;
    int S = 0;
    signed char* slices = nullptr;    // S matrices kpad x n, column-major (K-major), ld = kpad
    int* expo = nullptr;              // n column exponents e_j
    OzMaps maps;
    bool maps_valid = false;
};

// ---- PTX helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t oz_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void oz_bar_init(uint32_t bar, uint32_t cnt) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt) : "memory");
}
__device__ __forceinline__ void oz_bar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void oz_bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void oz_tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void oz_mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void oz_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void oz_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void oz_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void oz_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void oz_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor of one K-major 128 x 32-byte digit tile written by a SWIZZLE_32B TMA box:
// 8-row atoms of 256 bytes, stride between atoms (SBO) 256, one atom along K (LBO unused), descriptor version 1
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// instruction descriptor: D = s32, A = B = signed int8, both K-major, M = 128, N = 128, dense, no saturation
#define OZ_IDESC ((2u << 4) | (1u << 7) | (1u << 10) | ((OZ_T >> 3) << 17) | ((OZ_T >> 4) << 24))

// =====================================================================================================================
// split: column exponents, then the digit matrices
// =====================================================================================================================
__global__ void __launch_bounds__(256)
oz_colexp_kernel(long long m, long long n, const double* __restrict__ J, long long ld, int* __restrict__ expo) {
    __shared__ double red[32];
    const long long j = blockIdx.x;
    const double* col = J + j * ld;
    double mx = 0.0;
    bool bad = false;
    for (long long k = threadIdx.x; k < m; k += blockDim.x) {
        const double a = fabs(col[k]);
        if (!(a <= 1.79769313486231570815e308)) bad = true;       // NaN or Inf
        mx = fmax(mx, a);
    }
    mx = block_nanmax(bad ? NAN : mx, red);
    if (threadIdx.x == 0) {
        int e = 0;
        if (mx != mx) e = INT_MIN;                                // poisoned column: the result row / column becomes NaN
        else if (mx > 0.0) e = ilogb(mx) + 2;                     // |x| 2^-e <= 1/2
        expo[j] = e;
    }
}

// each thread converts 16 consecutive rows of one column (one 16-byte write per digit matrix).  The rows come in through
// shared memory: a warp reads its 512 rows of the column with 8 fully coalesced 16-byte loads per lane (4 cache lines per
// instruction; reading 128 contiguous bytes per LANE costs 32 L1 tag look-ups per instruction and made the kernel
// L1-bound at 3.6 TB/s), parks them with one pad double per 16 and every lane picks up its 16 rows.
// Digit extraction without conversion instructions: adding 1.5 * 2^52 rounds x (|x| <= 64) to the nearest integer (ties to
// even, like rint) and leaves that integer in two's complement in the low mantissa bits, so the digit byte is the low
// byte of the sum and the rounded value is recovered by subtracting the constant again — 4 fp64 adds / multiplies per digit.
template <int S>
__global__ void __launch_bounds__(256)
oz_split_kernel(long long m, long long n, long long kpad, const double* __restrict__ J, long long ld,
                const int* __restrict__ expo, signed char* __restrict__ slices) {
    __shared__ double stage[8][512 + 32];
    const long long j = blockIdx.y;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const long long w0 = ((long long)blockIdx.x * 8 + wrp) * 512;       // first row of this warp's 512
    if (w0 >= kpad) return;                                             // whole warp
    const long long k0 = w0 + 16 * lane;
    const int e = expo[j];
    const double* col = J + j * ld;
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
    const bool direct = (e > -1000 && e < 1000);                       // 2^-e is a normal double
    const double scale = (e == INT_MIN) ? 0.0 : (direct ? scalbn(1.0, -e) : 1.0);
    double* st = stage[wrp];
    const bool fast = (w0 + 512 <= m) && ((reinterpret_cast<uintptr_t>(col + w0) & 15) == 0);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int i = 2 * (32 * t + lane);                              // element pair (i, i+1) of the warp's 512
        double2 v;
        if (fast) v = __ldg(reinterpret_cast<const double2*>(col + w0 + i));
        else { v.x = (w0 + i < m) ? col[w0 + i] : 0.0; v.y = (w0 + i + 1 < m) ? col[w0 + i + 1] : 0.0; }
        const int q = i + (i >> 4);
        st[q] = v.x; st[q + 1] = v.y;
    }
    __syncwarp();
    if (k0 >= kpad) return;
    uint32_t w[S][4];
#pragma unroll
    for (int p = 0; p < S; ++p) { w[p][0] = 0u; w[p][1] = 0u; w[p][2] = 0u; w[p][3] = 0u; }
    double xin[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) xin[t] = st[17 * lane + t];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        double f = (direct || e == INT_MIN) ? xin[t] * scale : scalbn(xin[t], -e);     // |f| <= 1/2
#pragma unroll
        for (int p = 0; p < S; ++p) {
            const double x = f * 128.0;                 // exact
            const double r = x + MAGIC;                 // rounds to the nearest integer in [-64, 64]
            const double d = r - MAGIC;                 // that integer as a double (exact)
            w[p][t >> 2] |= ((uint32_t)__double2loint(r) & 0xffu) << (8 * (t & 3));
            f = x - d;                                  // exact, |f| <= 1/2
        }
    }
    const size_t plane = (size_t)kpad * (size_t)n;
#pragma unroll
    for (int p = 0; p < S; ++p)
        *reinterpret_cast<uint4*>(slices + (size_t)p * plane + (size_t)j * kpad + k0) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
}

// =====================================================================================================================
// the tile kernel
// =====================================================================================================================
struct OzPass { int dlo, dhi, nsl; };      // digits sums d = p + q covered by this pass; slices 1..nsl are loaded

template <int S>
__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_syrk_kernel(const __grid_constant__ OzMaps maps, long long n, long long rows_per_split, long long kpad,
               const int* __restrict__ expo, double* __restrict__ out, long long ldc, long long slab) {
    extern __shared__ unsigned char oz_raw[];
    unsigned char* sm = (unsigned char*)(((uintptr_t)oz_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(sm + OZ_STAGES * OZ_STAGE_BYTES);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * OZ_STAGES + 2);
    const uint32_t bar_full = oz_smem(bars), bar_empty = oz_smem(bars + OZ_STAGES);
    const uint32_t bar_tfull = oz_smem(bars + 2 * OZ_STAGES), bar_tempty = oz_smem(bars + 2 * OZ_STAGES + 1);
    const int tid = threadIdx.x, wrp = tid >> 5, lane = tid & 31;

    // tile (bi <= bj) from the linear index over the upper triangle, column-major enumeration
    long long t = blockIdx.x, bj = 0;
    while ((bj + 1) * (bj + 2) / 2 <= t) ++bj;
    const long long bi = t - bj * (bj + 1) / 2;
    const bool diag = (bi == bj);
    const long long k_begin = (long long)blockIdx.y * rows_per_split;
    long long k_end = k_begin + rows_per_split;
    if (k_end > kpad) k_end = kpad;
    const int nk = (int)((k_end - k_begin) / OZ_BK);

    constexpr int NPASS = (S + 3) / 4;
    OzPass pass[NPASS];
    {   // the pass with the largest d (smallest terms) first
        int dhi = S + 1;
        for (int i = 0; i < NPASS; ++i) {
            const int dlo = (dhi - 3 > 2) ? dhi - 3 : 2;
            pass[i].dlo = dlo; pass[i].dhi = dhi;
            pass[i].nsl = (dhi - 1 < S) ? dhi - 1 : S;
            dhi = dlo - 1;
        }
    }

    if (tid == 0) {
        for (int s = 0; s < OZ_STAGES; ++s) { oz_bar_init(bar_full + 8 * s, 1); oz_bar_init(bar_empty + 8 * s, 1); }
        oz_bar_init(bar_tfull, 1);
        oz_bar_init(bar_tempty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wrp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(oz_smem(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    oz_fence_before();
    __syncthreads();
    oz_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (wrp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int it = 0;
            for (int ch = 0; ch < nk; ch += OZ_CHUNK_STEPS) {
                const int ch_end = (ch + OZ_CHUNK_STEPS < nk) ? ch + OZ_CHUNK_STEPS : nk;
                for (int ps = 0; ps < NPASS; ++ps) {
                    const int nsl = pass[ps].nsl;
                    const uint32_t bytes = (uint32_t)(nsl * OZ_SLICE_BYTES * (diag ? 1 : 2));
                    for (int ks = ch; ks < ch_end; ++ks, ++it) {
                        const int stg = it % OZ_STAGES;
                        if (it >= OZ_STAGES) oz_bar_wait(bar_empty + 8 * stg, (uint32_t)(((it / OZ_STAGES) - 1) & 1));
                        const uint32_t sbase = oz_smem(sm + stg * OZ_STAGE_BYTES);
                        oz_bar_expect(bar_full + 8 * stg, bytes);
                        const int kc = (int)(k_begin + (long long)ks * OZ_BK);
                        for (int p = 0; p < nsl; ++p) {
                            oz_tma_2d(sbase + p * OZ_SLICE_BYTES, &maps.m[p], kc, (int)(bi * OZ_T), bar_full + 8 * stg);
                            if (!diag)
                                oz_tma_2d(sbase + (OZ_MAXS + p) * OZ_SLICE_BYTES, &maps.m[p], kc, (int)(bj * OZ_T), bar_full + 8 * stg);
                        }
                    }
                }
            }
        }
    } else if (wrp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            int it = 0, pi = 0;
            for (int ch = 0; ch < nk; ch += OZ_CHUNK_STEPS) {
              const int ch_end = (ch + OZ_CHUNK_STEPS < nk) ? ch + OZ_CHUNK_STEPS : nk;
              for (int ps = 0; ps < NPASS; ++ps, ++pi) {
                if (pi > 0) { oz_bar_wait(bar_tempty, (uint32_t)((pi - 1) & 1)); oz_fence_after(); }
                const int dlo = pass[ps].dlo, dhi = pass[ps].dhi;
                for (int ks = ch; ks < ch_end; ++ks, ++it) {
                    const int stg = it % OZ_STAGES;
                    oz_bar_wait(bar_full + 8 * stg, (uint32_t)((it / OZ_STAGES) & 1));
                    oz_fence_after();
                    const uint32_t sbase = oz_smem(sm + stg * OZ_STAGE_BYTES);
                    for (int d = dlo; d <= dhi; ++d) {
                        const uint32_t acc = tmem + (uint32_t)((d - dlo) * OZ_T);
                        const int plo = (d - S > 1) ? d - S : 1, phi = (d - 1 < S) ? d - 1 : S;
                        for (int p = plo; p <= phi; ++p) {
                            const int q = d - p;
                            const uint64_t da = oz_desc(sbase + (uint32_t)((p - 1) * OZ_SLICE_BYTES));
                            const uint64_t db = oz_desc(sbase + (uint32_t)(((diag ? 0 : OZ_MAXS) + (q - 1)) * OZ_SLICE_BYTES));
                            oz_mma_i8(acc, da, db, OZ_IDESC, (ks > ch || p > plo) ? 1u : 0u);
                        }
                    }
                    oz_commit(bar_empty + 8 * stg);          // frees the stage once these MMAs have read it
                }
                oz_commit(bar_tfull);                        // accumulators of this pass are complete
              }
            }
        }
    } else {
        // =============================== epilogue: TMEM -> fp64 ===============================
        const int we = wrp - 2;                  // 0..7
        const int quarter = wrp & 3;             // TMEM lanes this warp may read: 32 * (warp index % 4)
        const int half = we >> 2;                // columns [64 half, 64 half + 64)
        double acc[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) acc[c] = 0.0;
        int pi = 0;
        for (int ch = 0; ch < nk; ch += OZ_CHUNK_STEPS)
        for (int ps = 0; ps < NPASS; ++ps, ++pi) {
            oz_bar_wait(bar_tfull, (uint32_t)(pi & 1));
            oz_fence_after();
            for (int d = pass[ps].dhi; d >= pass[ps].dlo; --d) {          // smallest terms first
                const double sc = scalbn(1.0, -OZ_W * d);
                const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((d - pass[ps].dlo) * OZ_T + half * 64);
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t r[16];
                    oz_ld16(taddr + c0, r);
                    oz_ld_wait();
#pragma unroll
                    for (int c = 0; c < 16; ++c) acc[c0 + c] = fma((double)(int)r[c], sc, acc[c0 + c]);
                }
            }
            oz_fence_before();
            __syncwarp();
            if (lane == 0) oz_bar_arrive(bar_tempty);
        }
        // scale by 2^(e_i + e_j) and store the tile of this K chunk
        const long long row = bi * OZ_T + quarter * 32 + lane;
        if (row < n) {
            const int ei = expo[row];
            double* dst = out + (size_t)blockIdx.y * (size_t)slab;
#pragma unroll
            for (int c = 0; c < 64; ++c) {
                const long long col = bj * OZ_T + half * 64 + c;
                if (col < n) {
                    const int ej = expo[col];
                    const double v = (ei == INT_MIN || ej == INT_MIN) ? NAN : scalbn(acc[c], ei + ej);
                    dst[row + col * ldc] = v;
                }
            }
        }
    }
    oz_fence_before();
    __syncthreads();
    if (wrp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// =====================================================================================================================
// host side
// =====================================================================================================================
typedef CUresult (*oz_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static oz_encode_fn oz_get_encode() {
    static oz_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (oz_encode_fn)p;
    }
    return fn;
}

void oz_plan_destroy(OzPlan* p) {
    if (!p) return;
    cudaFree(p->slices);
    cudaFree(p->expo);
    delete p;
}

static int oz_plan_ensure(lso_ctx* ctx, OzPlan** pp, int64_t m, int64_t n, int S) {
    OzPlan* p = *pp;
    if (p && (p->n != n || p->S != S || p->m_cap < m)) { oz_plan_destroy(p); p = nullptr; *pp = nullptr; }
    if (p) return LSO_OK;
    p = new (std::nothrow) OzPlan();
    if (!p) return lso_set_error(ctx, LSO_ERR_ALLOC, "host allocation failed");
    p->n = n; p->S = S; p->m_cap = m;
    p->kpad = roundup64(m, 128);
    const size_t bytes = (size_t)S * (size_t)p->kpad * (size_t)n;
    cudaError_t e = cudaMalloc(&p->slices, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->expo, (size_t)n * sizeof(int));
    if (e != cudaSuccess) {
        cudaGetLastError();
        oz_plan_destroy(p);
        return lso_set_error(ctx, LSO_ERR_ALLOC, "Ozaki digit matrices (%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    oz_encode_fn enc = oz_get_encode();
    if (!enc) { oz_plan_destroy(p); return lso_set_error(ctx, LSO_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver"); }
    for (int s = 0; s < S; ++s) {
        cuuint64_t gdim[2] = {(cuuint64_t)p->kpad, (cuuint64_t)n};
        cuuint64_t gstr[1] = {(cuuint64_t)p->kpad};
        cuuint32_t box[2] = {OZ_BK, OZ_T};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&p->maps.m[s], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, p->slices + (size_t)s * (size_t)p->kpad * (size_t)n, gdim,
                         gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { oz_plan_destroy(p); return lso_set_error(ctx, LSO_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r); }
    }
    for (int s = S; s < OZ_MAXS; ++s) p->maps.m[s] = p->maps.m[0];
    *pp = p;
    return LSO_OK;
}

__global__ void oz_slab_reduce_kernel(long long count, int nslab, const double* __restrict__ part, long long slab_stride,
                                      double* __restrict__ C) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        double a = 0.0;
        for (int z = 0; z < nslab; ++z) a += part[z * slab_stride + i];
        C[i] = a;
    }
}

// C (upper 128-tiles, ldc) = J'J for the m x n column-major J, through int8 digit matrices.  `part` / `part_cap` are the
// split-K slabs of the Cholesky plan (each ldc * n doubles); with one chunk the result goes straight into C.
int oz_syrk_upper(lso_ctx* ctx, OzPlan** pp, int S, int64_t m, int64_t n, const double* d_J, int64_t ld, double* C,
                  int64_t ldc, double* part, int64_t part_cap) {
    LSO_REQUIRE(ctx, S >= 2 && S <= OZ_MAXS, "Ozaki syrk: 2 <= slices <= 8");
    LSO_TRY(oz_plan_ensure(ctx, pp, m, n, S));
    OzPlan* p = *pp;
    const int64_t kpad = roundup64(m, 128);             // rows in use (the plan may hold more)
    LSO_REQUIRE(ctx, kpad == p->kpad, "Ozaki syrk: the row count changed; destroy and recreate the workspace");
    oz_colexp_kernel<<<(unsigned)n, 256, 0, ctx->stream>>>(m, n, d_J, ld, p->expo);
    LSO_CHECK_LAUNCH(ctx);
    {
        dim3 grid((unsigned)cdiv64(kpad, 8 * 512), (unsigned)n);
        switch (S) {
            case 2: oz_split_kernel<2><<<grid, 256, 0, ctx->stream>>>(m, n, kpad, d_J, ld, p->expo, p->slices); break;
            case 3: oz_split_kernel<3><<<grid, 256, 0, ctx->stream>>>(m, n, kpad, d_J, ld, p->expo, p->slices); break;
            case 4: oz_split_kernel<4><<<grid, 256, 0, ctx->stream>>>(m, n, kpad, d_J, ld, p->expo, p->slices); break;
            case 5: oz_split_kernel<5><<<grid, 256, 0, ctx->stream>>>(m, n, kpad, d_J, ld, p->expo, p->slices); break;
            case 6: oz_split_kernel<6><<<grid, 256, 0, ctx->stream>>>(m, n, kpad, d_J, ld, p->expo, p->slices); break;
            case 7: oz_split_kernel<7><<<grid, 256, 0, ctx->stream>>>(m, n, kpad, d_J, ld, p->expo, p->slices); break;
            default: oz_split_kernel<8><<<grid, 256, 0, ctx->stream>>>(m, n, kpad, d_J, ld, p->expo, p->slices); break;
        }
        LSO_CHECK_LAUNCH(ctx);
    }
    // split K only to fill the machine (a CTA drains its int32 accumulators into fp64 registers every OZ_KCHUNK_MAX rows,
    // so any row count is safe); same rule as the DMMA syrk: use the plan's slabs
    int64_t ksplit = part_cap > 1 ? part_cap : 1;
    while (ksplit > 1 && cdiv64(kpad, ksplit) < 64 * OZ_BK) --ksplit;
    const int64_t ntb = cdiv64(n, OZ_T), ntiles = ntb * (ntb + 1) / 2;
    const int64_t rows_per_split = roundup64(cdiv64(kpad, ksplit), OZ_BK);
    ksplit = cdiv64(kpad, rows_per_split);
    const int64_t slab = ldc * n;
    double* out = (ksplit > 1) ? part : C;
    static bool attr_done[LSO_MAX_DEVICES][OZ_MAXS + 1] = {};
    bool& done = attr_done[ctx->device % LSO_MAX_DEVICES][S];
    dim3 grid((unsigned)ntiles, (unsigned)ksplit);
#define OZ_LAUNCH(SS)                                                                                                   \
    do {                                                                                                                \
        if (!done) {                                                                                                    \
            LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(oz_syrk_kernel<SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES)); \
            done = true;                                                                                                \
        }                                                                                                               \
        oz_syrk_kernel<SS><<<grid, OZ_THREADS, OZ_SMEM_BYTES, ctx->stream>>>(p->maps, n, rows_per_split, kpad, p->expo, out, ldc, slab); \
    } while (0)
    switch (S) {
        case 2: OZ_LAUNCH(2); break;
        case 3: OZ_LAUNCH(3); break;
        case 4: OZ_LAUNCH(4); break;
        case 5: OZ_LAUNCH(5); break;
        case 6: OZ_LAUNCH(6); break;
        case 7: OZ_LAUNCH(7); break;
        default: OZ_LAUNCH(8); break;
    }
#undef OZ_LAUNCH
    LSO_CHECK_LAUNCH(ctx);
    if (ksplit > 1) {
        int64_t g = cdiv64(slab, 256);
        if (g > (int64_t)ctx->num_sms * 8) g = (int64_t)ctx->num_sms * 8;
        oz_slab_reduce_kernel<<<(unsigned)g, 256, 0, ctx->stream>>>(slab, (int)ksplit, part, slab, C);
        LSO_CHECK_LAUNCH(ctx);
    }
    return LSO_OK;
}

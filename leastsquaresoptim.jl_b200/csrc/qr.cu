// qr.cu — communication-avoiding blocked Householder QR for tall dense systems (Q-a..Q-d in
// SURVEY.md Appendix B).  Replaces `ldiv!(qr!(qrm, ColumnNorm()), u)` at
// src/solver/dense_qr.jl:37,83 for the full-rank case (the damped LM system is full rank by
// construction, levenberg_marquardt.jl:85).
//
// Only R and Q'b are needed for the least-squares solve, so the right-hand side rides along as an
// extra column and V is discarded after each panel:
//   for each panel of QB columns:
//     level 0 : every QH-row block is factorised independently by one CTA (registers)
//     level l : the QB-row heads (R factors) of QG blocks of level l-1 are stacked and factorised
//               the same way (TSQR reduction tree), until one R remains; all levels run in ONE launch,
//               pipelined step by step (a parent starts its step j once its children have finished theirs)
//     then every level's block reflectors  I - V T' V'  are applied to the trailing columns, one
//     pass over the trailing matrix per level (level 0 touches every row once; level l touches
//     1/QG^l of the rows).  No global reduction, no grid-wide sync.
// The trailing update is the flop carrier (4*QB flop per 8-byte element) and runs on the fp64
// tensor pipe (mma.sync m8n8k4 -> DMMA); ALL its global traffic is cp.async.bulk (TMA unit, UBLKCP)
// loads and stores through an mbarrier ring, issued by a dedicated producer warp.  tcgen05.mma has no
// f64 kind, see DESIGN.md.
#include "qr.cuh"
#include <math.h>
#include <stdlib.h>

// Row layout of one tree level ("head first").  The level covers matrix rows [r0, r0 + rows), rows a multiple of
// QB.  It is cut into nb = ceil(rows / QH) blocks; block b owns
//     its HEAD : the QB rows  r0 + QB*b ...                       (receives the block's R factor)
//     its BODY : the QH-QB rows  r0 + QB*nb + (QH-QB)*b ...       (clipped to the level's row range)
// so the heads of all blocks form the contiguous range [r0, r0 + QB*nb), which is exactly the next level's row
// range: every level reads and writes two contiguous runs per column, and the final R lands in rows r0..r0+QB.
struct TileMap {
    long long r0;     // first active matrix row of this panel
    long long nb;     // blocks at this level
    long long rows;   // rows covered by this level (multiple of QB)
};
#define QBODY (QH - QB)

// matrix row of tile row p of block blk; false if the row lies beyond the level's range (reads as zero)
__device__ __forceinline__ bool tm_row(const TileMap& tm, long long blk, int p, long long& mrow) {
    if (p < QB) { mrow = tm.r0 + QB * blk + p; return true; }
    const long long off = QB * tm.nb + QBODY * blk + (p - QB);
    mrow = tm.r0 + off;
    return off < tm.rows;
}
// number of valid body rows of block blk (0 .. QBODY, multiple of QB)
__device__ __forceinline__ int tm_body_rows(const TileMap& tm, long long blk) {
    long long nv = tm.rows - QB * tm.nb - QBODY * blk;
    return (int)(nv < 0 ? 0 : (nv > QBODY ? QBODY : nv));
}

// =================================================================================================
// panel factorisation (qr_tree_kernel_t below): Householder QR of QH x QB blocks held in registers, the whole TSQR
// tree of a panel in one launch.  The T factor of the compact WY form is recovered after the loop from
// T^{-1} = diag(1/tau) + striu(V'V).
// =================================================================================================
__device__ __forceinline__ long long gtimer_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// Reflector scalars of one Householder step (dlarfg with an unnormalised v):
//     nrm = sqrt(alpha^2 + s),  beta = -sign(alpha) nrm,  vjj = alpha - beta,  t = 1 / (nrm (nrm + |alpha|)).
// Fast path: MUFU seeds + two Goldschmidt / Newton steps, no branches on the chain; results are within a couple of
// ulp, which perturbs H = I - t v v' by O(eps) like dlarfg's own rounding.  Arguments outside the safe range (zero
// column, subnormal or huge norm) take the library routines behind one rarely taken branch.
__device__ __forceinline__ void leaf_scalars(double alpha, double s, double& beta, double& tj, double& vjj) {
    const double q = fma(alpha, alpha, s);
    const bool safe = (s > 0.0) && (q > 1e-280) && (q < 1e280);
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(q));
    double g = q * y0, h = 0.5 * y0;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);                                   // g = sqrt(q)
    double nrm = g;
    const double d = fma(fabs(alpha), nrm, q);          // nrm (nrm + |alpha|)
    double z;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(d));
    double e = fma(-d, z, 1.0);
    z = fma(z, e, z);
    e = fma(-d, z, 1.0);
    z = fma(z, e, z);
    e = fma(-d, z, 1.0);
    z = fma(z, e, z);
    tj = z;
    if (!safe) {
        if (s == 0.0) { beta = alpha; tj = 0.0; vjj = 1.0; return; }     // dlarfg: xnorm == 0 -> H = I
        nrm = sqrt(q);
        tj = 1.0 / fma(fabs(alpha), nrm, q);
    }
    beta = -copysign(nrm, alpha);
    vjj = alpha - beta;                                 // = sign(alpha) (|alpha| + nrm)
}

#define AM_VS_BYTES (QB * QS * 8)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}


// ---- panel factorisation: the whole TSQR tree of a panel in ONE launch, levels pipelined step by step ----
// 256 threads per block.  The step time of a Householder factorisation held in registers is set by the length of
// one warp's dependent instruction stream and by the shared-memory pipe (every pivot-column value a thread needs
// costs 8 bytes of LDS return bandwidth), so the layout minimises "pivot-column values per thread" and reads them
// ONCE per step:
//     lane = (column pair cp = lane & 15, half hh = lane >> 4); warp w, half hh  ->  row group rg = 2 w + hh (16 groups)
//     a thread holds 16 rows ("slots") of TWO adjacent columns (xa: column 2 cp, xb: column 2 cp + 1)
//         slots 0, 1    head rows  rg, rg + 16          (the 32 head rows end up holding R)
//         slots 2..15   level 0 : body rows 32 + 14 rg ... + 13  (one contiguous run; always below the diagonal)
//                       level>0 : slot 2 s + h = row rg + 16 h of segment s = 1..7 (child s's R factor)
//     and loads the 16 pivot-column values of its row group into registers once per step (dot product AND update).
//
// Tree levels run CONCURRENTLY.  The block of level l > 0 stacks the R factors of 8 children; row i of a child's R is
// final after the child's step i, and the parent's step j only touches rows <= j of every child triangle (the rest
// of column j is zero).  So a child publishes row j of its R right after its step j, and the parent pulls row j + 3
// of its 8 children at its own step j: the levels of the tree are skewed by a few steps instead of running one
// after the other (QB + 3 (L-1) steps instead of QB L).  Rows travel through an L2-resident mailbox in 16-byte
// units {lo32, tag, hi32, tag} (the NCCL "LL" idea: the data carries its own flag; every 8-byte half is single-copy
// atomic, the tag is unique per launch and row), so neither side executes a fence: the child fires and forgets,
// the parent polls the data itself.  Blocks are numbered children first AND a block's id is the order in which its
// CTA actually started (atomic ticket), so a waiting parent never occupies a slot that one of its children still needs.
#define LEAF_THREADS 256
#define LEAF_NG 16          /* row groups */
#define LEAF_NX 16          /* rows (slots) per thread */
#define LEAF_BR 14          /* body rows per thread at level 0 */
#define LEAF_D 3            /* a parent requests row j + LEAF_D of its children at step j */
#define LEAF_NSTG 4         /* staging buffers for pulled rows */
#define TREE_DSMEM_BYTES (AM_VS_BYTES + 64)   /* block image + mbarrier */

struct TreeParams {
    int nlev;
    int start[QR_MAX_LEVELS + 1];     // first blockIdx of each level
    TileMap tm[QR_MAX_LEVELS];
    double* V[QR_MAX_LEVELS];
    double* T[QR_MAX_LEVELS];
    uint4* mail;                      // mailbox: [block][row][column] {lo32, tag, hi32, tag}
    unsigned base;                    // tag of row r in this launch = base + r + 1
    int* zero_ptr;                    // child counters of the fused trailing update that follows: zeroed here
    int zero_n;
    unsigned* ticket;                 // start-order counter (monotonic over launches) and its value before this launch
    unsigned ticket_base;
};

struct LeafSmem {
    double colbuf[2][LEAF_NG][LEAF_NX];   // pivot column (parity double buffer), exchanged WITHIN each half-warp
    double rowbuf[2][QB];                 // pivot-row entry of every column (parity double buffer)
    double red[2][8][QB];                 // per-warp partial dot products
    double stage[LEAF_NSTG][8][QB];       // rows pulled from the children (level > 0): [row % NSTG][segment][column]
    double Zs[QB][QB + 1];                // Zs[j][k] = v_k' v_j  (k < j)  =  U[k][j],  U = T^{-1}
    double Tm[QB][QB + 1];                // T
    double Wm[QB][QB + 1];                // scratch of the blocked inversion
    double taus[QB], rdiag[QB];
};

struct LeafCtx {
    int lane, wrp, cp, hh, rg;
    bool lazy, publish;
    uint4* mymail;            // this block's mailbox rows, already offset to column 2 cp  (publish)
    unsigned base;
    // puller state (level > 0): thread t pulls column (t & 31) of segment (t >> 5)
    const uint4* src;         // the child's mailbox, offset to that column (nullptr: no such child => zeros)
    uint4 pend;               // row requested during the previous step
    int pseg, pcol;
};

__device__ __forceinline__ uint4 ld_volatile_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u4(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 leaf_pack(double x, unsigned tag) {
    return make_uint4((unsigned)__double2loint(x), tag, (unsigned)__double2hiint(x), tag);
}
// request row `row` of the child (non-blocking)
__device__ __forceinline__ void leaf_request(LeafCtx& c, int row) {
    if (c.src != nullptr) c.pend = ld_volatile_u4(c.src + row * QB);
}
// complete the request for row `row`: poll until both halves carry this launch's tag for that row
__device__ __forceinline__ double leaf_complete(LeafCtx& c, int row) {
    if (c.src == nullptr) return 0.0;
    const unsigned tag = c.base + (unsigned)row + 1u;
    while (c.pend.y != tag || c.pend.w != tag) {
        __nanosleep(20);
        c.pend = ld_volatile_u4(c.src + row * QB);
    }
    return __hiloint2double((int)c.pend.z, (int)c.pend.x);
}

// one Householder step; COMP = j & 1 selects which of the thread's two columns can be the pivot column
template <int COMP, bool LAZY, bool PUB>
__device__ __forceinline__ void leaf_step(const int j, double (&xa)[LEAF_NX], double (&xb)[LEAF_NX], LeafSmem& sm, LeafCtx& c) {
    const int lane = c.lane, wrp = c.wrp, cp = c.cp, hh = c.hh, rg = c.rg;
    const int par = j & 1;
    const int rgj = j & (LEAF_NG - 1), ji = j >> 4;       // pivot row j: row group rgj, head slot ji
    if (LAZY) {
        // row j+2 (requested during the previous step) goes to the staging buffer; row j+3 is requested now
        if (j + LEAF_D - 1 < QB) sm.stage[(j + LEAF_D - 1) & (LEAF_NSTG - 1)][c.pseg][c.pcol] = leaf_complete(c, j + LEAF_D - 1);
        if (j + LEAF_D < QB) leaf_request(c, j + LEAF_D);
    }
    double v[LEAF_NX];
    {
        const double* cb = sm.colbuf[par][rg];
#pragma unroll
        for (int i = 0; i < LEAF_NX; i += 2) {
            const double2 t = *reinterpret_cast<const double2*>(cb + i);
            v[i] = t.x; v[i + 1] = t.y;
        }
    }
    double da0 = 0.0, da1 = 0.0, db0 = 0.0, db1 = 0.0;
#pragma unroll
    for (int i = 0; i < LEAF_NX; i += 2) {
        da0 = fma(v[i], xa[i], da0); da1 = fma(v[i + 1], xa[i + 1], da1);
        db0 = fma(v[i], xb[i], db0); db1 = fma(v[i + 1], xb[i + 1], db1);
    }
    double da = da0 + da1, db = db0 + db1;
    da += __shfl_xor_sync(0xffffffffu, da, 16);
    db += __shfl_xor_sync(0xffffffffu, db, 16);
    if (hh == 0) *reinterpret_cast<double2*>(&sm.red[par][wrp][2 * cp]) = make_double2(da, db);
    __syncthreads();                                   // the only block-wide barrier of the step
    if (LAZY && j + 1 < QB && rg == ((j + 1) & (LEAF_NG - 1))) {
        // row j+1 of the 8 children moves from the staging buffer into this row group's slots (it is all zero left of
        // column j+1, so the rest of this step leaves it alone, and the publication of column j+1 below includes it)
        const double* st = &sm.stage[(j + 1) & (LEAF_NSTG - 1)][0][2 * cp];
        if ((j + 1) >> 4) {
#pragma unroll
            for (int sg = 0; sg < 8; ++sg) { const double2 t = *reinterpret_cast<const double2*>(st + sg * QB); xa[2 * sg + 1] = t.x; xb[2 * sg + 1] = t.y; }
        } else {
#pragma unroll
            for (int sg = 0; sg < 8; ++sg) { const double2 t = *reinterpret_cast<const double2*>(st + sg * QB); xa[2 * sg] = t.x; xb[2 * sg] = t.y; }
        }
    }
    double sa, sb;
    {
        const double* rp = &sm.red[par][4 * hh][2 * cp];
        const double2 p0 = *reinterpret_cast<const double2*>(rp);
        const double2 p1 = *reinterpret_cast<const double2*>(rp + QB);
        const double2 p2 = *reinterpret_cast<const double2*>(rp + 2 * QB);
        const double2 p3 = *reinterpret_cast<const double2*>(rp + 3 * QB);
        sa = (p0.x + p1.x) + (p2.x + p3.x);
        sb = (p0.y + p1.y) + (p2.y + p3.y);
        sa += __shfl_xor_sync(0xffffffffu, sa, 16);
        sb += __shfl_xor_sync(0xffffffffu, sb, 16);
    }
    const double s_j = __shfl_sync(0xffffffffu, COMP ? sb : sa, j >> 1);      // sum over rows > j of a_j^2
    const double alpha = sm.rowbuf[par][j];
    const double2 rk = *reinterpret_cast<const double2*>(&sm.rowbuf[par][2 * cp]);
    double beta, tj, vjj;
    leaf_scalars(alpha, s_j, beta, tj, vjj);
    const double wza = fma(vjj, rk.x, sa), wzb = fma(vjj, rk.y, sb);   // v_j' a_k (k > j)  or  v_k' v_j (k < j)
    if (2 * cp + 1 > j) {
        const double cb_ = tj * wzb;
#pragma unroll
        for (int i = 0; i < LEAF_NX; ++i) xb[i] = fma(-cb_, v[i], xb[i]);
        if (rg == rgj) { if (ji == 0) xb[0] = fma(-cb_, vjj, xb[0]); else xb[1] = fma(-cb_, vjj, xb[1]); }
        if (2 * cp > j) {
            const double ca_ = tj * wza;
#pragma unroll
            for (int i = 0; i < LEAF_NX; ++i) xa[i] = fma(-ca_, v[i], xa[i]);
            if (rg == rgj) { if (ji == 0) xa[0] = fma(-ca_, vjj, xa[0]); else xa[1] = fma(-ca_, vjj, xa[1]); }
        }
    }
    if (rg == rgj) {
        if (PUB) {        // row j of R is final: hand it to the parent (fire and forget)
            const double ra = ji ? xa[1] : xa[0], rb = ji ? xb[1] : xb[0];
            const unsigned tag = c.base + (unsigned)j + 1u;
            uint4* dst = c.mymail + j * QB;
            st_volatile_u4(dst, leaf_pack((2 * cp > j) ? ra : ((2 * cp == j) ? beta : 0.0), tag));
            st_volatile_u4(dst + 1, leaf_pack((2 * cp + 1 > j) ? rb : ((2 * cp + 1 == j) ? beta : 0.0), tag));
        }
        if (cp == (j >> 1)) {                           // explicit diagonal entry of V
            if (COMP == 0) { if (ji == 0) xa[0] = vjj; else xa[1] = vjj; }
            else           { if (ji == 0) xb[0] = vjj; else xb[1] = vjj; }
        }
    }
    if (wrp == 0 && hh == 0) {
        if (2 * cp < j) sm.Zs[j][2 * cp] = wza;
        if (2 * cp + 1 < j) sm.Zs[j][2 * cp + 1] = wzb;
        if (lane == 0) { sm.taus[j] = tj; sm.rdiag[j] = beta; }
    }
    const int j1 = j + 1;
    if (j1 < QB) {
        const int rg1 = j1 & (LEAF_NG - 1), ji1 = j1 >> 4;
        if (cp == (j1 >> 1)) {    // publish the next pivot column: zeros at and above row j+1
            double* cn = sm.colbuf[par ^ 1][rg];
            const bool k0 = (0 > ji1) || (0 == ji1 && rg > rg1), k1 = (1 > ji1) || (1 == ji1 && rg > rg1);
            if (COMP == 0) {      // next pivot column is this thread's column b
                *reinterpret_cast<double2*>(cn) = make_double2(k0 ? xb[0] : 0.0, k1 ? xb[1] : 0.0);
#pragma unroll
                for (int i = 2; i < LEAF_NX; i += 2) *reinterpret_cast<double2*>(cn + i) = make_double2(xb[i], xb[i + 1]);
            } else {
                *reinterpret_cast<double2*>(cn) = make_double2(k0 ? xa[0] : 0.0, k1 ? xa[1] : 0.0);
#pragma unroll
                for (int i = 2; i < LEAF_NX; i += 2) *reinterpret_cast<double2*>(cn + i) = make_double2(xa[i], xa[i + 1]);
            }
        }
        if (rg == rg1)            // next pivot row (read after the next step's barrier)
            *reinterpret_cast<double2*>(&sm.rowbuf[par ^ 1][2 * cp]) = ji1 ? make_double2(xa[1], xb[1]) : make_double2(xa[0], xb[0]);
    }
    __syncwarp();
}

template <bool TIMING>
__global__ void __launch_bounds__(LEAF_THREADS, 2)
qr_tree_kernel_t(double* __restrict__ A, long long ld, long long c0, TreeParams tp, long long* __restrict__ tbuf) {
    __shared__ __align__(16) LeafSmem sm;

    const int tid = threadIdx.x;
    // Logical block id = order in which the CTAs actually START (a ticket), not blockIdx: the children of a block have
    // smaller logical ids, so by the time a parent exists every child has been given an SM slot and makes progress,
    // whatever order the hardware dispatches CTAs in (the same device as the dynamic tile ids of decoupled look-back scans).
    __shared__ unsigned s_ticket;
    // dynamic shared memory: the block as a [QB][QS] image (loaded and stored by the TMA unit, coalesced), + one mbarrier
    extern __shared__ __align__(128) unsigned char tree_dsm[];
    double* vt = (double*)tree_dsm;
    const uint32_t ld_bar = smem_u32(tree_dsm + AM_VS_BYTES);
    if (tid == 0) {
        s_ticket = atomicAdd(tp.ticket, 1u) - tp.ticket_base;
        mbar_init(ld_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    __syncthreads();
    const int bid = (int)s_ticket;
    int lev = 0;
    while (lev + 1 < tp.nlev && bid >= tp.start[lev + 1]) ++lev;
    const long long blk = bid - tp.start[lev];
    const TileMap tm = tp.tm[lev];
    const bool timed = TIMING && bid == (int)gridDim.x - 1;
    if (timed && tid == 32) tbuf[210] = clock64();
    // trace of selected blocks (globaltimer): start, loop start, loop end, exit
    int trace_slot = -1;
    if (TIMING && gridDim.x > 300) {
        const int sel[9] = {0, 147, 295, 296, 394, tp.start[1], tp.start[2] - 1, tp.start[2], (int)gridDim.x - 1};
        for (int q = 0; q < 9; ++q) if (bid == sel[q]) trace_slot = 220 + 4 * q;
    }
    if (trace_slot >= 0 && tid == 0) tbuf[trace_slot] = gtimer_ns();
    for (int i = blockIdx.x * LEAF_THREADS + tid; i < tp.zero_n; i += gridDim.x * LEAF_THREADS) tp.zero_ptr[i] = 0;
#define LEAF_T(slot) do { if (timed && tid == 32) tbuf[(slot)] = clock64(); } while (0)

    LeafCtx c;
    c.lane = tid & 31; c.wrp = tid >> 5; c.cp = c.lane & 15; c.hh = c.lane >> 4; c.rg = 2 * c.wrp + c.hh;
    c.lazy = lev > 0;
    c.publish = lev + 1 < tp.nlev;
    c.base = tp.base;
    c.src = nullptr; c.pend = make_uint4(0u, 0u, 0u, 0u); c.pseg = tid >> 5; c.pcol = tid & 31;
    const int cp = c.cp, rg = c.rg;
    double* __restrict__ cola = A + (c0 + 2 * cp) * ld;
    double* __restrict__ colb = cola + ld;
    c.mymail = tp.mail + (long long)bid * (QB * QB) + 2 * cp;
    double* __restrict__ heada = cola + tm.r0 + QB * blk;     // head row 0 of this block, column a
    const int body0 = QB + LEAF_BR * rg;              // level 0: first body row (tile coordinates)

    double xa[LEAF_NX], xb[LEAF_NX];
    if (!c.lazy) {
        // the block comes in as two bulk copies per column (head 256 B, body <= 1792 B), then every thread picks its
        // 2 x 16 entries out of shared memory
        const int nv = tm_body_rows(tm, blk);
        if (c.wrp == 0) {
            if (c.lane == 0) mbar_expect_tx(ld_bar, (uint32_t)(QB * (QB + nv)) * 8u);
            __syncwarp();
            const double* colp = A + (c0 + c.lane) * ld + tm.r0;
            const uint32_t dst = smem_u32(vt + c.lane * QS);
            bulk_g2s(dst, colp + QB * blk, QB * 8u, ld_bar);
            if (nv > 0) bulk_g2s(dst + QB * 8u, colp + QB * tm.nb + QBODY * blk, (uint32_t)nv * 8u, ld_bar);
        }
        mbar_wait(ld_bar, 0);
        const double* ta = vt + (2 * cp) * QS;
        const double* tb = ta + QS;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int p = rg + LEAF_NG * i;
            xa[i] = ta[p];
            xb[i] = tb[p];
        }
#pragma unroll
        for (int i = 0; i < LEAF_BR; i += 2) {
            const int p = body0 + i;
            const bool ok = (p - QB) < nv;
            double2 va = make_double2(0.0, 0.0), vb = make_double2(0.0, 0.0);
            if (ok) { va = *reinterpret_cast<const double2*>(ta + p); vb = *reinterpret_cast<const double2*>(tb + p); }
            xa[2 + i] = va.x; xa[3 + i] = va.y;
            xb[2 + i] = vb.x; xb[3 + i] = vb.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < LEAF_NX; ++i) { xa[i] = 0.0; xb[i] = 0.0; }
        // puller: segment 0 is child `blk` of the level below (its head is this block's head), segment s >= 1 is child
        // nb + 7 blk + s - 1 (this block's body rows 32 s ... 32 s + 31)
        const long long child = (c.pseg == 0) ? blk : tm.nb + (QG - 1) * blk + (c.pseg - 1);
        if (child < tp.tm[lev - 1].nb) c.src = tp.mail + (long long)(tp.start[lev - 1] + child) * (QB * QB) + c.pcol;
        leaf_request(c, 0);
        sm.stage[0][c.pseg][c.pcol] = leaf_complete(c, 0);
        leaf_request(c, 1);
        sm.stage[1][c.pseg][c.pcol] = leaf_complete(c, 1);
        leaf_request(c, 2);
        __syncthreads();
        if (rg == 0) {
            const double* st = &sm.stage[0][0][2 * cp];
#pragma unroll
            for (int sg = 0; sg < 8; ++sg) { const double2 t = *reinterpret_cast<const double2*>(st + sg * QB); xa[2 * sg] = t.x; xb[2 * sg] = t.y; }
        }
    }
    // publish pivot column 0 (zeros at and above the diagonal) and pivot row 0
    if (cp == 0) {
        double* cb = sm.colbuf[0][rg];
        *reinterpret_cast<double2*>(cb) = make_double2(rg > 0 ? xa[0] : 0.0, xa[1]);
#pragma unroll
        for (int i = 2; i < LEAF_NX; i += 2) *reinterpret_cast<double2*>(cb + i) = make_double2(xa[i], xa[i + 1]);
    }
    if (rg == 0) *reinterpret_cast<double2*>(&sm.rowbuf[0][2 * cp]) = make_double2(xa[0], xb[0]);
    __syncwarp();
    LEAF_T(0);
    if (trace_slot >= 0 && tid == 0) tbuf[trace_slot + 1] = gtimer_ns();

    // Reflectors are kept UNNORMALISED: v = a_j + sign(alpha) ||a_j|| e_j, H = I - t v v', t = 1 / (||a_j|| (||a_j|| + |alpha|)).
    // One block-wide barrier per step; the owners of column j+1 publish it from inside step j.
#define LEAF_LOOP(LZ, PB)                                                   \
    _Pragma("unroll 1") for (int jj = 0; jj < QB / 2; ++jj) {                \
        LEAF_T(1 + 2 * jj);                                                  \
        leaf_step<0, LZ, PB>(2 * jj, xa, xb, sm, c);                         \
        LEAF_T(2 + 2 * jj);                                                  \
        leaf_step<1, LZ, PB>(2 * jj + 1, xa, xb, sm, c);                     \
    }
    if (!c.lazy) { if (c.publish) { LEAF_LOOP(false, true) } else { LEAF_LOOP(false, false) } }
    else         { if (c.publish) { LEAF_LOOP(true, true) } else { LEAF_LOOP(true, false) } }
#undef LEAF_LOOP
    __syncthreads();
    LEAF_T(200);
    if (trace_slot >= 0 && tid == 0) tbuf[trace_slot + 2] = gtimer_ns();

    // ---- V (explicit diagonal entry, zeros above) to the workspace as ONE bulk store of its shared-memory image;
    //      R head back into the matrix ----
    {
        double* Va = vt + (2 * cp) * QS;
        double* Vbp = Va + QS;
        const int ca = 2 * cp, cbn = 2 * cp + 1;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = rg + LEAF_NG * i;                       // head row
            Va[r] = (r >= ca) ? xa[i] : 0.0;
            Vbp[r] = (r >= cbn) ? xb[i] : 0.0;
            if (!c.publish) {
                // Only the ROOT writes its R into the matrix.  The head rows of the blocks below it are the same matrix
                // rows (block b of every level heads at r0 + QB b); their R factors travel through the mailbox and nobody
                // reads them from the matrix, and writing them would race with the root's write a few steps later.
                heada[r] = (r < ca) ? xa[i] : ((r == ca) ? sm.rdiag[ca] : 0.0);
                heada[ld + r] = (r < cbn) ? xb[i] : ((r == cbn) ? sm.rdiag[cbn] : 0.0);
            }
        }
        if (!c.lazy) {
#pragma unroll
            for (int i = 0; i < LEAF_BR; i += 2) {
                *reinterpret_cast<double2*>(Va + body0 + i) = make_double2(xa[2 + i], xa[3 + i]);
                *reinterpret_cast<double2*>(Vbp + body0 + i) = make_double2(xb[2 + i], xb[3 + i]);
            }
        } else {
#pragma unroll
            for (int i = 2; i < LEAF_NX; ++i) {
                const int r = QB * (i >> 1) + rg + LEAF_NG * (i & 1);
                Va[r] = xa[i];
                Vbp[r] = xb[i];
            }
        }
        fence_async_smem();                  // generic-proxy writes -> visible to the bulk-copy engine
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(tp.V[lev] + blk * (long long)(QB * QS), smem_u32(vt), AM_VS_BYTES);
            bulk_commit();
        }
    }

    LEAF_T(202);
    // ---- T = U^{-1},  U = diag(1/t) + striu(V'V)  (U[k][i] = Zs[i][k], k < i).  Recursive blocked inversion of the upper
    //      triangular U:  [U11 U12; 0 U22]^{-1} = [T11, -T11 U12 T22; 0, T22], block size 1, 2, 4, 8, 16. ----
    {
        for (int e = tid; e < QB * QB; e += LEAF_THREADS) {
            const int r = e >> 5, cc = e & 31;
            sm.Tm[r][cc] = (r == cc) ? sm.taus[r] : 0.0;
        }
        __syncthreads();
#define LEAF_TINV_LEVEL(B)                                                                                               \
        {                                                                                                                \
            /* element (i, jn) of every pair's off-diagonal block: 16 B elements in all */                               \
            const int pr = tid / ((B) * (B)), rem = tid % ((B) * (B)), i = rem / (B), jn = rem % (B);                    \
            const int r0 = 2 * (B) * pr;                                                                                 \
            const bool act = tid < (QB / 2) * (B);                                                                       \
            if (act) { /* W = U12 T22 */                                                                                 \
                double acc = 0.0;                                                                                        \
                _Pragma("unroll") for (int k = 0; k < (B); ++k)                                                          \
                    acc = fma(sm.Zs[r0 + (B) + k][r0 + i], (k <= jn) ? sm.Tm[r0 + (B) + k][r0 + (B) + jn] : 0.0, acc);   \
                sm.Wm[r0 + i][r0 + (B) + jn] = acc;                                                                      \
            }                                                                                                            \
            __syncthreads();                                                                                             \
            if (act) { /* T12 = -T11 W */                                                                                \
                double acc = 0.0;                                                                                        \
                _Pragma("unroll") for (int k = 0; k < (B); ++k)                                                          \
                    acc = fma((k >= i) ? sm.Tm[r0 + i][r0 + k] : 0.0, sm.Wm[r0 + k][r0 + (B) + jn], acc);                \
                sm.Tm[r0 + i][r0 + (B) + jn] = -acc;                                                                     \
            }                                                                                                            \
            __syncthreads();                                                                                             \
        }
        LEAF_TINV_LEVEL(1) LEAF_TINV_LEVEL(2) LEAF_TINV_LEVEL(4) LEAF_TINV_LEVEL(8) LEAF_TINV_LEVEL(16)
#undef LEAF_TINV_LEVEL
        LEAF_T(203);
        double* __restrict__ Tb = tp.T[lev] + blk * (long long)(QB * QWS);     // Tb[c * QWS + k] = T[k][c]  (smem image of the update kernels)
        for (int e = tid; e < QB * QB; e += LEAF_THREADS) {
            const int cc = e >> 5, k = e & 31;
            Tb[cc * QWS + k] = sm.Tm[k][cc];
        }
    }
    if (tid == 0) bulk_wait0();             // the image must stay in shared memory until the copy engine is done with it
    LEAF_T(201);
    if (trace_slot >= 0 && tid == 0) tbuf[trace_slot + 3] = gtimer_ns();
#undef LEAF_T
}

// =================================================================================================
// trailing update, plain-FMA version (debug / cross-check path; ctx option qr_apply = 0)
// =================================================================================================
#define AF_SMEM_BYTES ((QB * QS + QB * 33 + QCT * QS + 2 * QB * (QCT + 1)) * 8)
__global__ void __launch_bounds__(256, 1)
qr_apply_fma_kernel(double* __restrict__ A, long long ld, long long ctrail, int ntiles, int tiles_per_cta,
                    TileMap tm, const double* __restrict__ V, const double* __restrict__ T) {
    extern __shared__ double asmem[];
    double* Vs = asmem;                // [QB][QS]
    double* Ts = Vs + QB * QS;         // Ts[i*33 + k] = T[k][i]
    double* Xs = Ts + QB * 33;         // [QCT][QS]
    double* W = Xs + QCT * QS;         // [QB][QCT+1]
    double* W2 = W + QB * (QCT + 1);
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const long long blk = blockIdx.x;
    const double* __restrict__ Vb = V + blk * (long long)(QB * QS);
    const double* __restrict__ Tb = T + blk * (long long)(QB * QWS);
    for (int e = tid; e < QB * QS; e += 256) Vs[e] = Vb[e];
    for (int e = tid; e < QB * QB; e += 256) { const int k = e & 31, i = e >> 5; Ts[i * 33 + k] = Tb[i * QWS + k]; }
    const int t0 = blockIdx.y * tiles_per_cta;
    int t1 = t0 + tiles_per_cta;
    if (t1 > ntiles) t1 = ntiles;
    for (int tile = t0; tile < t1; ++tile) {
        const long long cbase = ctrail + (long long)tile * QCT;
        __syncthreads();
        for (int col = wrp; col < QCT; col += 8) {
            const double* __restrict__ src = A + (cbase + col) * ld;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int p = q * 32 + lane;
                long long mrow;
                const bool ok = tm_row(tm, blk, p, mrow);
                Xs[col * QS + p] = ok ? src[mrow] : 0.0;
            }
        }
        __syncthreads();
        // W = V' X   (QB x QCT)
        for (int o = tid; o < QB * QCT; o += 256) {
            const int i = o & 31, c = o >> 5;
            double acc = 0.0;
            for (int r = 0; r < QH; ++r) acc = fma(Vs[i * QS + r], Xs[c * QS + r], acc);
            W[i * (QCT + 1) + c] = acc;
        }
        __syncthreads();
        // W2 = T' W
        for (int o = tid; o < QB * QCT; o += 256) {
            const int i = o & 31, c = o >> 5;
            double acc = 0.0;
            for (int k = 0; k <= i; ++k) acc = fma(Ts[i * 33 + k], W[k * (QCT + 1) + c], acc);
            W2[i * (QCT + 1) + c] = acc;
        }
        __syncthreads();
        // X -= V W2 ; thread = row
        {
            const int r = tid;
            long long mrow;
            const bool ok = tm_row(tm, blk, r, mrow);
            for (int c = 0; c < QCT; ++c) {
                double acc = Xs[c * QS + r];
                for (int i = 0; i < QB; ++i) acc = fma(-Vs[i * QS + r], W2[i * (QCT + 1) + c], acc);
                if (ok) A[(cbase + c) * ld + mrow] = acc;
            }
        }
    }
}

// =================================================================================================
// trailing update, first-generation tensor-pipe version (ctx option qr_apply = 1; kept as a cross-check):
// persistent CTAs, producer warp + 8 consumer warps, results stored by the consumers.
// =================================================================================================
#define AM_NST 2
#define AM_XS_BYTES (QCT * QS * 8)
#define QWP (QB + 8)    /* stride of the per-warp partial buffers: 16-byte stores of 8 lanes hit 32 distinct banks */
#define AM_SMEM_BYTES (AM_VS_BYTES + AM_NST * AM_XS_BYTES + 8 * QCT * QWP * 8 + 2 * QCT * QWS * 8 + QB * QWS * 8 + 64)

template <bool ATIMING>
__global__ void __launch_bounds__(288, 1)
qr_apply_mma_kernel_t(double* __restrict__ A, long long ld, long long ctrail, int ntiles, long long nblocks,
                    TileMap tm, const double* __restrict__ V, const double* __restrict__ T, long long* __restrict__ tbuf) {
    long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = 0;
    const long long tstart = ATIMING ? clock64() : 0;
    const long long tstart_ns = ATIMING ? gtimer_ns() : 0;
#define AP_T(slot) do { if (ATIMING && blockIdx.x == 0 && tid == 0) { const long long now = clock64(); tacc[slot] += now - tprev; tprev = now; } } while (0)
    extern __shared__ __align__(128) unsigned char amsm[];
    double* Vs = (double*)amsm;                                           // [QB][QS]
    double* Xs = (double*)(amsm + AM_VS_BYTES);                           // [AM_NST][QCT][QS]
    double* Wp = (double*)(amsm + AM_VS_BYTES + AM_NST * AM_XS_BYTES);    // [8][QCT][QWP] partial (V'X)' per warp
    double* Wsum = Wp + 8 * QCT * QWP;                                    // [QCT][QWS]
    double* Wfin = Wsum + QCT * QWS;                                      // [QCT][QWS]   -(T' V'X)
    double* Ts = Wfin + QCT * QWS;                                        // Ts[i*QWS + k] = T[k][i]
    uint64_t* bars = (uint64_t*)(Ts + QB * QWS);                          // full[], done[], vfull, vfree
    const uint32_t bar_full = smem_u32(bars), bar_done = smem_u32(bars + AM_NST), bar_v = smem_u32(bars + 2 * AM_NST),
                   bar_vfree = smem_u32(bars + 2 * AM_NST + 1);

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const long long jtot = nblocks * ntiles;
    const long long q_begin = (jtot * blockIdx.x) / gridDim.x;
    const long long q_end = (jtot * (blockIdx.x + 1)) / gridDim.x;
    const int njobs = (int)(q_end - q_begin);

    if (tid == 0) {
        for (int s = 0; s < AM_NST; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_done + 8 * s, 256);
        }
        mbar_init(bar_v, 1);
        mbar_init(bar_vfree, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    __syncthreads();
    if (njobs <= 0) return;

    if (wrp == 8) {
        // =========================== producer warp: TMA-unit bulk loads only ===========================
        long long cur_blk = -1;
        int seg_i = -1;
        for (int u = 0; u < njobs; ++u) {
            const long long q = q_begin + u;
            const long long blk = q / ntiles;
            const int tile = (int)(q - blk * ntiles);
            const int s = u % AM_NST;
            if (blk != cur_blk) {
                if (seg_i >= 0) mbar_wait(bar_vfree, (uint32_t)(seg_i & 1));   // consumers are done with the old V
                cur_blk = blk;
                ++seg_i;
                if (lane == 0) {
                    mbar_expect_tx(bar_v, AM_VS_BYTES);
                    bulk_g2s(smem_u32(Vs), V + blk * (long long)(QB * QS), AM_VS_BYTES, bar_v);
                }
            }
            if (u >= AM_NST) mbar_wait(bar_done + 8 * s, (uint32_t)(((u / AM_NST) - 1) & 1));   // stage drained
            double* xst = Xs + s * QCT * QS;
            const int nv = tm_body_rows(tm, blk);
            if (nv < QBODY) {
                for (int e = lane; e < QCT; e += 32)
                    for (int r = QB + nv; r < QH; ++r) xst[e * QS + r] = 0.0;
            }
            __syncwarp();
            if (lane == 0) mbar_expect_tx(bar_full + 8 * s, (uint32_t)(QCT * (QB + nv)) * 8u);
            __syncwarp();
            const long long cbase = ctrail + (long long)tile * QCT;
            const uint32_t xs = smem_u32(xst);
            {
                const int col = lane & (QCT - 1), seg = lane >> 4;      // 2 copies per column: head, body
                const double* colp = A + (cbase + col) * ld + tm.r0;
                if (seg == 0) bulk_g2s(xs + (uint32_t)(col * QS) * 8u, colp + QB * blk, QB * 8u, bar_full + 8 * s);
                else if (nv > 0)
                    bulk_g2s(xs + (uint32_t)(col * QS + QB) * 8u, colp + QB * tm.nb + QBODY * blk, (uint32_t)nv * 8u, bar_full + 8 * s);
            }
        }
        return;
    }

    // =============================== consumer warps ===============================
    // warp w owns tile rows [32w, 32w+32): its K-slice in GEMM1 and its output rows in GEMM2.  Both products
    // are formed transposed (tile columns on the MMA M axis) so that every accumulator fragment is two
    // consecutive ROWS of one column: 16-byte shared loads and 16-byte global stores.
    const int g = lane >> 2, t = lane & 3;
    long long cur_blk = -1;
    int seg_i = -1;
    long long wrow0 = 0;      // matrix row of this warp's first tile row
    bool wvalid = true;
    for (int u = 0; u < njobs; ++u) {
        const long long q = q_begin + u;
        const long long blk = q / ntiles;
        const int tile = (int)(q - blk * ntiles);
        const int s = u % AM_NST;
        if (blk != cur_blk) {
            cur_blk = blk;
            ++seg_i;
            const double* __restrict__ Tb = T + blk * (long long)(QB * QWS);
            for (int e = tid; e < QB * QB; e += 256) { const int k = e & 31, i = e >> 5; Ts[i * QWS + k] = Tb[i * QWS + k]; }
            wvalid = tm_row(tm, blk, 32 * wrp, wrow0);
            mbar_wait(bar_v, (uint32_t)(seg_i & 1));
        }
        const bool last_of_seg = (u + 1 == njobs) || ((q + 1) / ntiles != blk);
        if (ATIMING && blockIdx.x == 0 && tid == 0) { if (u == 0) tprev = clock64(); else AP_T(9); }
        mbar_wait(bar_full + 8 * s, (uint32_t)((u / AM_NST) & 1));
        AP_T(0);
        const double* Xst = Xs + s * QCT * QS;

        // ---- GEMM1 (transposed): partial (V'X)'_w = X[32w:32w+32, :]' V[32w:32w+32, :]   (QCT x QB) ----
        {
            double c1[2][4][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) c1[mi][ni][0] = c1[mi][ni][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int k0 = 32 * wrp + 4 * ks + t;
                double a[2], b[4];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) a[mi] = Xst[(8 * mi + g) * QS + k0];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = Vs[(8 * ni + g) * QS + k0];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma(c1[mi][ni], a[mi], b[ni]);
            }
            double* wp = Wp + wrp * QCT * QWP;
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    *reinterpret_cast<double2*>(wp + (8 * mi + g) * QWP + 8 * ni + 2 * t) =
                        make_double2(c1[mi][ni][0], c1[mi][ni][1]);
        }
        // accumulator init for GEMM2 (this warp's rows of the staged tile), issued now so the loads are in flight
        // during the reduction phases; afterwards this thread no longer reads the staged tile
        double c2[2][4][2];
        {
            const int rbase = 32 * wrp + 2 * t;
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const double2 v = *reinterpret_cast<const double2*>(Xst + (8 * mi + g) * QS + rbase + 8 * ni);
                    c2[mi][ni][0] = v.x; c2[mi][ni][1] = v.y;
                }
        }
        AP_T(1);
        consumer_sync();
        mbar_arrive(bar_done + 8 * s);          // every warp has finished GEMM1: the staged tile is free
        AP_T(2);
        // ---- reduce the 8 partials: Wsum[c][k] = (V'X)[k][c]  (tree order: fp64 adds have ~23-cycle latency) ----
        {
            const int c = tid >> 4, k = tid & 15;
            const double* w0 = Wp + c * QWP + k;
            double p[8], r[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) { p[w] = w0[w * QCT * QWP]; r[w] = w0[w * QCT * QWP + 16]; }
            Wsum[c * QWS + k] = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
            Wsum[c * QWS + k + 16] = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        }
        AP_T(3);
        consumer_sync();
        AP_T(4);
        // ---- Wfin' = -Wsum' T  on the tensor pipe: 2 x 4 fragments of 8 x 8, one per warp, K = 32 ----
        {
            const int mi = wrp & 1, ni = wrp >> 1;
            double ct[2][2] = {{0.0, 0.0}, {0.0, 0.0}};      // two accumulators (even / odd k-steps) to shorten the chain
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const double a = Wsum[(8 * mi + g) * QWS + 4 * ks + t];
                const double b = Ts[(8 * ni + g) * QWS + 4 * ks + t];
                dmma(ct[ks & 1], a, b);
            }
            *reinterpret_cast<double2*>(Wfin + (8 * mi + g) * QWS + 8 * ni + 2 * t) =
                make_double2(-(ct[0][0] + ct[1][0]), -(ct[0][1] + ct[1][1]));
        }
        AP_T(5);
        consumer_sync();
        AP_T(6);
        AP_T(7);
        // ---- GEMM2 (transposed): X[32w:32w+32, :]' += Wfin' V[32w:32w+32, :]' ; results go straight to HBM ----
        {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int k0 = 4 * ks + t;
                double a[2], b[4];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) a[mi] = Wfin[(8 * mi + g) * QWS + k0];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = Vs[k0 * QS + 32 * wrp + 8 * ni + g];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma(c2[mi][ni], a[mi], b[ni]);
            }
            if (ATIMING && blockIdx.x == 0 && tid == 0) { if (c2[0][0][0] == 1.2345e-300) tacc[11] = 1; }
            AP_T(8);
            if (last_of_seg) mbar_arrive(bar_vfree);     // V of this block is dead for this thread
            if (wvalid) {
                double* __restrict__ dst = A + (ctrail + (long long)tile * QCT + g) * ld + wrow0 + 2 * t;
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni)
                        *reinterpret_cast<double2*>(dst + (long long)(8 * mi) * ld + 8 * ni) =
                            make_double2(c2[mi][ni][0], c2[mi][ni][1]);
            }
        }
    }
    if (ATIMING && blockIdx.x == 0 && tid == 0) { AP_T(9); for (int i = 0; i < 10; ++i) tbuf[i] = tacc[i]; tbuf[10] = njobs; tbuf[11] = clock64() - tstart; tbuf[12] = gtimer_ns() - tstart_ns; }
#undef AP_T
}

// =================================================================================================
// trailing update, ping-pong version (ctx option qr_apply = 2, the default).
// Two independent consumer groups of 4 warps work on alternate tiles, so that one group's reduction / T-multiply /
// barrier phases run underneath the other group's DMMA phases, and NO consumer thread touches global memory:
//   producer warp : cp.async.bulk loads of V|T (per block) and of X tiles into a 4-stage ring, and cp.async.bulk
//                   STORES of the finished tiles straight from the ring (results are written back in place)
//   consumer warp w of a group owns tile rows [64w, 64w+64): its K-slice of W' = X'V (GEMM1) and its output rows of
//                   X' += Wfin' V' (GEMM2).  A warp only ever touches its own row slice of the staged tile.
// =================================================================================================
#define PP_NST 4
static_assert(PP_NST % 2 == 0, "a stage must always be used by the same consumer group: its full / out barriers are waited on by exactly one group per use, so that no phase is ever skipped");
#define PP_GW 4                                     /* warps per consumer group (8 measured 8 % slower: more barrier skew) */
#define PP_WR (QH / PP_GW)                           /* tile rows owned by one consumer warp */
#define PP_GT (32 * PP_GW)                          /* threads per consumer group */
#define PP_THREADS (32 * (2 * PP_GW + 1))           /* two consumer groups + the producer warp */
#define PP_T_BYTES (QB * QWS * 8)
#define PP_WS_BYTES (QCT * QWS * 8)
#define PP_SMEM_BYTES (AM_VS_BYTES + PP_T_BYTES + PP_NST * AM_XS_BYTES + 4 * PP_WS_BYTES + 128)

__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(PP_GT) : "memory"); }

// One launch can cover ONE tree level or ALL of them ("fused"): the job list of a CTA is then its slice of level 0,
// followed by its slice of level 1, ...  A job of level l > 0 (block b, tile t) needs the head rows that its up to
// 8 children at level l-1 wrote for tile t; the producer warp counts finished children per (parent block, tile) in
// global memory (children are published once their bulk stores have COMPLETED) and waits on that counter before it
// loads the tile.  Every CTA finishes its jobs of one level before touching the next, so the waits cannot cycle
// (all CTAs are resident: grid <= number of SMs, one CTA per SM).
struct ApplyLevels {
    int nlev;
    TileMap tm[QR_MAX_LEVELS];
    const double* V[QR_MAX_LEVELS];
    const double* T[QR_MAX_LEVELS];
    int* cnt[QR_MAX_LEVELS];          // level l >= 1: [nb_l * ntiles] finished-children counters (zero at launch)
};

// walks the job list of one CTA: level-major, block-major within a level
struct JobWalk {
    int lev, tile, left;              // left = jobs remaining in the current level, including the current one
    long long blk;
};
__device__ __forceinline__ void jw_enter_level(JobWalk& w, const ApplyLevels& L, int ntiles) {
    // skip levels in which this CTA has no job
    for (; w.lev < L.nlev; ++w.lev) {
        const long long jtot = L.tm[w.lev].nb * ntiles;
        const long long qb = (jtot * blockIdx.x) / gridDim.x, qe = (jtot * (blockIdx.x + 1)) / gridDim.x;
        if (qe > qb) {
            w.left = (int)(qe - qb);
            w.blk = qb / ntiles;
            w.tile = (int)(qb - w.blk * ntiles);
            return;
        }
    }
    w.left = 0;
}
__device__ __forceinline__ void jw_next(JobWalk& w, const ApplyLevels& L, int ntiles) {
    if (--w.left > 0) {
        if (++w.tile == ntiles) { w.tile = 0; ++w.blk; }
    } else {
        ++w.lev;
        jw_enter_level(w, L, ntiles);
    }
}
__device__ __forceinline__ int ld_acquire_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void bulk_wait1() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }

template <bool ATIMING>
__global__ void __launch_bounds__(PP_THREADS, 1)
qr_apply_pp_kernel_t(double* __restrict__ A, long long ld, long long ctrail, int ntiles, ApplyLevels L,
                     long long* __restrict__ tbuf) {
    long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = 0;
    const long long tstart = ATIMING ? clock64() : 0;
    const long long tstart_ns = ATIMING ? gtimer_ns() : 0;
#define PP_TM(slot) do { if (ATIMING && blockIdx.x == 0 && tid == 0) { const long long now = clock64(); tacc[slot] += now - tprev; tprev = now; } } while (0)
    extern __shared__ __align__(128) unsigned char ppsm[];
    double* Vs = (double*)ppsm;                                            // [QB][QS]
    double* Ts = (double*)(ppsm + AM_VS_BYTES);                            // Ts[i*QWS + k] = T[k][i]  (contiguous after Vs)
    double* Xs = (double*)(ppsm + AM_VS_BYTES + PP_T_BYTES);               // [PP_NST][QCT][QS]
    unsigned char* wbase = ppsm + AM_VS_BYTES + PP_T_BYTES + PP_NST * AM_XS_BYTES;
    uint64_t* bars = (uint64_t*)(wbase + 4 * PP_WS_BYTES);
    const uint32_t bar_full = smem_u32(bars), bar_out = smem_u32(bars + PP_NST), bar_v = smem_u32(bars + 2 * PP_NST),
                   bar_vfree = smem_u32(bars + 2 * PP_NST + 1);

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    int njobs = 0;
    for (int l = 0; l < L.nlev; ++l) {
        const long long jtot = L.tm[l].nb * ntiles;
        njobs += (int)((jtot * (blockIdx.x + 1)) / gridDim.x - (jtot * blockIdx.x) / gridDim.x);
    }

    if (tid == 0) {
        for (int s = 0; s < PP_NST; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_out + 8 * s, PP_GT);
        }
        mbar_init(bar_v, 1);
        mbar_init(bar_vfree, 2 * PP_GT);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    __syncthreads();
    if (njobs <= 0) return;

    JobWalk jw;
    jw.lev = 0; jw.tile = 0; jw.left = 0; jw.blk = 0;
    jw_enter_level(jw, L, ntiles);

    if (wrp == 2 * PP_GW) {
        // =========================== producer warp: all global traffic, via the TMA unit ===========================
        int seg_i = -1;
        int cur_lev = -1;
        long long cur_blk = -1;
        long long hist_blk[PP_NST];
        int hist_tile[PP_NST], hist_lev[PP_NST], hist_left[PP_NST];
        int next_store = 0;                     // jobs [0, next_store) of this CTA have had their stores issued
        // jobs whose stores have been issued but which have not been published to their parent yet (published in
        // batches: one completion wait + one release per batch keeps the fence off the per-tile path)
        int pq_lev[8], pq_tile[8], npq = 0;
        long long pq_blk[8];
        const int pcol = lane & (QCT - 1), pseg = lane >> 4;      // 2 copies per column: head, body
        const bool chained = L.nlev > 1;

        // tell the parent blocks that the first `cnt` queued jobs have written their head rows (their bulk stores have
        // completed in every lane: bulk groups are per thread)
        auto publish = [&](int cnt) {
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                for (int q = 0; q < cnt; ++q) {
                    if (pq_lev[q] + 1 < L.nlev) {
                        const long long nbp = L.tm[pq_lev[q] + 1].nb;
                        const long long pb = (pq_blk[q] < nbp) ? pq_blk[q] : (pq_blk[q] - nbp) / (QG - 1);
                        atomicAdd(L.cnt[pq_lev[q] + 1] + pb * ntiles + pq_tile[q], 1);
                    }
                }
            }
            for (int q = cnt; q < npq; ++q) { pq_lev[q - cnt] = pq_lev[q]; pq_tile[q - cnt] = pq_tile[q]; pq_blk[q - cnt] = pq_blk[q]; }
            npq -= cnt;
        };
        // issue the stores of the oldest unstored job (its stage holds the finished tile once bar_out completes)
        auto store_next = [&]() {
            const int j = next_store, s = j % PP_NST;
            mbar_wait(bar_out + 8 * s, (uint32_t)((j / PP_NST) & 1));
            const long long sb = hist_blk[s];
            const TileMap& stm = L.tm[hist_lev[s]];
            const long long cbase = ctrail + (long long)hist_tile[s] * QCT;
            const uint32_t xs = smem_u32(Xs + s * QCT * QS);
            const int snv = tm_body_rows(stm, sb);
            double* colp = A + (cbase + pcol) * ld + stm.r0;
            if (pseg == 0) bulk_s2g(colp + QB * sb, xs + (uint32_t)(pcol * QS) * 8u, QB * 8u);
            else if (snv > 0) bulk_s2g(colp + QB * stm.nb + QBODY * sb, xs + (uint32_t)(pcol * QS + QB) * 8u, (uint32_t)snv * 8u);
            bulk_commit();
            if (chained) {
                // everything but the group committed just now is complete after wait_group 1.  Level-0 jobs far from the
                // end of the CTA's list are published lazily (8 at a time); upper-level jobs and the last jobs of a level,
                // which some parent is about to wait for, one store step after their own
                const bool urgent = hist_lev[s] > 0 || hist_left[s] <= 8;
                if (npq > 0 && (urgent || npq == 8)) { bulk_wait1(); publish(npq); }
                pq_lev[npq] = hist_lev[s]; pq_tile[npq] = hist_tile[s]; pq_blk[npq] = sb; ++npq;
            }
            ++next_store;
        };

        for (int u = 0; u < njobs; ++u) {
            const int s = u % PP_NST;
            if (next_store <= u - PP_NST) {      // the stage must have been stored and read by the copy engine
                while (next_store <= u - PP_NST) store_next();
                bulk_wait_read0();
            }
            const TileMap& tm = L.tm[jw.lev];
            const long long blk = jw.blk;
            const int tile = jw.tile;
            if (ATIMING && lane == 0 && jw.lev != cur_lev) {
                if (blockIdx.x == gridDim.x / 2) tbuf[16 + jw.lev] = gtimer_ns() - tstart_ns;
                atomicMax((unsigned long long*)&tbuf[26 + jw.lev], (unsigned long long)gtimer_ns());      // slowest CTA to enter the level
                if (jw.lev == 0) atomicMin((unsigned long long*)&tbuf[24], (unsigned long long)gtimer_ns());
            }
            if (jw.lev != cur_lev || blk != cur_blk) {
                if (seg_i >= 0) mbar_wait(bar_vfree, (uint32_t)(seg_i & 1));   // both groups are done with the old V
                cur_lev = jw.lev;
                cur_blk = blk;
                ++seg_i;
                if (lane == 0) {
                    mbar_expect_tx(bar_v, AM_VS_BYTES + PP_T_BYTES);
                    bulk_g2s(smem_u32(Vs), L.V[jw.lev] + blk * (long long)(QB * QS), AM_VS_BYTES, bar_v);
                    bulk_g2s(smem_u32(Ts), L.T[jw.lev] + blk * (long long)(QB * QWS), PP_T_BYTES, bar_v);
                }
            }
            if (chained && jw.lev > 0) {
                // wait until every child of this block has written its head rows of this tile
                const long long nbc = L.tm[jw.lev - 1].nb;
                long long nch = nbc - tm.nb - (QG - 1) * blk;
                nch = 1 + (nch < 0 ? 0 : (nch > QG - 1 ? QG - 1 : nch));
                const int* cp = L.cnt[jw.lev] + blk * ntiles + tile;
                if (ld_acquire_s32(cp) < (int)nch) {
                    // before spinning, hand over everything this CTA still owes (a child may be one of its own last jobs)
                    while (next_store < u) store_next();
                    bulk_wait0();
                    publish(npq);
                    while (ld_acquire_s32(cp) < (int)nch) __nanosleep(100);
                }
                asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy acquire -> async-proxy loads below
            }
            double* xst = Xs + s * QCT * QS;
            const int nv = tm_body_rows(tm, blk);
            if (nv < QBODY) {
                for (int e = lane; e < QCT; e += 32)
                    for (int r = QB + nv; r < QH; ++r) xst[e * QS + r] = 0.0;
            }
            __syncwarp();
            if (lane == 0) mbar_expect_tx(bar_full + 8 * s, (uint32_t)(QCT * (QB + nv)) * 8u);
            __syncwarp();
            {
                const long long cbase = ctrail + (long long)tile * QCT;
                const uint32_t xs = smem_u32(xst);
                const double* colp = A + (cbase + pcol) * ld + tm.r0;
                if (pseg == 0) bulk_g2s(xs + (uint32_t)(pcol * QS) * 8u, colp + QB * blk, QB * 8u, bar_full + 8 * s);
                else if (nv > 0)
                    bulk_g2s(xs + (uint32_t)(pcol * QS + QB) * 8u, colp + QB * tm.nb + QBODY * blk, (uint32_t)nv * 8u, bar_full + 8 * s);
            }
            hist_blk[s] = blk;
            hist_tile[s] = tile;
            hist_lev[s] = jw.lev;
            hist_left[s] = jw.left;
            jw_next(jw, L, ntiles);
        }
        while (next_store < njobs) store_next();
        bulk_wait0();      // all stores have completed (not just been read) before the CTA retires
        if (chained) publish(npq);
        if (ATIMING && lane == 0) {
            if (blockIdx.x == gridDim.x / 2) tbuf[16 + L.nlev] = gtimer_ns() - tstart_ns;
            atomicMax((unsigned long long*)&tbuf[25], (unsigned long long)gtimer_ns());
            atomicAdd((unsigned long long*)&tbuf[30], (unsigned long long)gtimer_ns());
            atomicAdd((unsigned long long*)&tbuf[31], 1ull);
            atomicMin((unsigned long long*)&tbuf[23], (unsigned long long)gtimer_ns());
        }
        return;
    }

    // =============================== consumer groups ===============================
    const int grp = wrp / PP_GW, w4 = wrp % PP_GW;
    const int gtid = tid % PP_GT;
    const int g = lane >> 2, t = lane & 3;
    double* Wsum = (double*)(wbase + (2 * grp) * PP_WS_BYTES);          // [QCT][QWS]  (V'X)'
    double* Wfin = (double*)(wbase + (2 * grp + 1) * PP_WS_BYTES);      // [QCT][QWS]  -(T'V'X)'
    long long cur_blk = -1;
    int cur_lev = -1;
    int seg_i = -1;
    for (int u = 0; u < njobs; ++u) {
        if (jw.lev != cur_lev || jw.blk != cur_blk) {
            // EVERY consumer thread waits for EVERY V, also for blocks in which its group has no tile: a parity wait
            // is only meaningful against the phase that immediately precedes it, so no phase of bar_v may be skipped
            cur_lev = jw.lev; cur_blk = jw.blk; ++seg_i;
            mbar_wait(bar_v, (uint32_t)(seg_i & 1));
        }
        const bool last_of_seg = (jw.left == 1) || (jw.tile + 1 == ntiles);     // the next job has another V
        jw_next(jw, L, ntiles);
        if ((u & 1) != grp) {
            if (last_of_seg) mbar_arrive(bar_vfree);     // this thread's jobs of the segment are all behind it
            continue;
        }
        const int s = u % PP_NST;
        if (ATIMING && blockIdx.x == 0 && tid == 0) { if (u == 0) tprev = clock64(); else PP_TM(9); }
        mbar_wait(bar_full + 8 * s, (uint32_t)((u / PP_NST) & 1));
        PP_TM(0);
        double* Xst = Xs + s * QCT * QS;

        // this warp's rows of the staged tile become the GEMM2 accumulators now: after GEMM1 nobody reads the slice any
        // more, so the split-K partials of GEMM1 are parked in it (no separate partial buffers -> a 4th stage fits)
        double c2[2][PP_WR / 8][2];
        const int rbase = PP_WR * w4 + 2 * t;
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < PP_WR / 8; ++ni) {
                const double2 v = *reinterpret_cast<const double2*>(Xst + (8 * mi + g) * QS + rbase + 8 * ni);
                c2[mi][ni][0] = v.x; c2[mi][ni][1] = v.y;
            }
        // ---- GEMM1 (transposed): partial (V'X)'_w = X[64w:64w+64, :]' V[64w:64w+64, :]   (QCT x QB) ----
        {
            double c1[2][4][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) c1[mi][ni][0] = c1[mi][ni][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < PP_WR / 4; ++ks) {
                const int k0 = PP_WR * w4 + 4 * ks + t;
                double a[2], b[4];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) a[mi] = Xst[(8 * mi + g) * QS + k0];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = Vs[(8 * ni + g) * QS + k0];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma(c1[mi][ni], a[mi], b[ni]);
            }
            __syncwarp();       // every lane has finished reading the warp's slice
            // partial element (column c = 8 mi + g, reflector k = 8 ni + 2 t) -> Xst[c][PP_WR w + k]
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    *reinterpret_cast<double2*>(Xst + (8 * mi + g) * QS + PP_WR * w4 + 8 * ni + 2 * t) =
                        make_double2(c1[mi][ni][0], c1[mi][ni][1]);
        }
        PP_TM(1);
        group_sync(grp);
        PP_TM(2);
        // ---- reduce the PP_GW partials: Wsum[c][k] = (V'X)[k][c] ----
        {
            constexpr int NK = (QCT * QB) / PP_GT;           // outputs per thread: 4 (128 threads) or 2 (256 threads)
            const int c = gtid / (QB / NK), k0r = (gtid % (QB / NK)) * NK;
            const double* w0 = Xst + c * QS + k0r;
            double2 p[PP_GW][NK / 2];
#pragma unroll
            for (int w = 0; w < PP_GW; ++w)
#pragma unroll
                for (int h = 0; h < NK / 2; ++h) p[w][h] = *reinterpret_cast<const double2*>(w0 + PP_WR * w + 2 * h);
#pragma unroll
            for (int st = 1; st < PP_GW; st <<= 1)
#pragma unroll
                for (int w = 0; w < PP_GW; w += 2 * st)
#pragma unroll
                    for (int h = 0; h < NK / 2; ++h) { p[w][h].x += p[w + st][h].x; p[w][h].y += p[w + st][h].y; }
#pragma unroll
            for (int h = 0; h < NK / 2; ++h) *reinterpret_cast<double2*>(Wsum + c * QWS + k0r + 2 * h) = p[0][h];
        }
        PP_TM(3);
        group_sync(grp);
        PP_TM(4);
        // ---- Wfin' = -Wsum' T  on the tensor pipe: the 2 x 4 fragments (8 x 8 each) are spread over the group's warps ----
        {
            constexpr int NF = 8 / PP_GW;                    // fragments per warp: 2 (4 warps) or 1 (8 warps)
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const int fr = w4 * NF + f, mi = fr & 1, ni = fr >> 1;
                double ct[4][2];                             // 4 accumulators: dependent chains of 2 DMMAs
#pragma unroll
                for (int q = 0; q < 4; ++q) ct[q][0] = ct[q][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const double b = Ts[(8 * ni + g) * QWS + 4 * ks + t];
                    const double a = Wsum[(8 * mi + g) * QWS + 4 * ks + t];
                    dmma(ct[ks & 3], a, b);
                }
                *reinterpret_cast<double2*>(Wfin + (8 * mi + g) * QWS + 8 * ni + 2 * t) =
                    make_double2(-((ct[0][0] + ct[1][0]) + (ct[2][0] + ct[3][0])), -((ct[0][1] + ct[1][1]) + (ct[2][1] + ct[3][1])));
            }
        }
        PP_TM(5);
        group_sync(grp);
        PP_TM(6);
        // ---- GEMM2 (transposed): X[64w:64w+64, :]' += Wfin' V[64w:64w+64, :]' ; written back into the stage ----
        {
            PP_TM(7);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int k0 = 4 * ks + t;
                double a[2], b[PP_WR / 8];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) a[mi] = Wfin[(8 * mi + g) * QWS + k0];
#pragma unroll
                for (int ni = 0; ni < PP_WR / 8; ++ni) b[ni] = Vs[k0 * QS + PP_WR * w4 + 8 * ni + g];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < PP_WR / 8; ++ni) dmma(c2[mi][ni], a[mi], b[ni]);
            }
            if (ATIMING && blockIdx.x == 0 && tid == 0) { if (c2[0][0][0] == 1.2345e-300) tacc[11] = 1; }
            PP_TM(8);
            if (last_of_seg) mbar_arrive(bar_vfree);     // V of this block is dead for this thread
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < PP_WR / 8; ++ni)
                    *reinterpret_cast<double2*>(Xst + (8 * mi + g) * QS + rbase + 8 * ni) =
                        make_double2(c2[mi][ni][0], c2[mi][ni][1]);
        }
        fence_async_smem();                  // generic-proxy writes -> visible to the bulk-copy engine
        mbar_arrive(bar_out + 8 * s);
    }
    if (ATIMING && blockIdx.x == 0 && tid == 0) {
        PP_TM(9);
        for (int i = 0; i < 10; ++i) tbuf[i] = tacc[i];
        tbuf[10] = (njobs + 1) / 2;
        tbuf[11] = clock64() - tstart;
        tbuf[12] = gtimer_ns() - tstart_ns;
    }
#undef PP_TM
}

// =================================================================================================
// upper-triangular solve (single CTA; n is at most a few thousand on this path).  Left-looking by 32 x 32 diagonal
// blocks: the already-solved part is folded into the block's right-hand side by all 32 warps with coalesced loads
// (TRANS = 0: lanes = the block's 32 rows, warps stride over the solved columns; TRANS = 1: warp = one column of the
// block, lanes stride over the solved rows), then warp 0 solves the diagonal block with shuffles.
// =================================================================================================
#define TRI_THREADS 1024
template <int TRANS>
__global__ void __launch_bounds__(TRI_THREADS, 1)
tri_solve_kernel(int n, const double* __restrict__ R, long long ld, const double* c, double* xout) {
    extern __shared__ double tsm[];
    double* xs = tsm;                 // [n] right-hand side, overwritten by the solution
    double* D = tsm + n;              // [32][33] diagonal block
    double* red = D + 32 * 33;        // [32][33] partial sums
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    for (int i = tid; i < n; i += TRI_THREADS) xs[i] = c[i];
    const int nb = (n + 31) / 32;
    for (int bi = 0; bi < nb; ++bi) {
        const int jb = TRANS ? bi : (nb - 1 - bi);
        const int j0 = jb * 32;
        const int w = (n - j0 < 32) ? (n - j0) : 32;
        __syncthreads();
        {   // diagonal block into smem: D[r*33 + cc] = R[j0+r][j0+cc]
            const int r = tid & 31, cc = tid >> 5;
            D[r * 33 + cc] = (r < w && cc < w) ? R[(long long)(j0 + cc) * ld + j0 + r] : ((r == cc) ? 1.0 : 0.0);
        }
        double acc = 0.0;
        if (!TRANS) {
            // s_i = sum_{k >= j0 + 32} R[i][k] x[k],  i = j0 + lane ; warp strides over k
            if (lane < w) {
                const double* __restrict__ p = R + j0 + lane;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int k = j0 + 32 + wrp;
                for (; k + 96 < n; k += 128) {
                    a0 = fma(p[(long long)k * ld], xs[k], a0);
                    a1 = fma(p[(long long)(k + 32) * ld], xs[k + 32], a1);
                    a2 = fma(p[(long long)(k + 64) * ld], xs[k + 64], a2);
                    a3 = fma(p[(long long)(k + 96) * ld], xs[k + 96], a3);
                }
                for (; k < n; k += 32) a0 = fma(p[(long long)k * ld], xs[k], a0);
                acc = (a0 + a1) + (a2 + a3);
            }
            red[wrp * 33 + lane] = acc;
        } else {
            // s_j = sum_{i < j0} R[i][j] z[i],  j = j0 + wrp ; lanes stride over i
            if (wrp < w) {
                const double* __restrict__ p = R + (long long)(j0 + wrp) * ld;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int i = lane;
                for (; i + 96 < j0; i += 128) {
                    a0 = fma(p[i], xs[i], a0);
                    a1 = fma(p[i + 32], xs[i + 32], a1);
                    a2 = fma(p[i + 64], xs[i + 64], a2);
                    a3 = fma(p[i + 96], xs[i + 96], a3);
                }
                for (; i < j0; i += 32) a0 = fma(p[i], xs[i], a0);
                acc = (a0 + a1) + (a2 + a3);
            }
            red[lane * 33 + wrp] = acc;      // red[part][entry]: the same layout as TRANS = 0
        }
        __syncthreads();
        if (tid < 32) {
            double s = 0.0;
#pragma unroll 8
            for (int q = 0; q < 32; ++q) s += red[q * 33 + lane];
            double xi = (lane < w) ? xs[j0 + lane] - s : 0.0;
            const double inv = 1.0 / D[lane * 33 + lane];
            if (!TRANS) {
                for (int jj = w - 1; jj >= 0; --jj) {
                    const double xj = __shfl_sync(0xffffffffu, xi, jj) * __shfl_sync(0xffffffffu, inv, jj);
                    if (lane == jj) xi = xj;
                    else if (lane < jj) xi = fma(-D[lane * 33 + jj], xj, xi);
                }
            } else {   // R' is lower triangular: L[i][j] = R[j][i]
                for (int jj = 0; jj < w; ++jj) {
                    const double xj = __shfl_sync(0xffffffffu, xi, jj) * __shfl_sync(0xffffffffu, inv, jj);
                    if (lane == jj) xi = xj;
                    else if (lane > jj && lane < w) xi = fma(-D[jj * 33 + lane], xj, xi);
                }
            }
            if (lane < w) xs[j0 + lane] = xi;
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += TRI_THREADS) xout[i] = xs[i];
}

// =================================================================================================
// upper-triangular solve on MANY CTAs (ctx option "trisolve" = 1, default): CTA s owns the s-th 32 x 32 diagonal block in
// solve order (last block first for R x = c, first block first for R'z = c) and that block row / column of R.  It folds
// the already-solved blocks into its right-hand side AS THEY ARE PUBLISHED, solves its diagonal block in one warp and
// publishes its 32 values.  Values travel through an L2-resident mailbox as 16-byte units {lo32, tag, hi32, tag} (the
// data carries its own flag, as in the panel tree), so the dependent chain per block is one L2 round trip + a 32-step warp
// substitution instead of a pass of ONE CTA over the whole triangle: n = 1 000: 213 -> ~60 us, n = 4 000: 1.2 -> ~0.25 ms.
// A CTA's logical id is the order in which it STARTED (ticket), and it only waits for smaller ids: no deadlock whatever
// the dispatch order or residency.  Same operations per entry as the single-CTA kernel up to the order of the row sums.
// =================================================================================================
#define TMC_THREADS 256
__device__ __forceinline__ double tmc_poll(const uint4* p, unsigned tag) {
    uint4 v = ld_volatile_u4(p);
    while (v.y != tag || v.w != tag) {
        __nanosleep(20);
        v = ld_volatile_u4(p);
    }
    return __hiloint2double((int)v.z, (int)v.x);
}
template <int TRANS>
__global__ void __launch_bounds__(TMC_THREADS)
tri_solve_mc_kernel(int n, const double* __restrict__ R, long long ld, const double* c, double* xout, uint4* mail,
                    unsigned tag, unsigned* __restrict__ ticket, unsigned ticket_base) {
    __shared__ double D[32 * 33];
    __shared__ double red[8][33];
    __shared__ unsigned s_tk;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    if (tid == 0) s_tk = atomicAdd(ticket, 1u) - ticket_base;
    __syncthreads();
    const int s = (int)s_tk;
    const int nb = (n + 31) / 32;
    const int jb = TRANS ? s : nb - 1 - s;
    const int j0 = jb * 32;
    const int w = (n - j0 < 32) ? (n - j0) : 32;
    for (int e = tid; e < 1024; e += TMC_THREADS) {
        const int r = e & 31, cc = e >> 5;
        D[r * 33 + cc] = (r < w && cc < w) ? R[(long long)(j0 + cc) * ld + j0 + r] : ((r == cc) ? 1.0 : 0.0);
    }
    const double cval = (wrp == 0 && lane < w) ? c[j0 + lane] : 0.0;
    __syncthreads();
    // warp 0: its row (R x = c) or column (R'z = c) of the diagonal block, each entry divided by the diagonal entry of the
    // unknown it multiplies, so that a substitution step needs the RESIDUAL of row jj, not x_jj (padding: identity)
    double dn[32];
    double dinv_l = 1.0;
    if (wrp == 0) {
        dinv_l = 1.0 / D[lane * 33 + lane];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
            const double e = TRANS ? D[jj * 33 + lane] : D[lane * 33 + jj];
            dn[jj] = e * __shfl_sync(0xffffffffu, dinv_l, jj);
        }
    }
    // entries of R this thread multiplies with block k's values (4 per block); loaded one block ahead of the values
    auto load_r = [&](int k, double (&r)[4]) {
        const int k0 = (TRANS ? k : nb - 1 - k) * 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int cq = wrp + 8 * q;
            if (!TRANS) r[q] = (lane < w && k0 + cq < n) ? R[(long long)(k0 + cq) * ld + j0 + lane] : 0.0;   // row j0+lane, column k0+cq
            else        r[q] = (cq < w) ? R[(long long)(j0 + cq) * ld + k0 + lane] : 0.0;                      // row k0+lane, column j0+cq
        }
    };
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    double rc[4], rn[4];
    if (s > 0) load_r(0, rc);
    for (int k = 0; k < s; ++k) {
        if (k + 1 < s) load_r(k + 1, rn);
        const int k0 = (TRANS ? k : nb - 1 - k) * 32;
        const double xv = tmc_poll(mail + k0 + lane, tag);              // lane holds value k0 + lane of block k
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double xq = TRANS ? xv : __shfl_sync(0xffffffffu, xv, wrp + 8 * q);
            acc[q] = fma(rc[q], xq, acc[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) rc[q] = rn[q];
    }
    if (!TRANS) {
        red[wrp][lane] = (acc[0] + acc[1]) + (acc[2] + acc[3]);         // partial of row j0 + lane
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double a = acc[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) red[0][wrp + 8 * q] = a;                     // sum for column j0 + wrp + 8q
        }
    }
    __syncthreads();
    if (wrp != 0) return;
    double ssum;
    if (!TRANS) {
        ssum = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) ssum += red[q][lane];
    } else {
        ssum = red[0][lane];
    }
    // substitution in one warp, the lane's row of the (column-scaled) block in registers: one shuffle + one FMA per step
    double xi = (lane < w) ? cval - ssum : 0.0;
    if (!TRANS) {
#pragma unroll
        for (int jj = 31; jj >= 0; --jj) {
            const double rj = __shfl_sync(0xffffffffu, xi, jj);         // residual of row jj: x_jj = rj / D[jj][jj]
            if (lane < jj) xi = fma(-dn[jj], rj, xi);
        }
    } else {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
            const double rj = __shfl_sync(0xffffffffu, xi, jj);
            if (lane > jj) xi = fma(-dn[jj], rj, xi);
        }
    }
    xi *= dinv_l;
    if (lane >= w) xi = 0.0;
    st_volatile_u4(mail + j0 + lane, leaf_pack(xi, tag));               // fire and forget: the readers poll the data
    if (lane < w) xout[j0 + lane] = xi;
}

int tri_solve(lso_ctx* ctx, int64_t n, const double* d_R, int64_t ld, const double* d_c, double* d_x, int trans) {
    if (n == 0) return LSO_OK;
    LSO_REQUIRE(ctx, n <= 24000, "triangular solve: n too large");
    if (ctx->opt_trisolve && n > 32) {
        if (!ctx->d_trimail) {
            LSO_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_trimail, (size_t)24032 * sizeof(uint4)));
            LSO_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->d_trimail, 0, (size_t)24032 * sizeof(uint4), ctx->stream));
            ctx->tri_tag = 0;
        }
        if (++ctx->tri_tag == 0) {       // tag wrap: start over with a clean mailbox (stream-ordered)
            LSO_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->d_trimail, 0, (size_t)24032 * sizeof(uint4), ctx->stream));
            ctx->tri_tag = 1;
        }
        const unsigned nb = (unsigned)cdiv64(n, 32);
        unsigned* ticket = ctx->d_counters + 8;
        const unsigned base = ctx->tri_ticket_base;
        ctx->tri_ticket_base += nb;      // unsigned wrap-around is harmless: ids are differences
        if (trans) tri_solve_mc_kernel<1><<<nb, TMC_THREADS, 0, ctx->stream>>>((int)n, d_R, ld, d_c, d_x, (uint4*)ctx->d_trimail, ctx->tri_tag, ticket, base);
        else tri_solve_mc_kernel<0><<<nb, TMC_THREADS, 0, ctx->stream>>>((int)n, d_R, ld, d_c, d_x, (uint4*)ctx->d_trimail, ctx->tri_tag, ticket, base);
        LSO_CHECK_LAUNCH(ctx);
        return LSO_OK;
    }
    size_t smem = (size_t)(n + 2 * 32 * 33) * 8;
    static bool attr_done_dev[LSO_MAX_DEVICES] = {};      // function attributes are per device
    bool& attr_done = attr_done_dev[ctx->device % LSO_MAX_DEVICES];
    if (!attr_done) {
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(tri_solve_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(tri_solve_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_done = true;
    }
    if (trans) tri_solve_kernel<1><<<1, TRI_THREADS, smem, ctx->stream>>>((int)n, d_R, ld, d_c, d_x);
    else tri_solve_kernel<0><<<1, TRI_THREADS, smem, ctx->stream>>>((int)n, d_R, ld, d_c, d_x);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

// =================================================================================================
// plan + driver
// =================================================================================================
int qr_plan_create(lso_ctx* ctx, int64_t M, int64_t N, QRPlan* plan) {
    *plan = QRPlan();
    plan->M = M;
    plan->N = N;
    plan->Npad = roundup64(N > 0 ? N : 1, QB);
    plan->Nc = plan->Npad + QCT;
    plan->ld = roundup64(M, QB) + QH;
    size_t bytes = (size_t)plan->ld * plan->Nc * sizeof(double);
    cudaError_t e = cudaMalloc(&plan->A, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return lso_set_error(ctx, LSO_ERR_ALLOC, "QR workspace: cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    LSO_CHECK_CUDA(ctx, cudaMemsetAsync(plan->A, 0, bytes, ctx->stream));
    int64_t nb = cdiv64(M > 0 ? M : 1, QH);
    int L = 0;
    for (;;) {
        LSO_REQUIRE(ctx, L < QR_MAX_LEVELS, "too many TSQR levels");
        QRLevel& lv = plan->lev[L];
        lv.nblocks = nb;
        size_t vb = (size_t)nb * QB * QS * sizeof(double), tb = (size_t)nb * QB * QWS * sizeof(double);
        for (int b = 0; b < 2; ++b) {
            e = cudaMalloc(&lv.V[b], vb);
            if (e == cudaSuccess) e = cudaMalloc(&lv.T[b], tb);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return lso_set_error(ctx, LSO_ERR_ALLOC, "QR reflector workspace: %s", cudaGetErrorString(e));
            }
            LSO_CHECK_CUDA(ctx, cudaMemsetAsync(lv.V[b], 0, vb, ctx->stream));
            LSO_CHECK_CUDA(ctx, cudaMemsetAsync(lv.T[b], 0, tb, ctx->stream));
        }
        ++L;
        if (nb <= 1) break;
        nb = cdiv64(nb * QB, QH);
    }
    plan->nlevels = L;
    {
        int64_t tot = 0;
        for (int l = 0; l < L; ++l) tot += plan->lev[l].nblocks;
        LSO_CHECK_CUDA(ctx, cudaMalloc(&plan->mail, (size_t)tot * QB * QB * sizeof(uint4)));
        LSO_CHECK_CUDA(ctx, cudaMemsetAsync(plan->mail, 0, (size_t)tot * QB * QB * sizeof(uint4), ctx->stream));
        plan->prog_base = 0;
        int64_t upper = 0;
        for (int l = 1; l < L; ++l) upper += plan->lev[l].nblocks;
        const size_t cb = (size_t)(upper > 0 ? upper : 1) * (size_t)(plan->Nc / QCT) * sizeof(int);
        LSO_CHECK_CUDA(ctx, cudaMalloc(&plan->apply_cnt, cb));
        LSO_CHECK_CUDA(ctx, cudaMalloc(&plan->ticket, sizeof(unsigned)));
        LSO_CHECK_CUDA(ctx, cudaMemsetAsync(plan->ticket, 0, sizeof(unsigned), ctx->stream));
        plan->ticket_base = 0;
    }
    {   // the panel stream gets the highest priority: its (small, latency-bound) kernels must be placed as soon as
        // they are ready, underneath / ahead of the bulk trailing update
        int prio_lo = 0, prio_hi = 0;
        LSO_CHECK_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        LSO_CHECK_CUDA(ctx, cudaStreamCreateWithPriority(&plan->panel_stream, cudaStreamNonBlocking, prio_hi));
    }
    LSO_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&plan->ev_start, cudaEventDisableTiming));
    const int64_t npanels = plan->Npad / QB;
    plan->ev_leaf.resize(npanels);
    plan->ev_rest.resize(npanels);
    plan->ev_next.resize(npanels);
    for (int64_t k = 0; k < npanels; ++k) {
        LSO_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&plan->ev_leaf[k], cudaEventDisableTiming));
        LSO_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&plan->ev_rest[k], cudaEventDisableTiming));
        LSO_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&plan->ev_next[k], cudaEventDisableTiming));
    }
    static bool attr_done_dev[LSO_MAX_DEVICES] = {};      // function attributes are per device
    bool& attr_done = attr_done_dev[ctx->device % LSO_MAX_DEVICES];
    if (!attr_done) {
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(qr_apply_fma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM_BYTES));
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(qr_apply_mma_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AM_SMEM_BYTES));
        // same (maximum) shared-memory carveout for both kernels so a leaf CTA can join an SM that runs an update CTA
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(qr_apply_mma_kernel_t<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(qr_apply_mma_kernel_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AM_SMEM_BYTES));
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(qr_apply_pp_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM_BYTES));
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(qr_apply_pp_kernel_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM_BYTES));
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(qr_tree_kernel_t<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(qr_tree_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TREE_DSMEM_BYTES));
        LSO_CHECK_CUDA(ctx, cudaFuncSetAttribute(qr_tree_kernel_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TREE_DSMEM_BYTES));
        attr_done = true;
    }
    return qr_plan_tune(ctx, plan);
}

void qr_plan_destroy(QRPlan* plan) {
    if (!plan) return;
    if (plan->panel_stream) { cudaStreamSynchronize(plan->panel_stream); cudaStreamDestroy(plan->panel_stream); }
    if (plan->ev_start) cudaEventDestroy(plan->ev_start);
    for (cudaEvent_t ev : plan->ev_leaf) cudaEventDestroy(ev);
    for (cudaEvent_t ev : plan->ev_rest) cudaEventDestroy(ev);
    for (cudaEvent_t ev : plan->ev_next) cudaEventDestroy(ev);
    cudaFree(plan->A);
    cudaFree(plan->mail);
    cudaFree(plan->apply_cnt);
    cudaFree(plan->ticket);
    for (int l = 0; l < plan->nlevels; ++l)
        for (int b = 0; b < 2; ++b) {
            cudaFree(plan->lev[l].V[b]);
            cudaFree(plan->lev[l].T[b]);
        }
    *plan = QRPlan();
}

static long long* g_apply_tbuf = nullptr;    // debug: per-phase clock64 sums of CTA 0 / warp 0 of the first level-0 update
extern "C" int lso_debug_apply_timing(lso_ctx* ctx, long long* h_out /* 16 */) {
    if (!g_apply_tbuf) {
        LSO_CHECK_CUDA(ctx, cudaMalloc(&g_apply_tbuf, 32 * sizeof(long long)));
        LSO_CHECK_CUDA(ctx, cudaMemset(g_apply_tbuf, 0, 32 * sizeof(long long)));
        return LSO_OK;
    }
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaMemcpy(h_out, g_apply_tbuf, 32 * sizeof(long long), cudaMemcpyDeviceToHost));
    LSO_CHECK_CUDA(ctx, cudaMemset(g_apply_tbuf, 0, 32 * sizeof(long long)));      // re-arm
    LSO_CHECK_CUDA(ctx, cudaMemset(g_apply_tbuf + 24, 0xff, sizeof(long long)));
    LSO_CHECK_CUDA(ctx, cudaMemset(g_apply_tbuf + 23, 0xff, sizeof(long long)));
    return LSO_OK;
}
static long long* g_leaf_tbuf = nullptr;     // debug: per-phase clock64 stamps of the single-block leaf (LSO_LEAF_TIMING)
extern "C" int lso_debug_leaf_timing(lso_ctx* ctx, long long* h_out /* 256 */) {
    if (!g_leaf_tbuf) {
        LSO_CHECK_CUDA(ctx, cudaMalloc(&g_leaf_tbuf, 256 * sizeof(long long)));
        LSO_CHECK_CUDA(ctx, cudaMemset(g_leaf_tbuf, 0, 256 * sizeof(long long)));
        return LSO_OK;
    }
    LSO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    LSO_CHECK_CUDA(ctx, cudaMemcpy(h_out, g_leaf_tbuf, 256 * sizeof(long long), cudaMemcpyDeviceToHost));
    return LSO_OK;
}

struct PanelLevels {
    int L = 0;
    int64_t nblk[QR_MAX_LEVELS];
    TileMap tm[QR_MAX_LEVELS];
};

static void panel_levels(const QRPlan* plan, int64_t c0, PanelLevels& pl) {
    int L = 0;
    int64_t rows = roundup64(plan->M, QB) - c0;
    if (plan->band > 0) rows = std::min<int64_t>(rows, plan->band * (c0 + QB) - c0);   // the rest is zero in these columns
    for (;;) {
        const int64_t nb = cdiv64(rows, QH);
        pl.nblk[L] = nb;
        pl.tm[L].r0 = c0;
        pl.tm[L].nb = nb;
        pl.tm[L].rows = rows;
        ++L;
        if (nb <= 1) break;
        rows = nb * QB;
    }
    pl.L = L;
}

struct TLMark { cudaEvent_t ev; const char* what; int64_t k; int stream; };
static std::vector<TLMark> g_tl;
static bool g_tl_on = false;
static void tl_mark(cudaStream_t st, const char* what, int64_t k, int stream_id) {
    if (!g_tl_on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    g_tl.push_back({e, what, k, stream_id});
}
static int launch_leaf_chain(lso_ctx* ctx, QRPlan* plan, int64_t c0, const PanelLevels& pl, int buf, cudaStream_t st,
                             int zero_n = 0) {
    TreeParams tp;
    tp.nlev = pl.L;
    int total = 0;
    for (int l = 0; l < pl.L; ++l) {
        tp.start[l] = total;
        tp.tm[l] = pl.tm[l];
        tp.V[l] = plan->lev[l].V[buf];
        tp.T[l] = plan->lev[l].T[buf];
        total += (int)pl.nblk[l];
    }
    tp.start[pl.L] = total;
    tp.mail = plan->mail;
    if (plan->prog_base > 0xfffff000u) {      // counter wrap: start over (stream-ordered after every earlier launch)
        int64_t tot = 0;
        for (int l = 0; l < plan->nlevels; ++l) tot += plan->lev[l].nblocks;
        LSO_CHECK_CUDA(ctx, cudaMemsetAsync(plan->mail, 0, (size_t)tot * QB * QB * sizeof(uint4), st));
        plan->prog_base = 0;
    }
    plan->prog_base += 64;            // counters of earlier launches are all below the new base
    tp.base = plan->prog_base;
    tp.zero_ptr = plan->apply_cnt;
    tp.ticket = plan->ticket;
    tp.ticket_base = plan->ticket_base;
    plan->ticket_base += (unsigned)total;       // unsigned wrap-around is harmless: ids are differences
    tp.zero_n = zero_n;
    if (g_leaf_tbuf && (pl.nblk[0] <= 64 || (getenv("LSO_TREE_TRACE") && c0 == QB)))
        qr_tree_kernel_t<true><<<total, LEAF_THREADS, TREE_DSMEM_BYTES, st>>>(plan->A, plan->ld, c0, tp, g_leaf_tbuf);
    else
        qr_tree_kernel_t<false><<<total, LEAF_THREADS, TREE_DSMEM_BYTES, st>>>(plan->A, plan->ld, c0, tp, nullptr);
    LSO_CHECK_LAUNCH(ctx);
    return LSO_OK;
}

// all tree levels of the panel's update in ONE launch (levels chained through per-tile child counters)
static int launch_apply_fused(lso_ctx* ctx, QRPlan* plan, const PanelLevels& pl, int buf, int64_t cfirst, int ntiles, cudaStream_t st,
                              int lfirst = 0) {
    if (ntiles <= 0 || lfirst >= pl.L) return LSO_OK;
    ApplyLevels L;
    L.nlev = pl.L - lfirst;
    size_t off = 0;
    for (int l = lfirst; l < pl.L; ++l) {
        const int q = l - lfirst;
        L.tm[q] = pl.tm[l];
        L.V[q] = plan->lev[l].V[buf];
        L.T[q] = plan->lev[l].T[buf];
        L.cnt[q] = (q == 0) ? nullptr : plan->apply_cnt + off;
        if (q > 0) off += (size_t)pl.nblk[l] * ntiles;
    }
    const int64_t jtot = pl.nblk[lfirst] * ntiles;
    const int grid = (int)(jtot < ctx->num_sms ? jtot : ctx->num_sms);
    lso_prof_mark(ctx);
    if (g_apply_tbuf && cfirst == 2 * QB)
        qr_apply_pp_kernel_t<true><<<grid, PP_THREADS, PP_SMEM_BYTES, st>>>(plan->A, plan->ld, cfirst, ntiles, L, g_apply_tbuf);
    else
        qr_apply_pp_kernel_t<false><<<grid, PP_THREADS, PP_SMEM_BYTES, st>>>(plan->A, plan->ld, cfirst, ntiles, L, nullptr);
    lso_prof_mark(ctx);
    LSO_CHECK_LAUNCH(ctx);
    tl_mark(st, lfirst == 0 ? "apply (all levels) end" : "apply (upper levels, fused) end", cfirst / QB - 1, 0);
    return LSO_OK;
}

// apply the panel's block reflectors (all tree levels) to columns [cfirst, cfirst + ntiles*QCT)
static int launch_apply(lso_ctx* ctx, QRPlan* plan, const PanelLevels& pl, int buf, int64_t cfirst, int ntiles,
                        cudaStream_t st, bool mark, int lend = QR_MAX_LEVELS, int sm_cap = 0) {
    if (ntiles <= 0) return LSO_OK;
    const int64_t max_ctas = (sm_cap > 0 && sm_cap < ctx->num_sms) ? sm_cap : ctx->num_sms;
    for (int l = 0; l < pl.L && l < lend; ++l) {
        if (ctx->opt_qr_apply == 0) {
            int64_t chunks = cdiv64((int64_t)ctx->num_sms * 2, pl.nblk[l]);
            if (chunks > ntiles) chunks = ntiles;
            if (chunks < 1) chunks = 1;
            int tiles_per = (int)cdiv64(ntiles, chunks);
            dim3 grid((unsigned)pl.nblk[l], (unsigned)cdiv64(ntiles, tiles_per));
            qr_apply_fma_kernel<<<grid, 256, AF_SMEM_BYTES, st>>>(plan->A, plan->ld, cfirst, ntiles, tiles_per, pl.tm[l],
                                                                 plan->lev[l].V[buf], plan->lev[l].T[buf]);
        } else {
            int64_t jtot = pl.nblk[l] * ntiles;
            int grid = (int)(jtot < max_ctas ? jtot : max_ctas);
            if (mark) lso_prof_mark(ctx);
            if (ctx->opt_qr_apply >= 2) {
                ApplyLevels L;
                L.nlev = 1;
                L.tm[0] = pl.tm[l];
                L.V[0] = plan->lev[l].V[buf];
                L.T[0] = plan->lev[l].T[buf];
                L.cnt[0] = nullptr;
                if (g_apply_tbuf && l == 0 && cfirst == 2 * QB)
                    qr_apply_pp_kernel_t<true><<<grid, PP_THREADS, PP_SMEM_BYTES, st>>>(plan->A, plan->ld, cfirst, ntiles, L, g_apply_tbuf);
                else
                    qr_apply_pp_kernel_t<false><<<grid, PP_THREADS, PP_SMEM_BYTES, st>>>(plan->A, plan->ld, cfirst, ntiles, L, nullptr);
            } else if (g_apply_tbuf && l == 0 && cfirst <= 2 * QB)
                qr_apply_mma_kernel_t<true><<<grid, 288, AM_SMEM_BYTES, st>>>(plan->A, plan->ld, cfirst, ntiles, pl.nblk[l], pl.tm[l],
                                                                           plan->lev[l].V[buf], plan->lev[l].T[buf], g_apply_tbuf);
            else
                qr_apply_mma_kernel_t<false><<<grid, 288, AM_SMEM_BYTES, st>>>(plan->A, plan->ld, cfirst, ntiles, pl.nblk[l], pl.tm[l],
                                                                            plan->lev[l].V[buf], plan->lev[l].T[buf], nullptr);
            if (mark) lso_prof_mark(ctx);
        }
        LSO_CHECK_LAUNCH(ctx);
        static const char* lvl_names[QR_MAX_LEVELS] = {"apply L0 end", "apply L1 end", "apply L2 end", "apply L3 end", "apply L4 end",
                                                       "apply L5 end", "apply L6 end", "apply L7 end", "apply L8 end", "apply L9 end"};
        tl_mark(st, lvl_names[l], cfirst / QB - 1, st == ctx->stream ? 0 : 1);
    }
    return LSO_OK;
}

// Look-ahead schedule.  Main stream U carries the bulk trailing updates; the panel stream P carries, for each
// panel k, the update of just the next panel's columns followed by that panel's factorisation tree, so the
// latency-bound leaf kernels of panel k+1 run underneath the tensor-pipe-bound update of panel k (they fit on the
// same SMs: 188 KB + 15 KB of shared memory).  The bulk update of panel k is held back until the narrow update of
// the next panel's columns has been placed (two update kernels cannot share an SM).  V/T workspaces alternate
// between two buffers.  ctx option "qr_lookahead" = 0 runs everything in order on the main stream.
static void tl_dump(cudaStream_t U, cudaStream_t P) {
    if (!g_tl_on || g_tl.empty()) return;
    cudaStreamSynchronize(U);
    cudaStreamSynchronize(P);
    for (size_t i = 1; i < g_tl.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_tl[0].ev, g_tl[i].ev);
        fprintf(stderr, "TL %8.3f ms  %c  panel %3lld  %s\n", ms, g_tl[i].stream ? 'P' : 'U', (long long)g_tl[i].k, g_tl[i].what);
    }
    for (auto& m : g_tl) cudaEventDestroy(m.ev);
    g_tl.clear();
    g_tl_on = false;
}

static int qr_factor_run(lso_ctx* ctx, QRPlan* plan, int reserve, int64_t k_begin = 0, int64_t k_end = -1) {
    const int64_t M = plan->M;
    g_tl_on = (getenv("LSO_QR_TIMELINE") != nullptr) && plan->M > 50000;
    const int64_t npanels = std::min<int64_t>(plan->Npad / QB, cdiv64(M, QB));   // no rows left beyond that
    if (npanels <= 0) return LSO_OK;
    if (k_end < 0 || k_end > npanels) k_end = npanels;
    {   // algorithmic flops of this factorisation (bench.py roofline): the level-0 update of every panel, real columns only
        PanelLevels pl;
        for (int64_t k = k_begin; k < k_end; ++k) {
            panel_levels(plan, k * QB, pl);
            const int64_t trailing = plan->N + 1 - (k + 1) * QB;
            if (trailing > 0) ctx->stat_qr_update_flops += 4.0 * QB * (double)pl.tm[0].rows * (double)trailing;
        }
        const double Mq = (double)M, Nq = (double)plan->N;
        if (k_begin == 0) ctx->stat_qr_flops += 2.0 * Mq * Nq * Nq - 2.0 * Nq * Nq * Nq / 3.0;
    }
    cudaStream_t U = ctx->stream, P = plan->panel_stream;
    const int LA = QB / QCT;      // tiles that make up the next panel's columns
    PanelLevels cur, nxt;
    if (!ctx->opt_qr_lookahead || k_begin > 0 || k_end < npanels) {
        tl_mark(U, "start", 0, 0);
        for (int64_t k = k_begin; k < k_end; ++k) {
            const int64_t c0 = k * QB, ctrail = c0 + QB;
            panel_levels(plan, c0, cur);
            // qr_apply = 3: all levels in one launch; 4: levels 0 and 1 one launch each, the (tiny, latency-bound) levels
            // >= 2 chained in one launch
            const bool fused = ctx->opt_qr_apply == 3 && cur.L > 1;
            const bool fused_upper = ctx->opt_qr_apply == 4 && cur.L > 3;
            const int nt = (int)((plan->Nc - ctrail) / QCT);
            int zero_n = 0;
            if (fused) for (int l = 1; l < cur.L; ++l) zero_n += (int)cur.nblk[l] * nt;
            if (fused_upper) for (int l = 3; l < cur.L; ++l) zero_n += (int)cur.nblk[l] * nt;
            LSO_TRY(launch_leaf_chain(ctx, plan, c0, cur, 0, U, zero_n));
            tl_mark(U, "leaf chain end", k, 0);
            if (fused) {
                LSO_TRY(launch_apply_fused(ctx, plan, cur, 0, ctrail, nt, U));
            } else if (fused_upper) {
                const int64_t save = ctx->opt_qr_apply;
                ctx->opt_qr_apply = 2;
                const int st_ = launch_apply(ctx, plan, cur, 0, ctrail, nt, U, true, 2);
                ctx->opt_qr_apply = save;
                LSO_TRY(st_);
                LSO_TRY(launch_apply_fused(ctx, plan, cur, 0, ctrail, nt, U, 2));
            } else {
                LSO_TRY(launch_apply(ctx, plan, cur, 0, ctrail, nt, U, true));
            }
        }
        tl_dump(U, plan->panel_stream);
        return LSO_OK;
    }
    LSO_CHECK_CUDA(ctx, cudaEventRecord(plan->ev_start, U));
    LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(P, plan->ev_start, 0));
    tl_mark(U, "start", 0, 0);
    panel_levels(plan, 0, cur);
    LSO_TRY(launch_leaf_chain(ctx, plan, 0, cur, 0, P));
    LSO_CHECK_CUDA(ctx, cudaEventRecord(plan->ev_leaf[0], P));
    for (int64_t k = 0; k < npanels; ++k) {
        const int64_t c0 = k * QB, ctrail = c0 + QB;
        const int buf = (int)(k & 1);
        const int ntiles = (int)((plan->Nc - ctrail) / QCT);
        const bool has_next = (k + 1 < npanels);
        if (has_next) {
            if (k > 0) LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(P, plan->ev_rest[k - 1], 0));
            tl_mark(P, "applyNext begin", k, 1);
            LSO_TRY(launch_apply(ctx, plan, cur, buf, ctrail, LA, P, false));
            LSO_CHECK_CUDA(ctx, cudaEventRecord(plan->ev_next[k], P));
            tl_mark(P, "applyNext end / leaf(k+1) begin", k, 1);
            panel_levels(plan, ctrail, nxt);
            LSO_TRY(launch_leaf_chain(ctx, plan, ctrail, nxt, buf ^ 1, P));
            LSO_CHECK_CUDA(ctx, cudaEventRecord(plan->ev_leaf[k + 1], P));
            tl_mark(P, "leaf(k+1) end", k, 1);
            LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(U, plan->ev_next[k], 0));
            tl_mark(U, "applyRest begin", k, 0);
            // the next panel's tree runs under this launch: leave `reserve` SMs free for it
            LSO_TRY(launch_apply(ctx, plan, cur, buf, ctrail + LA * QCT, ntiles - LA, U, true, QR_MAX_LEVELS, ctx->num_sms - reserve));
            tl_mark(U, "applyRest end", k, 0);
        } else {
            LSO_CHECK_CUDA(ctx, cudaStreamWaitEvent(U, plan->ev_leaf[k], 0));
            LSO_TRY(launch_apply(ctx, plan, cur, buf, ctrail, ntiles, U, true));
        }
        LSO_CHECK_CUDA(ctx, cudaEventRecord(plan->ev_rest[k], U));
        if (has_next) cur = nxt;
    }
    tl_dump(U, P);
    return LSO_OK;
}

// panels [k_begin, k_end) only (the caller runs the ranges in order; used to interleave two factorisations panel by panel)
int qr_factor_range(lso_ctx* ctx, QRPlan* plan, int64_t k_begin, int64_t k_end) {
    const int64_t sa = ctx->opt_qr_apply, sl = ctx->opt_qr_lookahead;
    if (plan->sched_tuned && ctx->opt_qr_tune && !plan->sched_lookahead) ctx->opt_qr_apply = plan->sched_apply;
    ctx->opt_qr_lookahead = 0;
    const int st = qr_factor_run(ctx, plan, 0, k_begin, k_end);
    ctx->opt_qr_apply = sa;
    ctx->opt_qr_lookahead = sl;
    return st;
}
int64_t qr_num_panels(const QRPlan* plan) { return std::min<int64_t>(plan->Npad / QB, cdiv64(plan->M, QB)); }

int qr_factor(lso_ctx* ctx, QRPlan* plan) {
    if (!plan->sched_tuned || !ctx->opt_qr_tune) return qr_factor_run(ctx, plan, 0);
    const int64_t sa = ctx->opt_qr_apply, sl = ctx->opt_qr_lookahead;
    ctx->opt_qr_apply = plan->sched_apply;
    ctx->opt_qr_lookahead = plan->sched_lookahead;
    const int st = qr_factor_run(ctx, plan, plan->sched_reserve);
    ctx->opt_qr_apply = sa;
    ctx->opt_qr_lookahead = sl;
    return st;
}

// synthetic fill for the tuning runs: values in (-1, 1) from a counter hash, every entry of the first M rows
__global__ void qr_tune_fill_kernel(double* __restrict__ A, long long ld, long long M, long long ncols) {
    const long long col = blockIdx.y;
    if (col >= ncols) return;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < ld; r += (long long)gridDim.x * blockDim.x) {
        unsigned long long z = (unsigned long long)(col * 0x9E3779B97F4A7C15ull) ^ (unsigned long long)(r * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull);
        z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
        const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
        A[col * ld + r] = (r < M) ? 2.0 * u - 1.0 : 0.0;
    }
}

int qr_plan_tune(lso_ctx* ctx, QRPlan* plan) {
    plan->sched_tuned = false;
    if (!ctx->opt_qr_tune) return LSO_OK;
    PanelLevels p0;
    panel_levels(plan, 0, p0);
    int64_t tree_ctas = 0;
    for (int l = 0; l < p0.L; ++l) tree_ctas += p0.nblk[l];
    // only plans whose panel tree is small enough to run beside the update are latency-bound enough to matter
    // (and only moderately wide ones: above 2048 columns the twelve trial factorisations cost seconds and the per-panel
    // latency no longer dominates)
    if (tree_ctas > ctx->num_sms || plan->Npad / QB < 4 || plan->Npad > 2048) return LSO_OK;
    struct Cand { int apply, la, reserve; };
    const int reserve = (int)std::min<int64_t>((tree_ctas + 1) / 2, ctx->num_sms / 2);
    const Cand cand[4] = {{2, 0, 0}, {3, 0, 0}, {4, 0, 0}, {2, 1, reserve}};
    cudaEvent_t e0, e1;
    LSO_CHECK_CUDA(ctx, cudaEventCreate(&e0));
    LSO_CHECK_CUDA(ctx, cudaEventCreate(&e1));
    const int64_t sa = ctx->opt_qr_apply, sl = ctx->opt_qr_lookahead;
    const double sflops = ctx->stat_qr_flops, suflops = ctx->stat_qr_update_flops;
    const int64_t launches = ctx->launches;
    const int64_t prof = ctx->opt_profile;
    ctx->opt_profile = 0;
    int best = 0;
    int st = LSO_OK;
    for (int c = 0; c < 4 && st == LSO_OK; ++c) {
        float ms_best = 1e30f;
        for (int rep = 0; rep < 3 && st == LSO_OK; ++rep) {
            dim3 grid((unsigned)std::min<int64_t>(cdiv64(plan->ld, 256), 64), (unsigned)plan->Nc);
            qr_tune_fill_kernel<<<grid, 256, 0, ctx->stream>>>(plan->A, plan->ld, plan->M, plan->Nc);
            ctx->opt_qr_apply = cand[c].apply;
            ctx->opt_qr_lookahead = cand[c].la;
            cudaEventRecord(e0, ctx->stream);
            st = qr_factor_run(ctx, plan, cand[c].reserve);
            cudaEventRecord(e1, ctx->stream);
            if (cudaEventSynchronize(e1) != cudaSuccess) { st = lso_set_error(ctx, LSO_ERR_CUDA, "QR schedule tuning run failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < ms_best) ms_best = ms;        // the first run of a schedule warms the instruction caches
        }
        plan->tune_ms[c] = ms_best;
        if (ms_best < plan->tune_ms[best]) best = c;
    }
    ctx->opt_qr_apply = sa;
    ctx->opt_qr_lookahead = sl;
    ctx->opt_profile = prof;
    ctx->stat_qr_flops = sflops;
    ctx->stat_qr_update_flops = suflops;
    ctx->launches = launches;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    LSO_TRY(st);
    LSO_CHECK_CUDA(ctx, cudaMemsetAsync(plan->A, 0, (size_t)plan->ld * plan->Nc * sizeof(double), ctx->stream));
    plan->sched_apply = cand[best].apply;
    plan->sched_lookahead = cand[best].la;
    plan->sched_reserve = cand[best].reserve;
    plan->sched_tuned = true;
    if (getenv("LSO_QR_TUNE_VERBOSE"))
        fprintf(stderr, "[lsob200] QR plan %lld x %lld (band %lld): schedule ms {per-level %.3f, fused %.3f, fused-upper %.3f, look-ahead(reserve %d) %.3f} -> %d\n",
                (long long)plan->M, (long long)plan->N, (long long)plan->band, plan->tune_ms[0], plan->tune_ms[1], plan->tune_ms[2], reserve, plan->tune_ms[3], best);
    return LSO_OK;
}

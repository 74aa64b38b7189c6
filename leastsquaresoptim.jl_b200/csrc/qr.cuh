// qr.cuh — shared declarations of the communication-avoiding Householder QR (qr.cu) used by the
// QR solve, the TSQR multi-GPU reduction and the tests.
#pragma once
#include "common.cuh"

// Tiling constants of the CAQR:
//   QB  panel width (columns factorised together; reflectors per block reflector)
//   QH  rows of one leaf block (one CTA factorises a QH x QB block out of shared memory/registers)
//   QG  fan-in of the reduction tree (QH / QB heads are stacked into the next level's block; the rows of a
//       level are ordered "heads first", see TileMap in qr.cu)
//   QS  padded column stride (doubles) of V blocks in the workspace and in shared memory:
//       QS = QH + 4 makes every DMMA fragment load (4 rows x 8 cols, or 8 rows x 4 cols) hit 16
//       distinct 8-byte banks per half-warp.
//   QCT columns per trailing-update tile
#define QB 32
#define QH 256
#define QG (QH / QB)
#define QS (QH + 4)
#define QCT 16
#define QWS (QB + 4)
#define QR_MAX_LEVELS 10

struct QRLevel {
    int64_t nblocks;     // blocks at this level (for the widest panel, r0 = 0)
    double* V[2];        // nblocks * QB*QS doubles, double-buffered over panels (look-ahead)
    double* T[2];        // nblocks * QB*QB doubles
};

struct QRPlan {
    int64_t M = 0;       // rows of the matrix being factorised
    int64_t N = 0;       // columns to factorise
    int64_t Npad = 0;    // N rounded up to QB (extra columns are zero => identity reflectors)
    int64_t Nc = 0;      // total columns incl. right-hand-side tile: Npad + QCT
    int64_t ld = 0;      // leading dimension: roundup(M, QB) + QH zero rows of padding
    int64_t band = 0;    // > 0: row rho has no entry left of column rho / band (interleaved stack of triangles), so
                         // panel j only involves the rows above band * QB * (j + 1); 0 = dense
    double* A = nullptr; // ld x Nc, column-major
    int nlevels = 0;
    uint4* mail = nullptr;      // mailbox of the fused panel-tree kernel: [block][row][column] {lo32, tag, hi32, tag}
    int* apply_cnt = nullptr;   // finished-children counters of the fused (all levels in one launch) trailing update
    unsigned* ticket = nullptr; // start-order counter of the panel-tree kernel (logical block ids)
    unsigned ticket_base = 0;
    unsigned prog_base = 0;     // tag base of the current launch (row r carries tag base + r + 1)
    QRLevel lev[QR_MAX_LEVELS];
    // launch schedule chosen by qr_plan_tune for small (latency-bound) plans: which update mode, whether the next panel's
    // tree runs under the bulk update on SMs reserved for it
    int sched_apply = 2, sched_lookahead = 0, sched_reserve = 0;
    bool sched_tuned = false;
    float tune_ms[4] = {0.f, 0.f, 0.f, 0.f};
    // look-ahead: panel factorisations run on a second stream, overlapped with the previous trailing update
    cudaStream_t panel_stream = nullptr;
    cudaEvent_t ev_start = nullptr;
    std::vector<cudaEvent_t> ev_leaf, ev_rest, ev_next;   // per panel
};

int qr_plan_create(lso_ctx* ctx, int64_t M, int64_t N, QRPlan* plan);
void qr_plan_destroy(QRPlan* plan);
// Factorise plan->A in place: on return the leading N x N upper triangle holds R and column Npad
// rows 0..N-1 hold Q'b (the right-hand side that was stored in column Npad on entry).
int qr_factor(lso_ctx* ctx, QRPlan* plan);
// panels [k_begin, k_end) only, on ctx->stream; ranges must be run in order and cover [0, qr_num_panels)
int qr_factor_range(lso_ctx* ctx, QRPlan* plan, int64_t k_begin, int64_t k_end);
int64_t qr_num_panels(const QRPlan* plan);
// Small plans (a panel tree that fits on a third of the SMs) are latency-bound: time the launch schedules on synthetic
// data once and keep the fastest (results are bit-identical across schedules: every tile job does the same arithmetic).
int qr_plan_tune(lso_ctx* ctx, QRPlan* plan);
// x = R^{-1} c  (TRANS=0)  or  x = R^{-T} c (TRANS=1) for the upper-triangular n x n R at d_R (ld).
int tri_solve(lso_ctx* ctx, int64_t n, const double* d_R, int64_t ld, const double* d_c, double* d_x, int trans);

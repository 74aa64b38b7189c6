// chol.cuh — normal-equations path: J'J (syrk) + Cholesky + two triangular solves.
#pragma once
#include "common.cuh"

struct OzPlan;
void oz_plan_destroy(OzPlan* p);
// C (upper tiles) = J'J through int8 digit matrices on tcgen05 (ozaki.cu)
int oz_syrk_upper(lso_ctx* ctx, OzPlan** pp, int S, int64_t m, int64_t n, const double* d_J, int64_t ld, double* C,
                  int64_t ldc, double* part, int64_t part_cap);

struct CholPlan {
    int64_t n = 0;
    int64_t ldc = 0;          // roundup(n, 32)
    double* C = nullptr;      // ldc x n : J'J (+damp), overwritten by the upper factor R (R'R = C)
    double* rhs = nullptr;    // n : J'y.  C and rhs are contiguous ([C | rhs]) so one all-reduce covers both
    double* part = nullptr;   // split-K partial tiles
    int64_t part_cap = 0;     // number of n*ldc slabs available in part
    int* d_info = nullptr;
    double* packed = nullptr; // 2 x packed_len: [upper(J'J) by columns | J'y], the all-reduce buffer (+ the test hook's running sum)
    int64_t packed_len = 0;   // n(n+1)/2 + n
    OzPlan* oz = nullptr;     // digit matrices / tensor maps of the tcgen05 syrk (ctx option "syrk" = 2), created on first use
    bool kept = false;        // packed holds [upper(J'J) | J'y] of the last solve (before damping)
};

int chol_plan_create(lso_ctx* ctx, int64_t n, CholPlan* p);
void chol_plan_destroy(CholPlan* p);
int chol_solve(lso_ctx* ctx, CholPlan* p, int64_t m, int64_t n, const double* d_J, int64_t ld, const double* d_y,
               const double* d_damp, double* d_x, int sharded);
int chol_solve_kept(lso_ctx* ctx, CholPlan* p, const double* d_damp, double* d_x);
int chol_solve_emulated(lso_ctx* ctx, CholPlan* p, int P, int64_t ms, int64_t n, const double* d_J, int64_t ld,
                        const double* d_y, const double* d_damp, double* d_x);
// C (upper tiles) = J'J for the m x n column-major J
int syrk_upper(lso_ctx* ctx, CholPlan* p, int64_t m, int64_t n, const double* d_J, int64_t ld);
// in-place upper Cholesky of p->C ; returns LAPACK info (0 ok, k>0: leading minor k not positive definite)
int potrf_upper(lso_ctx* ctx, CholPlan* p, int* info_out);

// csc.cuh — device image of a SparseMatrixCSC{Float64,Int64}: CSC arrays (int32, 0-based) for the
// adjoint product / column norms and a CSR mirror (pattern built once per sparsity pattern, values
// refreshed by a gather after every g!) for the forward product, so both products are deterministic
// gathers with no floating-point atomics.
#pragma once
#include "common.cuh"

struct lso_csc {
    lso_ctx* ctx = nullptr;
    int64_t m = 0, n = 0, nnz = 0;
    int* d_colptr = nullptr;   // n+1
    int* d_rowidx = nullptr;   // nnz
    double* d_val = nullptr;   // nnz (CSC order)
    int* d_rowptr = nullptr;   // m+1
    int* d_colidx = nullptr;   // nnz (CSR order)
    int* d_perm = nullptr;     // nnz: CSR position -> CSC position
    double* d_valr = nullptr;  // nnz (CSR order)
    bool csr_dirty = true;
};

int csc_refresh_csr(lso_csc* A);

"""One Cholesky re-solve from the kept Gram matrix (potrf + two triangular solves) for an ncu launch list.  usage: n"""
import sys
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
m = n + 2000
ctx = L.Context.default(0)
A = L.DenseMatrix(ctx, m, n)
check(lib().lso_synth_dense_matrix(ctx.handle, m, n, 0, 20240608, A.ptr, A.ld), ctx.handle)
y = L.DeviceVector(ctx, m); check(lib().lso_synth_vector(ctx.handle, m, 0, 77, 1.0, y.ptr), ctx.handle)
dtd, x = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
A.colsumabs2(dtd); L.api._lm_damping(ctx, dtd, 0.1)
ws = L.DenseCholeskyAllocatedSolver(ctx, m, n, True)
ws.ldiv(x, A, y, dtd); ctx.sync()
ws.solve_kept(x, dtd); ctx.sync()

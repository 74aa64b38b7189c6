"""One LM(QR)-style damped solve at the bench shape, for ncu (launch list / full capture)."""
import sys
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
m, n = (100000, 1000) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
reps = 2 if len(sys.argv) < 4 else int(sys.argv[3])
ctx = L.Context.default(0)
A = L.DenseMatrix(ctx, m, n)
check(lib().lso_synth_dense_matrix(ctx.handle, m, n, 0, 20240608, A.ptr, A.ld), ctx.handle)
y = L.DeviceVector(ctx, m); check(lib().lso_synth_vector(ctx.handle, m, 0, 77, 1.0, y.ptr), ctx.handle)
dtd, x = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
ws = L.DenseQRAllocatedSolver(ctx, m, n, True)
for _ in range(reps):
    A.colsumabs2(dtd)
    L.api._lm_damping(ctx, dtd, 0.1)
    ws.ldiv(x, A, y, dtd)
ctx.sync()
print("ok", x.norm())

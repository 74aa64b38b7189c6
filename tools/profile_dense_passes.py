"""ncu driver: the HBM-bound dense passes of an LM iteration at 100 000 x 1 000 (colsumabs2, colsumabs2 + J'f fused, J'f,
J*x, ||J d - f||^2 fused)."""
import sys
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
m, n = 100000, 1000
ctx = L.Context.default(0)
A = L.DenseMatrix(ctx, m, n)
check(lib().lso_synth_dense_matrix(ctx.handle, m, n, 0, 20240608, A.ptr, A.ld), ctx.handle)
y, fp = L.DeviceVector(ctx, m), L.DeviceVector(ctx, m)
check(lib().lso_synth_vector(ctx.handle, m, 0, 77, 1.0, y.ptr), ctx.handle)
dtd, g, x = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
check(lib().lso_synth_vector(ctx.handle, n, 0, 78, 1.0, x.ptr), ctx.handle)
for _ in range(3):
    A.colsumabs2(dtd)
    A.colsumabs2_and_grad(dtd, g, y)
    A.mul_t(g, y, 1.0, 0.0)
    A.mul(fp, x, 1.0, 0.0)
    A.predicted_ssr(x, y, fp)
ctx.sync()
print("ok")

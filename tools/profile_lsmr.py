"""Small driver for ncu: a few SpMV / SpM'V / colsumabs2 launches at the config-3 shape."""
import sys
import numpy as np
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
m, n, k = 5_000_000, 500_000, 200
ctx = L.Context.default(0)
colptr = np.zeros(n + 1, dtype=np.int64); rowval = np.zeros(n * k, dtype=np.int64)
check(lib().lso_synth_csc_pattern(m, n, k, 20240609, colptr.ctypes.data, rowval.ctypes.data))
J = L.CSCMatrix(ctx, m, n, colptr - 1, rowval - 1)
aval = L.DeviceVector(ctx, n * k)
check(lib().lso_synth_vector(ctx.handle, n * k, 0, 99, 1.0, aval.ptr), ctx.handle)
check(lib().lso_csc_set_values_dev(J.handle, aval.ptr), ctx.handle)
x, g, d = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
f, t = L.DeviceVector(ctx, m), L.DeviceVector(ctx, m)
check(lib().lso_synth_vector(ctx.handle, n, 0, 7, 1.0, x.ptr), ctx.handle)
check(lib().lso_synth_vector(ctx.handle, m, 0, 8, 1.0, f.ptr), ctx.handle)
for _ in range(3):
    J.mul(t, x, 1.0, 0.0); J.mul_t(g, f, 1.0, 0.0); J.colsumabs2(d)
ctx.sync()
print("ok")

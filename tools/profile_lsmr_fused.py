"""ncu driver: a fixed number of fused-LSMR iterations at the config-3 shape (5M x 500k, nnz 1e8), damped.
    python tools/profile_lsmr_fused.py [iters] [spmv option]"""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
m, n, k = 5_000_000, 500_000, 200
ctx = L.Context.default(0)
if len(sys.argv) > 2:
    ctx.set_option("spmv", int(sys.argv[2]))
colptr = np.zeros(n + 1, dtype=np.int64); rowval = np.zeros(n * k, dtype=np.int64)
check(lib().lso_synth_csc_pattern(m, n, k, 20240609, colptr.ctypes.data, rowval.ctypes.data))
J = L.CSCMatrix(ctx, m, n, colptr - 1, rowval - 1)
aval = L.DeviceVector(ctx, n * k)
check(lib().lso_synth_vector(ctx.handle, n * k, 0, 99, 1.0, aval.ptr), ctx.handle)
check(lib().lso_csc_set_values_dev(J.handle, aval.ptr), ctx.handle)
f, dtd, dx = L.DeviceVector(ctx, m), L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
check(lib().lso_synth_vector(ctx.handle, m, 0, 8, 1.0, f.ptr), ctx.handle)
ws = L.LSMRDampenedAllocatedSolver(ctx, m, n)
for rep in range(2):
    J.colsumabs2(dtd)
    check(lib().lso_lm_damping(ctx.handle, n, dtd.ptr, 1e-6, 1e32, 0.1), ctx.handle)
    it, istop = C.c_int64(), C.c_int()
    check(lib().lso_lsmr_solve(ws._h, J.handle, None, 0, f.ptr, dtd.ptr, dx.ptr, 0.0, 0.0, 0.0, iters, C.byref(it), C.byref(istop)), ctx.handle)
ctx.sync()
print("ok", it.value, istop.value, ws.stats())

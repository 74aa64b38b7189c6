"""globaltimer trace of selected blocks of the panel-tree kernel (panel 1 of a 100000 x 1000 damped solve)."""
import ctypes as C, os, sys
import numpy as np
os.environ["LSO_TREE_TRACE"] = "1"
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
ctx = L.Context.default(0)
fn = lib().lso_debug_leaf_timing
fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_void_p]
buf = np.zeros(256, dtype=np.int64)
fn(ctx.handle, buf.ctypes.data)           # arm
m, n = 100000, 1000
A = L.DenseMatrix(ctx, m, n)
check(lib().lso_synth_dense_matrix(ctx.handle, m, n, 0, 20240608, A.ptr, A.ld), ctx.handle)
y = L.DeviceVector(ctx, m); check(lib().lso_synth_vector(ctx.handle, m, 0, 77, 1.0, y.ptr), ctx.handle)
dtd, x = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
A.colsumabs2(dtd); L.api._lm_damping(ctx, dtd, 0.1)
ws = L.DenseQRAllocatedSolver(ctx, m, n, True)
for _ in range(2):
    ws.ldiv(x, A, y, dtd)
fn(ctx.handle, buf.ctypes.data)
names = ["L0 #0", "L0 #147", "L0 #295", "L0 #296", "L0 #394", "L1 first", "L1 last", "L2 first", "root"]
t0 = min(int(buf[220 + 4 * q]) for q in range(9) if buf[220 + 4 * q])
print("block        start   loop-start  loop-end   exit   (us after the first traced block started)")
for q, nm in enumerate(names):
    v = [(int(buf[220 + 4 * q + i]) - t0) / 1e3 for i in range(4)]
    print(f"{nm:10s} {v[0]:8.1f} {v[1]:10.1f} {v[2]:9.1f} {v[3]:8.1f}")

import ctypes as C, sys
import numpy as np
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import lib
ctx = L.Context.default(0)
fn = lib().lso_debug_leaf_timing
fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_void_p]
buf = np.zeros(256, dtype=np.int64)
fn(ctx.handle, buf.ctypes.data)           # arm
m, n = (4000, 64) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
rng = np.random.default_rng(0)
ws = L.DenseQRAllocatedSolver(ctx, m, n, True)
x = L.DeviceVector(ctx, n)
J = L.DenseMatrix(ctx, m, n, rng.standard_normal((m, n))); y = L.DeviceVector(ctx, m, rng.standard_normal(m)); d = L.DeviceVector(ctx, n, np.ones(n))
for _ in range(3):
    ws.ldiv(x, J, y, d)
fn(ctx.handle, buf.ctypes.data)
t = buf
print("loop cycles", t[200] - t[0], "=> per step", (t[200] - t[0]) / 32, "| epilogue", t[201] - t[200])
print("per-step:", [int(t[2 + k] - t[1 + k]) for k in range(31)])
print("prologue", t[0] - t[210], "| V/R stores", t[202] - t[200], "| T inversion", t[203] - t[202], "| T store", t[201] - t[203])

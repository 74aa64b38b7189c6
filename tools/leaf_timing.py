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
m, n = 4000, 64
rng = np.random.default_rng(0)
ws = L.DenseQRAllocatedSolver(ctx, m, n, True)
x = L.DeviceVector(ctx, n)
J = L.DenseMatrix(ctx, m, n, rng.standard_normal((m, n))); y = L.DeviceVector(ctx, m, rng.standard_normal(m)); d = L.DeviceVector(ctx, n, np.ones(n))
for _ in range(3):
    ws.ldiv(x, J, y, d)
fn(ctx.handle, buf.ctypes.data)
t = buf
print("prologue (loads)->loop start: n/a ; total loop", t[200] - t[0], "epilogue", t[201] - t[200])
ph = np.zeros(6)
for j in range(32):
    b = 6 * j
    nxt = t[1 + 6 * (j + 1)] if j < 31 else t[200]
    ph += [t[2 + b] - t[1 + b], t[3 + b] - t[2 + b], t[4 + b] - t[3 + b], t[5 + b] - t[4 + b], t[6 + b] - t[5 + b], nxt - t[6 + b]]
print("avg cycles/step: publish %.0f | dot %.0f | barrier %.0f | reduce+scalars %.0f | update %.0f | loop overhead %.0f | sum %.0f" % (*(ph / 32), ph.sum() / 32))

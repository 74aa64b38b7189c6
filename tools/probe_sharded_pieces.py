"""Single-GPU timings of the pieces of the row-sharded QR at P ranks: one local QR of m/P rows, and the banded stack QR
(emulated-shards hook: P local QRs + the stack; the stack time is the difference)."""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
ctx = L.Context.default(0)
m, n = 100000, 1000
A = L.DenseMatrix(ctx, m, n)
check(lib().lso_synth_dense_matrix(ctx.handle, m, n, 0, 20240608, A.ptr, A.ld), ctx.handle)
y = L.DeviceVector(ctx, m); check(lib().lso_synth_vector(ctx.handle, m, 0, 77, 1.0, y.ptr), ctx.handle)
dtd, x = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
A.colsumabs2(dtd); L.api._lm_damping(ctx, dtd, 0.1)
def timeit(fn, reps=4):
    fn(); ctx.sync(); ts = []
    for _ in range(reps):
        ctx.sync(); t0 = time.perf_counter(); fn(); ctx.sync(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3
for P in (1, 2, 4, 8):
    ms = m // P
    wl = L.DenseQRAllocatedSolver(ctx, ms, n, damped=False)
    rank = C.c_int()
    t_local = timeit(lambda: wl.ldiv(x, A, y, None))         # first ms rows of A (ld = m)
    t_emul = timeit(lambda: check(lib().lso_debug_qr_solve_emulated_shards(wl._h, P, A.ptr, A.ld, y.ptr, dtd.ptr, x.ptr, C.byref(rank)), ctx.handle))
    print(f"P={P}: local QR of {ms} x {n}: {t_local:.2f} ms; P local QRs + banded stack QR: {t_emul:.2f} ms => stack part ~ {t_emul - P * t_local:.2f} ms", flush=True)
    del wl

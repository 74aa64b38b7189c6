"""Time of the triangular solves: multi-CTA (trisolve=1) vs single CTA (0), through lso_chol_solve_kept (potrf + R'z = c + R x = z)
and the QR finish; differences between the two settings isolate the solves."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
ctx = L.Context.default(0)
for n in (1000, 4000, 10000):
    m = n + 2000
    A = L.DenseMatrix(ctx, m, n)
    check(lib().lso_synth_dense_matrix(ctx.handle, m, n, 0, 20240608, A.ptr, A.ld), ctx.handle)
    y = L.DeviceVector(ctx, m); check(lib().lso_synth_vector(ctx.handle, m, 0, 77, 1.0, y.ptr), ctx.handle)
    dtd, x = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
    A.colsumabs2(dtd); L.api._lm_damping(ctx, dtd, 0.1)
    ws = L.DenseCholeskyAllocatedSolver(ctx, m, n, True)
    ws.ldiv(x, A, y, dtd); ctx.sync()
    res = {}
    for opt in (0, 1, 0, 1):
        ctx.set_option("trisolve", opt)
        ts = []
        for _ in range(5):
            ctx.sync(); t0 = time.perf_counter(); ws.solve_kept(x, dtd) if hasattr(ws, "solve_kept") else ws.ldiv(x, A, y, dtd, same_J=True); ctx.sync(); ts.append(time.perf_counter() - t0)
        res.setdefault(opt, []).append(min(ts) * 1e3)
    print(f"n={n}: chol kept re-solve (potrf + 2 triangular solves) single-CTA {min(res[0]):.3f} ms, multi-CTA {min(res[1]):.3f} ms => per solve {(min(res[0]) - min(res[1])) / 2 * 1e3:.0f} us saved", flush=True)
    del ws, A

"""QR solve time vs rows for the update-kernel launch modes (qr_apply = 2: one launch per level, 3: all levels chained
in one launch, 4: levels >= 2 chained)."""
import sys, time
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
ctx = L.Context.default(0)
n = 1000
for m in (3000, 6000, 12500, 25000, 50000, 100000):
    A = L.DenseMatrix(ctx, m, n)
    check(lib().lso_synth_dense_matrix(ctx.handle, m, n, 0, 20240608, A.ptr, A.ld), ctx.handle)
    y = L.DeviceVector(ctx, m); check(lib().lso_synth_vector(ctx.handle, m, 0, 77, 1.0, y.ptr), ctx.handle)
    dtd, x = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
    A.colsumabs2(dtd); L.api._lm_damping(ctx, dtd, 0.1)
    ws = L.DenseQRAllocatedSolver(ctx, m, n, True)
    out = []
    for mode in (2, 3, 4):
        ctx.set_option("qr_apply", mode)
        ws.ldiv(x, A, y, dtd); ctx.sync()
        ts = []
        for _ in range(4):
            ctx.sync(); t0 = time.perf_counter(); ws.ldiv(x, A, y, dtd); ctx.sync(); ts.append(time.perf_counter() - t0)
        out.append(min(ts) * 1e3)
    print(f"m={m:7d}: mode 2 {out[0]:6.2f} ms | mode 3 {out[1]:6.2f} ms | mode 4 {out[2]:6.2f} ms", flush=True)
    del ws, A, y

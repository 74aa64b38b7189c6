"""Does row-chunking pay when J is ALREADY on the device?  The host-fed chunked factorisation (chunks on their own streams,
stack QR pipelined behind) is fed from a second device copy of J (cudaMemcpyDefault: the 'transfer' is a 0.3 ms D2D copy),
and compared with the direct solve of the resident LM step."""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
import lsob200 as L, bench
from lsob200._lib import check, lib
ctx = L.Context.default(0)
m, n = 100000, 1000
prob = bench.DeviceProblem(L, ctx, m, n, 0, bench.SEED)
x = L.DeviceVector(ctx, n).copyto(prob.x0)
nls = L.LeastSquaresProblem(x=x, y=L.DeviceVector(ctx, m), f_=prob.f_, g_=prob.g_, J=L.DenseMatrix(ctx, m, n), device_callbacks=True, ctx=ctx)
anls = L.allocate(nls, L.LevenbergMarquardt(L.QR()))
prob.f_(anls.fcur, x); prob.g_(anls.J, x); ctx.sync()
J2 = L.DenseMatrix(ctx, m, n); f2 = L.DeviceVector(ctx, m)
check(lib().lso_vec_copy(ctx.handle, m * n, J2.ptr, anls.J.ptr), ctx.handle)
f2.copyto(anls.fcur)
dx = np.zeros(n)
dtd, dxd = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
def direct():
    anls.J.colsumabs2_and_grad(dtd, dxd, anls.fcur)
    L.api._lm_damping(ctx, dtd, 0.1)
    anls.solver.ldiv(dxd, anls.J, anls.fcur, dtd)
for _ in range(2): direct()
ts = []
for _ in range(5):
    ctx.sync(); t0 = time.perf_counter(); direct(); ctx.sync(); ts.append(time.perf_counter() - t0)
ref = dxd.download()
print(f"direct resident solve (colsumabs2 + damping + QR): {min(ts) * 1e3:.2f} ms", flush=True)
for chunks, twin in [(2, 2), (3, 2), (4, 2), (4, 3), ([0.3, 0.3, 0.25, 0.15], 2), (5, 3), (6, 3), (8, 3)]:
    ctx.set_option("qr_twin", twin)
    hs = L.HostStep(anls, chunks=chunks)
    for _ in range(2): hs.run(J2.ptr, f2.ptr, 10.0, dx)
    ts = []
    for _ in range(5):
        ctx.sync(); t0 = time.perf_counter(); hs.run(J2.ptr, f2.ptr, 10.0, dx); ts.append(time.perf_counter() - t0)
    print(f"chunks {chunks} workspaces {twin + 1}: {min(ts) * 1e3:.2f} ms per step incl. D2D 'transfer', tail and D2H; dx vs direct {np.linalg.norm(dx - ref) / np.linalg.norm(ref):.1e}", flush=True)
    del hs

"""End-to-end LM step from pinned host memory at 100 000 x 1 000 for different chunkings of the host-fed factorisation
(chunks, twin workspace on/off).  Prints ms per step (device events are not needed: the step ends with a D2H read)."""
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
hJ, hf = C.c_void_p(), C.c_void_p()
check(lib().lso_host_alloc_pinned(ctx.handle, m * n * 8, C.byref(hJ)), ctx.handle)
check(lib().lso_host_alloc_pinned(ctx.handle, m * 8, C.byref(hf)), ctx.handle)
check(lib().lso_download(ctx.handle, hJ, anls.J.ptr, m * n * 8), ctx.handle)
check(lib().lso_download(ctx.handle, hf, anls.fcur.ptr, m * 8), ctx.handle)
dx = np.zeros(n)
ref = None
cases = [(1, 2), ([0.3, 0.3, 0.25, 0.15], 2), ([0.2, 0.3, 0.3, 0.2], 2), ([0.25, 0.3, 0.25, 0.2], 2), ([0.2, 0.3, 0.3, 0.2], 3),
         ([0.15, 0.25, 0.25, 0.2, 0.15], 2), ([0.15, 0.25, 0.25, 0.2, 0.15], 3), ([0.2, 0.25, 0.25, 0.18, 0.12], 3), ([0.25, 0.25, 0.25, 0.25], 3),
         ([0.1, 0.2, 0.25, 0.25, 0.2], 3), ([0.28, 0.28, 0.26, 0.18], 2), ([0.3, 0.28, 0.24, 0.18], 2), ([0.32, 0.3, 0.24, 0.14], 2),
         ([0.3, 0.3, 0.27, 0.13], 2), ([0.33, 0.33, 0.34], 2), ([0.4, 0.35, 0.25], 2), ([0.38, 0.36, 0.26], 2), (None, 2)]
if len(sys.argv) > 1 and sys.argv[1] == "nopipe": ctx.set_option("qr_shard_pipeline", 0)
for chunks, twin in cases:
    ctx.set_option("qr_twin", twin)
    hs = L.HostStep(anls, chunks=chunks)
    for _ in range(2): hs.run(hJ.value, hf.value, 10.0, dx)
    ts = []
    for _ in range(6):
        ctx.sync(); t0 = time.perf_counter(); hs.run(hJ.value, hf.value, 10.0, dx); ts.append(time.perf_counter() - t0)
    if ref is None: ref = dx.copy()
    print(f"chunks {chunks} extra workspaces {twin}: {min(ts) * 1e3:.2f} ms per step (median {np.median(ts) * 1e3:.2f}); rows {hs.chunk_rows}; dx vs unchunked {np.linalg.norm(dx - ref) / np.linalg.norm(ref):.1e}", flush=True)
    del hs

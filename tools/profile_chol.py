"""One damped Cholesky solve (syrk + potrf) for ncu / timing.  usage: profile_chol.py m n reps [syrk option] [ozaki digits]
With option "profile" the syrk part (for syrk = 2: digit split + tcgen05 tile kernel + reduce) is timed by itself."""
import sys, time
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
m, n = (100000, 1000) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
reps = 2 if len(sys.argv) < 4 else int(sys.argv[3])
ctx = L.Context.default(0)
if len(sys.argv) > 4: ctx.set_option('syrk', int(sys.argv[4]))
if len(sys.argv) > 5: ctx.set_option('ozaki_slices', int(sys.argv[5]))
A = L.DenseMatrix(ctx, m, n)
check(lib().lso_synth_dense_matrix(ctx.handle, m, n, 0, 20240608, A.ptr, A.ld), ctx.handle)
y = L.DeviceVector(ctx, m); check(lib().lso_synth_vector(ctx.handle, m, 0, 77, 1.0, y.ptr), ctx.handle)
dtd, x = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
ws = L.DenseCholeskyAllocatedSolver(ctx, m, n, True)
A.colsumabs2(dtd); L.api._lm_damping(ctx, dtd, 0.1)
ws.ldiv(x, A, y, dtd); ctx.sync()
ctx.set_option("profile", 1); ctx.profile_read()
ts = []
for _ in range(reps):
    ctx.sync(); t0 = time.perf_counter(); ws.ldiv(x, A, y, dtd); ctx.sync(); ts.append(time.perf_counter() - t0)
print("chol solve ms", [round(t * 1e3, 3) for t in ts], "syrk-equivalent TFLOP/s %.2f" % (m * n * (n + 1) / min(ts) / 1e12), "x norm", x.norm())
ms, cnt = ctx.profile_read()
print("syrk part: %.3f ms per solve => %.2f TFLOP/s fp64-equivalent" % (ms / max(cnt, 1), m * n * (n + 1) / (ms / max(cnt, 1) * 1e-3) / 1e12))

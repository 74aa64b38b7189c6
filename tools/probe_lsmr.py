"""GPU probe for BASELINE.json configs[2]: sparse CSC J (default 5M x 500k, 200 nnz/col => ~20 nnz/row, nnz = 1e8),
LevenbergMarquardt(LSMR()): SpMV / SpM'V / colsumabs2 bandwidth and LM(LSMR) step time.  Prints JSON lines."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib

m = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 200
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = L.Context.default(0)
import os
if os.environ.get('LSO_SPMV'):
    ctx.set_option('spmv', int(os.environ['LSO_SPMV']))
t0 = time.perf_counter()
colptr = np.zeros(n + 1, dtype=np.int64)
rowval = np.zeros(n * k, dtype=np.int64)
check(lib().lso_synth_csc_pattern(m, n, k, 20240609, colptr.ctypes.data, rowval.ctypes.data))
t1 = time.perf_counter()
J = L.CSCMatrix(ctx, m, n, colptr - 1, rowval - 1)
t2 = time.perf_counter()
nnz = n * k
print(json.dumps({"m": m, "n": n, "nnz": nnz, "pattern_s": t1 - t0, "csc_create_s": t2 - t1}), flush=True)
del colptr, rowval

# values A (hash), model r = t + c t^2 - b with t = A x
aval = L.DeviceVector(ctx, nnz)
check(lib().lso_synth_vector(ctx.handle, nnz, 0, 99, 1.0, aval.ptr), ctx.handle)
check(lib().lso_csc_set_values_dev(J.handle, aval.ptr), ctx.handle)
x, xs, y, g, dtd = (L.DeviceVector(ctx, n) for _ in range(5))
f, t, b = L.DeviceVector(ctx, m), L.DeviceVector(ctx, m), L.DeviceVector(ctx, m)
check(lib().lso_synth_vector(ctx.handle, n, 0, 7, 1.0, xs.ptr), ctx.handle)
check(lib().lso_synth_vector(ctx.handle, m, 0, 8, 1.0, f.ptr), ctx.handle)


def timeit(fn, reps=5):
    fn(); ctx.sync()
    ts = []
    for _ in range(reps):
        ctx.sync(); a = time.perf_counter(); fn(); ctx.sync(); ts.append(time.perf_counter() - a)
    return min(ts)


res = {}
tn = timeit(lambda: J.mul(t, xs, 1.0, 0.0))
tt = timeit(lambda: J.mul_t(g, f, 1.0, 0.0))
tc = timeit(lambda: J.colsumabs2(dtd))
res["spmv_ms"] = tn * 1e3
res["spmv_gbs"] = (12 * nnz + 8 * (m + n) + 4 * m) / tn / 1e9
res["spmtv_ms"] = tt * 1e3
res["spmtv_gbs"] = (12 * nnz + 8 * (m + n) + 4 * n) / tt / 1e9
res["colsumabs2_ms"] = tc * 1e3
res["colsumabs2_gbs"] = (8 * nnz + 12 * n) / tc / 1e9
print(json.dumps(res), flush=True)

# LM(LSMR) on the sparse polynomial model, device callbacks
cmod = 0.1
J.mul(t, xs, 1.0, 0.0)
# b = t + c t^2 at x*  (+ noise): reuse the dense residual kernel pieces through vector ops
tt2 = L.DeviceVector(ctx, m).mul_(t, t)
b.copyto(t).axpy(cmod, tt2)
noise = L.DeviceVector(ctx, m)
check(lib().lso_synth_vector(ctx.handle, m, 0, 12, 1e-3, noise.ptr), ctx.handle)
b.axpy(1.0, noise)
pert = L.DeviceVector(ctx, n)
check(lib().lso_synth_vector(ctx.handle, n, 0, 13, 0.1, pert.ptr), ctx.handle)
x.copyto(xs).axpy(1.0, pert)
del tt2, noise, pert


def f_(out, xx):
    J0.mul(t, xx, 1.0, 0.0)                       # t = A x (A = base values)
    out.mul_(t, t).rmul(cmod).axpy(1.0, t).axpy(-1.0, b)


# two operators share the pattern: J0 holds A (for f!), J holds the Jacobian values diag(1+2ct) A
J0 = J
Jac = L.CSCMatrix(ctx, m, n, *(lambda ip, idx: (ip, idx))(*__import__("oracle.synth_ref", fromlist=["x"]).csc_pattern(m, n, k, 20240609)))


def g_(JJ, xx):
    J0.mul(t, xx, 1.0, 0.0)
    check(lib().lso_synth_csc_jacobian(JJ.handle, aval.ptr, t.ptr, cmod), ctx.handle)


nls = L.LeastSquaresProblem(x=x, y=L.DeviceVector(ctx, m), f_=f_, g_=g_, J=Jac, device_callbacks=True, ctx=ctx)
anls = L.allocate(nls, L.LevenbergMarquardt(L.LSMR()))
run = L.LMRun(anls)
its = []
for s in range(steps):
    ctx.sync(); a = time.perf_counter()
    run.iterate()
    ctx.sync(); dt = time.perf_counter() - a
    its.append({"step_ms": dt * 1e3, "lsmr_iters": anls.solver.last_iters, "istop": anls.solver.last_istop, "ssr": run.ssr})
    print(json.dumps(its[-1]), flush=True)

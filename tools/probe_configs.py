"""Single-GPU timings of the other BASELINE.json configs (parity for these shapes is covered at reduced size in tests/):
  C4 shard : dense J 250 000 x 4 000 (one of the 8 row shards of 2M x 4k), LevenbergMarquardt(Cholesky())
  C5       : bounded fit n = 10 000, m = 200 000, Dogleg(QR())
Prints one JSON line per config with per-step times."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
import lsob200 as L

ctx = L.Context.default(0)
which = sys.argv[1:] or ["c4", "c5"]


def run(name, m, n, optimizer, steps, bounds=False, seed=20240611):
    prob = bench.DeviceProblem(L, ctx, m, n, 0, seed)
    x = L.DeviceVector(ctx, n).copyto(prob.x0)
    nls = L.LeastSquaresProblem(x=x, y=L.DeviceVector(ctx, m), f_=prob.f_, g_=prob.g_, J=L.DenseMatrix(ctx, m, n),
                                device_callbacks=True, ctx=ctx)
    kw = {}
    if bounds:
        xs, x0 = prob.xstar.download(), prob.x0.download()
        lo, hi = np.full(n, -np.inf), np.full(n, np.inf)
        idx = np.arange(n) % 5 == 0                       # 20 % of the coordinates are boxed around x*
        lo[idx] = np.minimum(xs[idx] - 0.05, x0[idx])
        hi[idx] = np.maximum(xs[idx] + 0.05, x0[idx])
        kw = dict(lower=lo, upper=hi)
    ctx.sync()
    t0 = time.perf_counter()
    r = L.optimize_(L.allocate(nls, optimizer), iterations=steps, **kw)
    ctx.sync()
    dt = time.perf_counter() - t0
    print(json.dumps({"config": name, "m": m, "n": n, "optimizer": r.optimizer, "iterations": r.iterations,
                      "s_per_iteration_incl_setup": dt / max(r.iterations, 1), "ssr": r.ssr, "converged": r.converged,
                      "f_calls": r.f_calls, "g_calls": r.g_calls}), flush=True)


if "c4" in which:
    run("C4 shard (1/8 of 2M x 4k), LM(Cholesky)", 250_000, 4_000, L.LevenbergMarquardt(L.Cholesky()), 4)
if "c5" in which:
    run("C5 bounded n=10k m=200k, Dogleg(QR)", 200_000, 10_000, L.Dogleg(L.QR()), 3, bounds=True)

"""Sparse products at the config-3 size (5M x 500k, 200 entries per column) on the uniform random pattern and on banded
patterns of decreasing window: shows what bounds the kernels — L2 sectors of the gathers on the random pattern, HBM once
the gathers have locality.  Prints JSON lines (GB/s algorithmic: 12 B per entry + 8 B per vector element)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib

m, n, k = 5_000_000, 500_000, 200
ctx = L.Context.default(0)
nnz = n * k
aval = L.DeviceVector(ctx, nnz)
check(lib().lso_synth_vector(ctx.handle, nnz, 0, 99, 1.0, aval.ptr), ctx.handle)
x, g = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
f, t = L.DeviceVector(ctx, m), L.DeviceVector(ctx, m)
check(lib().lso_synth_vector(ctx.handle, n, 0, 7, 1.0, x.ptr), ctx.handle)
check(lib().lso_synth_vector(ctx.handle, m, 0, 8, 1.0, f.ptr), ctx.handle)


def timeit(fn, reps=7):
    fn(); ctx.sync()
    ts = []
    for _ in range(reps):
        ctx.sync(); a = time.perf_counter(); fn(); ctx.sync(); ts.append(time.perf_counter() - a)
    return min(ts)


for window in (0, 1_000_000, 100_000, 20_000, 4_000):
    colptr = np.zeros(n + 1, dtype=np.int64)
    rowval = np.zeros(nnz, dtype=np.int64)
    if window == 0:
        check(lib().lso_synth_csc_pattern(m, n, k, 20240609, colptr.ctypes.data, rowval.ctypes.data))
    else:
        check(lib().lso_synth_csc_pattern_banded(m, n, k, window, 20240609, colptr.ctypes.data, rowval.ctypes.data))
    J = L.CSCMatrix(ctx, m, n, colptr - 1, rowval - 1)
    check(lib().lso_csc_set_values_dev(J.handle, aval.ptr), ctx.handle)
    tn = timeit(lambda: J.mul(t, x, 1.0, 0.0))
    tt = timeit(lambda: J.mul_t(g, f, 1.0, 0.0))
    bytes_ = 12.0 * nnz + 8.0 * (m + n)
    print(json.dumps({"pattern": "uniform random rows" if window == 0 else f"banded, window {window} rows",
                      "J_v_ms": tn * 1e3, "J_v_GBs": bytes_ / tn / 1e9, "Jt_u_ms": tt * 1e3, "Jt_u_GBs": bytes_ / tt / 1e9,
                      "frac_of_6551": [bytes_ / tn / 1e9 / 6551, bytes_ / tt / 1e9 / 6551]}), flush=True)
    del J

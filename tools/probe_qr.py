"""GPU probe: peaks + timing of the dense solves at a given shape (scratch tool, prints JSON lines)."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib

ctx = L.Context.default(0)
out = C.c_double()
check(lib().lso_bench_fp64_mma_peak(ctx.handle, 20000, C.byref(out)), ctx.handle); dmma = out.value
check(lib().lso_bench_fp64_fma_peak(ctx.handle, 20000, C.byref(out)), ctx.handle); dfma = out.value
check(lib().lso_bench_hbm_copy(ctx.handle, 2 << 30, 5, C.byref(out)), ctx.handle); hbm = out.value
pat = {}
for mode in (1, 2, 3):
    check(lib().lso_bench_fp64_mma_pattern(ctx.handle, 20000, mode, C.byref(out)), ctx.handle); pat[f"dmma_pattern_{8 * 2 ** (mode - 1)}warps"] = out.value
for mode in (12, 14, 18):
    check(lib().lso_bench_fp64_mma_pattern(ctx.handle, 20000, mode, C.byref(out)), ctx.handle); pat[f"dmma_8warps_depdist{mode - 10}"] = out.value
print(json.dumps({"dmma_tflops": dmma, "dfma_tflops": dfma, "hbm_copy_gbs": hbm, **pat}), flush=True)

import torch
def timeit(fn, reps=5):
    fn(); ctx.sync()
    ts = []
    for _ in range(reps):
        ctx.sync(); t0 = time.perf_counter(); fn(); ctx.sync(); ts.append(time.perf_counter() - t0)
    return min(ts), float(np.median(ts))

shapes = [(100000, 1000)] if len(sys.argv) < 3 else [(int(sys.argv[1]), int(sys.argv[2]))]
for (m, n) in shapes:
    A = L.DenseMatrix(ctx, m, n)
    check(lib().lso_synth_dense_matrix(ctx.handle, m, n, 0, 20240608, A.ptr, A.ld), ctx.handle)
    y = L.DeviceVector(ctx, m); check(lib().lso_synth_vector(ctx.handle, m, 0, 77, 1.0, y.ptr), ctx.handle)
    dtd, g, x, fp = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n), L.DeviceVector(ctx, n), L.DeviceVector(ctx, m)
    A.colsumabs2(dtd)
    L.api._lm_damping(ctx, dtd, 0.1)
    res = {"m": m, "n": n}
    for kind in ("qr", "chol"):
        ws = (L.DenseQRAllocatedSolver if kind == "qr" else L.DenseCholeskyAllocatedSolver)(ctx, m, n, True)
        ctx.launch_count(True)
        ws.ldiv(x, A, y, dtd); res[kind + "_launches"] = ctx.launch_count(True)
        res[kind + "_ms"] = [1e3 * t for t in timeit(lambda: ws.ldiv(x, A, y, dtd))]
        res[kind + "_x_norm"] = float(np.linalg.norm(x.download()))
        del ws
    ws = L.DenseQRAllocatedSolver(ctx, m, n, True)
    for la in (0, 1):
        ctx.set_option("qr_lookahead", la)
        res[f"qr_lookahead{la}_ms"] = [1e3 * t for t in timeit(lambda: ws.ldiv(x, A, y, dtd))]
    del ws
    flops_qr = 2 * (m + n) * n * n - 2 * n ** 3 / 3
    res["qr_tflops"] = flops_qr / (res["qr_ms"][0] * 1e-3) / 1e12
    res["syrk_tflops_equiv"] = m * n * (n + 1) / (res["chol_ms"][0] * 1e-3) / 1e12
    res["colsumabs2_ms"] = [1e3 * t for t in timeit(lambda: A.colsumabs2(dtd))]
    res["colsumabs2_gbs"] = 8 * m * n / (res["colsumabs2_ms"][0] * 1e-3) / 1e9
    res["fused_cs_grad_ms"] = [1e3 * t for t in timeit(lambda: A.colsumabs2_and_grad(dtd, g, y))]
    res["pred_ssr_ms"] = [1e3 * t for t in timeit(lambda: A.predicted_ssr(x, y, fp))]
    res["pred_ssr_gbs"] = 8 * m * n / (res["pred_ssr_ms"][0] * 1e-3) / 1e9
    print(json.dumps(res), flush=True)

"""Parity at the NAMED sizes of BASELINE.json configs[2..4] (the bench shape configs[1] is in test_gpu_solvers.py):

  C3  sparse CSC J 5 000 000 x 500 000, 200 entries per column (~20 per row, nnz = 1e8): SpMV / SpMᵀV / colsumabs2 against
      scipy, one damped LSMR solve (btol = 0.5, iterative_lsmr.jl:238-259) against the oracle: same iterations / istop
  C4  one row shard 250 000 x 4 000 of the 2M x 4k LM(Cholesky) problem against the oracle (dsyrk + dpotrf + dpotrs),
      and the 8-shard algorithm (packed partial products summed in rank order) emulated on one device at 8 x 31 250 rows
  C5  the first Gauss-Newton solve of Dogleg(QR()) at 200 000 x 10 000 against the frozen dgelsy result
      (tests/golden/c5_first_solve.npz, made by oracle/make_golden_c5.py)

Inputs come from the counter-based generators (bit-identical on host and device, tests/test_gpu_synth.py)."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import reference_port as O
from oracle import synth_ref as S

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def test_c3_sparse_products_and_lsmr_at_named_size(ctx):
    import lsob200 as L
    from lsob200._lib import check, lib
    m, n, k = 5_000_000, 500_000, 200
    nnz = n * k
    colptr = np.zeros(n + 1, dtype=np.int64)
    rowval = np.zeros(nnz, dtype=np.int64)
    check(lib().lso_synth_csc_pattern(m, n, k, 20240609, colptr.ctypes.data, rowval.ctypes.data))
    J = L.CSCMatrix(ctx, m, n, colptr - 1, rowval - 1)
    vals = S.vector(nnz, 99)
    J.set_values(vals)
    A = sp.csc_matrix((vals, (rowval - 1).astype(np.int32), (colptr - 1).astype(np.int32)), shape=(m, n))
    del rowval, colptr
    xh, yh = S.vector(n, 7), S.vector(m, 8)
    x, y, g, dtd = L.DeviceVector(ctx, n, xh), L.DeviceVector(ctx, m, yh), L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
    t = L.DeviceVector(ctx, m)
    J.mul(t, x, 1.0, 0.0)
    assert rel(t.download(), A @ xh) <= 1e-13
    J.mul_t(g, y, 1.0, 0.0)
    assert rel(g.download(), A.T @ yh) <= 1e-13
    J.colsumabs2_and_grad(dtd, g, y)
    csq = np.asarray(A.multiply(A).sum(axis=0)).ravel()
    assert rel(dtd.download(), csq) <= 1e-13 and rel(g.download(), A.T @ yh) <= 1e-13
    # damped solve as LM issues it: damp = clamp(colsumabs2, 1e-6 mean, 1e32 mean) / Δ, Δ = 10 (LM:84-86, :42)
    damp = np.clip(csq, 1e-6 * csq.mean(), 1e32 * csq.mean()) * (1.0 / 10.0)
    xr, nmul_r, it_r, istop_r = O.lsmr_ldiv(A, yh, damp.copy())
    ws = L.LSMRDampenedAllocatedSolver(ctx, m, n)
    d, dx = L.DeviceVector(ctx, n, damp), L.DeviceVector(ctx, n)
    _, nmul = ws.ldiv(dx, J, y, d)
    assert (ws.last_iters, ws.last_istop, nmul) == (it_r, istop_r, nmul_r)
    assert rel(dx.download(), xr) <= 2e-5
    launches, syncs = ws.stats()
    assert syncs <= ws.last_iters + 1
    # tight run: converged LSMR agrees with the oracle's converged LSMR to 1e-10
    xt, _, it_t, _ = O.lsmr_ldiv(A, yh, damp.copy(), atol=1e-14, btol=1e-14, conlim=0.0)
    d.upload(damp)
    iters, istop = C.c_int64(), C.c_int()
    check(lib().lso_lsmr_solve(ws._h, J.handle, None, 0, y.ptr, d.ptr, dx.ptr, 1e-14, 1e-14, 0.0, 0, C.byref(iters),
                               C.byref(istop)), ctx.handle)
    assert abs(iters.value - it_t) <= 2
    assert rel(dx.download(), xt) <= 1e-10


def _device_problem(ctx, m, n, seed, row0=0):
    import bench
    import lsob200 as L
    prob = bench.DeviceProblem(L, ctx, m, n, row0, seed)
    x = L.DeviceVector(ctx, n).copyto(prob.x0)
    f, J = L.DeviceVector(ctx, m), L.DenseMatrix(ctx, m, n)
    prob.f_(f, x)
    prob.g_(J, x)
    return prob, x, f, J


def test_c4_cholesky_shard_and_emulated_8_shards_at_named_size(ctx):
    """dense_cholesky.jl:43-59 on one 250 000 x 4 000 row shard of configs[3] against the oracle; then the same rows cut
    into 8 shards and solved by the sharded algorithm (emulated on one device): identical δ to 1e-10."""
    import lsob200 as L
    from lsob200._lib import check, lib
    m, n = 250_000, 4_000
    prob, x, f, J = _device_problem(ctx, m, n, 20240607 + 4)
    Jh, fh = J.download(), f.download()
    dtd = np.einsum("ij,ij->j", Jh, Jh)
    damp = np.clip(dtd, 1e-6 * dtd.mean(), 1e32 * dtd.mean()) * (1.0 / 10.0)
    xr = O.chol_ldiv(Jh, fh, damp.copy())
    del Jh
    d, dx = L.DeviceVector(ctx, n, damp), L.DeviceVector(ctx, n)
    ws = L.DenseCholeskyAllocatedSolver(ctx, m, n, damped=True)
    ws.ldiv(dx, J, f, d)
    e1 = rel(dx.download(), xr)
    assert e1 <= 1e-10, e1
    x1 = dx.download().copy()
    ws8 = L.DenseCholeskyAllocatedSolver(ctx, m // 8, n, damped=True)
    check(lib().lso_debug_chol_solve_emulated_shards(ws8._h, 8, J.ptr, J.ld, f.ptr, d.ptr, dx.ptr), ctx.handle)
    assert rel(dx.download(), xr) <= 1e-10
    assert rel(dx.download(), x1) <= 1e-11


def test_c5_first_dogleg_qr_solve_at_named_size(ctx):
    """dense_qr.jl:30-42 at 200 000 x 10 000 against the frozen LAPACK dgelsy result."""
    import lsob200 as L
    path = os.path.join(G, "c5_first_solve.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/c5_first_solve.npz not generated (oracle/make_golden_c5.py, ~1 h of CPU)")
    gold = np.load(path)
    m, n = int(gold["m"]), int(gold["n"])
    prob, x, f, J = _device_problem(ctx, m, n, int(gold["seed"]))
    assert abs(np.sqrt(f.sumabs2()) - float(gold["f_norm"])) <= 1e-12 * float(gold["f_norm"])    # same inputs as the oracle's
    ws = L.DenseQRAllocatedSolver(ctx, m, n, damped=False)
    dx = L.DeviceVector(ctx, n)
    ws.ldiv(dx, J, f)
    assert ws.last_rank == int(gold["rank"]) == n
    e = rel(dx.download(), gold["delta"])
    assert e <= 1e-10, e

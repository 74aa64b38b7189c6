"""The synthetic workloads of bench.py: GPU generators are bit-identical to the numpy replica, and the LM(QR) /
LM(Cholesky) runs on the synthetic dense model match the oracle (iteration counts, minimizer, per-solve δ)."""
import ctypes as C

import numpy as np
import pytest

from oracle import reference_port as O
from oracle import synth_ref as S

pytestmark = pytest.mark.gpu


def _device_problem(ctx, m, n, seed):
    import bench
    import lsob200 as L
    return bench.DeviceProblem(L, ctx, m, n, 0, seed)


@pytest.mark.parametrize("m,n,row0", [(1000, 7, 0), (4097, 33, 12345)])
def test_generators_bit_identical(ctx, m, n, row0):
    import lsob200 as L
    from lsob200._lib import check, lib
    A = L.DenseMatrix(ctx, m, n)
    check(lib().lso_synth_dense_matrix(ctx.handle, m, n, row0, 99, A.ptr, A.ld), ctx.handle)
    assert np.array_equal(A.download(), S.dense_matrix(m, n, 99, row_offset=row0))
    v = L.DeviceVector(ctx, m)
    check(lib().lso_synth_vector(ctx.handle, m, row0, 5, 0.25, v.ptr), ctx.handle)
    assert np.array_equal(v.download(), S.vector(m, 5, 0.25, offset=row0))
    colptr = np.zeros(n + 1, dtype=np.int64)
    rowval = np.zeros(n * 5, dtype=np.int64)
    check(lib().lso_synth_csc_pattern(m, n, 5, 7, colptr.ctypes.data, rowval.ctypes.data))
    ip, idx = S.csc_pattern(m, n, 5, 7)
    assert np.array_equal(colptr - 1, ip) and np.array_equal(rowval - 1, idx)
    assert np.all(np.diff(rowval.reshape(n, 5), axis=1) > 0)


@pytest.mark.parametrize("solver", ["qr", "cholesky"])
def test_lm_on_synthetic_model_matches_oracle(ctx, solver):
    """Reduced-size instance of BASELINE.json configs[1]/[3]: same (x0, f!, g!) on both sides."""
    import lsob200 as L
    m, n, seed = 6000, 96, 20240609
    prob = _device_problem(ctx, m, n, seed)
    model = S.DenseModel(m, n, seed, c=0.1, noise=1e-3)
    assert np.array_equal(prob.A.download(), model.A)
    assert np.linalg.norm(prob.b.download() - model.b) <= 1e-13 * np.linalg.norm(model.b)
    x = L.DeviceVector(ctx, n).copyto(prob.x0)
    nls = L.LeastSquaresProblem(x=x, y=L.DeviceVector(ctx, m), f_=prob.f_, g_=prob.g_, J=L.DenseMatrix(ctx, m, n),
                                device_callbacks=True, ctx=ctx)
    sol = {"qr": L.QR, "cholesky": L.Cholesky}[solver]
    rg = L.optimize_(nls, L.LevenbergMarquardt(sol()), record_steps=True)
    ro = O.levenberg_marquardt(model.f, model.g, model.x0, np.zeros((m, n), order="F"), m, solver=solver, record=True)
    assert rg.converged and ro.converged
    assert rg.iterations == ro.iterations and (rg.f_calls, rg.g_calls) == (ro.f_calls, ro.g_calls)
    xg = rg.minimizer.download()
    assert np.linalg.norm(xg - ro.minimizer) <= 1e-9 * np.linalg.norm(ro.minimizer)
    assert abs(rg.ssr - ro.ssr) <= 1e-9 * ro.ssr
    for dg, do in list(zip(rg.deltas, ro.deltas))[:3]:
        assert np.linalg.norm(dg - do) <= 1e-10 * np.linalg.norm(do)


def test_host_step_matches_device_path(ctx):
    """e2e path (HostStep: J, f from pinned host buffers) gives the same δ as the resident path."""
    import lsob200 as L
    from lsob200._lib import check, lib
    m, n = 3000, 64
    prob = _device_problem(ctx, m, n, 77)
    x = L.DeviceVector(ctx, n).copyto(prob.x0)
    nls = L.LeastSquaresProblem(x=x, y=L.DeviceVector(ctx, m), f_=prob.f_, g_=prob.g_, J=L.DenseMatrix(ctx, m, n),
                                device_callbacks=True, ctx=ctx)
    anls = L.allocate(nls, L.LevenbergMarquardt(L.QR()))
    prob.f_(anls.fcur, x)
    prob.g_(anls.J, x)
    Jh, fh = anls.J.download(), anls.fcur.download()
    dtd = np.einsum("ij,ij->j", Jh, Jh)
    damp = np.clip(dtd, 1e-6 * dtd.mean(), 1e32 * dtd.mean()) * (1 / 10.0)
    dref, _ = O.qr_ldiv(Jh, fh, damp)
    dx = np.zeros(n)
    sc = L.HostStep(anls).run(Jh.ctypes.data, fh.ctypes.data, 10.0, dx)
    assert np.linalg.norm(dx - dref) <= 1e-10 * np.linalg.norm(dref)
    r = Jh @ dref - fh
    assert abs(sc["predicted_ssr"] - r @ r) <= 1e-10 * (r @ r)
    assert abs(sc["ssr"] - fh @ fh) <= 1e-12 * (fh @ fh)

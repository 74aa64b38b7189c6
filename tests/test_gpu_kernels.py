"""GPU parity: vector duck-type kernels, dense J passes and CSC products vs numpy on seeded inputs (fp64,
tolerance 1e-13 relative: summation order differs, arithmetic does not)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
RTOL = 1e-13


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(a - b)
    return d / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("n", [1, 7, 1000, 100003, 2_000_000])
def test_vector_family(ctx, n):
    import lsob200
    from lsob200 import DeviceVector, wdot, wnorm
    rng = np.random.default_rng(n)
    xh, yh, wh = rng.standard_normal(n), rng.standard_normal(n), rng.uniform(0.1, 2, n)
    x, y, w = DeviceVector(ctx, n, xh), DeviceVector(ctx, n, yh), DeviceVector(ctx, n, wh)
    assert abs(x.sumabs2() - np.sum(xh * xh)) <= RTOL * np.sum(xh * xh)
    assert abs(x.sum() - xh.sum()) <= 1e-12 * np.abs(xh).sum()
    assert abs(x.norm() - np.linalg.norm(xh)) <= RTOL * np.linalg.norm(xh)
    assert x.maxabs() == np.abs(xh).max()
    assert abs(x.dot(y) - xh @ yh) <= 1e-12 * np.abs(xh * yh).sum()
    assert abs(wdot(x, y, w) - np.sum(wh * xh * yh)) <= 1e-12 * np.abs(wh * xh * yh).sum()
    assert abs(wnorm(x, w) - np.sqrt(np.sum(wh * xh * xh))) <= RTOL * wnorm(x, w)
    z = x.similar().copyto(x)
    z.axpy(-1.5, y)
    assert np.array_equal(z.download(), xh + (-1.5) * yh) or rel(z.download(), xh - 1.5 * yh) < 1e-15
    z.rmul(0.25)
    assert rel(z.download(), 0.25 * (xh - 1.5 * yh)) < 1e-15
    z.div_(x, w)
    assert np.array_equal(z.download(), xh / wh)
    z.mul_(x, w)
    assert np.array_equal(z.download(), xh * wh)
    z.copyto(w).sqrt_()
    assert np.array_equal(z.download(), np.sqrt(wh))
    z.copyto(x).clamp(-0.5, 0.25)
    assert np.array_equal(z.download(), np.clip(xh, -0.5, 0.25))
    z.fill(3.0)
    assert np.all(z.download() == 3.0)


def test_nan_semantics(ctx):
    from lsob200 import DeviceVector, IsFiniteException
    xh = np.array([1.0, np.nan, -3.0, np.inf])
    x = DeviceVector(ctx, 4, xh)
    assert np.isnan(x.maxabs())           # Julia maximum(abs, x) propagates NaN
    with pytest.raises(IsFiniteException):
        x.check_finite()
    DeviceVector(ctx, 3, [1.0, 2.0, 3.0]).check_finite()


def test_box_and_projected_gradient(ctx):
    from lsob200 import DeviceVector
    from lsob200.api import _box_project, _maxabs_projected_gradient
    from oracle import reference_port as O
    rng = np.random.default_rng(5)
    n = 1001
    xh = rng.uniform(-1, 1, n)
    lo, hi = xh - rng.uniform(0, 0.3, n), xh + rng.uniform(0, 0.3, n)
    lo[::7] = xh[::7]
    hi[::5] = xh[::5]
    dh, gh = rng.standard_normal(n), rng.standard_normal(n)
    x, d, g = DeviceVector(ctx, n, xh), DeviceVector(ctx, n, dh), DeviceVector(ctx, n, gh)
    dlo, dhi = DeviceVector(ctx, n, lo), DeviceVector(ctx, n, hi)
    _box_project(ctx, d, x, dlo, dhi)
    assert np.array_equal(d.download(), np.maximum(np.minimum(dh, xh - lo), xh - hi))
    assert _maxabs_projected_gradient(ctx, g, x, dlo, dhi) == O.maxabs_projected_gradient(gh, xh, lo, hi)
    assert _maxabs_projected_gradient(ctx, g, x, None, None) == np.abs(gh).max()


def test_lm_damping(ctx):
    from lsob200 import DeviceVector
    from lsob200.api import _lm_damping
    rng = np.random.default_rng(11)
    dtd = rng.uniform(0, 1, 777) ** 8
    dtd[3] = 0.0
    v = DeviceVector(ctx, 777, dtd)
    _lm_damping(ctx, v, 1 / 7.0)
    mean = dtd.sum() / 777
    ref = np.clip(dtd, 1e-6 * mean, 1e32 * mean) * (1 / 7.0)
    assert rel(v.download(), ref) < 1e-14


@pytest.mark.parametrize("m,n", [(1, 1), (5, 3), (257, 33), (1000, 1000), (30011, 129), (200000, 40)])
def test_dense_passes(ctx, m, n):
    from lsob200 import DenseMatrix, DeviceVector
    rng = np.random.default_rng(m * 31 + n)
    Jh = np.asfortranarray(rng.standard_normal((m, n)) * np.exp2(rng.integers(-6, 7, n)))
    fh, dh = rng.standard_normal(m), rng.standard_normal(n)
    J, f, d = DenseMatrix(ctx, m, n, Jh), DeviceVector(ctx, m, fh), DeviceVector(ctx, n, dh)
    assert np.array_equal(J.download(), Jh)
    out, g, y = DeviceVector(ctx, n), DeviceVector(ctx, n), DeviceVector(ctx, m)
    J.colsumabs2(out)
    assert rel(out.download(), np.einsum("ij,ij->j", Jh, Jh)) < RTOL
    J.mul_t(g, f)
    assert rel(g.download(), Jh.T @ fh) < 1e-12
    J.colsumabs2_and_grad(out, g, f)
    assert rel(out.download(), np.einsum("ij,ij->j", Jh, Jh)) < RTOL
    assert rel(g.download(), Jh.T @ fh) < 1e-12
    J.mul(y, d)
    assert rel(y.download(), Jh @ dh) < 1e-12
    # alpha / beta forms (mul! 5-arg, lsmr.jl:118,122)
    y.upload(fh)
    J.mul(y, d, 2.0, -0.5)
    assert rel(y.download(), 2.0 * (Jh @ dh) - 0.5 * fh) < 1e-12
    g.upload(dh)
    J.mul_t(g, f, -1.0, 3.0)
    assert rel(g.download(), -(Jh.T @ fh) + 3.0 * dh) < 1e-12
    fp = DeviceVector(ctx, m)
    ssr = J.predicted_ssr(d, f, fp)
    r = Jh @ dh - fh
    assert rel(fp.download(), r) < 1e-12
    assert abs(ssr - r @ r) <= 1e-12 * (r @ r)
    assert abs(J.predicted_ssr(d, f, None) - ssr) == 0.0


@pytest.mark.parametrize("m,n,density", [(9, 6, 0.4), (500, 60, 0.1), (20000, 3000, 0.004), (3000, 20000, 0.004)])
def test_csc_products(ctx, m, n, density):
    from lsob200 import CSCMatrix, DeviceVector
    rng = np.random.default_rng(m + n)
    A = sp.random(m, n, density=density, random_state=m + n, format="csc")
    A.sort_indices()
    J = CSCMatrix.from_scipy(ctx, A)
    xh, yh = rng.standard_normal(n), rng.standard_normal(m)
    x, y = DeviceVector(ctx, n, xh), DeviceVector(ctx, m, yh)
    out = DeviceVector(ctx, n)
    J.colsumabs2(out)
    assert rel(out.download(), np.asarray(A.multiply(A).sum(axis=0)).ravel()) < RTOL
    J.mul(y, x, 1.0, 0.0)
    assert rel(y.download(), A @ xh) < 1e-13
    y.upload(yh)
    J.mul(y, x, -2.0, 0.5)
    assert rel(y.download(), -2.0 * (A @ xh) + 0.5 * yh) < 1e-13
    y.upload(yh)
    J.mul_t(x, y, 1.0, 0.0)
    assert rel(x.download(), A.T @ yh) < 1e-13
    x.upload(xh)
    J.mul_t(x, y, 3.0, -1.0)
    assert rel(x.download(), 3.0 * (A.T @ yh) - xh) < 1e-13
    # g! rewrote nonzeros(J): values-only refresh reaches both products
    A2 = A.copy()
    A2.data = rng.standard_normal(A.nnz)
    J.set_values(A2.data)
    x.upload(xh)
    J.mul(y, x, 1.0, 0.0)
    assert rel(y.download(), A2 @ xh) < 1e-13
    # determinism: bit-identical on repeat
    y1 = y.download().copy()
    J.mul(y, x, 1.0, 0.0)
    assert np.array_equal(y.download(), y1)


def test_device_central_difference_jacobian(ctx):
    """(f2) `autodiff = :central` (types.jl:54-58) with f! on the device: lso_fd_jacobian_central reproduces FiniteDiff's
    central differences (step cbrt(eps) max(1, |x_j|)) column by column, and an LM(QR) run that uses it converges like
    the run with the analytic g!."""
    import lsob200 as L
    from lsob200._lib import check, lib
    m, n = 500, 12
    rng = np.random.default_rng(8)
    Ah = np.asfortranarray(rng.standard_normal((m, n)))
    A = L.DenseMatrix(ctx, m, n, Ah)
    xs = rng.standard_normal(n)
    th = Ah @ xs
    b = L.DeviceVector(ctx, m, th + 0.1 * th * th)
    t = L.DeviceVector(ctx, m)

    def f_(out, x):          # r = t + c t^2 - b, t = A x   (device kernels only)
        check(lib().lso_synth_residual(ctx.handle, m, n, A.ptr, A.ld, x.ptr, b.ptr, 0.1, t.ptr, out.ptr), ctx.handle)

    x0 = xs + 0.05 * rng.standard_normal(n)
    x = L.DeviceVector(ctx, n, x0)
    J = L.DenseMatrix(ctx, m, n)
    nls = L.LeastSquaresProblem(x=x, y=L.DeviceVector(ctx, m), f_=f_, g_=None, J=J, device_callbacks=True, ctx=ctx)
    nls.g_(J, x)
    assert np.array_equal(x.download(), x0)                      # x restored
    Jfd = J.download()
    t0 = Ah @ x0
    Jan = (1.0 + 0.2 * t0)[:, None] * Ah
    assert np.abs(Jfd - Jan).max() <= 1e-8 * np.abs(Jan).max()
    # reference recipe on the host, same steps
    h = np.cbrt(np.finfo(float).eps) * np.maximum(1.0, np.abs(x0))
    col = 3
    xp, xm = x0.copy(), x0.copy()
    xp[col] += h[col]; xm[col] -= h[col]
    fp = (Ah @ xp) + 0.1 * (Ah @ xp) ** 2
    fm = (Ah @ xm) + 0.1 * (Ah @ xm) ** 2
    assert np.abs(Jfd[:, col] - (fp - fm) / (2 * h[col])).max() <= 1e-7 * np.abs(Jan[:, col]).max()
    r = L.optimize_(nls, L.LevenbergMarquardt(L.QR()))
    assert r.converged and np.linalg.norm(r.minimizer.download() - xs) <= 1e-6

"""GPU parity of the three solver plugins (`ldiv!`) against the oracle on identical (J, y, damp):
||δ_gpu − δ_ref|| / ||δ_ref|| <= 1e-10 per linear solve (BASELINE.json north_star)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import reference_port as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def make_J(m, n, seed, scaled=True):
    rng = np.random.default_rng(seed)
    J = rng.standard_normal((m, n))
    if scaled:
        J = J * np.exp2(rng.integers(-6, 7, n))
    return np.asfortranarray(J), rng.standard_normal(m), rng


SHAPES = [(2, 2), (9, 6), (40, 40), (100, 33), (300, 64), (777, 65), (3000, 100), (5000, 257), (20000, 96),
          (2304, 32), (70000, 40)]


@pytest.mark.parametrize("apply_kernel", [1, 0])
@pytest.mark.parametrize("m,n", SHAPES)
def test_qr_damped(ctx, m, n, apply_kernel):
    """ldiv!(x, J, y, damp, A::DenseQRAllocatedSolver) — dense_qr.jl:56-88"""
    from lsob200 import DenseMatrix, DenseQRAllocatedSolver, DeviceVector
    if apply_kernel == 0 and m * n > 400000:
        pytest.skip("plain-FMA cross-check kernel: small shapes only")
    ctx.set_option("qr_apply", apply_kernel)
    try:
        Jh, yh, rng = make_J(m, n, m * 7 + n)
        dtd = np.einsum("ij,ij->j", Jh, Jh)
        damp = np.clip(dtd, 1e-6 * dtd.mean(), 1e32 * dtd.mean()) / 10.0
        ws = DenseQRAllocatedSolver(ctx, m, n, damped=True)
        J, y, d, x = DenseMatrix(ctx, m, n, Jh), DeviceVector(ctx, m, yh), DeviceVector(ctx, n, damp), DeviceVector(ctx, n)
        for rep in range(2):      # workspace reuse: second call must give the same bits
            _, nmul = ws.ldiv(x, J, y, d)
            assert nmul == 1
            xg = x.download()
            if rep == 0:
                x0 = xg.copy()
        assert np.array_equal(xg, x0)
        xr, rank = O.qr_ldiv(Jh, yh, damp)
        assert rank == n and ws.last_rank == n
        assert rel(xg, xr) <= TOL, rel(xg, xr)
        assert np.array_equal(J.download(), Jh) and np.array_equal(y.download(), yh)   # J, y untouched
        # R factor: |R| matches LAPACK's unpivoted R of the augmented matrix up to row signs
        if n <= 100:
            aug = np.vstack([Jh, np.diag(np.sqrt(damp))])
            Rref = np.linalg.qr(aug, mode="r")
            Rg = ws.factor()
            assert rel(np.abs(Rg), np.abs(Rref)) <= 1e-11
    finally:
        ctx.set_option("qr_apply", 1)


@pytest.mark.parametrize("m,n", [(2, 2), (9, 6), (40, 40), (300, 64), (5000, 257), (20000, 96)])
def test_qr_undamped_full_rank(ctx, m, n):
    """ldiv!(x, J, y, A::DenseQRAllocatedSolver) — dense_qr.jl:30-42, full column rank"""
    from lsob200 import DenseMatrix, DenseQRAllocatedSolver, DeviceVector
    Jh, yh, rng = make_J(m, n, m * 3 + n + 1)
    ws = DenseQRAllocatedSolver(ctx, m, n, damped=False)
    J, y, x = DenseMatrix(ctx, m, n, Jh), DeviceVector(ctx, m, yh), DeviceVector(ctx, n)
    ws.ldiv(x, J, y)
    xr, rank = O.qr_ldiv(Jh, yh)
    assert rank == n
    assert rel(x.download(), xr) <= 1e-9 if m == n else rel(x.download(), xr) <= TOL


def test_qr_host_entry_point(ctx):
    """lso_qr_solve_host: what `ldiv!` on plain host Arrays binds (J with ld > m, host pointers)."""
    import ctypes as C
    from lsob200 import DenseQRAllocatedSolver
    from lsob200._lib import check, lib
    m, n, ld = 1234, 57, 1300
    Jh, yh, rng = make_J(m, n, 99)
    big = np.zeros((ld, n), order="F")
    big[:m] = Jh
    damp = rng.uniform(0.1, 1, n)
    ws = DenseQRAllocatedSolver(ctx, m, n, damped=True)
    x = np.zeros(n)
    rank = C.c_int()
    check(lib().lso_qr_solve_host(ws._h, big.ctypes.data, ld, yh.ctypes.data, damp.ctypes.data, x.ctypes.data,
                                  C.byref(rank)), ctx.handle)
    assert rel(x, O.qr_ldiv(Jh, yh, damp)[0]) <= TOL


@pytest.mark.parametrize("syrk_kernel", [1, 0])
@pytest.mark.parametrize("m,n", [(2, 2), (9, 6), (40, 40), (300, 64), (777, 129), (5001, 257), (20000, 96), (40000, 520)])
def test_cholesky_damped(ctx, m, n, syrk_kernel):
    """ldiv!(x, J, y, damp, A::DenseCholeskyAllocatedSolver) — dense_cholesky.jl:43-59"""
    from lsob200 import DenseCholeskyAllocatedSolver, DenseMatrix, DeviceVector
    ctx.set_option("syrk", syrk_kernel)
    try:
        Jh, yh, rng = make_J(m, n, m * 5 + n, scaled=False)
        dtd = np.einsum("ij,ij->j", Jh, Jh)
        damp = np.clip(dtd, 1e-6 * dtd.mean(), 1e32 * dtd.mean()) / 10.0
        ws = DenseCholeskyAllocatedSolver(ctx, m, n, damped=True)
        J, y, d, x = DenseMatrix(ctx, m, n, Jh), DeviceVector(ctx, m, yh), DeviceVector(ctx, n, damp), DeviceVector(ctx, n)
        ws.ldiv(x, J, y, d)
        xr = O.chol_ldiv(Jh, yh, damp.copy())
        assert rel(x.download(), xr) <= TOL, rel(x.download(), xr)
        if n <= 129:
            Rref = np.linalg.cholesky(Jh.T @ Jh + np.diag(damp)).T
            assert rel(ws.factor(), Rref) <= 1e-11
    finally:
        ctx.set_option("syrk", 1)


def test_cholesky_undamped_and_failure(ctx):
    """dense_cholesky.jl:29-35; not-PD / rank-deficient inputs surface as exceptions (info > 0), never silently."""
    from lsob200 import (DenseCholeskyAllocatedSolver, DenseMatrix, DeviceVector, PosDefException,
                         RankDeficientException)
    Jh, yh, rng = make_J(500, 20, 1, scaled=False)
    ws = DenseCholeskyAllocatedSolver(ctx, 500, 20, damped=False)
    J, y, x = DenseMatrix(ctx, 500, 20, Jh), DeviceVector(ctx, 500, yh), DeviceVector(ctx, 20)
    ws.ldiv(x, J, y)
    assert rel(x.download(), O.chol_ldiv(Jh, yh)) <= TOL
    Jh[:, 7] = 0.0
    J.upload(Jh)
    with pytest.raises(RankDeficientException):
        ws.ldiv(x, J, y)
    ws2 = DenseCholeskyAllocatedSolver(ctx, 500, 20, damped=True)
    d = DeviceVector(ctx, 20, -1e9 * np.ones(20))
    with pytest.raises(PosDefException):
        ws2.ldiv(x, J, y, d)


@pytest.mark.parametrize("kind", ["csc", "dense"])
@pytest.mark.parametrize("m,n,damped", [(9, 6, True), (9, 6, False), (400, 60, True), (400, 60, False),
                                        (20000, 3000, True), (5000, 300, False)])
def test_lsmr(ctx, kind, m, n, damped):
    """ldiv! for LSMRAllocatedSolver / LSMRDampenedAllocatedSolver (iterative_lsmr.jl:179-198, 238-259):
    same iteration count and istop as the oracle, iterate within 1e-10."""
    from lsob200 import CSCMatrix, DenseMatrix, DeviceVector, LSMRAllocatedSolver, LSMRDampenedAllocatedSolver
    rng = np.random.default_rng(m + n + damped)
    A = sp.random(m, n, density=min(1.0, 12.0 / n + 0.002), random_state=m + n, format="csc")
    A.sort_indices()
    if kind == "dense" and m * n > 3e6:
        pytest.skip("dense operator: small shapes only")
    yh = rng.standard_normal(m)
    damp = np.asarray(A.multiply(A).sum(axis=0)).ravel() / 10 + 1e-3
    Aor = A if kind == "csc" else A.toarray()
    xr, nmul_r, it_r, istop_r = O.lsmr_ldiv(Aor, yh, damp.copy() if damped else None)
    J = CSCMatrix.from_scipy(ctx, A) if kind == "csc" else DenseMatrix(ctx, m, n, A.toarray())
    y, x = DeviceVector(ctx, m, yh), DeviceVector(ctx, n)
    if damped:
        ws = LSMRDampenedAllocatedSolver(ctx, m, n)
        d = DeviceVector(ctx, n, damp)
        _, nmul = ws.ldiv(x, J, y, d)
        assert rel(d.download(), np.sqrt(damp)) < 1e-15     # damp <- sqrt(damp), iterative_lsmr.jl:252
    else:
        ws = LSMRAllocatedSolver(ctx, m, n)
        _, nmul = ws.ldiv(x, J, y)
    assert (ws.last_iters, ws.last_istop) == (it_r, istop_r)
    assert nmul == nmul_r
    assert rel(x.download(), xr) <= TOL
    assert np.array_equal(y.download(), yh)

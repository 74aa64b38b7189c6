"""Pins the oracle (oracle/reference_port.py) against every end-to-end assertion the reference's own test-suite
makes about this path and against independent implementations (scipy LSMR, LAPACK)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import problems as P
from oracle import reference_port as O


def _J(matrix, n, g):
    if matrix == "sparse":
        return P.dense_pattern_csc(n), P.sparse_adapter(g, n)
    return np.zeros((n, n), order="F"), g


@pytest.mark.parametrize("matrix,solver", [("dense", "qr"), ("dense", "lsmr"), ("sparse", "lsmr")])
@pytest.mark.parametrize("opt", ["dogleg", "lm"])
def test_minpack_ssr(matrix, solver, opt):
    """test/nonlinearsolvers.jl:505-537: r.ssr <= 1e-3 for the 21 problems in every combination."""
    for name, f, g, x0 in P.minpack_all():
        J, gg = _J(matrix, x0.size, g)
        r = O.optimize(f, gg, x0, J, x0.size, optimizer=opt, solver=solver)
        assert r.ssr <= 1e-3, (name, r.ssr)


@pytest.mark.parametrize("opt", ["dogleg", "lm"])
def test_minpack_cholesky(opt):
    """test/nonlinearsolvers.jl:585-595"""
    for name, f, g, x0 in P.minpack_cholesky():
        r = O.optimize(f, g, x0, np.zeros((x0.size, x0.size), order="F"), x0.size, optimizer=opt, solver="cholesky")
        assert r.converged and r.ssr <= 1e-3, (name, r.ssr)


@pytest.mark.parametrize("opt", ["dogleg", "lm"])
def test_factor_model(opt):
    """test/nonlinearleastsquares.jl:91-110: rank-deficient J'J, dense QR and sparse LSMR."""
    name, f, g, x0 = P.factor()
    J = np.ones((9, 6), order="F")
    r = O.optimize(f, g, x0, J, 9, optimizer=opt, solver="qr")
    assert r.ssr <= 12 and r.converged
    name, f, gs, x0, pat = P.factor_sparse_pattern()
    r = O.optimize(f, gs, x0, pat.copy(), 9, optimizer=opt, solver="lsmr")
    assert r.ssr <= 12 and r.converged


@pytest.mark.parametrize("opt", ["dogleg", "lm"])
def test_bounds(opt):
    """test/bounds.jl:11-36"""
    for name, f, g, x0, kw, xs in P.bounds_cases():
        r = O.optimize(f, g, x0, np.zeros((2, 2), order="F"), 2, optimizer=opt, solver="qr", **kw)
        assert r.converged, name
        if "active" in name and "inactive" not in name:
            assert r.g_converged, name
        assert np.linalg.norm(r.minimizer - xs) <= 1e-6, name


def test_defaults():
    """test/nonlinearsolvers.jl:619-628: dense -> Dogleg, sparse -> LevenbergMarquardt."""
    name, f, g, x0 = P.wood()
    r = O.optimize(f, g, x0, np.ones((4, 4), order="F"), 4, optimizer=None, solver=None)
    assert r.optimizer == "Dogleg"
    r = O.optimize(f, P.sparse_adapter(g, 4), x0, P.dense_pattern_csc(4), 4, optimizer=None, solver=None)
    assert r.optimizer == "LevenbergMarquardt"
    with pytest.raises(ValueError):
        O.optimize(f, g, x0, P.dense_pattern_csc(4), 4, optimizer="dogleg", solver="qr")


def test_readme_rosenbrock_counts():
    """README.md:13-18 run with analytic g!: iteration / call counts recorded in BASELINE.md §3."""
    name, f, g, x0 = P.readme_rosenbrock()
    r = O.optimize(f, g, x0, np.zeros((2, 2), order="F"), 2, optimizer="dogleg", solver="qr")
    assert (r.iterations, r.f_calls, r.g_calls) == (51, 52, 36) and r.x_converged and r.ssr == 0.0
    r = O.optimize(f, g, x0, np.zeros((2, 2), order="F"), 2, optimizer="lm", solver="qr")
    assert (r.iterations, r.f_calls, r.g_calls) == (56, 57, 51) and r.f_converged and r.ssr < 1e-20


@pytest.mark.parametrize("btol,atol", [(0.5, 1e-6), (1e-6, 1e-6), (1e-14, 1e-14)])
def test_lsmr_matches_scipy(btol, atol):
    """Literal lsmr.jl restatement vs scipy.sparse.linalg.lsmr: same iteration count, same iterate."""
    rng = np.random.default_rng(7)
    A = sp.random(400, 60, density=0.1, random_state=7, format="csc")
    b = rng.standard_normal(400)
    op = O._Precond(A, np.ones(60), None)
    x, it, istop = O.lsmr(np.zeros(60), op, b.copy(), atol=atol, btol=btol, conlim=1e8, maxiter=400)
    ref = spla.lsmr(A, b, atol=atol, btol=btol, conlim=1e8, maxiter=400)
    assert it == ref[2]
    assert np.linalg.norm(x - ref[0]) <= 1e-10 * np.linalg.norm(ref[0])


def test_dense_solvers_agree_with_normal_equations():
    rng = np.random.default_rng(3)
    J = rng.standard_normal((300, 20))
    y = rng.standard_normal(300)
    damp = rng.uniform(0.5, 2.0, 20)
    xq, rank = O.qr_ldiv(J, y, damp)
    xc = O.chol_ldiv(J, y, damp)
    xn = np.linalg.solve(J.T @ J + np.diag(damp), J.T @ y)
    assert rank == 20
    assert np.linalg.norm(xq - xn) <= 1e-12 * np.linalg.norm(xn)
    assert np.linalg.norm(xc - xn) <= 1e-12 * np.linalg.norm(xn)
    # undamped, rank deficient -> minimum-norm solution (pinv)
    J2 = np.hstack([J[:, :10], J[:, :10]])
    x2, rank2 = O.qr_ldiv(J2, y)
    assert rank2 == 10
    assert np.linalg.norm(x2 - np.linalg.pinv(J2) @ y) <= 1e-10 * np.linalg.norm(x2)
    with pytest.raises(O.RankDeficientException):
        O.chol_ldiv(np.hstack([J[:, :3], np.zeros((300, 1))]), y)


# ---- NIST StRD: the per-problem KNOWN ANSWERS the reference's test-suite holds (test/nonlinearfitting.jl) -----------
# (optimizer, dataset, start column) triples whose run does NOT end within 1e-3 of the certified values.  NIST's
# "start 1" columns of MGH09 / MGH10 are far from the solution (classed "higher difficulty" by NIST); the reference
# itself only counts successes there (`println("strd ...")`, :1471) and asserts no NaN (:1468).
NIST_MISSES = {("dogleg", "MGH09", 0), ("dogleg", "MGH10", 0), ("lm", "MGH10", 0)}
NIST_KW = dict(x_tol=1e-50, f_tol=1e-36, g_tol=1e-50)        # test/nonlinearfitting.jl:1466


@pytest.mark.parametrize("opt", ["dogleg", "lm"])
def test_nist_strd_certified_values_pin_the_oracle(opt):
    """The oracle's Dogleg(QR()) / LevenbergMarquardt(QR()) loops land on NIST's certified parameter values
    (norm(minimizer - solution) <= 1e-3, the reference's own criterion at :1467) from the NIST start columns, on all 16
    datasets the reference runs — except the three named hard starts.  This is what anchors the restatement to
    numbers the reference holds rather than to itself."""
    reached, total = 0, 0
    for name, f, g, starts, cert, m in P.nist_strd():
        for j, x0 in enumerate(starts):
            J = np.zeros((m, cert.size), order="F")
            with np.errstate(all="ignore"):
                r = O.optimize(f, g, x0.copy(), J, m, optimizer=opt, solver="qr", **NIST_KW)
            assert not np.isnan(np.mean(r.minimizer)), (name, j)          # the reference's @test (:1468)
            ok = np.linalg.norm(r.minimizer - cert) <= 1e-3
            assert ok == ((opt, name, j) not in NIST_MISSES), (opt, name, j, np.linalg.norm(r.minimizer - cert))
            reached += ok
            total += 1
    assert total == 33 and reached == total - sum(1 for o, _, _ in NIST_MISSES if o == opt)


def test_nist_models_have_exact_jacobians():
    """The reference differentiates `f(x, beta)` with ForwardDiff; the analytic Jacobians of tests/problems.py agree with
    central differences at every start column and at the certified point."""
    for name, f, g, starts, cert, m in P.nist_strd():
        n = cert.size
        for x0 in starts + [cert]:
            J, Jn = np.zeros((m, n)), np.zeros((m, n))
            g(J, x0)
            for k in range(n):
                h = 1e-6 * max(abs(x0[k]), 1e-3)
                xp, xm = x0.copy(), x0.copy()
                xp[k] += h
                xm[k] -= h
                fp, fm = np.zeros(m), np.zeros(m)
                f(fp, xp)
                f(fm, xm)
                Jn[:, k] = (fp - fm) / (2 * h)
            assert np.abs(J - Jn).max() <= 1e-6 * max(np.abs(Jn).max(), 1e-300), name

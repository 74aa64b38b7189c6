"""GPU parity of the sparse operator (stream SpMV / SpMᵀV kernels, fused passes) and of the fused device-resident LSMR
against scipy / the oracle: SparseArrays `mul!` both ways, `colsumabs2!` (utils.jl:146-151), the fused LM passes
(LM:82+102, LM:114-117), the wrappers of iterative_lsmr.jl:12-122 inside the SpMV epilogues, and lsmr.jl:53-238 with its
scalar recurrences on the device."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import reference_port as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def ragged(m, n, seed):
    """Pattern with everything the stream kernel special-cases: empty rows / columns (runs of them), one row and one
    column longer than a CTA slice (2048 entries), odd slice offsets, a dense block, single-entry segments."""
    rng = np.random.default_rng(seed)
    A = sp.random(m, n, density=min(1.0, 6.0 / n), random_state=seed, format="lil")
    A[m // 3, :] = rng.standard_normal(n)               # long row
    A[:, n // 2] = rng.standard_normal((m, 1))          # long column
    A[m // 2: m // 2 + 40, :] = 0.0                     # empty rows
    A[:, 5:25] = 0.0                                    # empty columns
    A[0, 0] = 1.5
    A = A.tocsc()
    A.eliminate_zeros()
    A.sort_indices()
    return A


@pytest.mark.parametrize("spmv", [2, 1, 0])
@pytest.mark.parametrize("m,n", [(9, 6), (333, 70), (5000, 2600), (2600, 5000), (40000, 3000)])
def test_stream_products_on_ragged_patterns(ctx, m, n, spmv):
    from lsob200 import CSCMatrix, DeviceVector
    ctx.set_option("spmv", spmv)
    try:
        A = ragged(m, n, 7 * m + n) if m >= 100 else sp.random(m, n, density=0.5, random_state=3, format="csc")
        A.sort_indices()
        rng = np.random.default_rng(m + 3 * n)
        J = CSCMatrix.from_scipy(ctx, A)
        xh, yh, fh = rng.standard_normal(n), rng.standard_normal(m), rng.standard_normal(m)
        x, y, f = DeviceVector(ctx, n, xh), DeviceVector(ctx, m, yh), DeviceVector(ctx, m, fh)
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
        x.upload(xh)
        # fused LM passes
        dtd, g = DeviceVector(ctx, n), DeviceVector(ctx, n)
        csq = np.asarray(A.multiply(A).sum(axis=0)).ravel()
        for rep in range(2):            # second call takes the cached-colsumabs2 branch
            J.colsumabs2_and_grad(dtd, g, f)
            assert rel(dtd.download(), csq) < 1e-13
            assert rel(g.download(), A.T @ fh) < 1e-13
        J.colsumabs2(dtd)
        assert rel(dtd.download(), csq) < 1e-13
        fp = DeviceVector(ctx, m)
        ssr = J.predicted_ssr(x, f, fp)
        r = A @ xh - fh
        assert rel(fp.download(), r) < 1e-13 and abs(ssr - r @ r) <= 1e-13 * (r @ r)
        assert J.predicted_ssr(x, f, None) == ssr
        # new values invalidate the colsumabs2 cache and the CSR mirror
        A2 = A.copy()
        A2.data = rng.standard_normal(A.nnz)
        J.set_values(A2.data)
        J.colsumabs2_and_grad(dtd, g, f)
        assert rel(dtd.download(), np.asarray(A2.multiply(A2).sum(axis=0)).ravel()) < 1e-13
        J.mul(y, x, 1.0, 0.0)
        assert rel(y.download(), A2 @ xh) < 1e-13
        y1 = y.download().copy()
        J.mul(y, x, 1.0, 0.0)
        assert np.array_equal(y.download(), y1)       # fixed summation order
    finally:
        ctx.set_option("spmv", 1)


def test_pattern_update_and_both_images(ctx):
    """lso_csc_update_pattern (g! that changes the pattern, test/nonlinearsolvers.jl:526-530) and a device g! writing
    the CSC and the CSR image directly (no mirror gather)."""
    from lsob200 import CSCMatrix, DeviceVector
    from lsob200._lib import check, lib
    rng = np.random.default_rng(5)
    m, n = 3000, 700
    A = sp.random(m, n, density=0.01, random_state=1, format="csc"); A.sort_indices()
    B = sp.random(m, n, density=0.03, random_state=2, format="csc"); B.sort_indices()      # more entries: re-allocation
    Cm = sp.random(m, n, density=0.002, random_state=3, format="csc"); Cm.sort_indices()    # fewer: buffers re-used
    J = CSCMatrix.from_scipy(ctx, A)
    xh, yh = rng.standard_normal(n), rng.standard_normal(m)
    x, y, g = DeviceVector(ctx, n, xh), DeviceVector(ctx, m, yh), DeviceVector(ctx, n)
    for M in (B, Cm, A):
        J.update_pattern(M.indptr, M.indices, M.data)
        J.mul(y, x, 1.0, 0.0)
        assert rel(y.download(), M @ xh) < 1e-13
        y.upload(yh)
        J.mul_t(g, y, 1.0, 0.0)
        assert rel(g.download(), M.T @ yh) < 1e-13
    # both images written on the device: J = diag(1 + 2 c t) A
    aval, aval_r, t = DeviceVector(ctx, A.nnz, A.data), DeviceVector(ctx, A.nnz), DeviceVector(ctx, m, yh)
    J.gather_csr(aval, aval_r)
    check(lib().lso_synth_csc_jacobian_both(J.handle, aval.ptr, aval_r.ptr, t.ptr, 0.1), ctx.handle)
    Jh = sp.diags(1.0 + 0.2 * yh) @ A
    J.mul(y, x, 1.0, 0.0)
    assert rel(y.download(), Jh @ xh) < 1e-13
    y.upload(yh)
    J.mul_t(g, y, 1.0, 0.0)
    assert rel(g.download(), Jh.T @ yh) < 1e-13


@pytest.mark.parametrize("m,n,damped", [(9, 6, True), (400, 60, False), (20000, 3000, True), (3000, 5000, True),
                                        (60000, 900, False)])
def test_fused_lsmr_equals_generic_and_oracle(ctx, m, n, damped):
    """The fused driver (3 launches per iteration, scalars on the device) and the generic driver (the reference's
    wrappers op for op, scalars on the host) stop at the same iteration with the same istop as the oracle; the launch
    and synchronisation counts of the fused driver are what DESIGN.md states."""
    from lsob200 import CSCMatrix, DeviceVector, LSMRAllocatedSolver, LSMRDampenedAllocatedSolver
    rng = np.random.default_rng(m + 2 * n + damped)
    A = ragged(m, n, m + n) if m >= 100 else sp.random(m, n, density=0.6, random_state=5, format="csc")
    A.sort_indices()
    yh = rng.standard_normal(m)
    damp = np.asarray(A.multiply(A).sum(axis=0)).ravel() / 10 + 1e-3
    xr, nmul_r, it_r, istop_r = O.lsmr_ldiv(A, yh, damp.copy() if damped else None)
    J = CSCMatrix.from_scipy(ctx, A)
    y = DeviceVector(ctx, m, yh)
    res = {}
    for fused in (1, 0):
        ctx.set_option("lsmr_fused", fused)
        try:
            ws = (LSMRDampenedAllocatedSolver if damped else LSMRAllocatedSolver)(ctx, m, n)
            x = DeviceVector(ctx, n)
            d = DeviceVector(ctx, n, damp) if damped else None
            _, nmul = ws.ldiv(x, J, y, d) if damped else ws.ldiv(x, J, y)
            res[fused] = (x.download(), ws.last_iters, ws.last_istop, nmul, ws.stats())
            if damped:
                assert rel(d.download(), np.sqrt(damp)) < 1e-15          # iterative_lsmr.jl:252
        finally:
            ctx.set_option("lsmr_fused", 1)
    for fused in (1, 0):
        xg, it, istop, nmul, _ = res[fused]
        assert (it, istop, nmul) == (it_r, istop_r, nmul_r), (fused, it, istop, it_r, istop_r)
        assert rel(xg, xr) <= 2e-5
    launches, syncs = res[1][4]
    it = res[1][1]
    assert syncs <= max(it, 1) + 1                       # at most one read-back per iteration (+ the one that sees `done`)
    assert np.array_equal(y.download(), yh)


def test_fused_lsmr_tight_and_user_preconditioner(ctx):
    """Run to full convergence (atol = btol = 1e-15) the fused LSMR is the least-squares solution to 1e-10; a user
    diagonal preconditioner (README.md:47) stays on the fused path, a callback preconditioner takes the generic one;
    both converge to the same solution."""
    from lsob200 import CSCMatrix, DeviceVector, LSMRDampenedAllocatedSolver
    from lsob200._lib import check, lib
    m, n = 6000, 500
    rng = np.random.default_rng(77)
    A = sp.random(m, n, density=0.03, random_state=9, format="csc"); A.sort_indices()
    yh = rng.standard_normal(m)
    damp = np.asarray(A.multiply(A).sum(axis=0)).ravel() / 10 + 1e-3
    xq, _ = O.qr_ldiv(A.toarray(), yh, damp)
    J, y = CSCMatrix.from_scipy(ctx, A), DeviceVector(ctx, m, yh)
    pvec = DeviceVector(ctx, n, 1.0 / np.sqrt(np.asarray(A.multiply(A).sum(axis=0)).ravel() + 2.0 * damp))

    def solve(pdiag=None, pfn=None):
        ws = LSMRDampenedAllocatedSolver(ctx, m, n)
        x, d = DeviceVector(ctx, n), DeviceVector(ctx, n, damp)
        iters, istop = C.c_int64(), C.c_int()
        check(lib().lso_lsmr_solve_ex(ws._h, J.handle, None, 0, y.ptr, d.ptr, x.ptr, 1e-15, 1e-15, 0.0, 0,
                                      pdiag, pfn, None, C.byref(iters), C.byref(istop)), ctx.handle)
        return x.download(), iters.value, ws.stats()

    x0, it0, st0 = solve()
    assert rel(x0, xq) <= 1e-9
    xr, _, it_r, _ = O.lsmr_ldiv(A, yh, damp.copy(), atol=1e-15, btol=1e-15, conlim=0.0)
    assert abs(it0 - it_r) <= 2 and rel(x0, xr) <= 1e-10
    x1, it1, st1 = solve(pdiag=pvec.ptr)
    assert rel(x1, xq) <= 1e-9 and st1[1] <= it1 + 1

    from lsob200.solvers import PRECOND_FN

    def cb(user, nn, d_in, d_out):
        check(lib().lso_vec_mul(ctx.handle, nn, d_out, d_in, pvec.ptr), ctx.handle)
        return 0
    keep = PRECOND_FN(cb)
    x2, it2, st2 = solve(pfn=C.cast(keep, C.c_void_p))
    assert rel(x2, xq) <= 1e-9
    assert abs(it2 - it1) <= 2 and rel(x2, x1) <= 1e-9


def test_lm_lsmr_with_user_preconditioner_and_pattern_change(ctx):
    """optimize! with LSMR(preconditioner) (README.md:47) and with a host g! that stores a DIFFERENT sparsity pattern on
    alternate calls (setindex!-style fill of a sparse J, test/nonlinearsolvers.jl:526-530): same run as the default."""
    import lsob200 as L
    import problems as P
    name, f, g, x0 = P.readme_rosenbrock()
    n = x0.size
    calls = {"k": 0}
    rows, cols = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")

    def g_sparse(J, x):
        Jd = np.zeros((n, n), order="F")
        g(Jd, x)
        calls["k"] += 1
        if calls["k"] % 2:       # all n*n entries stored (explicit zeros)
            Jn = sp.csc_matrix((Jd.ravel(), (rows.ravel(), cols.ravel())), shape=(n, n))
        else:                    # structural zeros dropped
            Jn = sp.csc_matrix(Jd)
        Jn.sort_indices()
        J.indptr, J.indices, J.data = Jn.indptr.copy(), Jn.indices.copy(), Jn.data.copy()

    def run(g_, J, solver):
        return L.optimize_(L.LeastSquaresProblem(x=x0.copy(), y=np.zeros(n), f_=f, g_=g_, J=J), L.LevenbergMarquardt(solver))

    r0 = run(P.sparse_adapter(g, n), P.dense_pattern_csc(n), L.LSMR())
    r1 = run(g_sparse, sp.csc_matrix(np.ones((n, n))), L.LSMR())
    assert r0.converged and r1.converged and calls["k"] >= 4
    assert (r1.iterations, r1.mul_calls) == (r0.iterations, r0.mul_calls)
    assert np.linalg.norm(r1.minimizer - r0.minimizer) <= 1e-9

    def pc(x, J, damp):          # preconditioner!(P, x, J, damp): the default one, built by the user
        v = L.DeviceVector(ctx, n)
        J.colsumabs2(v)
        h = v.download() + (damp.download() if damp is not None else 0.0)
        return L.DeviceVector(ctx, n, np.where(h > 0, 1.0 / np.sqrt(np.where(h > 0, h, 1.0)), 0.0))
    r2 = run(P.sparse_adapter(g, n), P.dense_pattern_csc(n), L.LSMR(pc))
    assert r2.converged and (r2.iterations, r2.mul_calls) == (r0.iterations, r0.mul_calls)
    assert np.linalg.norm(r2.minimizer - r0.minimizer) <= 1e-9

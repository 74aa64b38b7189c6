"""The tcgen05 path of `mul!(cholm, J', J)` (src/solver/dense_cholesky.jl:31,48): J'J formed from int8 digit matrices by
tcgen05.mma.kind::i8 into TMEM (ctx option "syrk" = 2, csrc/ozaki.cu) against fp64 references."""
import numpy as np
import pytest

from oracle import reference_port as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def make_J(m, n, seed, spread=6.0):
    rng = np.random.default_rng(seed)
    J = rng.standard_normal((m, n)) * 10.0 ** (spread * (rng.random(n) - 0.5))        # column scales over `spread` decades
    J *= 10.0 ** (2.0 * (rng.random((m, 1)) - 0.5))                                    # and rows over two more
    return np.asfortranarray(J), rng.standard_normal(m)


def gram_error(ctx, Jh, slices, damp=None):
    """max over (i, j) of |C_ij - (J'J)_ij| / (||J_i|| ||J_j||): C is read back through a solve's factor R
    (R'R = C [+ diag(damp)]), so the comparison is made on R'R."""
    import lsob200 as L
    m, n = Jh.shape
    ctx.set_option("syrk", 2)
    ctx.set_option("ozaki_slices", slices)
    try:
        ws = L.DenseCholeskyAllocatedSolver(ctx, m, n, damped=damp is not None)
        x = L.DeviceVector(ctx, n)
        ws.ldiv(x, L.DenseMatrix(ctx, m, n, Jh), L.DeviceVector(ctx, m, np.ones(m)),
                L.DeviceVector(ctx, n, damp) if damp is not None else None)
        R = ws.factor()
    finally:
        ctx.set_option("syrk", 1)
        ctx.set_option("ozaki_slices", 8)
    C = R.T @ R
    if damp is not None:
        C -= np.diag(damp)
    Jl = Jh.astype(np.longdouble)
    Cref = np.asarray(Jl.T @ Jl, dtype=np.float64)
    nrm = np.linalg.norm(Jh, axis=0)
    return np.max(np.abs(C - Cref) / np.outer(nrm, nrm))


@pytest.mark.parametrize("m,n", [(64, 8), (300, 40), (1000, 128), (5001, 257), (20000, 96), (40000, 520)])
def test_ozaki_cholesky_solve_matches_oracle(ctx, m, n):
    """dense_cholesky.jl:43-59 with the tcgen05 syrk: δ within 1e-10 of the oracle, like the DMMA path."""
    import lsob200 as L
    Jh, yh = make_J(m, n, 7 * m + n, spread=2.0)
    dtd = np.einsum("ij,ij->j", Jh, Jh)
    damp = np.clip(dtd, 1e-6 * dtd.mean(), 1e32 * dtd.mean()) / 10.0
    xr = O.chol_ldiv(Jh, yh, damp.copy())
    ctx.set_option("syrk", 2)
    try:
        ws = L.DenseCholeskyAllocatedSolver(ctx, m, n, damped=True)
        x = L.DeviceVector(ctx, n)
        ws.ldiv(x, L.DenseMatrix(ctx, m, n, Jh), L.DeviceVector(ctx, m, yh), L.DeviceVector(ctx, n, damp))
        e2 = rel(x.download(), xr)
    finally:
        ctx.set_option("syrk", 1)
    ws1 = L.DenseCholeskyAllocatedSolver(ctx, m, n, damped=True)
    x1 = L.DeviceVector(ctx, n)
    ws1.ldiv(x1, L.DenseMatrix(ctx, m, n, Jh), L.DeviceVector(ctx, m, yh), L.DeviceVector(ctx, n, damp))
    e1 = rel(x1.download(), xr)
    assert e2 <= 1e-10, (e2, e1)


def test_ozaki_gram_accuracy_by_digit_count(ctx):
    """The representation error is 2^(-7 S): with 8 digits J'J is at fp64 accuracy relative to ||J_i|| ||J_j||, and every
    digit removed costs a factor of 2^7 — which also proves all S(S+1)/2 digit products really contribute."""
    Jh, _ = make_J(6000, 200, 3)
    errs = {S: gram_error(ctx, Jh, S) for S in (8, 7, 6, 5, 4)}
    assert errs[8] <= 5e-15, errs
    for S in (7, 6, 5, 4):
        bound = 6000 ** 0.5 * 2.0 ** (-7 * S) * 8
        assert errs[S] <= bound, (S, errs)
        assert errs[S] >= 2.0 ** (-7 * S - 12), (S, errs)        # fewer digits must actually lose accuracy


def test_ozaki_many_rows_drains_the_integer_accumulators(ctx):
    """More than 32 768 rows per CTA: the int32 accumulators are drained into fp64 between chunks (no overflow), and the
    rows that do not fill the last 128-row block are padded with zeros."""
    Jh, _ = make_J(70001, 136, 11, spread=1.0)
    assert gram_error(ctx, Jh, 8) <= 5e-15
    # worst case for the integer sums: every digit at its maximum, all of the same sign
    # (the columns are identical, so the matrix is made positive definite by a damping term that is subtracted again)
    Jw = np.asfortranarray(np.full((66000, 130), 0.49999999999))
    assert gram_error(ctx, Jw, 8, damp=np.full(130, 66000 * 0.25)) <= 2e-14


def test_ozaki_nonfinite_input_is_not_silent(ctx):
    import lsob200 as L
    Jh, yh = make_J(500, 20, 1)
    Jh[17, 3] = np.nan
    ctx.set_option("syrk", 2)
    try:
        ws = L.DenseCholeskyAllocatedSolver(ctx, 500, 20, damped=False)
        x = L.DeviceVector(ctx, 20)
        with pytest.raises(L.LsoError):
            ws.ldiv(x, L.DenseMatrix(ctx, 500, 20, Jh), L.DeviceVector(ctx, 500, yh))
    finally:
        ctx.set_option("syrk", 1)


def test_ozaki_extreme_column_scales(ctx):
    """Columns 280 decades apart (their Gram entries still fit the double range) and a J whose odd leading dimension makes
    every second column 8-byte- but not 16-byte-aligned (the split's scalar-load path)."""
    import lsob200 as L
    m, n = 3001, 40
    rng = np.random.default_rng(9)
    Jh = np.asfortranarray(rng.standard_normal((m, n)))
    Jh[:, 3] *= 1e-140
    Jh[:, 7] *= 1e140
    Jh[:, 11] *= 1e-100
    yh = rng.standard_normal(m)
    dtd = np.einsum("ij,ij->j", Jh, Jh)
    damp = dtd / 10.0
    xr = O.chol_ldiv(Jh * 1.0, yh, damp.copy()) if np.all(np.isfinite(dtd)) else None
    ctx.set_option("syrk", 2)
    try:
        ws = L.DenseCholeskyAllocatedSolver(ctx, m, n, damped=True)
        x = L.DeviceVector(ctx, n)
        ws.ldiv(x, L.DenseMatrix(ctx, m, n, Jh), L.DeviceVector(ctx, m, yh), L.DeviceVector(ctx, n, damp))
        xg = x.download()
    finally:
        ctx.set_option("syrk", 1)
    # compare with the DMMA path on the same inputs (the oracle's LAPACK also just multiplies these columns out)
    ws1 = L.DenseCholeskyAllocatedSolver(ctx, m, n, damped=True)
    x1 = L.DeviceVector(ctx, n)
    ws1.ldiv(x1, L.DenseMatrix(ctx, m, n, Jh), L.DeviceVector(ctx, m, yh), L.DeviceVector(ctx, n, damp))
    scale = np.sqrt(dtd)                       # compare δ in the scaled variables (δ_j ‖J_j‖), where every column counts
    assert rel(xg * scale, x1.download() * scale) <= 1e-10
    if xr is not None:
        assert rel(xg * scale, xr * scale) <= 1e-10


def test_ozaki_error_bound_on_spiky_columns(ctx):
    """Digits are fixed point relative to the column maximum: with one dominant entry per column the error of (J'J)[i,j] is
    bounded by K * 2^(5 - 7S) * max|J_i| * max|J_j| in the worst case (dropped digit pairs p + q > S + 1 and the last digit's
    truncation), sqrt(K) of that for uncorrelated signs — check both at S = 8, K = 20 000."""
    m, n = 20000, 130
    rng = np.random.default_rng(21)
    Jh = rng.standard_normal((m, n))
    spikes = rng.integers(0, m, n)
    Jh[spikes, np.arange(n)] = 1e5 * (1.0 + rng.random(n))       # one entry 1e5 x larger than the rest of its column
    Jh = np.asfortranarray(Jh)
    err = gram_error(ctx, Jh, 8)
    assert err <= m * 2.0 ** (5 - 56), err
    assert err <= 30 * m ** 0.5 * 2.0 ** (5 - 56), err

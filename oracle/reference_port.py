"""ORACLE — CPU restatement of LeastSquaresOptim.jl's hot path (TEST INFRASTRUCTURE, not product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product path (leastsquaresoptim.jl_b200/) never does.

PARITY UNPINNED at per-solve level: the reference (Julia) cannot run in this image (no `julia`, no network),
and its own tests hold no per-solve golden vectors.  What this oracle IS pinned against:
  * every end-to-end assertion of the reference's test-suite that concerns this path (tests/test_oracle_pins.py):
    ssr <= 1e-3 on the 21 MINPACK problems x {dense,sparse} x {QR,LSMR} x {Dogleg,LM}
    (test/nonlinearsolvers.jl:505-537), converged && ssr <= 1e-3 with Cholesky (:573-595), the factor model
    (test/nonlinearleastsquares.jl:91-110), bounds (test/bounds.jl:11-36), default dispatch (:619-628);
  * scipy.sparse.linalg.lsmr (independent LSMR implementation) iteration-for-iteration;
  * LAPACK itself for the dense solves: the routines called here (dgelsy = dgeqp3 + dlaic1 rank + dtzrzf +
    dormqr/dormrz/dtrsm; dsyrk/dpotrf/dpotrs; dpstrf) are the very routines Julia's stdlib dispatches
    `qr!(·, ColumnNorm())`/`ldiv!`, `cholesky!` to — from OpenBLAS, the same library family Julia bundles.

Each function cites the reference file:line it follows (paths relative to the reference checkout).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg.lapack as lapack
import scipy.sparse as sp

MIN_DELTA, MAX_DELTA = 1e-16, 1e16          # src/types.jl:107-108
MIN_STEP_QUALITY = 1e-3                      # src/types.jl:109
MIN_DIAGONAL, MAX_DIAGONAL = 1e-6, 1e32      # src/types.jl:110-111
DECREASE_THRESHOLD, INCREASE_THRESHOLD = 0.25, 0.75   # src/optimizer/dogleg.jl:38-39
EPS = np.finfo(np.float64).eps


# ------------------------------------------------------------------------------------------------------
# src/utils/utils.jl:139-176
# ------------------------------------------------------------------------------------------------------
def colsumabs2(J):
    if sp.issparse(J):
        J = J.tocsc()
        out = np.zeros(J.shape[1])
        d2 = J.data * J.data
        for j in range(J.shape[1]):
            out[j] = d2[J.indptr[j]:J.indptr[j + 1]].sum()
        return out
    return np.einsum("ij,ij->j", J, J)


def wdot(x, y, w):
    return float(np.sum(w * x * y))


def wnorm(x, w):
    return math.sqrt(wdot(x, x, w))


def maxabs_projected_gradient(g, x, lower, upper):
    """src/utils/utils.jl:39-55"""
    haslower, hasupper = lower is not None and len(lower) > 0, upper is not None and len(upper) > 0
    if not (haslower or hasupper):
        return float(np.max(np.abs(g))) if not np.any(np.isnan(g)) else math.nan
    m = 0.0
    for i in range(len(g)):
        gi = g[i]
        if haslower and x[i] <= lower[i] and gi > 0:
            gi = 0.0
        elif hasupper and x[i] >= upper[i] and gi < 0:
            gi = 0.0
        a = abs(gi)
        if a > m:
            m = a
    return m


def assess_convergence(dx, maxabs_gr, ssr, trial_ssr, xtol, ftol, grtol, step_accepted):
    """src/utils/utils.jl:7-31"""
    x_c = f_c = g_c = False
    if step_accepted and abs(trial_ssr - ssr) <= ftol * (abs(ssr) + ftol):
        f_c = True
    elif _maxabs(dx) <= xtol:
        x_c = True
    elif maxabs_gr <= grtol:
        g_c = True
    return x_c, f_c, g_c, (x_c or f_c or g_c)


def _maxabs(v):
    a = np.abs(v)
    return math.nan if np.any(np.isnan(a)) else float(a.max())


# ------------------------------------------------------------------------------------------------------
# dense solvers
# ------------------------------------------------------------------------------------------------------
def qr_ldiv(J, y, damp=None):
    """src/solver/dense_qr.jl:30-42 (undamped) and :56-88 (damped).

    `ldiv!(qr!(qrm, ColumnNorm()), u)` [Julia stdlib] == LAPACK dgeqp3, then incremental condition estimation
    (dlaic1) with rcond = min(rows, cols) * eps to find the rank, complete orthogonal factorisation (dtzrzf) and
    the minimum-norm solution — i.e. LAPACK dgelsy without its input scaling.  Returns (x, rank)."""
    m, n = J.shape
    if damp is not None:
        qrm = np.zeros((m + n, n), order="F")
        qrm[:m, :] = J
        qrm[m + np.arange(n), np.arange(n)] = np.sqrt(damp)
        u = np.zeros(max(m + n, n))
        u[:m] = y
    else:
        qrm = np.array(J, order="F", dtype=np.float64)
        u = np.zeros(max(m, n))
        u[:m] = y
    rows = qrm.shape[0]
    rcond = min(rows, n) * EPS
    lwork = max(1, 4 * (rows + n) * 16 + 3 * n + 64)
    v, x, jpvt, rank, info = lapack.dgelsy(qrm, u.reshape(-1, 1), np.zeros(n, dtype=np.int32), rcond, lwork)
    if info != 0:
        raise RuntimeError(f"dgelsy info={info}")
    return np.array(x[:n, 0]), int(rank)


class PosDefException(Exception):
    pass


class RankDeficientException(Exception):
    pass


def chol_ldiv(J, y, damp=None):
    """src/solver/dense_cholesky.jl:29-35 (undamped, pivoted dpstrf, tol=0) and :43-59 (damped, dpotrf)."""
    C = np.asfortranarray(J.T @ J)
    x = J.T @ y
    n = C.shape[0]
    if damp is not None:
        C[np.arange(n), np.arange(n)] += damp
        c, info = lapack.dpotrf(C, lower=0)
        if info != 0:
            raise PosDefException(info)
        x, info = lapack.dpotrs(c, x, lower=0)
        return x
    c, piv, rank, info = lapack.dpstrf(C, lower=0, tol=0.0)
    if rank < n:
        raise RankDeficientException(info)
    p = piv - 1
    xp = x[p]
    U = np.triu(c)
    z = lapack.dtrtrs(U, xp, lower=0, trans=1)[0]
    w = lapack.dtrtrs(U, z, lower=0, trans=0)[0]
    out = np.empty(n)
    out[p] = w
    return out


# ------------------------------------------------------------------------------------------------------
# LSMR: src/utils/lsmr.jl:53-238 over the wrappers of src/solver/iterative_lsmr.jl
# ------------------------------------------------------------------------------------------------------
class _Precond:
    """PreconditionedMatrix(A, InverseDiagonal(P)) with A = J or DampenedMatrix(J, diag) — iterative_lsmr.jl:12-122.
    Augmented vectors (DampenedVector) are represented as one array [y; x]."""

    def __init__(self, J, P, diag=None):
        self.J, self.P, self.diag = J, P, diag
        self.m, self.n = J.shape

    def rows(self):
        return self.m + (self.n if self.diag is not None else 0)

    def mul(self, b, a, alpha, beta):
        """b <- alpha*A*a + beta*b   (:30-34 over :87-94)"""
        tmp = a * self.P
        if self.diag is not None:
            if beta != 1:
                b *= beta
            b[:self.m] += alpha * (self.J @ tmp)
            b[self.m:] = b[self.m:] + alpha * tmp * self.diag
        else:
            b[:] = alpha * (self.J @ tmp) + (beta * b if beta != 0 else 0.0)
        return b

    def mul_t(self, b, a, alpha, beta):
        """b <- alpha*A'*a + beta*b   (:36-51 over :95-109)"""
        tmp = self.J.T @ a[:self.m]
        if self.diag is not None:
            tmp = tmp + 1.0 * a[self.m:] * self.diag
        tmp2 = tmp * self.P
        if beta != 1:
            if beta == 0:
                b[:] = 0.0
            else:
                b *= beta
        b += alpha * tmp2
        return b

    def norm(self, b):
        if self.diag is not None:   # DampenedVector norm, :72
            return math.sqrt(np.linalg.norm(b[:self.m]) ** 2 + np.linalg.norm(b[self.m:]) ** 2)
        return float(np.linalg.norm(b))


def lsmr(x, A: _Precond, b, atol=1e-6, btol=1e-6, conlim=1e8, maxiter=None, lam=0.0):
    """src/utils/lsmr.jl:53-238.  x (initial guess) and b are modified in place.  Returns (x, iters, istop)."""
    m, n = A.rows(), A.n
    if maxiter is None:
        maxiter = max(m, n)
    ctol = 1.0 / conlim if conlim > 0 else 0.0
    u = A.mul(b, x, -1.0, 1.0)
    beta = A.norm(u)
    if beta > 0:
        u *= 1.0 / beta
    v = np.zeros(n)
    A.mul_t(v, u, 1.0, 0.0)
    alpha = float(np.linalg.norm(v))
    if alpha > 0:
        v *= 1.0 / alpha
    zetabar, alphabar, rho, rhobar, cbar, sbar = alpha * beta, alpha, 1.0, 1.0, 1.0, 0.0
    h = v.copy()
    hbar = np.zeros(n)
    betadd, betad, rhodold, tautildeold, thetatilde, zeta, d = beta, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0
    normA2, maxrbar, minrbar = alpha * alpha, 0.0, 1e100
    normb, istop, normr, normAr = beta, 0, beta, alpha * beta
    it = 0
    if normAr != 0:
        while it < maxiter:
            it += 1
            A.mul(u, v, 1.0, -alpha)
            beta = A.norm(u)
            if beta > 0:
                u *= 1.0 / beta
                A.mul_t(v, u, 1.0, -beta)
                alpha = float(np.linalg.norm(v))
                if alpha > 0:
                    v *= 1.0 / alpha
            alphahat = math.sqrt(alphabar ** 2 + lam ** 2)
            chat, shat = alphabar / alphahat, lam / alphahat
            rhoold = rho
            rho = math.sqrt(alphahat ** 2 + beta ** 2)
            c, s = alphahat / rho, beta / rho
            thetanew = s * alpha
            alphabar = c * alpha
            rhobarold, zetaold = rhobar, zeta
            thetabar, rhotemp = sbar * rho, cbar * rho
            rhobar = math.sqrt((cbar * rho) ** 2 + thetanew ** 2)
            cbar = cbar * rho / rhobar
            sbar = thetanew / rhobar
            zeta = cbar * zetabar
            zetabar = -sbar * zetabar
            hbar *= -thetabar * rho / (rhoold * rhobarold)
            hbar += h
            x += (zeta / (rho * rhobar)) * hbar
            h *= -thetanew / rho
            h += v
            betaacute, betacheck = chat * betadd, -shat * betadd
            betahat = c * betaacute
            betadd = -s * betaacute
            thetatildeold = thetatilde
            rhotildeold = math.sqrt(rhodold ** 2 + thetabar ** 2)
            ctildeold, stildeold = rhodold / rhotildeold, thetabar / rhotildeold
            thetatilde = stildeold * rhobar
            rhodold = ctildeold * rhobar
            betad = -stildeold * betad + ctildeold * betahat
            tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold
            taud = (zeta - thetatilde * tautildeold) / rhodold
            d = d + betacheck ** 2
            normr = math.sqrt(d + (betad - taud) ** 2 + betadd ** 2)
            normA2 = normA2 + beta ** 2
            normA = math.sqrt(normA2)
            normA2 = normA2 + alpha ** 2
            maxrbar = max(maxrbar, rhobarold)
            if it > 1:
                minrbar = min(minrbar, rhobarold)
            condA = max(maxrbar, rhotemp) / min(minrbar, rhotemp)
            normAr = abs(zetabar)
            normx = float(np.linalg.norm(x))
            test1 = normr / normb
            test2 = normAr / (normA * normr) if normA * normr != 0 else math.inf
            test3 = 1.0 / condA
            t1 = test1 / (1.0 + normA * normx / normb)
            rtol = btol + atol * normA * normx / normb
            if it >= maxiter: istop = 7; break
            if 1 + test3 <= 1: istop = 6; break
            if 1 + test2 <= 1: istop = 5; break
            if 1 + t1 <= 1: istop = 4; break
            if test3 <= ctol: istop = 3; break
            if test2 <= atol: istop = 2; break
            if test1 <= rtol: istop = 1; break
    return x, it, istop


def lsmr_ldiv(J, y, damp=None, atol=1e-6, btol=None, conlim=1e8, maxiter=None):
    """src/solver/iterative_lsmr.jl:179-198 (undamped) / :238-259 (damped; mutates damp <- sqrt(damp)).
    Returns (x, n_mul = 2*iters, iters, istop)."""
    m, n = J.shape
    x = np.zeros(n)
    P = colsumabs2(J)
    if damp is not None:
        P = P + damp
    with np.errstate(divide="ignore", invalid="ignore"):
        P = np.where(P > 0, 1.0 / np.sqrt(np.where(P > 0, P, 1.0)), 0.0)
    if damp is not None:
        np.sqrt(damp, out=damp)
        A = _Precond(J, P, damp)
        b = np.concatenate([np.asarray(y, dtype=np.float64), np.zeros(n)])
        x, it, istop = lsmr(x, A, b, atol=atol, btol=0.5 if btol is None else btol, conlim=conlim, maxiter=maxiter)
    else:
        A = _Precond(J, P, None)
        b = np.array(y, dtype=np.float64)
        x, it, istop = lsmr(x, A, b, atol=atol, btol=1e-6 if btol is None else btol, conlim=conlim, maxiter=maxiter)
    return x * P, 2 * it, it, istop


# ------------------------------------------------------------------------------------------------------
# optimizers
# ------------------------------------------------------------------------------------------------------
@dataclass
class Result:
    optimizer: str
    minimizer: np.ndarray
    ssr: float
    iterations: int
    converged: bool
    x_converged: bool
    f_converged: bool
    g_converged: bool
    f_calls: int
    g_calls: int
    mul_calls: int
    tr: list = field(default_factory=list)
    deltas: list = field(default_factory=list)
    solve_inputs: list = field(default_factory=list)


def _solve(solver, J, f, damp):
    if solver == "qr":
        Jd = J.toarray() if sp.issparse(J) else J
        x, _ = qr_ldiv(Jd, f, damp)
        return x, 1
    if solver == "cholesky":
        return chol_ldiv(J, f, damp), 1
    if solver == "lsmr":
        x, nmul, _, _ = lsmr_ldiv(J, f, damp)
        return x, nmul
    raise ValueError(solver)


def _check_bounds(x, lower, upper):
    n = len(x)
    lo = None if lower is None or len(lower) == 0 else np.asarray(lower, dtype=np.float64)
    hi = None if upper is None or len(upper) == 0 else np.asarray(upper, dtype=np.float64)
    if (lo is not None and lo.size != n) or (hi is not None and hi.size != n):
        raise ValueError("Bounds must either be empty or of the same length as the number of parameters.")
    if (lo is not None and not np.all(x >= lo)) or (hi is not None and not np.all(x <= hi)):
        raise ValueError("Initial guess must be within bounds.")
    return lo, hi


def levenberg_marquardt(f_, g_, x, J, m, solver="qr", x_tol=1e-8, f_tol=1e-8, g_tol=1e-8, iterations=1000, delta=10.0,
                        lower=None, upper=None, record=False, store_trace=False):
    """src/optimizer/levenberg_marquardt.jl:39-144"""
    x = np.array(x, dtype=np.float64)
    n = x.size
    lo, hi = _check_bounds(x, lower, upper)
    fcur, ftrial = np.zeros(m), np.zeros(m)
    decrease_factor = 2.0
    f_calls = g_calls = mul_calls = 0
    converged = x_c = f_c = g_c = False
    f_(fcur, x); f_calls += 1
    ssr = float(np.sum(fcur * fcur))
    maxabs_gr = math.inf
    need_jacobian = True
    it = 0
    tr = [(0, ssr, maxabs_gr)] if store_trace else []
    deltas, inputs = [], []
    while not converged and it < iterations:
        it += 1
        if not np.all(np.isfinite(x)):
            raise FloatingPointError("IsFiniteException")
        if need_jacobian:
            g_(J, x); g_calls += 1
            need_jacobian = False
        dtd = colsumabs2(J)
        dtd_mean = dtd.sum() / n
        dtd = np.clip(dtd, MIN_DIAGONAL * dtd_mean, MAX_DIAGONAL * dtd_mean)
        dtd = dtd * (1 / delta)
        if record:
            inputs.append((J.copy(), fcur.copy(), dtd.copy()))
        dx, lmiter = _solve(solver, J, fcur, dtd)
        if record:
            deltas.append(dx.copy())
        if lo is not None:
            dx = np.minimum(dx, x - lo)
        if hi is not None:
            dx = np.maximum(dx, x - hi)
        mul_calls += lmiter
        g = J.T @ fcur
        mul_calls += 1
        maxabs_gr = maxabs_projected_gradient(g, x, lo, hi)
        x -= dx
        f_(ftrial, x); f_calls += 1
        trial_ssr = float(np.sum(ftrial * ftrial))
        fpredict = J @ dx - fcur
        mul_calls += 1
        predicted_ssr = float(np.sum(fpredict * fpredict))
        predicted_reduction = abs(ssr - predicted_ssr)
        rho = (ssr - trial_ssr) / predicted_reduction if predicted_reduction > 0 else 0.0
        accepted = rho > MIN_STEP_QUALITY
        x_c, f_c, g_c, converged = assess_convergence(dx, maxabs_gr, ssr, trial_ssr, x_tol, f_tol, g_tol, accepted)
        if accepted:
            fcur[:] = ftrial
            ssr = trial_ssr
            t = 2.0 * rho - 1.0
            delta = min(delta / max(1 / 3, 1.0 - t * t * t), MAX_DELTA)
            decrease_factor = 2.0
            need_jacobian = True
        else:
            x += dx
            delta = max(delta / decrease_factor, MIN_DELTA)
            decrease_factor *= 2.0
        if store_trace:
            tr.append((it, ssr, maxabs_gr))
    return Result("LevenbergMarquardt", x, ssr, it, converged, x_c, f_c, g_c, f_calls, g_calls, mul_calls, tr, deltas, inputs)


def dogleg(f_, g_, x, J, m, solver="qr", x_tol=1e-8, f_tol=1e-8, g_tol=1e-8, iterations=1000, delta=1.0,
           lower=None, upper=None, record=False, store_trace=False):
    """src/optimizer/dogleg.jl:41-203"""
    x = np.array(x, dtype=np.float64)
    n = x.size
    lo, hi = _check_bounds(x, lower, upper)
    fcur, ftrial = np.zeros(m), np.zeros(m)
    reuse = False
    wnorm_dgn = wnorm_dgr = alpha = 0.0
    f_calls = g_calls = mul_calls = 0
    converged = x_c = f_c = g_c = False
    f_(fcur, x); f_calls += 1
    ssr = float(np.sum(fcur * fcur))
    maxabs_gr = math.inf
    it = 0
    tr = [(0, ssr, maxabs_gr)] if store_trace else []
    deltas, inputs = [], []
    dgn = dgr = dtd = None
    while not converged and it < iterations:
        it += 1
        if not np.all(np.isfinite(x)):
            raise FloatingPointError("IsFiniteException")
        if not reuse:
            g_(J, x); g_calls += 1
            dtd = np.clip(colsumabs2(J), MIN_DIAGONAL, MAX_DIAGONAL)
            if it == 1:
                wnorm_x = wnorm(x, dtd)
                if wnorm_x > 0:
                    delta *= wnorm_x
            dgr = J.T @ fcur
            mul_calls += 1
            maxabs_gr = maxabs_projected_gradient(dgr, x, lo, hi)
            dgr = dgr / dtd
            wnorm_dgr = wnorm(dgr, dtd)
            fp = J @ dgr
            mul_calls += 1
            with np.errstate(divide="ignore", invalid="ignore"):
                alpha = float(np.float64(wnorm_dgr ** 2) / np.float64(np.sum(fp * fp)))
            if record:
                inputs.append((J.copy(), fcur.copy(), None))
            dgn, ls_iter = _solve(solver, J, fcur, None)
            if record:
                deltas.append(dgn.copy())
            mul_calls += ls_iter
            wnorm_dgn = wnorm(dgn, dtd)
        if wnorm_dgn <= delta:
            dx = dgn.copy()
            wnorm_dx = wnorm_dgn
        elif wnorm_dgr * alpha >= delta:
            dx = dgr * (delta / wnorm_dgr)
            wnorm_dx = delta
        else:
            b_dot_a = alpha * wdot(dgr, dgn, dtd)
            a_sq = (alpha * wnorm_dgr) ** 2
            bma = a_sq - 2 * b_dot_a + wnorm_dgn ** 2
            c = b_dot_a - a_sq
            dd = math.sqrt(c ** 2 + bma * (delta ** 2 - a_sq))
            beta = (dd - c) / bma if c <= 0 else (delta ** 2 - a_sq) / (dd + c)
            dx = dgn * beta + alpha * (1 - beta) * dgr
            wnorm_dx = wnorm(dx, dtd)
        if lo is not None:
            dx = np.minimum(dx, x - lo)
        if hi is not None:
            dx = np.maximum(dx, x - hi)
        x -= dx
        f_(ftrial, x); f_calls += 1
        trial_ssr = float(np.sum(ftrial * ftrial))
        fpredict = J @ dx - fcur
        mul_calls += 1
        predicted_ssr = float(np.sum(fpredict * fpredict))
        predicted_reduction = abs(ssr - predicted_ssr)
        rho = (ssr - trial_ssr) / predicted_reduction if predicted_reduction > 0 else 0.0
        accepted = rho >= MIN_STEP_QUALITY
        x_c, f_c, g_c, converged = assess_convergence(dx, maxabs_gr, ssr, trial_ssr, x_tol, f_tol, g_tol, accepted)
        if accepted:
            reuse = False
            fcur[:] = ftrial
            ssr = trial_ssr
        else:
            reuse = True
            x += dx
        if rho < DECREASE_THRESHOLD:
            delta = max(MIN_DELTA, delta * 0.5)
        elif rho > INCREASE_THRESHOLD:
            delta = max(delta, 3.0 * wnorm_dx)
        if store_trace:
            tr.append((it, ssr, maxabs_gr))
    return Result("Dogleg", x, ssr, it, converged, x_c, f_c, g_c, f_calls, g_calls, mul_calls, tr, deltas, inputs)


def optimize(f_, g_, x0, J, m, optimizer="dogleg", solver=None, **kw):
    """`optimize!` with the defaults of src/types.jl:114-127."""
    if solver is None:
        solver = "lsmr" if sp.issparse(J) else "qr"
    if solver == "qr" and sp.issparse(J):
        raise ValueError("solver QR() is not available for sparse Jacobians. Choose between Cholesky() and LSMR()")
    if optimizer is None:
        optimizer = "lm" if solver == "lsmr" else "dogleg"
    if optimizer == "lm":
        return levenberg_marquardt(f_, g_, x0, J, m, solver=solver, **kw)
    return dogleg(f_, g_, x0, J, m, solver=solver, **kw)

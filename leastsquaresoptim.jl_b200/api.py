"""Host-side mirror of the reference's public API for the hot path (src/types.jl, src/optimizer/*.jl).

Same names, argument meaning and error behaviour as LeastSquaresOptim.jl:
  LeastSquaresProblem(x=..., f_=..., g_=..., J=..., y=..., output_length=...)     types.jl:7-68
  QR() / Cholesky() / LSMR()            solver markers                            types.jl:79-86
  Dogleg(solver) / LevenbergMarquardt(solver)                                     types.jl:89-98
  optimize_(nls, optimizer; x_tol, f_tol, g_tol, iterations, Δ, lower, upper)     `optimize!`  types.jl:207-209
  optimize(f, x0, optimizer)                                                      types.jl:182-184
  LeastSquaresResult                                                              types.jl:220-237

The trust-region outer loops below are literal restatements of levenberg_marquardt.jl:39-144 and
dogleg.jl:41-203 (control flow and scalars on the host, exactly what stays in Julia), but every vector /
matrix operation they issue runs on the device through the C ABI.  `f_` / `g_` are the user's `f!` / `g!`:
  * host mode  (default): f_(out: np.ndarray, x: np.ndarray), g_(J: np.ndarray | scipy csc, x) — as in Julia;
    the driver moves x down and f / J up around each call (this is the e2e path with host buffers);
  * device mode (`device_callbacks=True`): f_(out: DeviceVector, x: DeviceVector), g_(J, x) write HBM directly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from .device import Context, CSCMatrix, DenseMatrix, DeviceVector, wdot, wnorm
from .solvers import (DenseCholeskyAllocatedSolver, DenseQRAllocatedSolver, LSMRAllocatedSolver,
                      LSMRDampenedAllocatedSolver)

# shared constants, types.jl:107-111
MIN_DELTA = 1e-16
MAX_DELTA = 1e16
MIN_STEP_QUALITY = 1e-3
MIN_DIAGONAL = 1e-6
MAX_DIAGONAL = 1e32
# dogleg.jl:38-39
DECREASE_THRESHOLD = 0.25
INCREASE_THRESHOLD = 0.75


# ---- solver / optimizer markers (types.jl:79-98) ----------------------------------------------------
class AbstractSolver:
    pass


class QR(AbstractSolver):
    pass


class Cholesky(AbstractSolver):
    pass


class LSMR(AbstractSolver):
    """`LSMR(preconditioner!, P)` (types.jl:81-85, README.md:47).  `preconditioner` is passed to the allocated solver:
    None = default; a callable `(x, J, damp) -> DeviceVector` returning the vector of an InverseDiagonal; or a tuple
    `(update(x, J, damp), apply(out_ptr, in_ptr, n))` for a general `ldiv!(out, P, in)`."""

    def __init__(self, preconditioner=None):
        self.preconditioner = preconditioner


class AbstractOptimizer:
    def __init__(self, solver: Optional[AbstractSolver] = None):
        self.solver = solver


class Dogleg(AbstractOptimizer):
    pass


class LevenbergMarquardt(AbstractOptimizer):
    pass


def _is_sparse(J) -> bool:
    return hasattr(J, "tocsc") or isinstance(J, CSCMatrix)


def default_solver(solver, J):
    """types.jl:114-121"""
    if solver is not None:
        if isinstance(solver, QR) and _is_sparse(J):
            raise ValueError("solver QR() is not available for sparse Jacobians. Choose between Cholesky() and LSMR()")
        return solver
    return LSMR() if _is_sparse(J) else QR()


def default_optimizer(optimizer, solver):
    """types.jl:124-127"""
    if isinstance(optimizer, Dogleg):
        return Dogleg(solver)
    if isinstance(optimizer, LevenbergMarquardt):
        return LevenbergMarquardt(solver)
    return LevenbergMarquardt(LSMR()) if isinstance(solver, LSMR) else Dogleg(solver)


# ---- problem (types.jl:7-68) ---------------------------------------------------------------------------
class LeastSquaresProblem:
    def __init__(self, x=None, y=None, f_: Callable = None, g_: Callable = None, J=None, output_length: int = 0,
                 device_callbacks: bool = False, ctx: Context | None = None):
        if x is None:
            raise ValueError("initial x required")
        if f_ is None:
            raise ValueError("initial f! required")
        self.ctx = ctx or Context.default()
        self.device_callbacks = device_callbacks
        if device_callbacks:
            assert isinstance(x, DeviceVector) and isinstance(y, DeviceVector) and J is not None
            self.x, self.y, self.J = x, y, J
            if g_ is None:
                # autodiff = :central (types.jl:54-58) with f! on the device: J is produced on the device as well
                if not isinstance(J, DenseMatrix):
                    raise TypeError("the finite-difference Jacobian fills a dense J; give g! for a sparse J")
                g_ = _device_central_difference_jacobian(self.ctx, f_, J.m, J.n)
            self.f_, self.g_ = f_, g_
            m, n = J.shape
        else:
            self.x = np.array(x, dtype=np.float64)     # optimize! mutates nls.x in place (types.jl:189)
            if y is None:
                if output_length == 0:
                    if J is None:
                        raise ValueError("specify J or output_length")
                    output_length = J.shape[0]
                y = np.zeros(output_length)
            self.y = np.asarray(y, dtype=np.float64)
            if J is None:
                J = np.zeros((self.y.size, self.x.size), order="F")
            self.J = J
            self.f_ = f_
            if g_ is None:
                g_ = _central_difference_jacobian(f_, self.y.size)
            self.g_ = g_
            m, n = J.shape
        if len(self.x) != n:
            raise ValueError("DimensionMismatch: x must have length size(J, 2)")
        if len(self.y) != m:
            raise ValueError("DimensionMismatch: y must have length size(J, 1)")


def _device_central_difference_jacobian(ctx, f_, m, n):
    """(f2) `finite_difference_jacobian!(J, f!, x, cache)` (types.jl:56-58) for a device f!: lso_fd_jacobian_central calls
    f! back with device pointers; the Python f_(out: DeviceVector, x: DeviceVector) is wrapped accordingly."""
    import ctypes as C
    from ._lib import check, lib
    RES_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)
    work = DeviceVector(ctx, 2 * m)
    state = {}

    def _cb(user, d_x, d_out):
        try:
            f_(DeviceVector.view(ctx, d_out, m), DeviceVector.view(ctx, d_x, n))
            return 0
        except Exception as e:       # pragma: no cover
            state["error"] = e
            return 1
    cb = RES_FN(_cb)

    def g_(J, x):
        state.pop("error", None)
        st = lib().lso_fd_jacobian_central(ctx.handle, m, n, C.cast(cb, C.c_void_p), None, x.ptr, J.ptr, J.ld, work.ptr)
        if "error" in state:
            raise state["error"]
        check(st, ctx.handle)
    g_._keepalive = (cb, work)
    return g_


def _central_difference_jacobian(f_, m):
    """Stand-in for the FiniteDiff closure at types.jl:56-58 (host-only helper, off the hot path)."""
    def g_(J, x):
        n = x.size
        fp, fm = np.empty(m), np.empty(m)
        xx = x.copy()
        for j in range(n):
            h = np.cbrt(np.finfo(float).eps) * max(1.0, abs(x[j]))
            xx[j] = x[j] + h
            f_(fp, xx)
            xx[j] = x[j] - h
            f_(fm, xx)
            xx[j] = x[j]
            J[:, j] = (fp - fm) / (2 * h)
    return g_


@dataclass
class OptimizationState:
    iteration: int
    value: float
    g_norm: float


def format_trace(states, show_every: int = 1) -> str:
    """`show(os::OptimizationState)` for every state with `iteration % show_every == 0` (utils.jl:104-108, 124-127:
    `@printf "%6d   %14e   %14e\n" iteration value g_norm`)."""
    def e14(v):                         # Julia's Printf writes Inf / NaN where C writes inf / nan
        v = float(v)
        if math.isnan(v):
            return "NaN".rjust(14)
        if math.isinf(v):
            return ("Inf" if v > 0 else "-Inf").rjust(14)
        return "%14e" % v

    out = []
    for st in states:
        if st.iteration % max(int(show_every), 1) == 0:
            out.append("%6d   %s   %s\n" % (st.iteration, e14(st.value), e14(st.g_norm)))
    return "".join(out)


@dataclass
class LeastSquaresResult:
    optimizer: str
    minimizer: np.ndarray
    ssr: float
    iterations: int
    converged: bool
    x_converged: bool
    x_tol: float
    f_converged: bool
    f_tol: float
    g_converged: bool
    g_tol: float
    tr: list = field(default_factory=list)
    f_calls: int = 0
    g_calls: int = 0
    mul_calls: int = 0
    deltas: list = field(default_factory=list)   # per-solve δ (only when record_steps=True; parity tests)


# ---- device-side problem state -------------------------------------------------------------------------
class _Allocated:
    """LeastSquaresProblemAllocated (types.jl:141-157): device mirrors of x, y, J + optimizer/solver workspaces."""

    def __init__(self, nls: LeastSquaresProblem, optimizer: AbstractOptimizer, sharded: bool = False):
        self.nls = nls
        self.sharded = sharded
        self.optimizer = optimizer
        ctx = self.ctx = nls.ctx
        self.host = not nls.device_callbacks
        if self.host:
            m, n = nls.J.shape
            self.x = DeviceVector(ctx, n, nls.x)
            self.fcur = DeviceVector(ctx, m)
            self.sparse = _is_sparse(nls.J)
            if self.sparse:
                nls.J = nls.J.tocsc()
                nls.J.sort_indices()
                self.J = CSCMatrix.from_scipy(ctx, nls.J)
                self._pattern = (nls.J.indptr.copy(), nls.J.indices.copy())
            else:
                self.J = DenseMatrix(ctx, m, n)
            self._hx = np.empty(n)
            self._hf = np.empty(m)
        else:
            self.x, self.fcur, self.J = nls.x, nls.y, nls.J
            self.sparse = isinstance(self.J, CSCMatrix)
            m, n = self.J.shape
        self.m, self.n = m, n
        solver = optimizer.solver
        damped = isinstance(optimizer, LevenbergMarquardt)
        if sharded and getattr(ctx, "nranks", 1) <= 1:
            raise ValueError("sharded=True needs a communicator on the context (Context.comm_init)")
        sharded_lsmr = sharded and isinstance(solver, LSMR) and self.sparse
        if sharded and not sharded_lsmr and (isinstance(solver, LSMR) or self.sparse or
                                             (not damped and not isinstance(solver, QR))):
            # the row-sharded paths are LM(QR) / Dogleg(QR) (TSQR), LM(Cholesky) (one all-reduce of [J'J | J'f]) and LSMR on a
            # sparse J (one all-reduce of [J'u | ||u||²] per iteration); the undamped Cholesky has no sharded form
            raise ValueError("sharded=True is implemented for LevenbergMarquardt with QR() or Cholesky() and for Dogleg with "
                             "QR() on a dense J, and for LSMR() on a sparse (CSC) J")
        if isinstance(solver, QR):
            # row-sharded J: local QR of [J_k | y_k] needs the undamped m_k x n workspace; the sqrt(damp) rows join
            # the stack of R factors (lso_qr_solve_sharded)
            self.solver = DenseQRAllocatedSolver(ctx, m, n, damped and not sharded, sharded=sharded)
        elif isinstance(solver, Cholesky):
            if self.sparse:
                raise TypeError("MethodError: no Cholesky solver for sparse Jacobians (dense_cholesky.jl:19)")
            self.solver = DenseCholeskyAllocatedSolver(ctx, m, n, damped, sharded=sharded)
        elif isinstance(solver, LSMR):
            pc = getattr(solver, "preconditioner", None)
            m_total = 0
            if sharded_lsmr:
                if isinstance(pc, tuple):
                    raise ValueError("the row-sharded LSMR takes the default or a diagonal preconditioner")
                cnt = DeviceVector(ctx, 1, np.array([float(m)]))        # rows of the whole J: lsmr.jl:55 default maxiter
                ctx.allreduce(cnt)
                m_total = int(round(float(cnt.download()[0])))
            cls = LSMRDampenedAllocatedSolver if damped else LSMRAllocatedSolver
            self.solver = cls(ctx, m, n, pc, sharded=sharded_lsmr, m_total=m_total)
        else:
            raise TypeError(f"unknown solver {solver!r}")

    def workspace(self, key, make):
        """AllocatedLevenbergMarquardt / AllocatedDogleg vectors (levenberg_marquardt.jl:8-31, dogleg.jl:7-30):
        allocated once per problem and reused by every `optimize!` call on it."""
        if not hasattr(self, "_ws"):
            self._ws = {}
        if key not in self._ws:
            self._ws[key] = make()
        return self._ws[key]

    # f!(out, x) and g!(J, x) with the data movement each mode needs
    def f(self, out: DeviceVector, x: DeviceVector):
        if self.host:
            x.download(self._hx)
            self.nls.f_(self._hf, self._hx)
            out.upload(self._hf)
        else:
            self.nls.f_(out, x)

    def g(self, x: DeviceVector):
        if self.host:
            x.download(self._hx)
            self.nls.g_(self.nls.J, self._hx)
            if self.sparse:
                Jh = self.nls.J
                if not Jh.has_sorted_indices:
                    Jh.sort_indices()
                if Jh.nnz != self.J.nnz or not self._same_pattern(Jh):
                    # g! changed the sparsity pattern (setindex! into a sparse J, test/nonlinearsolvers.jl:526-530)
                    self.J.update_pattern(Jh.indptr, Jh.indices, Jh.data)
                    self._pattern = (Jh.indptr.copy(), Jh.indices.copy())
                else:
                    self.J.set_values(Jh.data)
            else:
                self.J.upload(self.nls.J)
        else:
            self.nls.g_(self.J, x)

    def _same_pattern(self, Jh) -> bool:
        pat = getattr(self, "_pattern", None)
        if pat is None:
            self._pattern = pat = (Jh.indptr.copy(), Jh.indices.copy())
            return True       # the device image was built from this very pattern in __init__
        return np.array_equal(pat[0], Jh.indptr) and np.array_equal(pat[1], Jh.indices)

    def finish(self, x: DeviceVector):
        if self.host:
            x.download(self.nls.x)
            self.fcur.download(self.nls.y)
            return self.nls.x
        return x


def _bounds(ctx, x: DeviceVector, lower, upper):
    n = len(x)
    lo = np.asarray(lower, dtype=np.float64) if lower is not None and len(lower) else None
    hi = np.asarray(upper, dtype=np.float64) if upper is not None and len(upper) else None
    if (lo is not None and lo.size != n) or (hi is not None and hi.size != n):
        raise ValueError("Bounds must either be empty or of the same length as the number of parameters.")
    xh = x.download()
    if (lo is not None and not np.all(xh >= lo)) or (hi is not None and not np.all(xh <= hi)):
        raise ValueError("Initial guess must be within bounds.")
    dlo = DeviceVector(ctx, n, lo) if lo is not None else None
    dhi = DeviceVector(ctx, n, hi) if hi is not None else None
    return dlo, dhi


def _box_project(ctx, dx, x, dlo, dhi):
    if dlo is None and dhi is None:
        return
    from ._lib import check, lib
    check(lib().lso_vec_box_project(ctx.handle, len(x), dx.ptr, x.ptr, dlo.ptr if dlo else None,
                                    dhi.ptr if dhi else None), ctx.handle)


def _maxabs_projected_gradient(ctx, g, x, dlo, dhi) -> float:
    import ctypes as C
    from ._lib import check, lib
    out = C.c_double()
    check(lib().lso_vec_maxabs_projected(ctx.handle, len(g), g.ptr, x.ptr, dlo.ptr if dlo else None,
                                         dhi.ptr if dhi else None, C.byref(out)), ctx.handle)
    return out.value


def _gradient_norm_async(ctx, g, x, dlo, dhi):
    """maxabs_projected_gradient (utils.jl:38-55) enqueued into the step's device scalars; read back by `_step_tail`."""
    from ._lib import check, lib
    check(lib().lso_lm_gradient_norm_async(ctx.handle, len(g), g.ptr, x.ptr, dlo.ptr if dlo else None,
                                           dhi.ptr if dhi else None), ctx.handle)


def _step_tail(ctx, J, dx, fcur, ftrial, fpredict, allreduce):
    """(f1) sum(abs2, ftrial), fpredict = J*dx - fcur with sum(abs2, fpredict), maximum(abs, dx) and the projected
    gradient norm with ONE host synchronisation (lso_lm_step_tail)."""
    import ctypes as C
    from ._lib import check, lib
    out = (C.c_double * 4)()
    if isinstance(J, CSCMatrix):
        dJ, ld, csc = None, 0, J.handle
    else:
        dJ, ld, csc = J.ptr, J.ld, None
    check(lib().lso_lm_step_tail(ctx.handle, J.m, J.n, dJ, ld, csc, dx.ptr, fcur.ptr, ftrial.ptr, fpredict.ptr,
                                 int(bool(allreduce)), C.addressof(out)), ctx.handle)
    return out[0], out[1], out[2], out[3]


def assess_convergence(dx, maxabs_gr, ssr, trial_ssr, xtol, ftol, grtol, step_accepted):
    """src/utils/utils.jl:7-31 — an if / elseif chain: at most one flag is set.  `dx` is the step (DeviceVector) or its
    max-norm already reduced on the device (float)."""
    x_c = f_c = g_c = False
    if step_accepted and abs(trial_ssr - ssr) <= ftol * (abs(ssr) + ftol):
        f_c = True
    elif (dx if isinstance(dx, float) else dx.maxabs()) <= xtol:
        x_c = True
    elif maxabs_gr <= grtol:
        g_c = True
    return x_c, f_c, g_c, (x_c or f_c or g_c)


def _lm_damping(ctx, dtd: DeviceVector, inv_delta: float):
    from ._lib import check, lib
    check(lib().lso_lm_damping(ctx.handle, len(dtd), dtd.ptr, MIN_DIAGONAL, MAX_DIAGONAL, inv_delta), ctx.handle)


# ---- LevenbergMarquardt (levenberg_marquardt.jl:39-144) ----------------------------------------------------
class LMRun:
    """State of one `optimize!` run with LevenbergMarquardt; `iterate()` is one pass of the `while` body
    (levenberg_marquardt.jl:72-140).  When the context carries a communicator (row-sharded J, one rank per GPU)
    the m-dimension reductions are all-reduced over NCCL: x, δ, dtd are replicated, f / J are this rank's rows."""

    def __init__(self, anls: "_Allocated", x_tol=1e-8, f_tol=1e-8, g_tol=1e-8, iterations=1000, Δ=10.0,
                 store_trace=False, lower=None, upper=None, record_steps=False):
        self.anls = anls
        ctx, n, m = anls.ctx, anls.n, anls.m
        self.ctx = ctx
        self.x_tol, self.f_tol, self.g_tol, self.iterations = x_tol, f_tol, g_tol, iterations
        self.store_trace, self.record_steps = store_trace, record_steps
        w = anls.workspace("lm", lambda: dict(dx=DeviceVector(ctx, n), dtd=DeviceVector(ctx, n),
                                              ftrial=DeviceVector(ctx, m), fpredict=DeviceVector(ctx, m),
                                              red=DeviceVector(ctx, 8)))
        self.dx, self.dtd, self.ftrial, self.fpredict, self.red = w["dx"], w["dtd"], w["ftrial"], w["fpredict"], w["red"]
        self.grad = anls.workspace("lm_grad", lambda: dict(g=DeviceVector(ctx, n)))["g"]
        self.dlo, self.dhi = _bounds(ctx, anls.x, lower, upper)
        self.sharded = bool(anls.sharded)       # ONE source of truth: the flag the problem was allocated with
        self.Δ = float(Δ)
        self.decrease_factor = 2.0
        self.f_calls = self.g_calls = self.mul_calls = 0
        self.converged = self.x_converged = self.f_converged = self.g_converged = False
        anls.f(anls.fcur, anls.x)
        self.f_calls += 1
        self.ssr = self._allsum(anls.fcur.sumabs2())
        self.maxabs_gr = math.inf
        self.need_jacobian = True
        self.it = 0
        self.tr = [OptimizationState(0, self.ssr, self.maxabs_gr)] if store_trace else []
        self.deltas = []

    def _allsum(self, *vals):
        """Sum scalars over ranks (identity on one GPU)."""
        if not self.sharded:
            return vals[0] if len(vals) == 1 else vals
        buf = np.zeros(8)
        buf[:len(vals)] = vals
        self.red.upload(buf)
        self.ctx.allreduce(self.red)
        out = self.red.download()
        return float(out[0]) if len(vals) == 1 else tuple(float(v) for v in out[:len(vals)])

    def iterate(self):
        anls, ctx = self.anls, self.ctx
        x, fcur, J = anls.x, anls.fcur, anls.J
        dx, dtd, ftrial, fpredict = self.dx, self.dtd, self.ftrial, self.fpredict
        self.it += 1
        x.check_finite()
        same_J = not self.need_jacobian       # a rejected step: J and fcur are what the previous solve saw (:77-87)
        if self.need_jacobian:
            anls.g(x)
            self.g_calls += 1
            self.need_jacobian = False
        # :82 colsumabs2!(dtd, J) and :102 mul!(dtd, J', fcur) both read J and fcur, which do not change in between:
        # one pass over J produces both (the gradient is parked in `grad` until the solver is done with dtd)
        J.colsumabs2_and_grad(dtd, self.grad, fcur)
        if self.sharded:
            ctx.allreduce(dtd)
            ctx.allreduce(self.grad)
        _lm_damping(ctx, dtd, 1 / self.Δ)                     # :84-86
        _, lmiter = anls.solver.ldiv(dx, J, fcur, dtd, same_J=same_J)        # :87
        if self.record_steps:
            self.deltas.append(dx.download())
        _box_project(ctx, dx, x, self.dlo, self.dhi)          # :89-98
        self.mul_calls += lmiter
        dtd.copyto(self.grad)                                 # :102 gradient J'f (computed above)
        self.mul_calls += 1
        _gradient_norm_async(ctx, dtd, x, self.dlo, self.dhi)  # :104 (read back with the other scalars below)
        x.axpy(-1.0, dx)                                      # :106
        anls.f(ftrial, x)
        self.f_calls += 1
        # :110 sum(abs2, ftrial), :114-117 ||J δ - f||², maximum(abs, δx) for utils.jl:21 — one synchronisation; the two
        # m-dimension sums are all-reduced over the ranks on the device when J is row-sharded
        trial_ssr, predicted_ssr, maxabs_dx, self.maxabs_gr = _step_tail(ctx, J, dx, fcur, ftrial, fpredict, self.sharded)
        self.mul_calls += 1
        ssr = self.ssr
        predicted_reduction = abs(ssr - predicted_ssr)
        ρ = (ssr - trial_ssr) / predicted_reduction if predicted_reduction > 0 else 0.0
        step_accepted = ρ > MIN_STEP_QUALITY
        self.x_converged, self.f_converged, self.g_converged, self.converged = assess_convergence(
            maxabs_dx, self.maxabs_gr, ssr, trial_ssr, self.x_tol, self.f_tol, self.g_tol, step_accepted)
        if step_accepted:
            fcur.copyto(ftrial)
            self.ssr = trial_ssr
            t = 2.0 * ρ - 1.0
            self.Δ = min(self.Δ / max(1 / 3, 1.0 - t * t * t), MAX_DELTA)
            self.decrease_factor = 2.0
            self.need_jacobian = True
        else:
            x.axpy(1.0, dx)
            self.Δ = max(self.Δ / self.decrease_factor, MIN_DELTA)
            self.decrease_factor *= 2.0
        if self.store_trace:
            self.tr.append(OptimizationState(self.it, self.ssr, self.maxabs_gr))
        return step_accepted

    def result(self):
        xmin = self.anls.finish(self.anls.x)
        return LeastSquaresResult("LevenbergMarquardt", xmin, self.ssr, self.it, self.converged, self.x_converged,
                                  self.x_tol, self.f_converged, self.f_tol, self.g_converged, self.g_tol, self.tr,
                                  self.f_calls, self.g_calls, self.mul_calls, self.deltas)


def _optimize_lm(anls: _Allocated, **kw):
    run = LMRun(anls, **kw)
    while not run.converged and run.it < run.iterations:
        run.iterate()
    return run.result()


# ---- Dogleg (dogleg.jl:41-203) ---------------------------------------------------------------------------------
class DoglegRun:
    """State of one `optimize!` run with Dogleg; `iterate()` is one pass of the `while` body (dogleg.jl:77-199)."""

    def __init__(self, anls: "_Allocated", x_tol=1e-8, f_tol=1e-8, g_tol=1e-8, iterations=1000, Δ=1.0,
                 store_trace=False, lower=None, upper=None, record_steps=False):
        self.anls = anls
        ctx, n, m = anls.ctx, anls.n, anls.m
        self.ctx = ctx
        self.x_tol, self.f_tol, self.g_tol, self.iterations = x_tol, f_tol, g_tol, iterations
        self.store_trace, self.record_steps = store_trace, record_steps
        # AllocatedDogleg (dogleg.jl:7-30): allocated once per problem, reused by every run on it
        w = anls.workspace("dogleg", lambda: dict(dgn=DeviceVector(ctx, n), dgr=DeviceVector(ctx, n), dx=DeviceVector(ctx, n),
                                                  dtd=DeviceVector(ctx, n), ftrial=DeviceVector(ctx, m),
                                                  fpredict=DeviceVector(ctx, m)))
        self.dgn, self.dgr, self.dx, self.dtd = w["dgn"], w["dgr"], w["dx"], w["dtd"]
        self.ftrial, self.fpredict = w["ftrial"], w["fpredict"]
        self.dlo, self.dhi = _bounds(ctx, anls.x, lower, upper)
        self.sharded = bool(anls.sharded)       # rows of J / f are this rank's shard: the m-dimension sums are all-reduced
        self.red = anls.workspace("dogleg_red", lambda: dict(r=DeviceVector(ctx, 8)))["r"]
        self.Δ = float(Δ)
        self.reuse = False
        self.wnorm_dgn = self.wnorm_dgr = 0.0
        self.α = 0.0
        self.f_calls = self.g_calls = self.mul_calls = 0
        self.converged = self.x_converged = self.f_converged = self.g_converged = False
        anls.f(anls.fcur, anls.x)
        self.f_calls += 1
        self.ssr = self._allsum(anls.fcur.sumabs2())
        self.maxabs_gr = math.inf
        self.it = 0
        self.tr = [OptimizationState(0, self.ssr, self.maxabs_gr)] if store_trace else []
        self.deltas = []

    def _allsum(self, val: float) -> float:
        """Sum a scalar over the ranks (identity on one GPU)."""
        if not self.sharded:
            return val
        buf = np.zeros(8)
        buf[0] = val
        self.red.upload(buf)
        self.ctx.allreduce(self.red)
        return float(self.red.download()[0])

    def iterate(self):
        import ctypes as C
        from ._lib import check, lib
        anls, ctx, n = self.anls, self.ctx, self.anls.n
        x, fcur, J = anls.x, anls.fcur, anls.J
        dgn, dgr, dx, dtd, ftrial, fpredict = self.dgn, self.dgr, self.dx, self.dtd, self.ftrial, self.fpredict
        dlo, dhi = self.dlo, self.dhi
        self.it += 1
        x.check_finite()
        if not self.reuse:
            anls.g(x)
            self.g_calls += 1
            J.colsumabs2(dtd)                               # :85
            if self.sharded:
                ctx.allreduce(dtd)
            dtd.clamp(MIN_DIAGONAL, MAX_DIAGONAL)           # :90  (absolute floor, unlike LM)
            if self.it == 1:
                wnorm_x = wnorm(x, dtd)
                if wnorm_x > 0:
                    self.Δ *= wnorm_x
            J.mul_t(dgr, fcur, 1.0, 0.0)                    # :99
            if self.sharded:
                ctx.allreduce(dgr)
            self.mul_calls += 1
            _gradient_norm_async(ctx, dgr, x, dlo, dhi)       # :101 (read back by the step tail; kept across reuse)
            dgr.div_(dgr, dtd)                              # :105  δgr = D⁻¹ g
            self.wnorm_dgr = wnorm(dgr, dtd)
            J.mul(fpredict, dgr, 1.0, 0.0)                  # :109
            self.mul_calls += 1
            denom = self._allsum(fpredict.sumabs2())
            w2 = self.wnorm_dgr ** 2
            self.α = w2 / denom if denom != 0 else (math.nan if w2 == 0 else math.inf)   # :111 (0/0 -> NaN)
            dgn.fill(0.0)                                   # :114
            _, ls_iter = anls.solver.ldiv(dgn, J, fcur)     # :115
            if self.record_steps:
                self.deltas.append(dgn.download())
            self.mul_calls += ls_iter
            self.wnorm_dgn = wnorm(dgn, dtd)
        # δx: Gauss-Newton inside / scaled Cauchy / dogleg blend (:120-145)
        out = C.c_double()
        check(lib().lso_dogleg_blend(ctx.handle, n, dx.ptr, dgn.ptr, dgr.ptr, dtd.ptr, self.Δ, self.α, self.wnorm_dgn,
                                     self.wnorm_dgr, C.byref(out)), ctx.handle)
        wnorm_dx = out.value
        _box_project(ctx, dx, x, dlo, dhi)                  # :148-157
        x.axpy(-1.0, dx)                                    # :160
        anls.f(ftrial, x)
        self.f_calls += 1
        # :168 sum(abs2, ftrial), :171-174 ||J δ - f||², maximum(abs, δx), projected gradient norm: one synchronisation
        trial_ssr, predicted_ssr, maxabs_dx, self.maxabs_gr = _step_tail(ctx, J, dx, fcur, ftrial, fpredict, self.sharded)
        self.mul_calls += 1
        ssr = self.ssr
        predicted_reduction = abs(ssr - predicted_ssr)
        ρ = (ssr - trial_ssr) / predicted_reduction if predicted_reduction > 0 else 0.0
        step_accepted = ρ >= MIN_STEP_QUALITY
        self.x_converged, self.f_converged, self.g_converged, self.converged = assess_convergence(
            maxabs_dx, self.maxabs_gr, ssr, trial_ssr, self.x_tol, self.f_tol, self.g_tol, step_accepted)
        if step_accepted:
            self.reuse = False
            fcur.copyto(ftrial)
            self.ssr = trial_ssr
        else:
            self.reuse = True
            x.axpy(1.0, dx)
        if ρ < DECREASE_THRESHOLD:
            self.Δ = max(MIN_DELTA, self.Δ * 0.5)
        elif ρ > INCREASE_THRESHOLD:
            self.Δ = max(self.Δ, 3.0 * wnorm_dx)
        if self.store_trace:
            self.tr.append(OptimizationState(self.it, self.ssr, self.maxabs_gr))
        return step_accepted

    def result(self):
        xmin = self.anls.finish(self.anls.x)
        return LeastSquaresResult("Dogleg", xmin, self.ssr, self.it, self.converged, self.x_converged, self.x_tol,
                                  self.f_converged, self.f_tol, self.g_converged, self.g_tol, self.tr, self.f_calls,
                                  self.g_calls, self.mul_calls, self.deltas)


def _optimize_dogleg(anls: _Allocated, **kw):
    run = DoglegRun(anls, **kw)
    while not run.converged and run.it < run.iterations:
        run.iterate()
    return run.result()


AUTO_CHUNK_SHARES = (0.30, 0.30, 0.25, 0.15)


def host_chunk_rows(m: int, chunks) -> list:
    """Row counts of the chunks in which a host-resident J crosses PCIe (`lso_qr_factor_keep_host_chunks`), in transfer
    order: an int = that many (near-)equal chunks, a sequence of positive shares = those fractions of the m rows.  Every
    row belongs to exactly one chunk and no chunk is empty."""
    if isinstance(chunks, (int, np.integer)):
        k = int(chunks)
        if k <= 1:
            return [int(m)]
        k = min(k, int(m))
        return [(m * (i + 1)) // k - (m * i) // k for i in range(k)]
    shares = [float(c) for c in chunks]
    if not shares or min(shares) <= 0:
        raise ValueError("chunk shares must be positive")
    edges = np.rint(np.cumsum([0.0] + shares) / float(sum(shares)) * m).astype(np.int64)
    edges[-1] = m
    rows = [int(b - a) for a, b in zip(edges[:-1], edges[1:])]
    if min(rows) <= 0:
        raise ValueError(f"chunk shares {shares} leave an empty chunk at m = {m}")
    return rows


class HostStep:
    """Hot-path body of one LevenbergMarquardt iteration driven from HOST buffers (the e2e path of bench.py and
    what the Julia glue does when `J` / `f` are plain host Arrays): H2D of J and f, colsumabs2! + damping (LM:82-86),
    the damped solve (:87), J'f and its max-norm (:102-104), ||Jδ - f||² (:114-117), then D2H of δ and the scalars."""

    def __init__(self, anls: "_Allocated", chunks: int | None = None):
        from ._lib import check, lib
        self.anls, self.ctx = anls, anls.ctx
        self.lib, self.check = lib(), check
        ctx, n, m = anls.ctx, anls.n, anls.m
        w = anls.workspace("lm", lambda: dict(dx=DeviceVector(ctx, n), dtd=DeviceVector(ctx, n),
                                              ftrial=DeviceVector(ctx, m), fpredict=DeviceVector(ctx, m),
                                              red=DeviceVector(ctx, 8)))
        self.dx, self.dtd, self.fpredict, self.red = w["dx"], w["dtd"], w["fpredict"], w["red"]
        self.grad = anls.workspace("lm_grad", lambda: dict(g=DeviceVector(ctx, n)))["g"]
        self.sharded = bool(anls.sharded)
        # Row chunks of the host-fed factorisation.  `chunks`: None = automatic, an int = that many (near-)equal chunks, a
        # list of fractions = those shares of the rows, in transfer order.  Automatic: four chunks with shrinking sizes when
        # J is tall enough that a chunk still has many more rows than the per-panel latency floor is worth — chunks are
        # factorised round-robin in two workspaces, so a small LAST chunk leaves little work after the last byte has landed
        # (measured at 100 000 x 1 000: see tools/probe_e2e_chunks.py) — else the plain upload + solve.
        auto = chunks is None
        if auto:
            ok = not self.sharded and isinstance(anls.solver, DenseQRAllocatedSolver) and m >= 40 * n and m >= 50000
            chunks = list(AUTO_CHUNK_SHARES) if ok else 1
        self.chunk_rows = host_chunk_rows(m, chunks)
        self.chunks = len(self.chunk_rows)
        if self.chunks > 1:
            self.chunk_solver = DenseQRAllocatedSolver(ctx, max(max(self.chunk_rows), n), n, damped=False)

    def run(self, hJ_ptr: int, hf_ptr: int, Δ: float, dx_host: np.ndarray):
        a, ctx, h = self.anls, self.ctx, self.ctx.handle
        J, fcur, dx, dtd = a.J, a.fcur, self.dx, self.dtd
        if self.chunks > 1:
            # J and f cross PCIe in row chunks, each chunk is factorised (undamped) while the next one is in flight; the
            # damping joins in the stacked finish, so colsumabs2! can wait for the whole J (levenberg_marquardt.jl:82-87)
            self.chunk_solver.factor_keep_host_chunks(self.chunk_rows, hJ_ptr, a.m, hf_ptr, J, fcur)
            J.colsumabs2_and_grad(dtd, self.grad, fcur)
            _lm_damping(ctx, dtd, 1 / Δ)
            self.chunk_solver.solve_kept(dx, dtd)
        else:
            self.check(self.lib.lso_upload_async(h, J.ptr, hJ_ptr, a.m * a.n * 8), h)
            self.check(self.lib.lso_upload_async(h, fcur.ptr, hf_ptr, a.m * 8), h)
            J.colsumabs2_and_grad(dtd, self.grad, fcur)          # LM:82 and LM:102 in one pass over J
            if self.sharded:
                ctx.allreduce(dtd)
                ctx.allreduce(self.grad)
            _lm_damping(ctx, dtd, 1 / Δ)
            a.solver.ldiv(dx, J, fcur, dtd)
        dtd.copyto(self.grad)
        _gradient_norm_async(ctx, dtd, a.x, None, None)
        # ssr = sum(abs2, fcur), ||J δ - f||², max|δ|, max|J'f|: one synchronisation (all-reduced over ranks when sharded)
        ssr, predicted_ssr, maxabs_dx, maxabs_gr = _step_tail(ctx, J, dx, fcur, fcur, self.fpredict, self.sharded)
        dx.download(dx_host)
        return {"ssr": ssr, "predicted_ssr": predicted_ssr, "maxabs_gr": maxabs_gr, "maxabs_dx": maxabs_dx}


def optimize_(nls: LeastSquaresProblem, optimizer: Optional[AbstractOptimizer] = None, **kwargs) -> LeastSquaresResult:
    """`optimize!(nls, optimizer; kwargs...)` — types.jl:207-209 + LeastSquaresProblemAllocated (:152-157)."""
    anls = nls if isinstance(nls, _Allocated) else allocate(nls, optimizer)
    # show_trace / show_every (utils.jl:97-128) are host-side display, off the path: the states are stored during the run
    # and printed in the reference's format afterwards (the reference prints them as it goes)
    show = bool(kwargs.pop("show_trace", False))
    every = int(kwargs.pop("show_every", 1) or 1)
    if not show:
        if isinstance(anls.optimizer, LevenbergMarquardt):
            return _optimize_lm(anls, **kwargs)
        return _optimize_dogleg(anls, **kwargs)
    keep = bool(kwargs.get("store_trace", False))
    kwargs["store_trace"] = True
    r = _optimize_lm(anls, **kwargs) if isinstance(anls.optimizer, LevenbergMarquardt) else _optimize_dogleg(anls, **kwargs)
    print(format_trace(r.tr, every), end="")
    if not keep:
        r.tr = []
    return r


def allocate(nls: LeastSquaresProblem, optimizer: Optional[AbstractOptimizer] = None, sharded: bool = False) -> "_Allocated":
    """`LeastSquaresProblemAllocated(nls, optimizer)` — types.jl:152-157: default solver/optimizer dispatch, then the
    optimizer and solver workspaces.  The result can be passed to `optimize_` repeatedly (no allocation per run)."""
    solver = default_solver(optimizer.solver if optimizer is not None else None, nls.J)
    optimizer = default_optimizer(optimizer, solver)
    return _Allocated(nls, optimizer, sharded=sharded)


def optimize(f: Callable, x0, optimizer: AbstractOptimizer, **kwargs) -> LeastSquaresResult:
    """`optimize(f, x, t)` — types.jl:182-184: wraps f into f!(out, x) = copyto!(out, f(x))."""
    x0 = np.array(x0, dtype=np.float64)
    m = np.atleast_1d(f(x0)).size

    def f_(out, x):
        out[:] = np.atleast_1d(f(x))

    return optimize_(LeastSquaresProblem(x=x0.copy(), f_=f_, output_length=m), optimizer, **kwargs)

"""Allocated solvers: the `AbstractAllocatedSolver` plugin surface of the reference, backed by the C ABI.

Each class mirrors one reference struct and its `ldiv!` methods; `ldiv` returns `(x, n_mul)` exactly like
the reference so `mul_calls` bookkeeping in the optimizers is unchanged.

  DenseQRAllocatedSolver        src/solver/dense_qr.jl:6-88
  DenseCholeskyAllocatedSolver  src/solver/dense_cholesky.jl:7-59
  LSMRAllocatedSolver           src/solver/iterative_lsmr.jl:161-198
  LSMRDampenedAllocatedSolver   src/solver/iterative_lsmr.jl:221-259
"""
from __future__ import annotations

import ctypes as C
import weakref

from ._lib import check, lib
from .device import Context, CSCMatrix, DenseMatrix, DeviceVector

LSO_ERR_UNSUPPORTED = -5

LSO_SOLVER_QR = 1
LSO_SOLVER_CHOLESKY = 2


class _DenseWorkspace:
    def __init__(self, ctx: Context, m: int, n: int, kind: int, damped: bool):
        self.ctx, self.m, self.n, self.damped = ctx, m, n, damped
        self._h = C.c_void_p()
        check(lib().lso_dense_ws_create(ctx.handle, m, n, kind, int(damped), C.byref(self._h)), ctx.handle)
        self._fin = weakref.finalize(self, lib().lso_dense_ws_destroy, self._h)

    def factor(self):
        """n x n upper-triangular factor of the last solve (tests)."""
        import numpy as np

        R = np.zeros((self.n, self.n), order="F")
        check(lib().lso_dense_ws_get_factor(self._h, R.ctypes.data), self.ctx.handle)
        return R


class DenseQRAllocatedSolver(_DenseWorkspace):
    """dense_qr.jl: Dogleg{QR} workspace (m x n, :25-28) or LevenbergMarquardt{QR} workspace ((m+n) x n, :50-54).

    `reuse` (f3; levenberg_marquardt.jl:77-87 re-solves with the same J and f after a rejected step).  Unless "off", a
    re-solve with a LARGER damping (what a rejected LM step is) re-damps the triangular factor of the previous solve
    (lso_qr_solve_redamp: QR of the banded 2n x n stack [R; sqrt(D_new - D_last)], no pass over J).  When that does not
    apply (the damping shrank):
      "lazy"    a fresh J is solved by the direct QR of [J; sqrt(D)]; the FIRST re-solve with the same J factors J once
                (lso_qr_factor_keep) and every re-solve costs only the banded 2n x n stack QR (lso_qr_solve_kept)
      "always"  every fresh J is factored undamped and finished through the stack
      "off"     every solve refactors, like the reference
    On a row-sharded workspace the gathered R factors are always kept: a re-solve has no local QR and no collective."""

    def __init__(self, ctx: Context, m: int, n: int, damped: bool, sharded: bool = False, reuse: str = "lazy"):
        super().__init__(ctx, m, n, LSO_SOLVER_QR, damped)
        self.last_rank = n
        self.sharded = sharded      # J, y are this rank's row shard: TSQR over the context's communicator
        self.reuse = reuse if m >= n else "off"
        self._kept = False
        self._have_damped = False
        self.solves_direct = self.solves_kept = self.factor_keeps = self.solves_redamped = 0

    def ldiv(self, x: DeviceVector, J: DenseMatrix, y: DeviceVector, damp: DeviceVector | None = None, same_J: bool = False):
        rank = C.c_int()
        h = self.ctx.handle
        dptr = damp.ptr if damp is not None else None
        use_keep = self.reuse != "off" and damp is not None
        if use_keep and same_J and self._have_damped:
            # a rejected step: same J and f, larger damping — re-damp the triangular factor of the last solve
            st = lib().lso_qr_solve_redamp(self._h, dptr, x.ptr, C.byref(rank))
            if st == 0:
                self.solves_redamped += 1
                self.last_rank = rank.value
                return x, 1
            if st != LSO_ERR_UNSUPPORTED:
                check(st, h)
        if not same_J:
            check(lib().lso_qr_kept_invalidate(self._h), h)
            self._kept = False
        self._have_damped = False
        done = False
        if use_keep and same_J and self._kept:
            st = lib().lso_qr_solve_kept(self._h, dptr, x.ptr, C.byref(rank))
            if st == 0:
                self.solves_kept += 1
                done = True
            elif st != LSO_ERR_UNSUPPORTED:      # (the panel-pipelined sharded solve keeps no gathered factors)
                check(st, h)
        if done:
            pass
        elif self.sharded:
            check(lib().lso_qr_solve_sharded(self._h, J.ptr, J.ld, y.ptr, dptr, x.ptr, C.byref(rank)), h)
            self._kept = True
            self.solves_direct += 1
        elif use_keep and (same_J or self.reuse == "always"):
            check(lib().lso_qr_factor_keep(self._h, J.ptr, J.ld, y.ptr), h)
            check(lib().lso_qr_solve_kept(self._h, dptr, x.ptr, C.byref(rank)), h)
            self._kept = True
            self.factor_keeps += 1
            self.solves_kept += 1
        else:
            check(lib().lso_qr_solve(self._h, J.ptr, J.ld, y.ptr, dptr, x.ptr, C.byref(rank)), h)
            self._kept = False
            self.solves_direct += 1
        self._have_damped = damp is not None
        self.last_rank = rank.value
        return x, 1


    # ---- host-fed, chunk-pipelined form (the end-to-end path when J lives in host memory) ----
    def factor_keep_host(self, m_total: int, hJ_ptr: int, ld_h: int, hy_ptr: int, J: DenseMatrix, y: DeviceVector):
        """QR of [J | y] from HOST memory in row chunks of this workspace's m rows, each chunk factorised while the next
        one is still crossing PCIe (lso_qr_factor_keep_host); J and y also land in the device arrays given."""
        check(lib().lso_qr_factor_keep_host(self._h, m_total, hJ_ptr, ld_h, hy_ptr, J.ptr, J.ld, y.ptr), self.ctx.handle)
        self._kept, self._have_damped = True, False

    def factor_keep_host_chunks(self, chunk_rows, hJ_ptr: int, ld_h: int, hy_ptr: int, J: DenseMatrix, y: DeviceVector):
        """The same with an explicit list of chunk sizes, sent in that order (lso_qr_factor_keep_host_chunks)."""
        rows = (C.c_int64 * len(chunk_rows))(*[int(r) for r in chunk_rows])
        check(lib().lso_qr_factor_keep_host_chunks(self._h, len(chunk_rows), C.addressof(rows), hJ_ptr, ld_h, hy_ptr, J.ptr, J.ld,
                                                   y.ptr), self.ctx.handle)
        self._kept, self._have_damped = True, False

    def solve_kept(self, x: DeviceVector, damp: DeviceVector | None):
        rank = C.c_int()
        check(lib().lso_qr_solve_kept(self._h, damp.ptr if damp is not None else None, x.ptr, C.byref(rank)), self.ctx.handle)
        self.last_rank = rank.value
        self._have_damped = damp is not None
        self.solves_kept += 1
        return x, 1


class DenseCholeskyAllocatedSolver(_DenseWorkspace):
    """dense_cholesky.jl:19-21 (one n x n workspace for both optimizers).  After a rejected step (same J and f, new
    damping) the kept J'J and J'f are re-used: no pass over J, no collective (`reuse=False` refactors like the reference)."""

    def __init__(self, ctx: Context, m: int, n: int, damped: bool, sharded: bool = False, reuse: bool = True):
        super().__init__(ctx, m, n, LSO_SOLVER_CHOLESKY, damped)
        self.sharded = sharded      # J, y are this rank's row shard: one all-reduce of the packed [upper(J'J) | J'y]
        self.reuse = reuse
        self._kept = False
        self.solves_direct = self.solves_kept = 0

    def ldiv(self, x: DeviceVector, J: DenseMatrix, y: DeviceVector, damp: DeviceVector | None = None, same_J: bool = False):
        dptr = damp.ptr if damp is not None else None
        if self.reuse and same_J and self._kept:
            check(lib().lso_chol_solve_kept(self._h, dptr, x.ptr), self.ctx.handle)
            self.solves_kept += 1
            return x, 1
        fn = lib().lso_chol_solve_sharded if self.sharded else lib().lso_chol_solve
        self._kept = False
        check(fn(self._h, J.ptr, J.ld, y.ptr, dptr, x.ptr), self.ctx.handle)
        self._kept = True
        self.solves_direct += 1
        return x, 1


PRECOND_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)


class _LSMRWorkspace:
    """`preconditioner` mirrors `LSMR(preconditioner!, P)` (types.jl:81-85, README.md:47):
      None                       the default diagonal preconditioner 1/sqrt(colsumabs2(J) + damp) (iterative_lsmr.jl:130-138)
      callable(x, J, damp) -> DeviceVector     `preconditioner!` producing the vector of an InverseDiagonal each solve
      (callable(x, J, damp), apply(out_ptr, in_ptr, n))   a general P: `apply` is `ldiv!(out, P, in)` on device pointers"""

    def __init__(self, ctx: Context, m: int, n: int, damped: bool, preconditioner=None, sharded: bool = False,
                 m_total: int = 0):
        self.ctx, self.m, self.n, self.damped = ctx, m, n, damped
        # sharded: J / y are this rank's rows (CSC), everything of length n is replicated (lso_lsmr_solve_sharded)
        self.sharded, self.m_total = bool(sharded), int(m_total)
        self._h = C.c_void_p()
        check(lib().lso_lsmr_ws_create(ctx.handle, m, n, int(damped), C.byref(self._h)), ctx.handle)
        self._fin = weakref.finalize(self, lib().lso_lsmr_ws_destroy, self._h)
        self.last_iters = 0
        self.last_istop = 0
        self.preconditioner = preconditioner

    def stats(self):
        """(kernel launches, host synchronisations) of the last solve."""
        a, b = C.c_int64(), C.c_int64()
        check(lib().lso_lsmr_ws_stats(self._h, C.byref(a), C.byref(b)), self.ctx.handle)
        return a.value, b.value

    def _solve(self, x, J, y, damp, atol, btol, conlim, maxiter):
        iters, istop = C.c_int64(), C.c_int()
        csc = J.handle if isinstance(J, CSCMatrix) else None
        dj = J.ptr if isinstance(J, DenseMatrix) else None
        ld = J.ld if isinstance(J, DenseMatrix) else 0
        pdiag, pfn, keep = None, None, None
        if self.preconditioner is not None:
            if isinstance(self.preconditioner, tuple):
                update, apply = self.preconditioner
                update(x, J, damp)                       # preconditioner!(P, x, J, damp)

                def _cb(user, n, d_in, d_out):
                    try:
                        apply(d_out, d_in, n)
                        return 0
                    except Exception:
                        return 1
                keep = PRECOND_FN(_cb)
                pfn = C.cast(keep, C.c_void_p)
            else:
                keep = self.preconditioner(x, J, damp)   # keep the vector alive until the solve has been enqueued and read back
                pdiag = keep.ptr
        if self.sharded:
            if csc is None or pfn is not None:
                raise ValueError("the row-sharded LSMR needs a CSCMatrix and a diagonal (or the default) preconditioner")
            check(lib().lso_lsmr_solve_sharded(self._h, csc, y.ptr, damp.ptr if damp is not None else None, x.ptr,
                                               atol, btol, conlim, maxiter, self.m_total, pdiag, C.byref(iters),
                                               C.byref(istop)), self.ctx.handle)
            self.last_iters, self.last_istop = iters.value, istop.value
            return x, 2 * iters.value
        check(lib().lso_lsmr_solve_ex(self._h, csc, dj, ld, y.ptr, damp.ptr if damp is not None else None, x.ptr,
                                      atol, btol, conlim, maxiter, pdiag, pfn, None, C.byref(iters), C.byref(istop)),
              self.ctx.handle)
        self.last_iters, self.last_istop = iters.value, istop.value
        return x, 2 * iters.value      # ch.mvps = 2 * iter (lsmr.jl:236)


class LSMRAllocatedSolver(_LSMRWorkspace):
    """iterative_lsmr.jl:161-198 — undamped, lsmr! defaults atol = btol = 1e-6, conlim = 1e8 (lsmr.jl:53-55)."""

    def __init__(self, ctx, m, n, preconditioner=None, sharded=False, m_total=0):
        super().__init__(ctx, m, n, False, preconditioner, sharded, m_total)

    def ldiv(self, x, J, y, damp=None, same_J=False):
        assert damp is None
        return self._solve(x, J, y, None, 1e-6, 1e-6, 1e8, 0)


class LSMRDampenedAllocatedSolver(_LSMRWorkspace):
    """iterative_lsmr.jl:221-259 — damped, btol = 0.5 (:255); `damp` is overwritten by sqrt(damp) (:252)."""

    def __init__(self, ctx, m, n, preconditioner=None, sharded=False, m_total=0):
        super().__init__(ctx, m, n, True, preconditioner, sharded, m_total)

    def ldiv(self, x, J, y, damp, same_J=False):
        return self._solve(x, J, y, damp, 1e-6, 0.5, 1e8, 0)

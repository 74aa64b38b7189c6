"""lsob200 — B200-native inner solver for LeastSquaresOptim.jl's per-iteration linear-algebra hot path.

Layout:
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (liblsob200.so, include/lsob200.h)
  _lib.py          ctypes binding generated from the header
  device.py        device vector / dense / CSC operator wrappers (the reference's duck-type interfaces)
  solvers.py       AbstractAllocatedSolver plugin surface (QR / Cholesky / LSMR workspaces + ldiv!)
  api.py           LeastSquaresProblem / optimize! / Dogleg / LevenbergMarquardt mirror (host control flow)
"""
from ._lib import (DimensionMismatch, IsFiniteException, LsoError, PosDefException, RankDeficientException, lib)
from .device import Context, CSCMatrix, DenseMatrix, DeviceVector, wdot, wnorm
from .solvers import (DenseCholeskyAllocatedSolver, DenseQRAllocatedSolver, LSMRAllocatedSolver,
                      LSMRDampenedAllocatedSolver)
from .api import (LSMR, QR, Cholesky, Dogleg, DoglegRun, LeastSquaresProblem, LeastSquaresResult, LevenbergMarquardt, LMRun, HostStep,
                  allocate, optimize, optimize_)

__all__ = [n for n in dir() if not n.startswith("_")]

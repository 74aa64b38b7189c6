"""Small Cholesky solves (potrf panel + multi-CTA triangular solves, both orientations) for compute-sanitizer racecheck / memcheck."""
import sys
import numpy as np
sys.path.insert(0, ".")
import lsob200 as L
sys.path.insert(0, "tests")
from oracle import reference_port as O
ctx = L.Context.default(0)
for m, n in ((300, 100), (900, 333), (2000, 1000)):
    rng = np.random.default_rng(n)
    Jh = np.asfortranarray(rng.standard_normal((m, n)))
    yh = rng.standard_normal(m)
    damp = np.einsum("ij,ij->j", Jh, Jh) / 7
    x = L.DeviceVector(ctx, n)
    L.DenseCholeskyAllocatedSolver(ctx, m, n, damped=True).ldiv(x, L.DenseMatrix(ctx, m, n, Jh), L.DeviceVector(ctx, m, yh), L.DeviceVector(ctx, n, damp))
    xr = O.chol_ldiv(Jh, yh, damp.copy())
    print(m, n, "rel err", np.linalg.norm(x.download() - xr) / np.linalg.norm(xr))

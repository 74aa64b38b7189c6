"""ORACLE (TEST INFRASTRUCTURE): golden vector for BASELINE.json configs[4] at its NAMED size — the first Gauss-Newton
solve of Dogleg(QR()) on the bounded synthetic fit, n = 10 000, m = 200 000 (dogleg.jl:115 -> dense_qr.jl:30-42 ->
[stdlib] qr!(·, ColumnNorm()) + ldiv! = LAPACK dgeqp3 + dlaic1 + dormqr + dtrtrs, i.e. dgelsy with rcond = min(m,n)·eps).

dgeqp3 is BLAS-2 bound: ~1 h on 8 cores for this shape, which is why the result is frozen here instead of being
recomputed by the test.  J(x0) and f(x0) come from the counter-based generators of oracle/synth_ref.py (bit-identical
to csrc/synth.cu, checked by tests/test_gpu_synth.py), so the GPU test rebuilds exactly the same inputs on the device.

    python oracle/make_golden_c5.py [m n]      ->  tests/golden/c5_first_solve.npz   (oracle-derived; no Julia here)
"""
import os
import sys
import time

import numpy as np
from scipy.linalg import lapack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth_ref as S  # noqa: E402

SEED = 20240607 + 5
C_MODEL, NOISE = 0.1, 1e-3


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "tests", "golden", "c5_first_solve.npz")
    t0 = time.time()
    A = S.dense_matrix(m, n, SEED)
    xstar = S.vector(n, SEED + 11)
    t = A @ xstar
    b = (t + C_MODEL * t * t) + NOISE * S.vector(m, SEED + 12)
    x0 = xstar + 0.1 * S.vector(n, SEED + 13)
    t = A @ x0
    f = (t + (C_MODEL * t) * t) - b
    w = 1.0 + (2.0 * C_MODEL) * t
    np.multiply(A, w[:, None], out=A)                  # J(x0) = diag(1 + 2 c t) A, in place (16 GB at full size)
    J = A
    g = J.T @ f                                        # kept for the optimality check below
    print(f"inputs built in {time.time() - t0:.0f} s", flush=True)
    u = np.zeros((max(m, n), 1), order="F")
    u[:m, 0] = f
    rcond = min(m, n) * np.finfo(float).eps
    lwork = max(1, 4 * (m + n) * 16 + 3 * n + 64)
    t1 = time.time()
    v, x, jpvt, rank, info = lapack.dgelsy(J, u, np.zeros(n, dtype=np.int32), rcond, lwork, overwrite_a=1, overwrite_b=1)
    dt = time.time() - t1
    assert info == 0
    delta = np.array(x[:n, 0])
    print(f"dgelsy {m}x{n}: {dt:.0f} s, rank {rank}", flush=True)
    np.savez_compressed(out, delta=delta, rank=np.int64(rank), m=np.int64(m), n=np.int64(n), seed=np.int64(SEED),
                        c=np.float64(C_MODEL), noise=np.float64(NOISE), dgelsy_seconds=np.float64(dt),
                        cores=np.int64(os.cpu_count()), grad_norm=np.float64(np.linalg.norm(g)),
                        f_norm=np.float64(np.linalg.norm(f)))
    print("wrote", out, flush=True)


if __name__ == "__main__":
    main()

"""ORACLE side of the synthetic workloads (TEST INFRASTRUCTURE): numpy replica of the counter-based generators
in leastsquaresoptim.jl_b200/csrc/synth.cu, bit-identical by construction (integer hash -> exact 52-bit
fraction -> power-of-two column scale), and the polynomial residual model of SURVEY.md §8d:

    r(x) = t + c t^2 - b,   t = A x,   J(x) = diag(1 + 2 c t) A

Used by tests/ (GPU-vs-CPU bit check of the generators, parity at reduced size) and by bench.py's
cpu_baseline / --impl reference legs.
"""
import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_K1, _C1 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0x632BE59BD9B4E019)
_K2, _C2 = np.uint64(0xC2B2AE3D27D4EB4F), np.uint64(0x165667B19E3779F9)


def _mix64(z):
    z = z ^ (z >> np.uint64(30))
    z = z * _M1
    z = z ^ (z >> np.uint64(27))
    z = z * _M2
    z = z ^ (z >> np.uint64(31))
    return z


def hash64(seed, i, j):
    with np.errstate(over="ignore"):
        i = np.asarray(i, dtype=np.uint64)
        j = np.asarray(j, dtype=np.uint64)
        return _mix64(np.uint64(seed) ^ _mix64(i * _K1 + _C1) ^ _mix64(j * _K2 + _C2))


def unif(h):
    return (h >> np.uint64(12)).astype(np.float64) * (1.0 / 2251799813685248.0) - 1.0


def colscale(seed, n):
    e = (hash64(seed + 1, np.uint64(0x5CA1E), np.arange(n, dtype=np.uint64)) % np.uint64(13)).astype(np.int64) - 6
    return np.exp2(e.astype(np.float64))


def dense_matrix(m, n, seed, row_offset=0, chunk=4096):
    """A[i, j] = unif(hash(seed, row_offset + i, j)) * colscale(seed, j); column-major."""
    A = np.empty((m, n), order="F")
    s = colscale(seed, n)
    rows = (np.arange(m, dtype=np.uint64) + np.uint64(row_offset))
    with np.errstate(over="ignore"):
        hi = _mix64(rows * _K1 + _C1)
        for j in range(n):
            hj = _mix64(np.uint64(j) * _K2 + _C2)
            A[:, j] = unif(_mix64(np.uint64(seed) ^ hi ^ hj)) * s[j]
    return A


def vector(n, seed, scale=1.0, offset=0):
    i = np.arange(n, dtype=np.uint64) + np.uint64(offset)
    return scale * unif(hash64(seed, i, np.uint64(0xFFFFFFFF)))


def csc_pattern(m, n, nnz_per_col, seed):
    """Stratified rows (sorted, distinct). Returns 0-based (indptr, indices)."""
    k = np.arange(nnz_per_col, dtype=np.int64)
    lo = (m * k) // nnz_per_col
    hi = (m * (k + 1)) // nnz_per_col
    jj = np.repeat(np.arange(n, dtype=np.uint64), nnz_per_col)
    kk = np.tile(k.astype(np.uint64), n)
    h = hash64(seed + 2, jj, kk)
    rows = np.tile(lo, n) + (h % np.tile((hi - lo).astype(np.uint64), n)).astype(np.int64)
    indptr = np.arange(n + 1, dtype=np.int64) * nnz_per_col
    return indptr, rows


class DenseModel:
    """The bench's dense synthetic nonlinear least-squares problem (configs 2, 4, 5)."""

    def __init__(self, m, n, seed, c=0.1, noise=1e-3, A=None):
        self.m, self.n, self.c = m, n, c
        self.A = dense_matrix(m, n, seed) if A is None else A
        self.xstar = vector(n, seed + 11)
        t = self.A @ self.xstar
        self.b = (t + c * t * t) + noise * vector(m, seed + 12)
        self.x0 = self.xstar + 0.1 * vector(n, seed + 13)
        self._t = np.empty(m)

    def f(self, out, x):
        t = self.A @ x
        self._t[:] = t
        out[:] = (t + (self.c * t) * t) - self.b

    def g(self, J, x):
        t = self.A @ x
        w = 1.0 + (2.0 * self.c) * t
        np.multiply(self.A, w[:, None], out=J)

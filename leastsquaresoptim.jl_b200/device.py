"""Device-resident vector / matrix wrappers: the Python mirror of the `B200Vector` / `B200Matrix` /
`B200SparseMatrixCSC` duck types the Julia glue defines (julia/LeastSquaresOptimB200.jl).  They own
HBM buffers through the C ABI and implement exactly the operator interface the reference requires of a
Jacobian (README.md:37-43: `mul!(y,A,x,α,β)`, `mul!(x,A',y,α,β)`, `colsumabs2!(x,A)`, `size`, `eltype`)
and of a vector (src/utils/lsmr.jl:30-44 plus what the optimizers call).
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from ._lib import check, lib


class Context:
    """One CUDA device + stream + scratch (lso_ctx). Single-caller, like the reference (one Julia thread)."""

    _default = {}

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().lso_ctx_create(device, C.byref(self._h)))
        self.device = device
        self._fin = weakref.finalize(self, lib().lso_ctx_destroy, self._h)

    @classmethod
    def default(cls, device: int = 0) -> "Context":
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    @property
    def handle(self):
        return self._h

    def sync(self):
        check(lib().lso_ctx_sync(self._h), self._h)

    def stream(self) -> int:
        return lib().lso_ctx_stream(self._h) or 0

    def set_option(self, key: str, value: int):
        check(lib().lso_ctx_set_option(self._h, key.encode(), int(value)), self._h)

    def launch_count(self, reset: bool = False) -> int:
        out = C.c_int64()
        check(lib().lso_ctx_launch_count(self._h, C.byref(out), int(reset)), self._h)
        return out.value

    def profile_read(self):
        """(total_ms, launches) of the event-bracketed dominant-kernel launches since the last read."""
        ms, cnt = C.c_double(), C.c_int64()
        check(lib().lso_ctx_profile_read(self._h, C.byref(ms), C.byref(cnt)), self._h)
        return ms.value, cnt.value

    def stat(self, key: str, reset: bool = False) -> float:
        out = C.c_double()
        check(lib().lso_ctx_stat(self._h, key.encode(), C.byref(out), int(reset)), self._h)
        return out.value

    def profile_read_collective(self):
        """(total_ms, calls) of the event-bracketed collectives (all-reduce / all-gather) since the last read."""
        ms, cnt = C.c_double(), C.c_int64()
        check(lib().lso_ctx_profile_read_collective(self._h, C.byref(ms), C.byref(cnt)), self._h)
        return ms.value, cnt.value

    # -- raw memory --
    def alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(lib().lso_dev_alloc(self._h, int(nbytes), C.byref(p)), self._h)
        return p.value

    def free(self, ptr: int):
        lib().lso_dev_free(self._h, ptr)

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        check(lib().lso_comm_init_rank(self._h, nranks, rank, buf), self._h)
        self.nranks, self.rank = nranks, rank

    nranks = 1
    rank = 0

    def allreduce(self, v: "DeviceVector"):
        """In-place sum over ranks (NCCL all-reduce on the context stream); identity without a communicator."""
        if self.nranks > 1:
            check(lib().lso_comm_allreduce_sum(self._h, v.ptr, v.n), self._h)
        return v

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().lso_comm_unique_id(buf))
        return buf.raw


def _np_ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class DeviceVector:
    """fp64 vector in HBM."""

    def __init__(self, ctx: Context, n: int, data=None):
        self.ctx = ctx
        self.n = int(n)
        self.ptr = ctx.alloc(max(self.n, 1) * 8)
        self._fin = weakref.finalize(self, lib().lso_dev_free, ctx.handle, self.ptr)
        if data is not None:
            self.upload(data)
        else:
            self.fill(0.0)

    @classmethod
    def view(cls, ctx: Context, ptr: int, n: int) -> "DeviceVector":
        """A non-owning vector over device memory someone else manages (callback arguments)."""
        v = cls.__new__(cls)
        v.ctx, v.n, v.ptr, v._fin = ctx, int(n), int(ptr), None
        return v

    def __len__(self):
        return self.n

    @property
    def dtype(self):
        return np.float64

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        assert a.size == self.n, "length mismatch"
        check(lib().lso_upload(self.ctx.handle, self.ptr, _np_ptr(a), self.n * 8), self.ctx.handle)
        return self

    def download(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.n, dtype=np.float64)
        assert out.size == self.n and out.dtype == np.float64 and out.flags.c_contiguous
        check(lib().lso_download(self.ctx.handle, _np_ptr(out), self.ptr, self.n * 8), self.ctx.handle)
        return out

    def similar(self) -> "DeviceVector":
        return DeviceVector(self.ctx, self.n)

    # ---- vector duck type (lsmr.jl:30-44 + optimizer usage) ----
    def fill(self, v: float):
        check(lib().lso_vec_fill(self.ctx.handle, self.n, self.ptr, float(v)), self.ctx.handle)
        return self

    def copyto(self, src: "DeviceVector"):
        assert src.n == self.n
        check(lib().lso_vec_copy(self.ctx.handle, self.n, self.ptr, src.ptr), self.ctx.handle)
        return self

    def rmul(self, a: float):
        check(lib().lso_vec_scal(self.ctx.handle, self.n, self.ptr, float(a)), self.ctx.handle)
        return self

    def axpy(self, a: float, x: "DeviceVector"):
        """self += a * x   (axpy!(a, x, self))"""
        assert x.n == self.n
        check(lib().lso_vec_axpy(self.ctx.handle, self.n, float(a), x.ptr, self.ptr), self.ctx.handle)
        return self

    def _scalar(self, fn, *args) -> float:
        out = C.c_double()
        check(fn(self.ctx.handle, self.n, *args, C.byref(out)), self.ctx.handle)
        return out.value

    def sum(self) -> float:
        return self._scalar(lib().lso_vec_sum, self.ptr)

    def sumabs2(self) -> float:
        return self._scalar(lib().lso_vec_sumabs2, self.ptr)

    def norm(self) -> float:
        return self._scalar(lib().lso_vec_nrm2, self.ptr)

    def maxabs(self) -> float:
        return self._scalar(lib().lso_vec_maxabs, self.ptr)

    def dot(self, y: "DeviceVector") -> float:
        return self._scalar(lib().lso_vec_dot, self.ptr, y.ptr)

    def clamp(self, lo: float, hi: float):
        check(lib().lso_vec_clamp(self.ctx.handle, self.n, self.ptr, float(lo), float(hi)), self.ctx.handle)
        return self

    def sqrt_(self):
        check(lib().lso_vec_sqrt(self.ctx.handle, self.n, self.ptr), self.ctx.handle)
        return self

    def div_(self, x: "DeviceVector", y: "DeviceVector"):
        """self = x ./ y   (map!(/, self, x, y))"""
        check(lib().lso_vec_div(self.ctx.handle, self.n, self.ptr, x.ptr, y.ptr), self.ctx.handle)
        return self

    def mul_(self, x: "DeviceVector", y: "DeviceVector"):
        check(lib().lso_vec_mul(self.ctx.handle, self.n, self.ptr, x.ptr, y.ptr), self.ctx.handle)
        return self

    def check_finite(self):
        bad = C.c_int64(-1)
        st = lib().lso_vec_check_finite(self.ctx.handle, self.n, self.ptr, C.byref(bad))
        if st == -6:
            from ._lib import IsFiniteException

            raise IsFiniteException(st, f"non-finite entry at index {bad.value + 1}")
        check(st, self.ctx.handle)


def wdot(x: DeviceVector, y: DeviceVector, w: DeviceVector) -> float:
    """src/utils/utils.jl:165-173"""
    out = C.c_double()
    check(lib().lso_vec_wdot(x.ctx.handle, x.n, x.ptr, y.ptr, w.ptr, C.byref(out)), x.ctx.handle)
    return out.value


def wnorm(x: DeviceVector, w: DeviceVector) -> float:
    """src/utils/utils.jl:176"""
    return float(np.sqrt(wdot(x, x, w)))


class DenseMatrix:
    """Column-major fp64 matrix in HBM (the device image of a Julia `Matrix{Float64}`)."""

    def __init__(self, ctx: Context, m: int, n: int, data=None):
        self.ctx = ctx
        self.m, self.n = int(m), int(n)
        self.ld = self.m
        self.ptr = ctx.alloc(self.m * self.n * 8)
        self._fin = weakref.finalize(self, lib().lso_dev_free, ctx.handle, self.ptr)
        if data is not None:
            self.upload(data)

    @property
    def shape(self):
        return (self.m, self.n)

    def upload(self, a):
        a = np.asfortranarray(a, dtype=np.float64)
        assert a.shape == (self.m, self.n)
        check(lib().lso_upload(self.ctx.handle, self.ptr, _np_ptr(a), self.m * self.n * 8), self.ctx.handle)
        return self

    def download(self) -> np.ndarray:
        out = np.empty((self.m, self.n), dtype=np.float64, order="F")
        check(lib().lso_download(self.ctx.handle, _np_ptr(out), self.ptr, self.m * self.n * 8), self.ctx.handle)
        return out

    # operator interface (README.md:37-43)
    def colsumabs2(self, out: DeviceVector):
        check(lib().lso_dense_colsumabs2(self.ctx.handle, self.m, self.n, self.ptr, self.ld, out.ptr), self.ctx.handle)

    def mul(self, y: DeviceVector, x: DeviceVector, alpha=1.0, beta=0.0):
        """y = alpha * J * x + beta * y"""
        check(lib().lso_dense_gemv_n(self.ctx.handle, self.m, self.n, float(alpha), self.ptr, self.ld, x.ptr,
                                     float(beta), y.ptr), self.ctx.handle)

    def mul_t(self, x: DeviceVector, y: DeviceVector, alpha=1.0, beta=0.0):
        """x = alpha * J' * y + beta * x"""
        check(lib().lso_dense_gemv_t(self.ctx.handle, self.m, self.n, float(alpha), self.ptr, self.ld, y.ptr,
                                     float(beta), x.ptr), self.ctx.handle)

    def colsumabs2_and_grad(self, dtd: DeviceVector, g: DeviceVector, f: DeviceVector):
        check(lib().lso_dense_colsumabs2_gemv_t(self.ctx.handle, self.m, self.n, self.ptr, self.ld, f.ptr, dtd.ptr,
                                                g.ptr), self.ctx.handle)

    def predicted_ssr(self, delta: DeviceVector, f: DeviceVector, fpredict: DeviceVector | None) -> float:
        out = C.c_double()
        check(lib().lso_dense_predicted_ssr(self.ctx.handle, self.m, self.n, self.ptr, self.ld, delta.ptr, f.ptr,
                                            fpredict.ptr if fpredict is not None else None, C.byref(out)),
              self.ctx.handle)
        return out.value


class CSCMatrix:
    """Device image of a SparseMatrixCSC{Float64,Int64}; built from a scipy.sparse.csc_matrix pattern."""

    def __init__(self, ctx: Context, m: int, n: int, colptr0: np.ndarray, rowidx0: np.ndarray, values=None):
        self.ctx = ctx
        self.m, self.n = int(m), int(n)
        colptr = np.ascontiguousarray(colptr0, dtype=np.int64) + 1     # Julia's 1-based Int64 arrays
        rowval = np.ascontiguousarray(rowidx0, dtype=np.int64) + 1
        self.nnz = int(rowval.size)
        self._h = C.c_void_p()
        check(lib().lso_csc_create(ctx.handle, self.m, self.n, self.nnz, _np_ptr(colptr), _np_ptr(rowval),
                                   C.byref(self._h)), ctx.handle)
        self._fin = weakref.finalize(self, lib().lso_csc_destroy, self._h)
        if values is not None:
            self.set_values(values)

    @classmethod
    def from_scipy(cls, ctx: Context, A) -> "CSCMatrix":
        A = A.tocsc()
        A.sort_indices()
        return cls(ctx, A.shape[0], A.shape[1], A.indptr, A.indices, A.data)

    @property
    def shape(self):
        return (self.m, self.n)

    @property
    def handle(self):
        return self._h

    def set_values(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.size == self.nnz
        check(lib().lso_csc_set_values_host(self._h, _np_ptr(v)), self.ctx.handle)

    def update_pattern(self, colptr0: np.ndarray, rowidx0: np.ndarray, values=None):
        """g! changed the sparsity pattern (setindex! into a sparse J, test/nonlinearsolvers.jl:526-530)."""
        colptr = np.ascontiguousarray(colptr0, dtype=np.int64) + 1
        rowval = np.ascontiguousarray(rowidx0, dtype=np.int64) + 1
        self.nnz = int(rowval.size)
        check(lib().lso_csc_update_pattern(self._h, self.nnz, _np_ptr(colptr), _np_ptr(rowval)), self.ctx.handle)
        if values is not None:
            self.set_values(values)

    def values_ptr(self) -> int:
        return lib().lso_csc_values(self._h)

    def values_csr_ptr(self) -> int:
        return lib().lso_csc_values_csr(self._h)

    def values_changed(self, both: bool = False):
        """After a device g! wrote `values_ptr()` (and, with both=True, `values_csr_ptr()` as well)."""
        fn = lib().lso_csc_values_changed_both if both else lib().lso_csc_values_changed
        check(fn(self._h), self.ctx.handle)

    def gather_csr(self, src: DeviceVector, dst: DeviceVector):
        """dst (CSR order) = src (CSC order) permuted; once per pattern for constant Jacobian factors."""
        check(lib().lso_csc_gather_csr(self._h, src.ptr, dst.ptr), self.ctx.handle)

    def colsumabs2(self, out: DeviceVector):
        check(lib().lso_csc_colsumabs2(self._h, out.ptr), self.ctx.handle)

    def mul(self, y: DeviceVector, x: DeviceVector, alpha=1.0, beta=0.0):
        check(lib().lso_csc_mul_n(self._h, float(alpha), x.ptr, float(beta), y.ptr), self.ctx.handle)

    def mul_t(self, x: DeviceVector, y: DeviceVector, alpha=1.0, beta=0.0):
        check(lib().lso_csc_mul_t(self._h, float(alpha), y.ptr, float(beta), x.ptr), self.ctx.handle)

    def colsumabs2_and_grad(self, dtd: DeviceVector, g: DeviceVector, f: DeviceVector):
        """colsumabs2!(dtd, J) (LM:82) and mul!(g, J', f) (LM:102) in one pass over the CSC image."""
        check(lib().lso_csc_colsumabs2_gemv_t(self._h, f.ptr, dtd.ptr, g.ptr), self.ctx.handle)

    def predicted_ssr(self, delta: DeviceVector, f: DeviceVector, fpredict: DeviceVector | None) -> float:
        """mul!(fpredict, J, δx, 1, 0); axpy!(-1, fcur, fpredict); sum(abs2, fpredict)  (LM:114-117) in one launch."""
        out = C.c_double()
        check(lib().lso_csc_predicted_ssr(self._h, delta.ptr, f.ptr, fpredict.ptr if fpredict is not None else None,
                                          C.byref(out)), self.ctx.handle)
        return out.value

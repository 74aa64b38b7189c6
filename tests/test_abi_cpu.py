"""CPU-side checks of the drop-in boundary: the header parses, the C-ABI library loads, exports every symbol
include/lsob200.h declares, and refuses to run (loudly) without a CUDA device — no compute calls here."""
import ctypes as C
import os
import subprocess

import pytest

import lsob200
from lsob200 import _lib


def test_header_declares_the_boundary():
    names = set(_lib.PROTOTYPES)
    for required in ("lso_ctx_create", "lso_qr_solve", "lso_qr_solve_host", "lso_chol_solve", "lso_lsmr_solve",
                     "lso_csc_create", "lso_csc_mul_n", "lso_csc_mul_t", "lso_csc_colsumabs2", "lso_dense_colsumabs2",
                     "lso_dense_gemv_n", "lso_dense_gemv_t", "lso_dense_predicted_ssr", "lso_vec_axpy", "lso_vec_wdot",
                     "lso_vec_box_project", "lso_vec_maxabs_projected", "lso_lm_damping", "lso_dogleg_blend",
                     "lso_comm_init_rank", "lso_comm_allreduce_sum", "lso_qr_solve_sharded"):
        assert required in names
    assert len(names) >= 70


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    for name in _lib.PROTOTYPES:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIBPATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    assert set(_lib.PROTOTYPES) <= exported
    assert lib.lso_version() >= 100


def test_library_is_sm100a_native():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIBPATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_device():
    lib = _lib.lib()
    cnt = C.c_int()
    lib.lso_device_count(C.byref(cnt))
    if cnt.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(lsob200.LsoError) as e:
        lsob200.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.abspath(_lib.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "reference_port" not in src and "import oracle" not in src and "from oracle" not in src, f


def test_host_chunk_rows_partition_every_row_once():
    """Host logic of the chunked upload (HostStep / lso_qr_factor_keep_host_chunks): the chunk sizes cover the m rows exactly,
    in order, with no empty chunk; bad shares are rejected before anything is enqueued."""
    import pytest
    from lsob200.api import AUTO_CHUNK_SHARES, host_chunk_rows
    assert host_chunk_rows(100000, 1) == [100000]
    assert host_chunk_rows(100000, list(AUTO_CHUNK_SHARES)) == [30000, 30000, 25000, 15000]
    for m, c in [(60001, 2), (60001, 3), (60001, 7), (7, 16), (100003, [0.3, 0.3, 0.25, 0.15]), (12345, [5, 1, 1])]:
        rows = host_chunk_rows(m, c)
        assert sum(rows) == m and min(rows) >= 1
        if isinstance(c, int):
            assert len(rows) == min(c, m) and max(rows) - min(rows) <= 1
    with pytest.raises(ValueError):
        host_chunk_rows(100, [0.5, 0.0, 0.5])
    with pytest.raises(ValueError):
        host_chunk_rows(2, [1, 1, 1])


def test_trace_lines_have_the_reference_format():
    """utils.jl:124-127: `@printf "%6d   %14e   %14e\n"`; show_every filters on iteration % show_every (utils.jl:104-108)."""
    from lsob200.api import OptimizationState, format_trace
    st = [OptimizationState(0, 24.2, float("inf")), OptimizationState(1, 4.731884e+00, 3.0), OptimizationState(2, 1e-30, 0.0)]
    assert format_trace(st) == ("     0     2.420000e+01              Inf\n"
                                "     1     4.731884e+00     3.000000e+00\n"
                                "     2     1.000000e-30     0.000000e+00\n")
    assert format_trace(st, 2).count("\n") == 2 and format_trace([], 1) == ""

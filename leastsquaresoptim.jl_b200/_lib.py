"""ctypes binding of liblsob200.so, generated from include/lsob200.h at import time.

The product path has NO CPU fallback: if the shared library is missing this module raises, and if no
CUDA device is visible `lso_ctx_create` fails with a clear message.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "lsob200.h")
LIBPATH = os.path.join(_HERE, "liblsob200.so")

_BASE = {
    "void": None,
    "int": C.c_int,
    "int64_t": C.c_int64,
    "uint64_t": C.c_uint64,
    "size_t": C.c_size_t,
    "double": C.c_double,
    "char": C.c_char,
}
_OPAQUE = ("lso_ctx", "lso_dense_ws", "lso_csc", "lso_lsmr_ws")


def _ctype(decl: str):
    """Map a C parameter / return declaration (without the name) to a ctypes type."""
    decl = decl.replace("const", " ").strip()
    stars = decl.count("*")
    base = decl.replace("*", " ").split()[0]
    if base in _OPAQUE:
        return C.c_void_p if stars == 1 else C.POINTER(C.c_void_p)
    if base == "void":
        if stars == 0:
            return None
        return C.c_void_p if stars == 1 else C.POINTER(C.c_void_p)
    if base == "char" and stars == 1:
        return C.c_char_p
    if base in ("lso_precond_fn", "lso_residual_fn"):
        return C.c_void_p          # function pointer: pass a ctypes CFUNCTYPE instance cast to c_void_p, or None
    t = _BASE[base]
    if stars == 0:
        return t
    # all numeric pointers are passed as raw addresses (device pointers, numpy data pointers, byref)
    return C.c_void_p


def parse_header(path: str = HEADER):
    """Return {name: (restype, [argtypes], [argnames])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = "\n".join(l for l in src.splitlines() if not l.strip().startswith("#"))
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(lso_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef") or not ret:
            continue
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                argtypes.append(_ctype(mm.group(1)))
                argnames.append(mm.group(2))
        protos[name] = (_ctype(ret), argtypes, argnames)
    return protos


PROTOTYPES = parse_header()


class LsoError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"lsob200 error {code}: {message}")
        self.code = code
        self.message = message


class DimensionMismatch(LsoError):
    pass


class PosDefException(LsoError):
    pass


class RankDeficientException(LsoError):
    pass


class IsFiniteException(LsoError):
    pass


def load_library(path: str = LIBPATH):
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C leastsquaresoptim.jl_b200/csrc`). There is no CPU fallback."
        )
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (restype, argtypes, _) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


def check(status: int, ctx=None):
    """Translate a status code into the exception the Julia glue would raise for it."""
    if status == 0:
        return
    msg = lib().lso_last_error(ctx)
    msg = msg.decode() if msg else ""
    if status > 0:
        if "RankDeficient" in msg:
            raise RankDeficientException(status, msg)
        raise PosDefException(status, msg)
    if status == -1:
        raise DimensionMismatch(status, msg)
    if status == -6:
        raise IsFiniteException(status, msg)
    raise LsoError(status, msg)

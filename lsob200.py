"""Import shim: the package directory is named after the reference repo (`leastsquaresoptim.jl_b200`), which
is not a valid Python identifier, so it is loaded here under the importable alias `lsob200`."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "leastsquaresoptim.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "lsob200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["lsob200"] = _mod
_spec.loader.exec_module(_mod)

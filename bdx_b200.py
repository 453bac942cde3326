"""Import shim: loads the package directory ``biodemux.jl_b200/`` (whose name is
not a valid Python identifier) under the module name ``bdx_b200``."""
import importlib.util as _u
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_pkg = _os.path.join(_here, "biodemux.jl_b200")
_spec = _u.spec_from_file_location("bdx_b200", _os.path.join(_pkg, "__init__.py"),
                                   submodule_search_locations=[_pkg])
_mod = _u.module_from_spec(_spec)
_sys.modules["bdx_b200"] = _mod
_spec.loader.exec_module(_mod)

"""Import shim: exposes the package in `faer-rs_b200/` (not a valid Python identifier) as `faer_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "faer-rs_b200")
_NAME = "faer_rs_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

_pkg = sys.modules[_NAME]
capi = _pkg.capi
linalg = _pkg.linalg
dist = _pkg.dist
solvers = _pkg.solvers
load = _pkg.load
PKG_DIR = _PKG_DIR

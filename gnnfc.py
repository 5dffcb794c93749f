"""Import shim: ``import gnnfc`` loads the package in ``gnn-formation-control_b200/``
(a directory name that is not a valid Python identifier)."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gnn-formation-control_b200")
_spec = importlib.util.spec_from_file_location(
    "gnnfc", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gnnfc"] = _mod
_spec.loader.exec_module(_mod)

"""Import shim: the product package lives in the directory ``h-tsp_b200/`` (the name the
project layout prescribes), which is not a valid Python identifier.  This module loads that
directory as the package ``htsp_b200`` so that ``import htsp_b200`` and
``from htsp_b200.path_solver import B200PathSolver`` work from the repo root."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "h-tsp_b200")
_spec = _ilu.spec_from_file_location(
    "htsp_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = _ilu.module_from_spec(_spec)
_sys.modules["htsp_b200"] = _mod
_spec.loader.exec_module(_mod)

"""Import shim: makes the package directory ``pelvistim-fem_b200/`` (hyphenated, as
the project is named) importable as ``pelvistim_fem_b200``.

``import pelvistim_fem_b200`` from the repo root executes
``pelvistim-fem_b200/__init__.py`` and resolves sub-modules
(``pelvistim_fem_b200.engine`` ...) from that directory.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pelvistim-fem_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)

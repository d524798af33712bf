"""Import shim: the package directory is named ``gpt-st_b200`` (not a valid Python identifier), so
``import gptst_b200`` loads it from there under the importable name ``gptst_b200``."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "gpt-st_b200")
_spec = _ilu.spec_from_file_location("gptst_b200", _os.path.join(_pkg_dir, "__init__.py"),
                                     submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["gptst_b200"] = _mod
_spec.loader.exec_module(_mod)

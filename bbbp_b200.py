"""Import shim: ``import bbbp_b200`` -> the package in ./bbbp-multi-modal-deep-ensemble-framework_b200/
(the directory name the build contract prescribes is not a valid Python identifier)."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "bbbp-multi-modal-deep-ensemble-framework_b200")
_spec = _u.spec_from_file_location("bbbp_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["bbbp_b200"] = _mod
_spec.loader.exec_module(_mod)

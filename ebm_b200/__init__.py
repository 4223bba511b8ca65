"""Import shim: ``import ebm_b200`` loads the package that lives in ``energybalancemodel.jl_b200/``
(the directory name carries the reference's name and is not a valid Python identifier)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_root = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "energybalancemodel.jl_b200")
_spec = _ilu.spec_from_file_location("ebm_b200", _os.path.join(_root, "__init__.py"), submodule_search_locations=[_root])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["ebm_b200"] = _mod
_spec.loader.exec_module(_mod)

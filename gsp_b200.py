"""Import shim: the package directory is named `geostatsprocesses.jl_b200` (not a valid module name),
so this file loads it under the name `gsp_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "geostatsprocesses.jl_b200")
_spec = importlib.util.spec_from_file_location("gsp_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gsp_b200"] = _mod
_spec.loader.exec_module(_mod)

"""Import shim: the package directory is named ``zlib.es_b200`` (after the reference,
zprodev/zlib.es), which is not a valid Python identifier.  ``import zles`` loads it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "zlib.es_b200")
_spec = importlib.util.spec_from_file_location("zles", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["zles"] = _mod
_spec.loader.exec_module(_mod)

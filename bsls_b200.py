"""Import alias: ``import bsls_b200`` loads the package that lives in the directory
``block-simplex-least-squares_b200/`` (a name Python cannot import directly)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "block-simplex-least-squares_b200")
_spec = importlib.util.spec_from_file_location("bsls_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["bsls_b200"] = _mod
_spec.loader.exec_module(_mod)

"""Import alias: ``import sifnn_b200`` loads the package whose directory name
(``land-surface-temperature-super-resolution-with-a-scale-invariance-free-neural-approach_b200``)
is fixed by the build contract but is not a valid Python identifier."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "land-surface-temperature-super-resolution-with-a-scale-invariance-free-neural-approach_b200")
_spec = importlib.util.spec_from_file_location("sifnn_b200", os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sifnn_b200"] = _mod
_spec.loader.exec_module(_mod)

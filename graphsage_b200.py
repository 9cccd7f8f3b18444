"""Import alias: `import graphsage_b200` loads the package that lives in the directory
`graphsage-pytorch_b200/` (a hyphen is not importable, the layout name is fixed)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "graphsage-pytorch_b200")
_spec = _ilu.spec_from_file_location("graphsage_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["graphsage_b200"] = _mod
_spec.loader.exec_module(_mod)

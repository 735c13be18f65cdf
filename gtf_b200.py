"""Import alias: the package directory is `gnn-track-finding_b200/` (a hyphenated name Python cannot
import directly); `import gtf_b200` loads it under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gnn-track-finding_b200")
_spec = importlib.util.spec_from_file_location("gtf_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gtf_b200"] = _mod
_spec.loader.exec_module(_mod)

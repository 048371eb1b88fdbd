"""Import alias: `import omfs_b200` gives the package that lives in `omfs-4d-video-gen_b200/`
(a directory name Python's import statement cannot spell)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("omfs-4d-video-gen_b200")
sys.modules[__name__] = _pkg

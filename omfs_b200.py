"""Import alias: `import omfs_b200` gives the package that lives in `omfs-4d-video-gen_b200/`
(a directory name Python's import statement cannot spell).

`omfs_b200.<module>` resolves to the SAME module object as `omfs-4d-video-gen_b200.<module>`: without that, a
dotted import (`from omfs_b200.runtime import X`) would execute the module a second time under the alias name and
leave the process with two copies of its state (two library handles, two OmfsError classes)."""
import importlib
import importlib.abc
import importlib.util
import os
import sys

_REAL = "omfs-4d-video-gen_b200"
_ALIAS = __name__

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(_ALIAS + "."):
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_ALIAS):])   # the one real module object

    def exec_module(self, module):
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
_pkg = importlib.import_module(_REAL)
sys.modules[_ALIAS] = _pkg

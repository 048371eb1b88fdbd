"""The C-ABI library loads and exports every symbol include/omfs_b200.h declares (no compute calls:
this runs without a GPU)."""
import ctypes
import os

import pytest


def test_library_exports_every_declared_symbol():
    import omfs_b200  # noqa: F401
    from omfs_b200 import runtime
    if not os.path.exists(runtime.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    names = runtime.declared_symbols()
    assert len(names) >= 30 and "omfs_session_render_host" in names and "omfs_composite" in names
    L = runtime.load_library()
    for n in names:
        assert hasattr(L, n), n
    assert L.omfs_abi_version() == 2


def test_argument_errors_do_not_need_a_gpu():
    import omfs_b200  # noqa: F401
    from omfs_b200 import runtime
    L = runtime.load_library()
    # bad sizes are rejected before any CUDA call, with a message
    rc = L.omfs_face_frames(1, 0, 0, None, None, None, None)
    assert rc == -1
    assert b"omfs_face_frames" in L.omfs_last_error()
    assert L.omfs_binning_workspace_bytes(0, 0, 0, 0, 0) == 0
    assert L.omfs_binning_workspace_bytes(4, 1000, 64, 64, 10000) > 0
    assert L.omfs_binning_sort_bits(1, 512, 512) == 42       # 42 bits at 512^2
    assert L.omfs_binning_sort_bits(1, 1024, 1024) == 44
    assert L.omfs_binning_sort_bits(64, 512, 512) == 48


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import omfs_b200  # noqa: F401
    from omfs_b200 import avatar, runtime, synthetic
    model, params, av, cam = synthetic.make_scene(n_gauss=100, n_frames=1, width=32, height=32, n_verts=162)
    with pytest.raises(runtime.OmfsError):
        runtime.Session(model, avatar.bake(av), 32, 32)


def _header_struct_fields(name):
    """Field names of `typedef struct <name> {...}` in include/omfs_b200.h, in declaration order."""
    import re
    from omfs_b200 import runtime
    text = re.sub(r"/\*.*?\*/", "", open(runtime.HEADER_PATH).read(), flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if not decl.startswith("const") else decl.split(None, 2)[2]
        for n in names.split(","):
            fields.append(re.sub(r"[\*\s]|\[.*\]", "", n))
    return fields


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The three structs that cross the C-ABI: field order in runtime.py and in INTEGRATION.md's stub equals the
    header's, and gcc's sizeof / offsetof equal ctypes' (the header is compiled as plain C, as a cgo/ctypes user would)."""
    import re
    import subprocess
    import omfs_b200  # noqa: F401
    from omfs_b200 import runtime
    pairs = [("omfs_model_desc", runtime.ModelDesc), ("omfs_session_config", runtime.SessionConfig),
             ("omfs_frames_desc", runtime.FramesDesc)]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    stub = open(os.path.join(root, "INTEGRATION.md")).read()
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "omfs_b200.h"', 'int main(void) {']
    for cname, cls in pairs:
        fields = _header_struct_fields(cname)
        assert fields == [f[0] for f in cls._fields_], cname
        src.append(f'printf("{cname} %zu", sizeof({cname}));')
        for f in fields:
            src.append(f'printf(" %zu", offsetof({cname}, {f}));')
        src.append('printf("\\n");')
        # the stub in INTEGRATION.md names the same fields, in the same order
        cls_src = re.search(r"class _\w+\(ctypes\.Structure\):\s+# %s\n(.*?)(?=\nclass |\ndef )" % cname, stub, flags=re.S).group(1)
        quoted = re.findall(r'"([a-z_0-9]+)"', cls_src)
        assert quoted == fields, (cname, quoted)
    src += ['return 0; }']
    c = tmp_path / "abi.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(c), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).strip().splitlines()
    for line, (cname, cls) in zip(out, pairs):
        nums = [int(x) for x in line.split()[1:]]
        assert nums[0] == ctypes.sizeof(cls), cname
        assert nums[1:] == [getattr(cls, f[0]).offset for f in cls._fields_], cname


def test_the_product_never_touches_the_oracle():
    """oracle/ is the checker: nothing in the product package or its C/CUDA sources imports, loads or links it, and
    bench.py reaches it only inside its two CPU-baseline functions."""
    import ast
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "omfs-4d-video-gen_b200")

    def oracle_imports(path):
        hits = []
        for node in ast.walk(ast.parse(open(path).read())):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            elif isinstance(node, ast.Constant) and isinstance(node.value, str) and "liboracle" in node.value:
                names = ["oracle"]
            hits += [(node.lineno, n) for n in names if n.split(".")[0] == "oracle"]
        return hits

    for path in glob.glob(os.path.join(pkg, "*.py")) + [os.path.join(root, "omfs_b200.py")]:
        assert oracle_imports(path) == [], path
    for path in glob.glob(os.path.join(pkg, "csrc", "*")) + glob.glob(os.path.join(root, "include", "*")):
        text = open(path).read()
        assert "omfs_oracle" not in text and "liboracle" not in text and "orc_" not in text, path
    # bench.py: the imports sit inside cpu_oracle_fps / run_reference (the cpu_baseline leg and --impl reference) only
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    allowed = set()
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name in ("cpu_oracle_fps", "run_reference")]:
        allowed |= {n.lineno for n in ast.walk(fn) if isinstance(n, (ast.Import, ast.ImportFrom))}
    lines = {ln for ln, _ in oracle_imports(os.path.join(root, "bench.py"))}
    assert lines and lines <= allowed, (lines, allowed)


def test_alias_imports_share_one_module_object():
    """`omfs_b200.<module>` (dotted) and `from omfs_b200 import <module>` must be the same object: a second copy of
    runtime would carry its own library handle and its own OmfsError class."""
    import importlib
    import sys
    import omfs_b200  # noqa: F401
    from omfs_b200.runtime import OmfsError
    from omfs_b200 import runtime
    real = importlib.import_module("omfs-4d-video-gen_b200.runtime")
    assert runtime is real and OmfsError is real.OmfsError and sys.modules["omfs_b200.runtime"] is real

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
    assert L.omfs_abi_version() == 1


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

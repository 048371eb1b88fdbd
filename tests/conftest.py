import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def small_scene():
    """A small FLAME-like scene the oracle renders in well under a second."""
    import omfs_b200  # noqa: F401
    from omfs_b200 import avatar, synthetic
    model, params, av, cam = synthetic.make_scene(n_gauss=6000, n_frames=3, width=160, height=112, n_verts=1202)
    return model, params, av, avatar.bake(av), cam

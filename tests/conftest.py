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


def record_parity(name: str, **values):
    """Measured parity figures of the GPU tests -> gpurun_out/parity_r2.json (merged back by gpurun; a copy of the
    builder's run is tracked as profiles/parity_r2.json).  One entry per test, overwritten on re-runs."""
    import json
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "parity_r2.json")
    data = {}
    if os.path.exists(path):
        try:
            data = json.load(open(path))
        except ValueError:
            data = {}
    data[name] = {k: (float(v) if isinstance(v, (int, float)) or hasattr(v, "item") else v) for k, v in values.items()}
    with open(path, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)

"""The record / pair-word layout helpers of runtime.py (pure numpy): what the GPU tests use to put the device layout
back into the published field order before comparing with the oracle."""
import numpy as np


def test_published_records_and_pair_words():
    import omfs_b200  # noqa: F401
    from omfs_b200 import runtime as rt
    rng = np.random.default_rng(0)
    ext = rng.uniform(0.0, 50.0, size=(2, 7, 2)).astype(np.float16)
    P0 = rng.normal(size=(2, 7, 4)).astype(np.float32)
    P2 = rng.normal(size=(2, 7, 4)).astype(np.float32)
    P0[..., 2] = np.ascontiguousarray(ext).view(np.float32).reshape(2, 7)      # two halves in the slot of one float
    p0, p2, e = rt.published_records(P0, P2)
    assert np.array_equal(p0[..., [0, 1, 3]], P0[..., [0, 1, 3]]) and np.array_equal(p0[..., 2], P2[..., 3])
    assert np.array_equal(p2[..., :3], P2[..., :3]) and not p2[..., 3].any()
    assert np.array_equal(e, ext.astype(np.float32))
    assert P0[0, 0, 2] != p0[0, 0, 2]                                           # inputs untouched
    idx = rng.integers(0, 1 << 28, size=100, dtype=np.uint32)
    hint = rng.integers(0, 16, size=100, dtype=np.uint32)
    vals = idx | (hint << np.uint32(rt.VAL_INDEX_BITS))
    assert np.array_equal(rt.pair_indices(vals), idx) and np.array_equal(rt.pair_hints(vals), hint)


def test_header_and_runtime_agree_on_the_index_width():
    import os
    import re
    import omfs_b200  # noqa: F401
    from omfs_b200 import runtime as rt
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "omfs_b200.h")).read()
    assert int(re.search(r"#define OMFS_VAL_INDEX_BITS (\d+)", hdr).group(1)) == rt.VAL_INDEX_BITS

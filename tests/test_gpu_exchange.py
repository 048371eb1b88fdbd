"""Frame exchange on one GPU: a second PROCESS opens the root's IPC handle and pushes its
frame block into its slot with the copy engine (omfs_push_frames); the root checks every byte."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = textwrap.dedent("""
    import sys
    import numpy as np
    sys.path.insert(0, {root!r})
    import omfs_b200  # noqa
    from omfs_b200 import runtime as rt, sharding
    handle = bytes.fromhex(sys.argv[1])
    slot = int(sys.argv[2])
    rank = int(sys.argv[3])
    world = int(sys.argv[4])
    if len(sys.argv) > 5:
        rt.check(rt.load_library().omfs_set_device(int(sys.argv[5])))
    g = sharding.PeerFrameGather(slot, rank, world, lambda obj: [handle] + [None] * (world - 1))
    src = rt.DeviceArray.from_numpy(((np.arange(slot, dtype=np.uint32) * (rank + 7)) % 251).astype(np.uint8))
    g.push(src.ptr, slot)
    rt.check(rt.load_library().omfs_device_sync())
    g.close()
    print("pushed", rank)
""")


def test_peer_push_through_ipc(tmp_path):
    import omfs_b200  # noqa: F401
    from omfs_b200 import runtime as rt, sharding
    world, slot = 3, 1 << 20
    captured = {}

    def exchange(obj):
        captured["handle"] = obj
        return [obj] + [None] * (world - 1)

    root = sharding.PeerFrameGather(slot, 0, world, exchange)
    root.buffer.zero()
    own = rt.DeviceArray.from_numpy(((np.arange(slot, dtype=np.uint32) * 7) % 251).astype(np.uint8))
    root.push(own.ptr, slot)
    script = tmp_path / "child.py"
    script.write_text(CHILD.format(root=ROOT))
    for r in (1, 2):
        out = subprocess.run([sys.executable, str(script), captured["handle"].hex(), str(slot), str(r), str(world)],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        assert f"pushed {r}" in out.stdout
    got = root.numpy()
    for r in range(world):
        want = ((np.arange(slot, dtype=np.uint32) * (r + 7)) % 251).astype(np.uint8)
        assert np.array_equal(got[r], want), r
    root.close()


def test_peer_push_between_two_gpus(tmp_path):
    """The bench's exchange as it really runs: the pushing process sits on ANOTHER GPU, so the bytes cross NVLink
    (peer-to-peer copy by the sender's copy engine) before the root checks every one of them."""
    import omfs_b200  # noqa: F401
    from omfs_b200 import runtime as rt, sharding
    n_dev = rt.load_library().omfs_device_count()
    if n_dev < 2:
        pytest.skip("needs two GPUs")
    rt.check(rt.load_library().omfs_set_device(0))
    world, slot = n_dev, (1 << 22) + 12     # an odd-sized slot: offsets are not 16-byte aligned
    captured = {}

    def exchange(obj):
        captured["handle"] = obj
        return [obj] + [None] * (world - 1)

    root = sharding.PeerFrameGather(slot, 0, world, exchange)
    root.buffer.zero()
    script = tmp_path / "child.py"
    script.write_text(CHILD.format(root=ROOT))
    for r in range(1, world):
        out = subprocess.run([sys.executable, str(script), captured["handle"].hex(), str(slot), str(r), str(world),
                              str(r)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        assert f"pushed {r}" in out.stdout
    got = root.numpy()
    assert not got[0].any()                                   # nobody wrote the root's own slot
    for r in range(1, world):
        want = ((np.arange(slot, dtype=np.uint32) * (r + 7)) % 251).astype(np.uint8)
        assert np.array_equal(got[r], want), r
    root.close()


def test_push_larger_than_slot_is_rejected():
    import omfs_b200  # noqa: F401
    from omfs_b200 import sharding
    root = sharding.PeerFrameGather(1024, 0, 1, lambda obj: [obj])
    with pytest.raises(ValueError):
        root.push(root.base, 2048)
    root.close()

"""torch.ops.omfs.* — what can be checked without a GPU: the operators are registered with the documented schema,
CPU tensors are rejected (no CPU path), unknown handles are reported."""
import pytest
import torch

import omfs_b200  # noqa: F401
from omfs_b200 import runtime, torch_ops  # noqa: F401


def test_operators_are_registered():
    for name in ("render", "render_image"):
        op = getattr(torch.ops.omfs, name)
        schema = str(op.default._schema)
        assert schema.startswith(f"omfs::{name}(int session, Tensor expr, Tensor rotation")
        assert "Tensor? dynamic_offset" in schema and schema.endswith("-> Tensor")


def test_cpu_tensors_are_rejected():
    z = torch.zeros(2, 3)
    with pytest.raises(runtime.OmfsError, match="no CPU implementation"):
        torch.ops.omfs.render(1, torch.zeros(2, 100), z, z, z, torch.zeros(2, 6), z, torch.zeros(1, 40), None)


def test_unknown_session_handle():
    with pytest.raises(runtime.OmfsError, match="no open session"):
        torch_ops.check(12345)

"""Export the teeth-augmented FLAME model (5143 vertices) that the reference pipeline's records and
GaussianAvatars checkpoints are written for, as the flame_model.npz this repository loads.

Run it INSIDE the upstream GaussianAvatars checkout the reference clones (train_ghost.py:33-38), in
the environment that already holds the licence-gated FLAME assets:

    cd gaussian_avatars_repo && python /path/to/tools/export_flame_with_teeth.py flame_model.npz

It only reads attributes of upstream's own `FlameHead` (which grafts the teeth on at construction
from upstream's mask asset) and re-lays them out; nothing of upstream is copied into this repo.
The output goes next to the trained model (`<model_path>/flame_model.npz`) or wherever
$OMFS_FLAME_MODEL points.
"""
from __future__ import annotations

import sys

import numpy as np


def main(out_path: str, n_shape: int = 300, n_expr: int = 100) -> None:
    from flame_model.flame import FlameHead  # upstream module, only available inside its checkout

    head = FlameHead(n_shape, n_expr, add_teeth=True)
    as_np = lambda t: t.detach().cpu().numpy()
    v = as_np(head.v_template).astype(np.float32)                    # (V,3)
    V = v.shape[0]
    sd = as_np(head.shapedirs).astype(np.float32)                    # (V,3,n_shape+n_expr)
    pd = as_np(head.posedirs).astype(np.float32)                     # (36, V*3)
    dirs = np.zeros((300 + n_expr, 3 * V), np.float32)
    dirs[:n_shape] = sd[:, :, :n_shape].reshape(3 * V, n_shape).T
    dirs[300:] = sd[:, :, n_shape:].reshape(3 * V, n_expr).T
    np.savez(out_path, v_template=v, faces=as_np(head.faces).astype(np.int32), shapedirs=dirs,
             posedirs=pd.reshape(36, 3 * V), j_regressor=as_np(head.J_regressor).astype(np.float32),
             lbs_weights=as_np(head.lbs_weights).astype(np.float32),
             parents=np.array([-1, 0, 1, 1, 1], np.int32))
    print(f"wrote {out_path}: {V} vertices, {int(as_np(head.faces).shape[0])} faces")


if __name__ == "__main__":
    if len(sys.argv) != 2:
        raise SystemExit(__doc__)
    main(sys.argv[1])

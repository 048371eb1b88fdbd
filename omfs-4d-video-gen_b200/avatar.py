"""Avatar bake: raw GaussianAvatars attributes -> the float4 SoA streams the kernels read.

Everything baked here is frame-invariant, so it is computed ONCE when the avatar is
loaded and never again per frame:

  xyzb     [N,4]    face-local xyz, parent-face index (int32 bit pattern in .w)
  scale_lo [N,4]    exp(log-scale) xyz  (the upstream scaling activation [UPSTREAM]),
                    lo = log2(sigmoid(opacity logit)) in .w
  rot      [N,4]    unit local quaternion wxyz (the upstream rotation activation)
  sh       [12,N,4] the 48 SH floats per Gaussian as 12 float4 planes: flat index
                    k*3+channel -> plane (flat>>2), lane (flat&3); a warp reading one
                    plane touches consecutive 16-byte words

The transcendental activations are evaluated in float64 and rounded once, which makes the
baked streams a fixed input of the per-frame path on both the CUDA side and the oracle.
Binding indices are carried as raw int32 bits and must round-trip bit-exact.
"""
from __future__ import annotations

import numpy as np


def bake_avatar(xyz, scaling, rotation, opacity, sh, binding) -> dict:
    xyz = np.asarray(xyz, dtype=np.float32)
    N = xyz.shape[0]
    binding = np.ascontiguousarray(binding, dtype=np.int32).reshape(N)
    xyzb = np.empty((N, 4), dtype=np.float32)
    xyzb[:, :3] = xyz
    xyzb[:, 3] = binding.view(np.float32)

    scale_lo = np.empty((N, 4), dtype=np.float32)
    scale_lo[:, :3] = np.exp(np.asarray(scaling, dtype=np.float64)).astype(np.float32)
    o = np.asarray(opacity, dtype=np.float64).reshape(N)
    # log2(sigmoid(o)) = -log2(1 + exp(-o)), stable for both signs
    lo = -(np.logaddexp(0.0, -o)) / np.log(2.0)
    scale_lo[:, 3] = lo.astype(np.float32)

    q = np.asarray(rotation, dtype=np.float64).reshape(N, 4)
    nrm = np.linalg.norm(q, axis=1, keepdims=True)
    nrm = np.where(nrm < 1e-12, 1.0, nrm)
    rot = (q / nrm).astype(np.float32)

    sh = np.asarray(sh, dtype=np.float32).reshape(N, 48)
    sh_planes = np.ascontiguousarray(sh.reshape(N, 12, 4).transpose(1, 0, 2))
    return {
        "xyzb": np.ascontiguousarray(xyzb),
        "scale_lo": np.ascontiguousarray(scale_lo),
        "rot": np.ascontiguousarray(rot),
        "sh": sh_planes,
        "n": N,
    }


def bake(avatar) -> dict:
    """Bake a synthetic.Avatar / flame_io-loaded avatar."""
    return bake_avatar(avatar.xyz, avatar.scaling, avatar.rotation, avatar.opacity, avatar.sh, avatar.binding)


def binding_of(baked: dict) -> np.ndarray:
    return np.ascontiguousarray(baked["xyzb"][:, 3]).view(np.int32)

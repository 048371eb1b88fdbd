"""On-disk formats either side of the hot path.

  * GaussianAvatars point-cloud PLY (`point_cloud/iteration_<N>/point_cloud.ply`) incl. the
    per-Gaussian `binding_0` column [UPSTREAM, unverifiable offline];
  * `flame_param.npz` / `flame_param/%05d.npz` / `canonical_flame_param.npz` records
    (/root/reference/02_Visual_Engine/flame_fitter.py:431-441, preprocess_video.py:314-354);
  * `transforms_{train,test,val}.json` (preprocess_video.py:372-401);
  * the FLAME linear model: an .npz, or the published pickle read directly (chumpy-wrapped arrays, scipy-sparse
    joint regressor, uint32 kinematic table) without chumpy installed (the real file is licence-gated; the tests
    build one of the same structure).
"""
from __future__ import annotations

import json
import os
import pickle
import re
import zipfile
from dataclasses import dataclass

import numpy as np

from . import cameras as cam_mod
from .synthetic import PARENTS, Avatar, FlameModel, FrameParams

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
    "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
    "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


# --------------------------------------------------------------------------------------- PLY
def read_ply(path: str) -> dict[str, np.ndarray]:
    """Vertex element of a PLY file (binary little/big endian or ascii), scalar properties only."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt = None
        count = None
        props: list[tuple[str, str]] = []
        in_vertex = False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            tok = line.decode("ascii", errors="replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    count = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties on the vertex element are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if count is None or not props:
            raise ValueError(f"{path}: no vertex element")
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=count, dtype=np.float64).reshape(count, len(props))
            return {n: data[:, i].astype(t) for i, (n, t) in enumerate(props)}
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(n, order + t) for n, t in props])
        rec = np.frombuffer(f.read(count * dt.itemsize), dtype=dt, count=count)
        return {n: np.ascontiguousarray(rec[n]) for n, _ in props}


def write_ply(path: str, columns: dict[str, np.ndarray]) -> None:
    names = list(columns)
    n = len(columns[names[0]])
    dt = np.dtype([(k, "<" + np.asarray(columns[k]).dtype.str[1:]) for k in names])
    rec = np.empty(n, dtype=dt)
    for k in names:
        rec[k] = columns[k]
    rev = {"f4": "float", "f8": "double", "i4": "int", "u4": "uint", "u1": "uchar", "i2": "short", "u2": "ushort", "i1": "char"}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\n")
        f.write(f"element vertex {n}\n".encode())
        for k in names:
            f.write(f"property {rev[np.asarray(columns[k]).dtype.str[1:]]} {k}\n".encode())
        f.write(b"end_header\n")
        f.write(rec.tobytes())


def load_avatar_ply(path: str) -> Avatar:
    """GaussianAvatars checkpoint PLY -> raw avatar.  f_rest is stored channel-major
    (f_rest_{c*15+k}), as upstream's save_ply flattens features_rest.transpose(1, 2)."""
    c = read_ply(path)
    need = ["x", "y", "z", "opacity", "scale_0", "rot_0", "f_dc_0"]
    for k in need:
        if k not in c:
            raise ValueError(f"{path}: missing PLY property '{k}'")
    n = len(c["x"])
    xyz = np.stack([c["x"], c["y"], c["z"]], axis=1).astype(np.float32)
    scaling = np.stack([c[f"scale_{i}"] for i in range(3)], axis=1).astype(np.float32)
    rot = np.stack([c[f"rot_{i}"] for i in range(4)], axis=1).astype(np.float32)
    sh = np.zeros((n, 16, 3), np.float32)
    sh[:, 0, :] = np.stack([c[f"f_dc_{i}"] for i in range(3)], axis=1)
    n_rest = len([k for k in c if k.startswith("f_rest_")])
    if n_rest % 3:
        raise ValueError(f"{path}: f_rest count {n_rest} is not a multiple of 3")
    per = n_rest // 3
    if per > 15:
        raise ValueError(f"{path}: SH degree above 3 is not supported")
    for ch in range(3):
        for k in range(per):
            sh[:, 1 + k, ch] = c[f"f_rest_{ch * per + k}"]
    if "binding_0" not in c:
        raise ValueError(f"{path}: no binding_0 column — the avatar is not bound to a mesh (--bind_to_mesh)")
    binding = np.asarray(c["binding_0"]).astype(np.int64)
    if np.any(binding < 0):
        raise ValueError(f"{path}: negative binding index")
    return Avatar(xyz, scaling, rot, np.asarray(c["opacity"], np.float32), sh, binding.astype(np.int32))


def save_avatar_ply(path: str, av: Avatar) -> None:
    n = av.n
    cols: dict[str, np.ndarray] = {}
    for i, k in enumerate("xyz"):
        cols[k] = av.xyz[:, i].astype(np.float32)
    for k in ("nx", "ny", "nz"):
        cols[k] = np.zeros(n, np.float32)
    for i in range(3):
        cols[f"f_dc_{i}"] = av.sh[:, 0, i].astype(np.float32)
    for ch in range(3):
        for k in range(15):
            cols[f"f_rest_{ch * 15 + k}"] = av.sh[:, 1 + k, ch].astype(np.float32)
    cols["opacity"] = av.opacity.astype(np.float32)
    for i in range(3):
        cols[f"scale_{i}"] = av.scaling[:, i].astype(np.float32)
    for i in range(4):
        cols[f"rot_{i}"] = av.rotation[:, i].astype(np.float32)
    cols["binding_0"] = av.binding.astype(np.int32)
    write_ply(path, cols)


# --------------------------------------------------------------------------------------- FLAME model
def save_flame_model(path: str, m: FlameModel) -> None:
    np.savez(path, v_template=m.v_template, faces=m.faces, shapedirs=m.shapedirs, posedirs=m.posedirs,
             j_regressor=m.j_regressor, lbs_weights=m.lbs_weights, parents=m.parents)


class _ChumpyArray:
    """Stand-in for `chumpy.ch.Ch` while unpickling: the published FLAME pickles (flame2023.pkl, generic_model.pkl)
    store v_template, shapedirs, posedirs, J and weights as chumpy objects, and chumpy is neither maintained nor
    installed next to a renderer.  Only the wrapped ndarray (`x`) is kept."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"x": state})

    def __array__(self, dtype=None, copy=None):
        x = self.__dict__.get("x")
        if x is None:  # any other chumpy node type: the first array it holds
            x = next((v for v in self.__dict__.values() if isinstance(v, np.ndarray)), None)
        if x is None:
            raise ValueError("chumpy object without array data in the FLAME pickle")
        return np.asarray(x, dtype=dtype)


class _FlameUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "chumpy" or module.startswith("chumpy."):
            return _ChumpyArray
        return super().find_class(module, name)


def load_flame_model(path: str) -> FlameModel:
    """.npz written by save_flame_model, or a FLAME pickle as the reference reads it (flame_fitter.py:79-108:
    keys v_template, shapedirs [V,3,400], posedirs [V*3... as (V,3,36)], J_regressor (scipy sparse or dense), weights,
    f, kintree_table), with or without chumpy-wrapped arrays."""
    if not os.path.exists(path):
        raise FileNotFoundError(f"FLAME model not found: {path}")
    if path.endswith(".npz"):
        d = np.load(path)
        return FlameModel(*(np.ascontiguousarray(d[k]) for k in
                            ("v_template", "faces", "shapedirs", "posedirs", "j_regressor", "lbs_weights", "parents")))
    with open(path, "rb") as f:
        d = _FlameUnpickler(f, encoding="latin1").load()
    v = np.asarray(d["v_template"], np.float32)
    V = v.shape[0]
    sd = np.asarray(d["shapedirs"], np.float32)            # (V,3,400)
    pd = np.asarray(d["posedirs"], np.float32)             # (V,3,36)
    jr = d["J_regressor"]
    jr = np.asarray(jr.todense() if hasattr(jr, "todense") else jr, np.float32)
    parents = PARENTS.copy()
    if "kintree_table" in d:                                # flame_fitter.py:104-106: parents = kintree[0], root = -1
        parents = np.asarray(d["kintree_table"]).astype(np.int64)[0].copy()
        parents[0] = -1
        parents = parents.astype(np.int32)
        if not np.array_equal(parents, PARENTS):
            raise ValueError(f"{path}: kinematic tree {parents.tolist()} is not FLAME's {PARENTS.tolist()} "
                             "(root, neck, jaw, two eyes); the skinning kernel is built for that chain")
    if sd.shape[:2] != (V, 3) or pd.reshape(V * 3, -1).shape[1] != 36 or jr.shape != (5, V):
        raise ValueError(f"{path}: not a FLAME model (v_template {v.shape}, shapedirs {sd.shape}, posedirs {pd.shape}, "
                         f"J_regressor {jr.shape})")
    return FlameModel(
        v_template=v, faces=np.asarray(d["f"]).astype(np.int32),
        shapedirs=np.ascontiguousarray(sd.reshape(V * 3, -1).T),
        posedirs=np.ascontiguousarray(pd.reshape(V * 3, -1).T),
        j_regressor=np.ascontiguousarray(jr), lbs_weights=np.asarray(d["weights"], np.float32), parents=parents)


def check_subject_matches_model(model: FlameModel, params: "FrameParams", avatar) -> None:
    """The three inputs of a render must describe ONE mesh.  The reference's records and a GaussianAvatars
    checkpoint are written for the 5143-vertex FLAME-with-teeth mesh (static_offset (1,5143,3), dynamic_offset
    (T,5143,3): flame_fitter.py:439, preprocess_video.py:329); the licence-gated flame2023.pkl holds the raw
    5023-vertex mesh.  Upstream grafts the teeth on at load time from its own mask asset, which is not
    redistributable and not under the reference tree, so the augmented model has to be exported once
    (tools/export_flame_with_teeth.py, run inside the upstream checkout) — a mismatch is named here instead of
    surfacing as an out-of-range binding or a failed size check deep in the session."""
    rec_v = int(params.static_offset.shape[1])
    problems = []
    if rec_v != model.n_verts:
        problems.append(f"the FLAME records cover {rec_v} vertices, the model has {model.n_verts}")
    if avatar.n and (int(avatar.binding.max()) >= model.n_faces or int(avatar.binding.min()) < 0):
        problems.append(f"the avatar binds Gaussians to faces {int(avatar.binding.min())}..{int(avatar.binding.max())}, "
                        f"the model has {model.n_faces} faces")
    if problems:
        hint = ""
        if model.n_verts == 5023:
            hint = (" — this is the raw FLAME mesh; records and checkpoints of the reference pipeline use the "
                    "5143-vertex FLAME-with-teeth mesh: export it with tools/export_flame_with_teeth.py and point "
                    "$OMFS_FLAME_MODEL at the resulting flame_model.npz")
        raise ValueError("FLAME model does not match the dataset/checkpoint: " + "; ".join(problems) + hint)


# --------------------------------------------------------------------------------------- parameters
PARAM_KEYS = ("shape", "expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation",
              "static_offset", "dynamic_offset")


_NPY_HEADER = re.compile(rb"\{'descr': '([<|=][a-zA-Z][0-9]+)', 'fortran_order': False, 'shape': \(([0-9, ]*)\), \}")


def _npy_member(name: str, b, out: dict) -> None:
    if not name.endswith(".npy") or bytes(b[:6]) != b"\x93NUMPY":
        raise ValueError(name)
    if b[6] == 1:
        hlen, off = int.from_bytes(b[8:10], "little"), 10
    else:
        hlen, off = int.from_bytes(b[8:12], "little"), 12
    m = _NPY_HEADER.match(bytes(b[off:off + hlen]).strip())
    if m is None:
        raise ValueError(name)
    shape = tuple(int(x) for x in m.group(2).split(b",") if x.strip())
    out[name[:-4]] = np.frombuffer(b, dtype=np.dtype(m.group(1).decode()), offset=off + hlen).reshape(shape).copy()


def _read_stored_zip(path: str) -> dict[str, np.ndarray]:
    """The members of an archive of STORED entries (what np.savez writes), found by walking the local file headers of
    one read of the file: no central directory, no ZipInfo objects.  Raises ValueError for anything it does not cover
    (compression, data descriptors, encryption); the caller falls back to zipfile / np.load."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    out, pos, n = {}, 0, len(buf)
    while pos + 4 <= n and bytes(buf[pos:pos + 4]) == b"PK\x03\x04":
        if pos + 30 > n:
            raise ValueError(path)
        flags, method = int.from_bytes(buf[pos + 6:pos + 8], "little"), int.from_bytes(buf[pos + 8:pos + 10], "little")
        csize, usize = int.from_bytes(buf[pos + 18:pos + 22], "little"), int.from_bytes(buf[pos + 22:pos + 26], "little")
        nlen, xlen = int.from_bytes(buf[pos + 26:pos + 28], "little"), int.from_bytes(buf[pos + 28:pos + 30], "little")
        if method != 0 or flags & 0x0009:
            raise ValueError(path)
        name = bytes(buf[pos + 30:pos + 30 + nlen]).decode("utf-8")
        extra = buf[pos + 30 + nlen:pos + 30 + nlen + xlen]
        if csize == 0xFFFFFFFF or usize == 0xFFFFFFFF:   # zip64: the sizes are in the extra field (np.savez forces it)
            x, found = 0, False
            while x + 4 <= len(extra):
                hid, hsz = int.from_bytes(extra[x:x + 2], "little"), int.from_bytes(extra[x + 2:x + 4], "little")
                if hid == 1 and hsz >= 16:
                    usize, csize = int.from_bytes(extra[x + 4:x + 12], "little"), int.from_bytes(extra[x + 12:x + 20], "little")
                    found = True
                    break
                x += 4 + hsz
            if not found:
                raise ValueError(path)
        if csize != usize:
            raise ValueError(path)
        start = pos + 30 + nlen + xlen
        if start + csize > n:
            raise ValueError(path)
        _npy_member(name, buf[start:start + csize], out)
        pos = start + csize
    if not out or bytes(buf[pos:pos + 4]) not in (b"PK\x01\x02", b"PK\x06\x06", b"PK\x05\x06"):
        raise ValueError(path)   # the walk must end at the central directory
    return out


def read_npz(path: str) -> dict[str, np.ndarray]:
    """np.load(path) as a dict, for the layout np.savez writes (stored .npy members, C order, plain little-endian
    numeric dtypes): the archive's local headers are walked in one read of the file and the .npy headers are matched
    instead of evaluated, which makes a per-frame FLAME record 0.15 ms instead of 1.3 ms to read — a 300-frame
    dataset is 300 of them.  Anything else (compressed members, object arrays, Fortran order, big-endian) goes
    through zipfile and then np.load."""
    try:
        return _read_stored_zip(path)
    except (ValueError, TypeError, UnicodeDecodeError):
        pass
    out = {}
    try:
        with zipfile.ZipFile(path) as z:
            for name in z.namelist():
                _npy_member(name, z.read(name), out)
        return out
    except (ValueError, TypeError, zipfile.BadZipFile):
        return dict(np.load(path, allow_pickle=True))


def load_flame_params(path: str, n_verts: int) -> FrameParams:
    return FrameParams.from_dict(read_npz(path), n_verts=n_verts)


def save_flame_params(path: str, p: FrameParams) -> None:
    np.savez(path, **p.as_dict())


def stack_frame_params(records: list[FrameParams]) -> FrameParams:
    """Per-frame records (T=1 each, as `flame_param/%05d.npz` holds) -> one batched record."""
    first = records[0]
    cat = lambda k: np.concatenate([getattr(r, k) for r in records], axis=0)
    return FrameParams(first.shape, cat("expr"), cat("rotation"), cat("neck_pose"), cat("jaw_pose"),
                       cat("eyes_pose"), cat("translation"), first.static_offset, cat("dynamic_offset"))


@dataclass
class DatasetFrame:
    file_path: str
    flame_param_path: str
    timestep_index: int
    camera_index: int
    camera: cam_mod.Camera


def load_transforms(data_dir: str, split: str = "train") -> list[DatasetFrame]:
    path = os.path.join(data_dir, f"transforms_{split}.json")
    if not os.path.exists(path):
        raise FileNotFoundError(f"transforms file not found: {path}")
    with open(path, "r") as f:
        t = json.load(f)
    out = []
    for i, fr in enumerate(t.get("frames", [])):
        w = int(fr.get("w", t.get("w", 512)))
        h = int(fr.get("h", t.get("h", 512)))
        fovx = float(fr.get("camera_angle_x", t.get("camera_angle_x", 0.3)))
        cam = cam_mod.camera_from_c2w(np.array(fr["transform_matrix"], dtype=np.float64), fovx, w, h)
        out.append(DatasetFrame(fr.get("file_path", f"images/{i:05d}_00.png"),
                                fr.get("flame_param_path", f"flame_param/{int(fr.get('timestep_index', i)):05d}.npz"),
                                int(fr.get("timestep_index", i)), int(fr.get("camera_index", 0)), cam))
    return out


def load_dataset_params(data_dir: str, frames: list[DatasetFrame], n_verts: int) -> FrameParams:
    """One FrameParams row per dataset frame, read from each frame's own flame_param file (what the
    reference's modified dataset points at, render_surgery.py:211-215)."""
    cache: dict[str, FrameParams] = {}
    recs = []
    for fr in frames:
        p = os.path.join(data_dir, fr.flame_param_path)
        if p not in cache:
            if not os.path.exists(p):
                raise FileNotFoundError(f"FLAME parameter file not found: {p}")
            cache[p] = load_flame_params(p, n_verts)
        rec = cache[p]
        if rec.n_frames == 1:
            recs.append(rec)
        else:  # a batched file: pick the row of this timestep
            t = fr.timestep_index
            recs.append(rec.slice(t, t + 1))
    return stack_frame_params(recs)


# --------------------------------------------------------------------------------------- synthetic dataset on disk
def write_synthetic_dataset(data_dir: str, model_dir: str, model: FlameModel, params: FrameParams, avatar: Avatar,
                            cam_c2w: np.ndarray, camera_angle_x: float, width: int, height: int,
                            iteration: int = 30000, write_images: bool = False) -> None:
    """Lay a synthetic subject out exactly as the reference's pipeline leaves a real one on disk:
    dataset dir (preprocess_video.py:246-416) + trained-model dir (render_surgery.py:271-287)."""
    os.makedirs(os.path.join(data_dir, "images"), exist_ok=True)
    os.makedirs(os.path.join(data_dir, "flame_param"), exist_ok=True)
    T = params.n_frames
    for t in range(T):
        save_flame_params(os.path.join(data_dir, "flame_param", f"{t:05d}.npz"), params.slice(t, t + 1))
    save_flame_params(os.path.join(data_dir, "flame_param.npz"), params)
    z = lambda *s: np.zeros(s, np.float32)
    canon = FrameParams(params.shape, z(1, params.expr.shape[1]), z(1, 3), z(1, 3), z(1, 3), z(1, 6), z(1, 3),
                        params.static_offset, z(1, params.static_offset.shape[1], 3))
    save_flame_params(os.path.join(data_dir, "canonical_flame_param.npz"), canon)
    frames = [{"file_path": f"images/{t:05d}_00.png", "flame_param_path": f"flame_param/{t:05d}.npz",
               "transform_matrix": np.asarray(cam_c2w).tolist(), "timestep_index": t, "camera_index": 0,
               "camera_angle_x": camera_angle_x, "w": width, "h": height,
               "fg_mask_path": f"fg_masks/{t:05d}_00.png"} for t in range(T)]
    # top-level intrinsics as preprocess_video.py:389-401 writes them
    fl_x = cam_mod.fov2focal(camera_angle_x, width)
    top = {"camera_angle_x": camera_angle_x, "camera_angle_y": cam_mod.focal2fov(fl_x, height), "fl_x": fl_x,
           "fl_y": fl_x, "cx": width / 2.0, "cy": height / 2.0, "w": width, "h": height, "frames": frames,
           "timestep_indices": list(range(T)), "camera_indices": [0]}
    if write_images:
        # placeholder ground-truth frames and foreground masks (the renderer never reads them; the single-frame
        # experiment and the validation report copy / compare them)
        from PIL import Image
        os.makedirs(os.path.join(data_dir, "fg_masks"), exist_ok=True)
        yy, xx = np.mgrid[0:height, 0:width]
        for t in range(T):
            img = np.stack([(xx + 3 * t) % 256, (yy + 5 * t) % 256, (xx + yy) % 256], -1).astype(np.uint8)
            Image.fromarray(img).save(os.path.join(data_dir, "images", f"{t:05d}_00.png"))
            Image.fromarray(np.full((height, width), 255, np.uint8)).save(
                os.path.join(data_dir, "fg_masks", f"{t:05d}_00.png"))
    split = max(1, T - T // 10)
    for name, fr in (("train", frames[:split]), ("test", frames[split:]), ("val", frames[split:])):
        with open(os.path.join(data_dir, f"transforms_{name}.json"), "w") as f:
            json.dump({**top, "frames": fr}, f, indent=2)
    write_ply(os.path.join(data_dir, "points3d.ply"), {"x": z(1), "y": z(1), "z": z(1)})
    save_avatar_ply(os.path.join(model_dir, "point_cloud", f"iteration_{iteration}", "point_cloud.ply"), avatar)
    save_flame_model(os.path.join(model_dir, "flame_model.npz"), model)

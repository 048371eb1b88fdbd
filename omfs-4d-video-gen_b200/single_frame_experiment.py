"""Single-frame experiment — the driver of BASELINE.json configs[1], with the reference's function names
(02_Visual_Engine/single_frame_experiment.py: build_single_frame_dataset :32-81, render_single_frame_and_save
:108-162, main :165-173).

The reference hard-codes its directories as module constants and shells out twice (train_ghost.py, then
render_surgery.py with zero offsets).  Here the directories are arguments with the same defaults, the render
is the in-process B200 path (`render_surgery.render_with_gaussians`), and training is out of scope
(SURVEY.md §2.1): `train_single_frame` only checks that a trained avatar is where the reference's trainer
would have left it.
"""
from __future__ import annotations

import json
import os
import shutil
import sys
from pathlib import Path

import numpy as np

from . import render_surgery

VISUAL_DIR = Path("02_Visual_Engine")
DATA_CONDA = VISUAL_DIR / "data_conda"
DATA_SINGLE = VISUAL_DIR / "data_single_frame"
MODEL_SINGLE = VISUAL_DIR / "output" / "model_single_frame"

_TOP_KEYS = ("camera_angle_x", "camera_angle_y", "fl_x", "fl_y", "cx", "cy", "w", "h")
_FRAME0 = {"images": "00000_00.png", "flame_param": "00000.npz", "fg_masks": "00000_00.png"}


def build_single_frame_dataset(data_conda: Path = DATA_CONDA, data_single: Path = DATA_SINGLE) -> Path:
    """`data_single` := frame 0 of `data_conda` (image, flame_param, fg_mask), one-frame transforms for all
    three splits, the batched flame_param.npz with a leading frame axis, and the canonical record."""
    src, dst = Path(data_conda), Path(data_single)
    if dst.exists():
        shutil.rmtree(dst)
    for sub, name in _FRAME0.items():
        (dst / sub).mkdir(parents=True, exist_ok=True)
        shutil.copy2(src / sub / name, dst / sub / name)
    with open(src / "transforms_train.json") as f:
        full = json.load(f)
    single = {k: full[k] for k in _TOP_KEYS}          # KeyError on an incomplete transforms file, as upstream
    single["frames"] = [full["frames"][0]]
    for split in ("train", "test", "val"):
        with open(dst / f"transforms_{split}.json", "w") as f:
            json.dump(single, f, indent=2)
    rec = dict(np.load(src / "flame_param" / _FRAME0["flame_param"], allow_pickle=True))
    batched = {k: (v if v.ndim == 1 or v.shape[0] == 1 else v[None, ...]) for k, v in rec.items()}
    np.savez(dst / "flame_param.npz", **batched)
    shutil.copy2(src / "canonical_flame_param.npz", dst / "canonical_flame_param.npz")
    print(f"[single_frame] Built {dst} (1 frame)")
    return dst


def train_single_frame(model_single: Path = MODEL_SINGLE) -> None:
    """Training the avatar is not part of this repository's path: accept a model directory that the reference's
    trainer (train_ghost.py) produced, fail the way the reference does otherwise."""
    pc = Path(model_single) / "point_cloud"
    if not pc.is_dir() or not any(p.name.startswith("iteration_") for p in pc.iterdir()):
        raise RuntimeError("Training failed")      # the reference's message when train_ghost.py returns non-zero
    print("[single_frame] Using the trained avatar in", model_single)


def _newest_renders(model_single: Path) -> Path | None:
    train_dir = Path(model_single) / "train"
    if not train_dir.exists():
        return None

    def iteration_of(p: Path) -> int:
        try:
            return int(p.name.split("_")[-1]) if p.name.startswith("ours_") else -1
        except (ValueError, IndexError):
            return -1

    for d in sorted(train_dir.iterdir(), key=iteration_of, reverse=True):
        r = d / "renders"
        if r.exists() and list(r.glob("*.png")):
            return r
    return None


def render_single_frame_and_save(model_single: Path = MODEL_SINGLE, data_single: Path = DATA_SINGLE,
                                 out_dir: Path = VISUAL_DIR) -> tuple[Path, Path]:
    """Render the one view with zero surgery offsets and save GT + render side by side
    (single_frame_render.png, single_frame_gt.png in `out_dir`).  Returns (render, gt)."""
    model_single, data_single = Path(model_single), Path(data_single)
    modified = render_surgery.create_modified_dataset(str(data_single), render_surgery.compute_offset(0.0, 1.0),
                                                      render_surgery.compute_offset(0.0, 1.0))
    try:
        render_surgery.render_with_gaussians(str(model_single), modified)
    except Exception as e:
        print(str(e)[-1500:])
        raise RuntimeError("Render failed") from e
    finally:
        shutil.rmtree(modified, ignore_errors=True)
    renders_dir = _newest_renders(model_single)
    if renders_dir is None:
        train_dir = model_single / "train"
        subdirs = [p.name for p in train_dir.iterdir()] if train_dir.exists() else []
        raise FileNotFoundError("No rendered frames found. Looked in model_single_frame/train/*/renders/. "
                                f"Subdirs: {subdirs}")
    out_dir = Path(out_dir).resolve()
    out_dir.mkdir(parents=True, exist_ok=True)
    render_dst, gt_dst = out_dir / "single_frame_render.png", out_dir / "single_frame_gt.png"
    shutil.copy2(sorted(renders_dir.glob("*.png"))[0], render_dst)
    shutil.copy2(data_single / "images" / _FRAME0["images"], gt_dst)
    print("\n--- OUTPUT FILES (absolute paths) ---")
    print(f"  Render: {render_dst}")
    print(f"  GT:     {gt_dst}")
    print("------------------------------------")
    return render_dst, gt_dst


def main(data_conda: Path = DATA_CONDA, data_single: Path = DATA_SINGLE, model_single: Path = MODEL_SINGLE,
         out_dir: Path = VISUAL_DIR) -> None:
    if not Path(data_conda).exists():
        print("Run the full conda pipeline first to create 02_Visual_Engine/data_conda")
        sys.exit(1)
    build_single_frame_dataset(data_conda, data_single)
    train_single_frame(model_single)
    render, gt = render_single_frame_and_save(model_single, data_single, out_dir)
    print(f"\nDone. Open: {gt} and {render}")


if __name__ == "__main__":
    main()

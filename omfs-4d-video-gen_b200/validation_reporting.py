"""Deterministic validation reporting — the reference's 02_Visual_Engine/validation_reporting.py kept under the
same names, arguments, file layout and exceptions (psnr :16-20, ssim_global :23-37, _bucket :40-45,
_find_latest_train_dir :48-55, generate_report :58-123, main :126-140), plus a device path.

The reference compares rendered PNGs with ground-truth photos on the host.  `frame_metrics_device` gives the same
two numbers (PSNR, global SSIM) for frames that are still in HBM — e.g. a planned render against the zero-offset
render of the same clip (the A/B the deterministic export exists for) — from ONE pass of omfs_frame_metrics over both
frame sets: the kernel accumulates the six moments per frame pair, the closed forms below finish them.
"""
from __future__ import annotations

import argparse
import json
import math
from pathlib import Path

import numpy as np

C1 = (0.01 * 255) ** 2
C2 = (0.03 * 255) ** 2
LUMA = (0.299, 0.587, 0.114)            # BT.601 weights the reference applies to RGB inputs (:24-27)
VIEW_BANDS = ("front", "profile", "rear")


# One implementation serves the host and the device path: both reduce a frame pair to the same moments —
#   [0] sum (a-b)^2 over every channel value, then with x, y the luma of a, b: [1] sum x, [2] sum y,
#   [3] sum x^2, [4] sum y^2, [5] sum xy — and `metrics_from_moments` finishes PSNR and the global SSIM from them.
# On the device the moments come from omfs_frame_metrics (one pass over frames in HBM); on the host from numpy.
def _luma(img: np.ndarray) -> np.ndarray:
    """Grey image the global SSIM is computed on: RGB inputs are weighted in their own dtype (the reference keeps
    float32 there), grey inputs pass through; the moments are then taken in float64."""
    img = np.asarray(img)
    if img.ndim == 3:
        img = LUMA[0] * img[:, :, 0] + LUMA[1] * img[:, :, 1] + LUMA[2] * img[:, :, 2]
    return img.astype(np.float64)


def frame_moments(a: np.ndarray, b: np.ndarray) -> tuple[np.ndarray, int, int]:
    """(moments[6], number of channel values, number of pixels) of one frame pair, in float64."""
    a, b = np.asarray(a), np.asarray(b)
    d = a.astype(np.float64) - b.astype(np.float64)
    x, y = _luma(a).ravel(), _luma(b).ravel()
    m = np.array([np.dot(d.ravel(), d.ravel()), x.sum(), y.sum(), np.dot(x, x), np.dot(y, y), np.dot(x, y)], np.float64)
    return m, int(d.size), int(x.size)


def metrics_from_moments(moments: np.ndarray, n_pixels: int, n_values: int | None = None) -> tuple[np.ndarray, np.ndarray]:
    """PSNR and global SSIM per frame pair from moments [T,6] (layout above).  `n_values` is the number of channel
    values behind moment 0 (3 per pixel for RGB frames, the default).  PSNR is 20 log10(255 / sqrt(MSE)) on the
    0-255 scale and 99.0 for identical frames (reference :16-20); SSIM is the single-window form with the usual
    constants (:23-37), written with E[x^2] - mu^2 for the variances."""
    m = np.atleast_2d(np.asarray(moments, dtype=np.float64))
    n = float(n_pixels)
    mse = m[:, 0] / float(3 * n_pixels if n_values is None else n_values)
    same = mse == 0.0
    p = np.where(same, 99.0, 20.0 * np.log10(255.0 / np.sqrt(np.where(same, 1.0, mse))))
    mu_x, mu_y = m[:, 1] / n, m[:, 2] / n
    var_x = m[:, 3] / n - mu_x * mu_x
    var_y = m[:, 4] / n - mu_y * mu_y
    cov = m[:, 5] / n - mu_x * mu_y
    s = ((2 * mu_x * mu_y + C1) * (2 * cov + C2)) / ((mu_x * mu_x + mu_y * mu_y + C1) * (var_x + var_y + C2))
    return p, s


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    """Reference :16-20: PSNR of two 0-255 arrays, 99.0 when they are identical."""
    m, n_values, n_pix = frame_moments(a, b)
    return float(metrics_from_moments(m, n_pix, n_values)[0][0])


def ssim_global(a: np.ndarray, b: np.ndarray) -> float:
    """Reference :23-37: one SSIM over the whole image (no windows), on the luma of RGB inputs."""
    m, n_values, n_pix = frame_moments(a, b)
    return float(metrics_from_moments(m, n_pix, n_values)[1][0])


def _bucket(progress: float) -> str:
    """View band of a frame at `progress` in [0, 1] of the clip (reference :40-45): the middle third-ish of the turn
    is the profile, the two ends are frontal, what lies between is called rear."""
    if 0.35 <= progress <= 0.65:
        return "profile"
    return "rear" if 0.20 <= progress <= 0.80 else "front"


def _find_latest_train_dir(model_path: Path) -> Path:
    """`train/ours_<N>` with the largest N (reference :48-55)."""
    train_dir = Path(model_path) / "train"
    if not train_dir.exists():
        raise FileNotFoundError(f"Missing train directory: {train_dir}")
    runs = {int(p.name.rsplit("_", 1)[1]): p for p in train_dir.iterdir() if p.is_dir() and p.name.startswith("ours_")}
    if not runs:
        raise FileNotFoundError(f"No ours_* directories in {train_dir}")
    return runs[max(runs)]


def _summarise(metrics: list[dict]) -> dict:
    """Row count and per-band mean PSNR / SSIM (reference :95-106); an empty band reports nulls."""
    bands: dict[str, list[dict]] = {name: [] for name in VIEW_BANDS}
    for row in metrics:
        bands[row["bucket"]].append(row)
    by_bucket = {}
    for name, rows in bands.items():
        by_bucket[name] = {"count": len(rows)}
        for key in ("psnr", "ssim"):
            by_bucket[name][key] = float(np.mean([r[key] for r in rows])) if rows else None
    return {"count": len(metrics), "by_bucket": by_bucket}


CHECKLIST = """# Human Review Checklist

- [ ] Jawline continuity in profile views.
- [ ] Ear geometry plausibility in left/right profile.
- [ ] Neck-head transition remains stable across motion.
- [ ] No visible shimmer/flicker in slow turns.
- [ ] Maxilla/mandible changes remain anatomically plausible.
"""


def _write_report(output_dir: Path, summary: dict, metrics: list[dict]) -> None:
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    scores_path = output_dir / "strict_scores.json"
    with open(scores_path, "w", encoding="utf-8") as f:
        json.dump({"summary": summary, "rows": metrics}, f, indent=2)
    checklist_path = output_dir / "human_review_checklist.md"
    checklist_path.write_text(CHECKLIST, encoding="utf-8")
    print(f"[validation_reporting] Wrote strict report: {scores_path}")
    print(f"[validation_reporting] Wrote checklist: {checklist_path}")


def _score_rows(exports: list[dict], score) -> list[dict]:
    """One report row per manifest entry; `score(index, source_name)` returns (psnr, ssim) or None to skip."""
    top = max((int(r.get("index", 0)) for r in exports), default=1)
    scored = []
    for entry in exports:
        i, name = int(entry["index"]), entry["source"]
        ps = score(i, name)
        if ps is None:
            continue
        where = i / max(1, top)
        scored.append({"index": i, "frame": name, "progress": where, "bucket": _bucket(where),
                       "psnr": float(ps[0]), "ssim": float(ps[1])})
    return scored


def generate_report(model_path: Path, deterministic_frames_dir: Path, output_dir: Path):
    """renders/ vs gt/ of the newest `ours_<N>` directory, for the frames of the deterministic manifest."""
    from PIL import Image
    newest = _find_latest_train_dir(Path(model_path))
    folders = {k: newest / k for k in ("renders", "gt")}
    if not all(f.exists() for f in folders.values()):
        raise FileNotFoundError(f"Missing renders/gt directories in {newest}")
    manifest = Path(deterministic_frames_dir) / "deterministic_indices_manifest.json"
    if not manifest.exists():
        raise FileNotFoundError(f"Missing deterministic manifest: {manifest}")
    exports = json.loads(manifest.read_text(encoding="utf-8")).get("exports", [])

    def rgb(path: Path) -> np.ndarray:
        return np.asarray(Image.open(path).convert("RGB"), dtype=np.float32)

    def score(_, name):
        pair = [folders[k] / name for k in ("renders", "gt")]
        if not all(q.exists() for q in pair):
            return None
        a, b = rgb(pair[0]), rgb(pair[1])
        return psnr(a, b), ssim_global(a, b)

    rows = _score_rows(exports, score)
    _write_report(output_dir, _summarise(rows), rows)


# ----------------------------------------------------------------------------- device path
def frame_metrics_device(d_a_u8: int, d_b_u8: int, n_frames: int, height: int, width: int, stream: int = 0):
    """PSNR / SSIM of two uint8 [T,H,W,3] frame sets resident in HBM (device pointers).  Returns (psnr[T], ssim[T])."""
    from . import runtime
    L = runtime.load_library()
    out = runtime.DeviceArray((n_frames, 6), np.float64)
    try:
        runtime.check(L.omfs_frame_metrics(n_frames, height, width, d_a_u8, d_b_u8, out.ptr, stream))
        moments = out.numpy()
    finally:
        out.free()
    return metrics_from_moments(moments, height * width)


def generate_report_device(d_a_u8: int, d_b_u8: int, n_frames: int, height: int, width: int, indices: list[int],
                           output_dir: Path, frame_names: list[str] | None = None) -> dict:
    """The strict report for an A/B of two frame sets in HBM (same JSON layout as generate_report; `indices` are
    the deterministic export's selected indices)."""
    p, s = frame_metrics_device(d_a_u8, d_b_u8, n_frames, height, width)
    exports = [{"index": int(i), "source": frame_names[i] if frame_names else f"{i:05d}.png"}
               for i in indices if 0 <= i < n_frames]
    rows = _score_rows(exports, lambda i, _: (p[i], s[i]))
    summary = _summarise(rows)
    _write_report(output_dir, summary, rows)
    return {"summary": summary, "rows": rows}


def main(argv: list[str] | None = None):
    parser = argparse.ArgumentParser(description="Generate deterministic validation report.")
    parser.add_argument("--model_path", required=True, type=Path)
    parser.add_argument("--deterministic_frames_dir", required=True, type=Path)
    parser.add_argument("--output_dir", type=Path, default=Path("02_Visual_Engine/output/model/eval_strict/reports"))
    args = parser.parse_args(argv)
    generate_report(args.model_path, args.deterministic_frames_dir, args.output_dir)


if __name__ == "__main__":
    main()

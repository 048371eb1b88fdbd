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


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    """20 log10(255 / sqrt(MSE)) on 0-255 float arrays; 99.0 when the images are identical."""
    mse = float(np.mean((a - b) ** 2))
    if mse == 0.0:
        return 99.0
    return 20.0 * math.log10(255.0 / math.sqrt(mse))


def ssim_global(a: np.ndarray, b: np.ndarray) -> float:
    """One SSIM over the whole image (no windows), on the BT.601 luma of RGB inputs."""
    if a.ndim == 3:
        a = (0.299 * a[:, :, 0] + 0.587 * a[:, :, 1] + 0.114 * a[:, :, 2])
    if b.ndim == 3:
        b = (0.299 * b[:, :, 0] + 0.587 * b[:, :, 1] + 0.114 * b[:, :, 2])
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    mu_x = a.mean()
    mu_y = b.mean()
    sig_x = ((a - mu_x) ** 2).mean()
    sig_y = ((b - mu_y) ** 2).mean()
    sig_xy = ((a - mu_x) * (b - mu_y)).mean()
    return float(((2 * mu_x * mu_y + C1) * (2 * sig_xy + C2)) / ((mu_x * mu_x + mu_y * mu_y + C1) * (sig_x + sig_y + C2)))


def _bucket(progress: float) -> str:
    if progress < 0.20 or progress > 0.80:
        return "front"
    if 0.35 <= progress <= 0.65:
        return "profile"
    return "rear"


def _find_latest_train_dir(model_path: Path) -> Path:
    train_dir = Path(model_path) / "train"
    if not train_dir.exists():
        raise FileNotFoundError(f"Missing train directory: {train_dir}")
    dirs = [p for p in train_dir.iterdir() if p.is_dir() and p.name.startswith("ours_")]
    if not dirs:
        raise FileNotFoundError(f"No ours_* directories in {train_dir}")
    return sorted(dirs, key=lambda p: int(p.name.split("_")[-1]), reverse=True)[0]


def _summarise(metrics: list[dict]) -> dict:
    summary = {"count": len(metrics), "by_bucket": {}}
    for bucket in ("front", "profile", "rear"):
        vals = [m for m in metrics if m["bucket"] == bucket]
        if not vals:
            summary["by_bucket"][bucket] = {"count": 0, "psnr": None, "ssim": None}
            continue
        summary["by_bucket"][bucket] = {
            "count": len(vals),
            "psnr": float(np.mean([v["psnr"] for v in vals])),
            "ssim": float(np.mean([v["ssim"] for v in vals])),
        }
    return summary


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
def metrics_from_moments(moments: np.ndarray, n_pixels: int) -> tuple[np.ndarray, np.ndarray]:
    """PSNR and global SSIM per frame pair from omfs_frame_metrics' moments [T,6] =
    (sum (a-b)^2 over 3*n_pixels channel values, sum x, sum y, sum x^2, sum y^2, sum xy), x/y = luma."""
    m = np.asarray(moments, dtype=np.float64)
    n = float(n_pixels)
    mse = m[:, 0] / (3.0 * n)
    with np.errstate(divide="ignore"):
        p = np.where(mse == 0.0, 99.0, 20.0 * np.log10(255.0 / np.sqrt(np.where(mse == 0.0, 1.0, mse))))
    mu_x, mu_y = m[:, 1] / n, m[:, 2] / n
    sig_x = m[:, 3] / n - mu_x * mu_x
    sig_y = m[:, 4] / n - mu_y * mu_y
    sig_xy = m[:, 5] / n - mu_x * mu_y
    s = ((2 * mu_x * mu_y + C1) * (2 * sig_xy + C2)) / ((mu_x * mu_x + mu_y * mu_y + C1) * (sig_x + sig_y + C2))
    return p, s


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

"""The reference's image-parity metric (02_Visual_Engine/validation_reporting.py:16-20), kept under
the same name so that the report code can import it from here."""
from __future__ import annotations

import math

import numpy as np


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    """20 log10(255 / sqrt(MSE)) on 0-255 float arrays; 99.0 when the images are identical."""
    mse = float(np.mean((a - b) ** 2))
    if mse == 0.0:
        return 99.0
    return 20.0 * math.log10(255.0 / math.sqrt(mse))

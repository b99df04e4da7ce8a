"""Blend weight masks (estimator/models/utils.py:51-60), cached per shape.

The mask is produced by the same OpenCV call the reference makes (third-party arithmetic is
called, not restated) so the count maps stay bit-identical; unlike the reference it is computed
once per (shape, border) instead of on every forward.
"""
from __future__ import annotations

from functools import lru_cache

import cv2
import numpy as np


@lru_cache(maxsize=32)
def _mask(h: int, w: int, border: float) -> np.ndarray:
    size = (h, w)
    mask = np.zeros(size, dtype=np.float32)
    sigma = int(size[0] / 16)
    k_size = int(2 * np.ceil(2 * int(size[0] / 16)) + 1)
    mask[int(border * size[0]):size[0] - int(border * size[0]), int(border * size[1]):size[1] - int(border * size[1])] = 1
    mask = cv2.GaussianBlur(mask, (int(k_size), int(k_size)), sigma)
    mask = (mask - mask.min()) / (mask.max() - mask.min())
    mask = mask.astype(np.float32)
    mask.setflags(write=False)
    return mask


def generatemask(size, border: float = 0.1) -> np.ndarray:
    """float32 [h,w] Gaussian-blurred box, min-max normalised (utils.py:51-60)."""
    return _mask(int(size[0]), int(size[1]), float(border))


def random_patch_mask(size, border: float = 0.15) -> np.ndarray:
    """patchrefiner.py:386: ``generatemask(raw patch, border=0.15) + 1e-3`` (float32 + Python float
    stays float32 in NumPy)."""
    return (generatemask(size, border) + 1e-3).astype(np.float32)

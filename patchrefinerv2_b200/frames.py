"""Frame ingest / egress around the hot path: what the reference's ``ImageDataset`` and ``Tester.run`` do to a
file on disk before and after ``model(mode='infer')``.  Host code, PyTorch + OpenCV,
exactly the calls the reference makes -- this is not the accelerated path.

  * read + colour convert + ``/255`` + bicubic(align_corners=True) resize to ``image_raw_shape``
    (estimator/datasets/general_dataset.py:52-59)
  * ``image_lr`` = the model's resizer applied to the full frame (general_dataset.py:218; on the device here)
  * uint16 PNG with multiplier 256 and an 8-bit preview (estimator/tester/tester.py:72-91)
"""
from __future__ import annotations

import os
from typing import Iterator, List, Tuple

import cv2
import numpy as np
import torch
import torch.nn.functional as F

IMAGE_EXT = (".png", ".jpg", ".jpeg", ".bmp", ".tif", ".tiff", ".webp")


def list_frames(rgb_image_dir: str) -> List[str]:
    """sorted(os.listdir(dir)) like general_dataset.py:180, restricted to image files."""
    if not os.path.isdir(rgb_image_dir):
        raise FileNotFoundError(f"rgb_image_dir {rgb_image_dir!r} is not a directory")
    return [f for f in sorted(os.listdir(rgb_image_dir)) if f.lower().endswith(IMAGE_EXT)]


def basename(name: str) -> str:
    """general_dataset.py:68-71."""
    for ext in (".jpg", ".png", ".jpeg"):
        name = name.replace(ext, "")
    return name


def read_frame(path: str, image_raw_shape) -> torch.Tensor:
    """-> fp32 RGB [3,H,W] in [0,1] at ``image_raw_shape`` (general_dataset.py:52-59: cv2 BGR->RGB, /255 in float64,
    F.interpolate bicubic align_corners=True, then ToTensor + .float())."""
    img = cv2.imread(path)
    if img is None:
        raise IOError(f"cannot read image {path!r}")
    if img.ndim == 2:
        img = cv2.cvtColor(img, cv2.COLOR_GRAY2BGR)
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB) / 255.0
    t = torch.tensor(img).unsqueeze(0).permute(0, 3, 1, 2)
    t = F.interpolate(t, tuple(int(x) for x in image_raw_shape), mode="bicubic", align_corners=True)
    return t[0].float().contiguous()


def iter_frames(rgb_image_dir: str, image_raw_shape) -> Iterator[Tuple[str, torch.Tensor]]:
    for f in list_frames(rgb_image_dir):
        yield basename(f), read_frame(os.path.join(rgb_image_dir, f), image_raw_shape)


def save_prediction(depth: torch.Tensor, work_dir: str, name: str, gray_scale: bool = False, coarse: torch.Tensor = None,
                    image_raw_shape=None) -> dict:
    """tester.py:72-99 without the matplotlib / kornia dependencies: ``<name>_uint16.png`` is bit-for-bit the
    reference's (``(depth*256).astype('uint16')``); the 8-bit previews use the same 0..100 percentile window
    (estimator/utils/colorize.py) through an OpenCV colour map instead of matplotlib's ``Spectral``."""
    os.makedirs(work_dir, exist_ok=True)
    d = depth.detach().squeeze().float().cpu().numpy()
    out = {"uint16": os.path.join(work_dir, f"{name}_uint16.png"), "preview": os.path.join(work_dir, f"{name}.png")}
    cv2.imwrite(out["uint16"], (d * 256).astype("uint16"))
    cv2.imwrite(out["preview"], _preview(d, gray_scale))
    if coarse is not None:
        c = coarse.detach().float()
        if image_raw_shape is not None:
            c = F.interpolate(c, tuple(int(x) for x in image_raw_shape), mode="bilinear")        # tester.py:93
        out["coarse"] = os.path.join(work_dir, f"{name}_coarse.png")
        cv2.imwrite(out["coarse"], _preview(c.squeeze().cpu().numpy(), gray_scale))
    return out


def _preview(d: np.ndarray, gray_scale: bool) -> np.ndarray:
    valid = d > -99
    vmin, vmax = (np.percentile(d[valid], 0), np.percentile(d[valid], 100)) if valid.any() else (0.0, 1.0)
    n = (d - vmin) / (vmax - vmin) if vmax > vmin else d * 0.0
    g = np.clip(n * 255.0, 0, 255).astype(np.uint8)
    if gray_scale:
        return 255 - g                                   # 'gray_r'
    return cv2.applyColorMap(g, cv2.COLORMAP_TURBO)

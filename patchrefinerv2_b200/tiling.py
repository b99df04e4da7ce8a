"""Host-side tiling geometry of the high-resolution inference path (pure Python / NumPy).

Mirrors, value for value, what the reference computes on the host:
  * ``prepare_tile_cfg``        estimator/models/baseline_pretrain.py:96-124
  * regular / shifted grids     estimator/models/baseline_pretrain.py:249-270
  * random patches              estimator/models/baseline_pretrain.py:158-161 (global ``random`` stream:
                                ``process_num`` row draws, then ONE column draw shared by them)
  * bbox -> feature coordinates estimator/models/baseline_pretrain.py:283-296 (int32 bbox times a
                                float32 factor obtained by rounding the Python double ``1/W*pw``)
  * patch schedule of a whole forward (m1 / m2 / rN)   estimator/models/patchrefiner.py:361-392
"""
from __future__ import annotations

import dataclasses
import random as _random
from typing import List, Optional, Sequence, Tuple

import numpy as np


def prepare_tile_cfg(patch_process_shape: Sequence[int], image_raw_shape: Sequence[int], patch_split_num: Sequence[int]) -> dict:
    """baseline_pretrain.py:96-124 (the divisibility asserts are commented out there too)."""
    ph, pw = int(patch_process_shape[0]), int(patch_process_shape[1])
    sh, sw = int(patch_split_num[0]), int(patch_split_num[1])
    raw = (int(image_raw_shape[0]) // sh, int(image_raw_shape[1]) // sw)
    return {
        "patch_split_num": patch_split_num,
        "patch_reensemble_shape": (ph * sh, pw * sw),
        "patch_raw_shape": raw,
        "image_raw_shape": image_raw_shape,
        "raw_h_split_point": [int(raw[0] * i) for i in range(sh)],
        "raw_w_split_point": [int(raw[1] * i) for i in range(sw)],
    }


def resizer_size(patch_process_shape: Sequence[int]) -> Tuple[int, int]:
    """Output size of the DA ``Resize`` (external/depth_anything/transform.py:43-125) built with
    keep_aspect_ratio=False, ensure_multiple_of=14, resize_method='minimal': each axis of the
    requested process shape rounded to the nearest multiple of 14."""
    ph, pw = patch_process_shape
    return int(np.round(ph / 14) * 14), int(np.round(pw / 14) * 14)


def bbox_feat_factor(image_raw_shape: Sequence[int], patch_process_shape: Sequence[int]) -> np.ndarray:
    """float32([1/W*pw, 1/H*ph, 1/W*pw, 1/H*ph]) with the products formed in Python doubles first
    (baseline_pretrain.py:289-293); NOT pw/W."""
    H, W = int(image_raw_shape[0]), int(image_raw_shape[1])
    ph, pw = int(patch_process_shape[0]), int(patch_process_shape[1])
    return np.array([1 / W * pw, 1 / H * ph, 1 / W * pw, 1 / H * ph], dtype=np.float64).astype(np.float32)


def bboxs_to_feat(bboxs: np.ndarray, image_raw_shape, patch_process_shape) -> np.ndarray:
    """[P,4] int32 -> [P,5] float32 (column 0 = arange(P); baseline_pretrain.py:294-296)."""
    bboxs = np.asarray(bboxs, dtype=np.int32).reshape(-1, 4)
    bf = bboxs.astype(np.float32) * bbox_feat_factor(image_raw_shape, patch_process_shape)[None, :]
    inds = np.arange(bboxs.shape[0], dtype=np.float32)[:, None]
    return np.concatenate([inds, bf.astype(np.float32)], axis=1)


def check_roi_regime(rois: np.ndarray, levels) -> None:
    """The gather kernels implement torchvision's roi_align (aligned=True, sampling_ratio=-1) in its ONE-sample-per-bin regime:
    adaptive grid ceil(roi_size / pooled_size) == 1, evaluated in float32 exactly as torchvision does.  ``levels`` =
    [(pooled_h, pooled_w, spatial_scale)].  A ROI as large as the whole map (patch_split_num [1, 1]) can round to a grid of 2;
    that regime is not implemented, so it is refused here instead of silently differing from the reference."""
    r = np.asarray(rois, dtype=np.float32).reshape(-1, 4)
    for ph, pw, scale in levels:
        sc = np.float32(scale)
        rw = (r[:, 2] * sc - np.float32(0.5)) - (r[:, 0] * sc - np.float32(0.5))
        rh = (r[:, 3] * sc - np.float32(0.5)) - (r[:, 1] * sc - np.float32(0.5))
        gw, gh = np.ceil(rw / np.float32(pw)), np.ceil(rh / np.float32(ph))
        if r.shape[0] and (gw.max() > 1 or gh.max() > 1):
            raise NotImplementedError(f"roi_align sampling grid {int(gh.max())}x{int(gw.max())} at a {ph}x{pw} level: only the one-sample-per-bin "
                                      "regime (ROI no larger than the pooled map, i.e. patch_split_num >= [1, 1] without float32 overshoot) is implemented")


@dataclasses.dataclass
class Stage:
    """One ``regular_tile`` / ``random_tile`` call: its patches in reference order."""
    kind: str                       # 'regular' | 'random'
    bboxs: np.ndarray               # [P,4] int32 (x0,y0,x1,y1) raw-frame pixels
    off_process: Tuple[int, int]    # canvas offset (regular stages)
    grid: Tuple[int, int]           # (n_h, n_w) of the process-canvas grid (regular stages)
    init: bool = False


def regular_stage(tile_cfg: dict, patch_process_shape, offset, offset_process, init: bool) -> Stage:
    """baseline_pretrain.py:249-287."""
    rh, rw = tile_cfg["patch_raw_shape"]
    H, W = tile_cfg["image_raw_shape"]
    oh, ow = offset
    assert ow >= 0 and oh >= 0
    nh, nw = (H - oh) // rh, (W - ow) // rw
    hs = [rh * i + oh for i in range(nh)]
    ws = [rw * j + ow for j in range(nw)]
    ph, pw = patch_process_shape
    oph, opw = offset_process
    assert oph >= 0 and opw >= 0
    nhp = (tile_cfg["patch_reensemble_shape"][0] - oph) // ph
    nwp = (tile_cfg["patch_reensemble_shape"][1] - opw) // pw
    if (nhp, nwp) != (nh, nw):
        # the reference would mis-index predictions[patch_select_idx] here; refuse instead
        raise ValueError(f"raw grid {nh}x{nw} and process grid {nhp}x{nwp} disagree for offset {offset}")
    bb = np.array([[w0, h0, w0 + rw, h0 + rh] for h0 in hs for w0 in ws], dtype=np.int32).reshape(-1, 4)
    return Stage("regular", bb, (oph, opw), (nh, nw), init)


def random_stage(tile_cfg: dict, process_num: int, rng=_random) -> Stage:
    """baseline_pretrain.py:158-161: draws come from the *global* ``random`` module by default."""
    rh, rw = tile_cfg["patch_raw_shape"]
    H, W = tile_cfg["image_raw_shape"]
    hs = [rng.randint(0, H - rh - 1) for _ in range(process_num)]
    ws = [rng.randint(0, W - rw - 1)]
    bb = np.array([[w0, h0, w0 + rw, h0 + rh] for h0 in hs for w0 in ws], dtype=np.int32).reshape(-1, 4)
    return Stage("random", bb, (0, 0), (0, 0), False)


def parse_cai_mode(cai_mode: str) -> Tuple[bool, int]:
    """patchrefiner.py:371,385,389: returns (shifted grids?, number of random patches requested)."""
    if cai_mode == "m1":
        return False, 0
    if cai_mode == "m2":
        return True, 0
    if cai_mode and cai_mode[0] == "r":
        return True, int(cai_mode[1:])
    raise NotImplementedError(f"cai_mode {cai_mode!r}")


def schedule(tile_cfg: dict, patch_process_shape, cai_mode: str, process_num: int, rng=_random) -> List[Stage]:
    """Every patch of one forward, in the reference's order (patchrefiner.py:361-392).  The random
    draws happen here, up front; the stream consumption is identical to the reference's because
    nothing else on its path touches ``random`` between the draws."""
    ph, pw = patch_process_shape
    rh, rw = tile_cfg["patch_raw_shape"]
    shifted, n_random = parse_cai_mode(cai_mode)
    stages = [regular_stage(tile_cfg, patch_process_shape, (0, 0), (0, 0), True)]
    if shifted:
        stages.append(regular_stage(tile_cfg, patch_process_shape, (0, rw // 2), (0, pw // 2), False))
        stages.append(regular_stage(tile_cfg, patch_process_shape, (rh // 2, 0), (ph // 2, 0), False))
        stages.append(regular_stage(tile_cfg, patch_process_shape, (rh // 2, rw // 2), (ph // 2, pw // 2), False))
    for _ in range(n_random // process_num):
        stages.append(random_stage(tile_cfg, process_num, rng))
    return stages


def shard_patches(n_patches: int, rank: int, world_size: int, frames: int = 1) -> np.ndarray:
    """Ownership mask over the flattened patch list; every rank computes the same full schedule first.
    One frame: round-robin, patch i belongs to rank i % world_size.  A batch of frames (``frames`` > 1, frame-major list):
    contiguous blocks of the same sizes (+-1 patch), so that a rank touches the fewest frames -- exactly its own frame when
    there are as many frames as ranks -- and only those frames need to be uploaded to it."""
    own = np.zeros(n_patches, dtype=np.uint8)
    if frames > 1:
        own[(n_patches * rank) // world_size:(n_patches * (rank + 1)) // world_size] = 1
    else:
        own[rank::world_size] = 1
    return own

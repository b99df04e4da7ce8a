"""Frame egress: depth metrics and the timing reporter around the tiled-inference path.

``compute_errors`` / ``compute_metrics`` follow estimator/utils/metric.py:11-52,88-149 (same crops, clamps, masks, metric
names and formulas; ``see`` = soft edge error :67-86), ``benchmark`` follows ``Tester.benchmark``
(estimator/tester/tester.py:325-404: 20 warm-up + 30 timed forwards per run, ``repeat_times`` runs, ``benchmark.txt``).  This is
host-side evaluation code (NumPy on the depth map the model returned), like the reference's; the per-run synchronisation is the
reference's as well, so its frames/s are comparable with the reference's own ``benchmark.txt``.
"""
from __future__ import annotations

import os
import time
from typing import Dict, Iterable, Optional

import numpy as np
import torch
import torch.nn.functional as F


def compute_errors(gt: np.ndarray, pred: np.ndarray) -> Dict[str, float]:
    """metric.py:11-52: a1/a2/a3, abs_rel, rmse, log_10, rmse_log, silog, sq_rel over the given (already masked) pixels."""
    thresh = np.maximum(gt / pred, pred / gt)
    a1, a2, a3 = (thresh < 1.25).mean(), (thresh < 1.25 ** 2).mean(), (thresh < 1.25 ** 3).mean()
    abs_rel = np.mean(np.abs(gt - pred) / gt)
    sq_rel = np.mean(((gt - pred) ** 2) / gt)
    rmse = np.sqrt(((gt - pred) ** 2).mean())
    rmse_log = np.sqrt(((np.log(gt) - np.log(pred)) ** 2).mean())
    err = np.log(pred) - np.log(gt)
    silog = np.sqrt(np.mean(err ** 2) - np.mean(err) ** 2) * 100
    log_10 = np.abs(np.log10(gt) - np.log10(pred)).mean()
    return dict(a1=a1, a2=a2, a3=a3, abs_rel=abs_rel, rmse=rmse, log_10=log_10, rmse_log=rmse_log, silog=silog, sq_rel=sq_rel)


def _shift_2d_replace(data: np.ndarray, dx: int, dy: int, constant=0) -> np.ndarray:
    """metric.py:54-66."""
    out = np.roll(data, dx, axis=1)
    if dx < 0:
        out[:, dx:] = constant
    elif dx > 0:
        out[:, 0:dx] = constant
    out = np.roll(out, dy, axis=0)
    if dy < 0:
        out[dy:, :] = constant
    elif dy > 0:
        out[0:dy, :] = constant
    return out


def soft_edge_error(pred: np.ndarray, gt: np.ndarray, radius: int = 1) -> np.ndarray:
    """metric.py:67-72: per pixel, the smallest |gt shifted by (i, j) - pred| over the (2r+1)^2 neighbourhood."""
    diffs = [np.abs(_shift_2d_replace(gt, i, j, 0) - pred) for i in range(-radius, radius + 1) for j in range(-radius, radius + 1)]
    return np.minimum.reduce(diffs)


def get_boundaries(disp: np.ndarray, th: float = 1.0, dilation: int = 10) -> np.ndarray:
    """metric.py:74-86: depth-discontinuity mask, dilated with OpenCV."""
    import cv2
    ey = np.logical_or(np.pad(np.abs(disp[1:, :] - disp[:-1, :]) > th, ((1, 0), (0, 0))), np.pad(np.abs(disp[:-1, :] - disp[1:, :]) > th, ((0, 1), (0, 0))))
    ex = np.logical_or(np.pad(np.abs(disp[:, 1:] - disp[:, :-1]) > th, ((0, 0), (1, 0))), np.pad(np.abs(disp[:, :-1] - disp[:, 1:]) > th, ((0, 0), (0, 1))))
    edges = np.logical_or(ey, ex).astype(np.float32)
    if dilation > 0:
        edges = cv2.dilate(edges, np.ones((dilation, dilation), np.uint8), iterations=1)
    return edges


def compute_metrics(gt: torch.Tensor, pred: torch.Tensor, interpolate: bool = True, garg_crop: bool = False, eigen_crop: bool = True, dataset: str = "nyu",
                    min_depth_eval: float = 0.1, max_depth_eval: float = 10, disp_gt_edges: Optional[torch.Tensor] = None,
                    additional_mask: Optional[torch.Tensor] = None) -> Dict[str, float]:
    """metric.py:88-149.  ``pred`` may live on the GPU (the reference's Tester passes the CPU tensor the model returned)."""
    if gt.shape[-2:] != pred.shape[-2:] and interpolate:
        pred = F.interpolate(pred, gt.shape[-2:], mode="bilinear", align_corners=False).squeeze()
    pred = pred.squeeze().cpu().numpy().copy()
    pred[pred < min_depth_eval] = min_depth_eval
    pred[pred > max_depth_eval] = max_depth_eval
    pred[np.isinf(pred)] = max_depth_eval
    pred[np.isnan(pred)] = min_depth_eval
    gt_depth = gt.squeeze().cpu().numpy()
    valid = np.logical_and(gt_depth > min_depth_eval, gt_depth < max_depth_eval)
    eval_mask = np.ones(valid.shape)
    if garg_crop or eigen_crop:
        gh, gw = gt_depth.shape
        eval_mask = np.zeros(valid.shape)
        if garg_crop:
            eval_mask[int(0.40810811 * gh):int(0.99189189 * gh), int(0.03594771 * gw):int(0.96405229 * gw)] = 1
        elif dataset == "kitti":
            eval_mask[int(0.3324324 * gh):int(0.91351351 * gh), int(0.0359477 * gw):int(0.96405229 * gw)] = 1
        else:
            eval_mask[45:471, 41:601] = 1
    valid = np.logical_and(valid, eval_mask)
    if additional_mask is not None:
        valid = np.logical_and(valid, additional_mask.squeeze().detach().cpu().numpy())
    metrics = compute_errors(gt_depth[valid], pred[valid])
    if disp_gt_edges is not None:
        mask = np.logical_and(valid.squeeze(), disp_gt_edges.squeeze().numpy())
        see = torch.tensor([0])
        if mask.sum() > 0:
            see = soft_edge_error(pred, gt_depth)[mask].mean()
        metrics["see"] = see
    return metrics


@torch.no_grad()
def benchmark(model, batches: Iterable[dict], cai_mode: str = "r32", process_num: int = 4, image_raw_shape=(2160, 3840), patch_split_num=(4, 4),
              repeat_times: int = 10, log_interval: int = 10, num_warmup: int = 20, total_iters: int = 50, work_dir: Optional[str] = None,
              shard: bool = False, log=print) -> Dict[str, float]:
    """``Tester.benchmark`` (tester.py:325-404): per run, ``total_iters`` forwards with a device synchronisation around each one,
    the first ``num_warmup`` untimed; ``repeat_times`` runs; average and variance of the runs' frames/s.  ``batches`` yields dicts
    with ``image_lr`` / ``image_hr`` (it is cycled when shorter than ``total_iters``).  Writes ``benchmark.txt`` into ``work_dir``."""
    batches = list(batches)
    if not batches:
        raise ValueError("benchmark needs at least one frame")
    tile_cfg = {"image_raw_shape": list(image_raw_shape), "patch_split_num": list(patch_split_num)}
    out = dict(unit="img / s")
    fps_runs = []
    for run in range(repeat_times):
        log(f"Run {run + 1}:")
        pure = 0.0
        for i in range(total_iters):
            b = batches[i % len(batches)]
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            model(mode="infer", cai_mode=cai_mode, process_num=process_num, tile_cfg=tile_cfg, image_lr=b["image_lr"], image_hr=b["image_hr"], shard=shard)
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i >= num_warmup:
                pure += dt
                if (i + 1) % log_interval == 0:
                    log(f"Done image [{i + 1:<3}/ {total_iters}], fps: {(i + 1 - num_warmup) / pure:.3f} img / s")
        fps = (total_iters - num_warmup) / pure
        log(f"Overall fps: {fps:.3f} img / s\n")
        out[f"overall_fps_{run + 1}"] = round(fps, 2)
        fps_runs.append(fps)
    out["average_fps"] = round(float(np.mean(fps_runs)), 2)
    out["fps_variance"] = round(float(np.var(fps_runs)), 4)
    log(f"Average fps of {repeat_times} evaluations: {out['average_fps']}")
    log(f"The variance of {repeat_times} evaluations: {out['fps_variance']}")
    if work_dir:
        os.makedirs(work_dir, exist_ok=True)
        eng = getattr(model, "_engine", None)
        lines = []
        if eng is not None and "fine" in eng and "coarse" in eng:          # FLOPs per frame from the engine's own layer table (instead of mmengine's tracer)
            st = getattr(model, "last_stats", {})
            ph, pw = model.patch_process_shape
            fl = eng["coarse"].flops(1, ph, pw) + st.get("patches", 0) * eng["fine"].flops(1, ph, pw)
            lines.append(f"Model Flops (ViT + DPT branches, per frame, {st.get('patches', 0)} patches): {fl / 1e12:.2f} TFLOP")
        lines.append(f"\n\n Average fps of {repeat_times} evaluations: {out['average_fps']}")
        lines.append(f"\n\n The variance of {repeat_times} evaluations: {out['fps_variance']}")
        with open(os.path.join(work_dir, "benchmark.txt"), "w") as fh:
            fh.write("".join(lines))
    return out

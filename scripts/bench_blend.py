#!/usr/bin/env python
"""Micro-benchmark of the CAI blend kernels on the BASELINE frame (2160x3840, 448x448 patches, m2 canvas + r32 stage).
Every launch is preceded by an L2 flush (256 MB write) and bracketed by its own CUDA event pair; the host queues all
launches before synchronising, so event deltas are kernel durations (+ the launch gap), not host latency.
Prints algorithmic GB/s (DESIGN.md section 4 byte counts) for the fast paths and the generic kernels."""
import json
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from patchrefinerv2_b200 import _lib, masks, ops, tiling  # noqa: E402


def time_launches(fn, n=20, flush=None):
    evs = []
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for _ in range(n):
        if flush is not None:
            flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    return {"median_us": t[len(t) // 2], "min_us": t[0]}


def main():
    dev = torch.device("cuda")
    shape, raw, split, mode, pn = (448, 448), (2160, 3840), (4, 4), "r32", 4
    ph, pw = shape
    tc = tiling.prepare_tile_cfg(shape, raw, split)
    random.seed(1)
    stages = tiling.schedule(tc, shape, mode, pn)
    bb = np.concatenate([s.bboxs for s in stages])
    grid, first = [], 0
    for s in stages:
        if s.kind == "regular":
            grid.append((s.off_process[0], s.off_process[1], s.grid[0], s.grid[1], first))
            first += s.bboxs.shape[0]
    preds = torch.rand(bb.shape[0], ph, pw, device=dev) * 10
    mask = torch.from_numpy(masks.generatemask(shape, 0.15).copy()).to(dev)
    rh, rw = tc["patch_raw_shape"]
    H, W = tc["image_raw_shape"]
    Hc, Wc = tc["patch_reensemble_shape"]
    rmask = torch.from_numpy(masks.random_patch_mask((rh, rw), 0.15).copy()).to(dev)
    starts = torch.from_numpy(np.ascontiguousarray(bb[first:, [1, 0]])).to(dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    n_rand = bb.shape[0] - first
    bytes_canvas = 4.0 * (first * ph * pw + ph * pw + 2 * Hc * Wc)
    bytes_raw = 4.0 * (2 * Hc * Wc + n_rand * ph * pw + rh * rw + 2 * H * W)
    avg_c, cnt_c = ops.blend_canvas(preds[:first], mask, grid, Hc, Wc)
    out = {"bytes_canvas": bytes_canvas, "bytes_raw": bytes_raw}
    for label, generic in (("generic", 1), ("fast", 0)):
        _lib.call("prv2_debug_blend_generic", generic)
        for fl_label, fl in (("flushed", flush), ("warm", None)):
            c = time_launches(lambda: ops.blend_canvas(preds[:first], mask, grid, Hc, Wc), flush=fl)
            rprep = ops.blend_raw_prepare(rmask, pw)
            r0 = time_launches(lambda: ops.blend_raw(avg_c, cnt_c, preds[first:], starts, rmask, ph, pw, rh, rw, H, W), flush=fl)
            print(f"   (raw stage without the prepared weight map: {r0})")
            r = time_launches(lambda: ops.blend_raw(avg_c, cnt_c, preds[first:], starts, rmask, ph, pw, rh, rw, H, W, prep=rprep), flush=fl)
            c["GBps"] = bytes_canvas / c["median_us"] / 1e3
            r["GBps"] = bytes_raw / r["median_us"] / 1e3
            out[f"{label}_{fl_label}"] = {"canvas": c, "raw": r}
            print(label, fl_label, "canvas", c, "raw", r, flush=True)
    # rN-stage variants (prv2_debug_blend_generic knobs: bit 1 = segment kernel off, bits 4-7 rows per CTA, bits 8-15 warps per CTA)
    rprep = ops.blend_raw_prepare(rmask, pw)
    ref = None
    for label, knob in (("tab", 2), ("seg", 0), ("seg_w8", 8 << 8), ("seg_w5", 5 << 8), ("seg_w4", 4 << 8), ("seg_w3", 3 << 8), ("seg_w2", 2 << 8)):
        _lib.call("prv2_debug_blend_generic", knob)
        got = ops.blend_raw(avg_c, cnt_c, preds[first:], starts, rmask, ph, pw, rh, rw, H, W, prep=rprep)
        if ref is None:
            ref = got
        same = bool(torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]))
        r = time_launches(lambda: ops.blend_raw(avg_c, cnt_c, preds[first:], starts, rmask, ph, pw, rh, rw, H, W, prep=rprep), flush=flush)
        r["GBps"] = bytes_raw / r["median_us"] / 1e3
        r["bit_identical_to_tab"] = same
        out["raw_" + label] = r
        print("raw", label, r, flush=True)
    _lib.call("prv2_debug_blend_generic", 0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

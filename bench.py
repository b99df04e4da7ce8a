#!/usr/bin/env python
"""Benchmark of the tiled high-resolution inference hot path (BASELINE.json metric):
synthetic 2160x3840 frames, CAI r32, PatchRefiner DAv2 ViT-L (coarse + refiner) + FusionUnet,
frames/s (+ patches/s) on N B200s, with the roofline fraction of the dominant kernel and the
reference's CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched by torchrun (one rank per GPU); patches of each frame are sharded over the ranks
and the packed partial canvases are combined by ONE NCCL sum-reduce ("strong" scaling of a frame).
The measured arm builds its config, seeded random-init weights and synthetic frame itself; `oracle/` is imported only by
the CPU / eager-GPU baseline legs and `--impl reference`.  A step = one frame through the public model API.  `value` has the frame resident in HBM and the
result left on the device; `e2e` copies the frame from pinned host memory and the depth map back to
the host inside the timed region.  Rank 0 prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (encoder, patch_process_shape, image_raw_shape, patch_split_num, cai_mode, process_num)
    "dav2_vitl_2160x3840_4x4_r32": ("vitl", (448, 448), (2160, 3840), (4, 4), "r32", 4),       # BASELINE configs[3]: the configuration the metric is quoted on
    "dav2_vitl_2160x3840_4x4_m2": ("vitl", (448, 448), (2160, 3840), (4, 4), "m2", 4),         # BASELINE configs[2]: 49 patches, canvas output
    "dav2_vitl_4320x7680_8x8_r128": ("vitl", (448, 448), (4320, 7680), (8, 8), "r128", 4),     # BASELINE configs[4] geometry: 353 patches / frame
    "dav2_vits_2160x3840_4x4_r32": ("vits", (448, 448), (2160, 3840), (4, 4), "r32", 4),
    "dav2_vits_1080x1920_2x2_m1": ("vits", (448, 448), (1080, 1920), (2, 2), "m1", 4),
    "dav2_vits_432x768_2x2_r4": ("vits", (224, 224), (432, 768), (2, 2), "r4", 2),
}
METRIC = "frames_per_sec_2160x3840_cai_r32"


def make_config(encoder, patch_process_shape, image_raw_shape, patch_split_num, max_depth=80.0):
    """Model config shaped like the reference's configs/patchrefiner_dav2/pr_u4k.py:10-53 (DAv2 coarse + DAv2 refiner + FusionUnet)."""
    feats, oc = {"vitl": (256, [256, 512, 1024, 1024]), "vitb": (128, [96, 192, 384, 768]), "vits": (64, [48, 96, 192, 384])}[encoder]
    half = feats // 2
    branch = lambda: dict(type="DA2", pretrained=None, model_cfg=dict(encoder=encoder, features=feats, out_channels=list(oc)))
    return dict(image_raw_shape=list(image_raw_shape), patch_process_shape=list(patch_process_shape), patch_split_num=list(patch_split_num),
                fusion_feat_level=6, min_depth=1e-3, max_depth=max_depth, strategy_refiner_target="offset_coarse",
                coarse_branch=branch(),
                refiner=dict(fine_branch=branch(),
                             fusion_model=dict(type="FusionUnet", input_chl=[half * 2] + [feats * 2] * 5, temp_chl=[half] + [feats] * 5,
                                               dec_chl=[feats] * 4 + [half])),
                sigloss=dict(type="SILogLoss"), pretrained=None, pre_norm_bbox=True, pretrain_coarse_model=None, pretrain_fine_model=None)


def random_state_dict(spec, seed=0):
    """Seeded random-init weights for every tensor of the model's own state dict (name -> shape): fan-in-scaled normals for
    matrices / conv kernels, near-one LayerNorm scales, small biases, LayerScale around 0.2 -- numerically healthy, not trained.
    The measured arm generates its inputs itself; `oracle/` is only used by the CPU / eager baseline legs."""
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in spec.items():
        shp = tuple(shp)
        u = lambda scale: (torch.rand(shp, generator=g) * 2 - 1) * scale
        if k.endswith("mask_token"):
            v = torch.zeros(shp)
        elif k.endswith(("cls_token", "pos_embed")):
            v = torch.randn(shp, generator=g) * 0.02
        elif k.endswith(".gamma"):
            v = 0.2 + u(0.1)
        elif len(shp) == 1 and k.endswith(".weight"):
            v = 1.0 + u(0.1)                                   # LayerNorm scales
        elif k.endswith(".bias"):
            v = u(0.05)
        else:
            fan_in = max(1, int(torch.tensor(shp[1:]).prod())) if len(shp) > 1 else shp[0]
            v = torch.randn(shp, generator=g) * (1.3 / math.sqrt(fan_in))
        sd[k] = v
    return sd


def synthetic_frame(image_raw_shape, seed=1):
    """fp32 RGB frame in [0,1], [1,3,H,W]: smooth low-frequency structure + noise (host tensor)."""
    import torch
    import torch.nn.functional as F
    H, W = image_raw_shape
    g = torch.Generator().manual_seed(seed)
    smooth = F.interpolate(torch.rand(1, 3, 9, 16, generator=g), (H, W), mode="bicubic", align_corners=True).clamp(0, 1)
    return (0.7 * smooth + 0.3 * torch.rand(1, 3, H, W, generator=g)).contiguous()


def profiled_traffic():
    """Average DRAM bytes per launch of the dominant kernel from the committed ncu launch list of this command
    (profiles/traffic.json, written by scripts/summarize_launches.py); None when no capture is committed."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            return None
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]), tf_sustained=float(p["bf16_tflops_sustained"]), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8 and f[0] == str(self.gpu_index):
                self.rows.append(f)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(rows), "reasons": reasons}


import contextlib


@contextlib.contextmanager
def stdout_to_stderr():
    """Whatever the reference (or NCCL) prints must not land on stdout: it carries exactly one JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def _oracle_parity_patches(cfg, sd, hr, k=3):
    """Oracle (CPU port, bit-identical to the reference by tests/test_oracle_vs_reference.py) predictions for `k` single
    patches of this workload: the checker for the `parity` block.  Returns (bboxs [k,4] int32, preds [k,ph,pw], coarse_roi
    [k,ph,pw], seconds)."""
    import torch
    from oracle import pr_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    orc = O.PatchRefinerOracle(cfg, sd)
    lr = O.resizer(tuple(cfg["patch_process_shape"]), hr)
    tc = orc.tile_cfg
    rh, rw = tc["patch_raw_shape"]
    H, W = tc["image_raw_shape"]
    t0 = time.perf_counter()
    with torch.no_grad():
        feats, coarse = orc.coarse_forward(lr)
        tt = {"coarse_prediction": coarse, "coarse_features": feats}
        # a grid-aligned patch, a random-stage style patch with odd offsets, and one touching the frame's bottom-right corner
        starts = [(0, rw), (137 % (H - rh), 211 % (W - rw)), (H - rh - 1, W - rw - 1)][:k]
        bbs, preds, rois = [], [], []
        for (y0, x0) in starts:
            bb = O.make_bboxs([y0], [x0], rh, rw)
            preds.append(orc._predict(hr[0], bb, tc, tt, 1, None)[0, 0])
            bf = O.bboxs_to_feat(bb, tc["image_raw_shape"], orc.patch_process_shape)
            d_roi, _ = O.coarse_postprocess_test(coarse, [], bf, orc.patch_process_shape[0])
            rois.append(d_roi[0, 0])
            bbs.append(bb)
    return torch.cat(bbs).int(), torch.stack(preds), torch.stack(rois), time.perf_counter() - t0


def parity_block(models, cfg, sd, hr, lr_dev, hr_dev, log=lambda *a: None):
    """Depth of the SAME patches from the GPU model (every precision mode in `models`) against the oracle, on the benchmarked
    configuration and weights.  rel = |got - want| / max(|want|, 1e-3) per pixel; offset = depth - coarse_roi (what the refiner
    produces), its error relative to max|offset_ref|.  rel = |got - want| / want over pixels with want >= 0.1 m; the rest (incl. pixels the
    reference clamps to exactly 0) are reported as absolute errors."""
    bbs, want, roi, secs = _oracle_parity_patches(cfg, sd, hr)
    off_ref = want - roi
    out = {"checker": "oracle port on the host (bit-identical to the reference by tests/test_oracle_vs_reference.py)",
           "patches": int(bbs.shape[0]), "bboxs": bbs.tolist(), "oracle_seconds": secs,
           "offset_ref_abs_max": float(off_ref.abs().max()), "depth_ref_range": [float(want.min()), float(want.max())]}
    floor = 0.1                                            # metres; below it (incl. pixels the reference clamps to exactly 0) errors are reported as absolute
    big = want >= floor
    out["rel_floor_m"] = floor
    out["pixels"] = int(want.numel())
    out["pixels_below_floor"] = int((~big).sum())
    for name, m in models.items():
        got, _ = m.predict_patches(lr_dev, hr_dev, bbs)
        got = got.float().cpu()
        err = (got - want).abs()
        rel = (err / want.clamp_min(floor))[big]
        off = ((got - roi) - off_ref).abs() / off_ref.abs().max().clamp_min(1e-6)
        out[f"{name}_max_rel"], out[f"{name}_mean_rel"] = float(rel.max()), float(rel.mean())
        out[f"{name}_p9999_rel"] = float(rel.flatten().kthvalue(max(1, int(rel.numel() * 0.9999))).values)
        out[f"{name}_pixels_over_1e-3"] = int((rel > 1e-3).sum())
        out[f"{name}_max_abs_m"] = float(err.max())
        out[f"{name}_max_abs_below_floor_m"] = float(err[~big].max()) if (~big).any() else 0.0
        out[f"{name}_offset_max_rel"], out[f"{name}_offset_mean_rel"] = float(off.max()), float(off.mean())
    out["tolerance"] = {"fp32": "p99.99 <= 1e-3 (max a few 1e-3: the tensor core's round-toward-zero fp32 accumulation, DESIGN.md)", "bf16": 5e-2}
    return out


def _build_reference(cfg, sd, device="cpu"):
    """The reference's own PatchRefiner (estimator/models/patchrefiner.py:54) from oracle/_ref (or /root/reference), through the
    import shim, with this bench's weights.  Raises when no reference tree is available."""
    import tempfile
    import torch
    from oracle import ref_shim
    if not ref_shim.reference_available():
        raise FileNotFoundError("no reference tree (oracle/_ref is built by __graft_entry__.build() where /root/reference exists)")
    cwd = os.getcwd()
    d = tempfile.mkdtemp(prefix="prv2_ref_")
    cp, fp = os.path.join(d, "c.pth"), os.path.join(d, "f.pth")
    torch.save({k[len("coarse_branch."):]: v for k, v in sd.items() if k.startswith("coarse_branch.")}, cp)
    torch.save({k[len("refiner_fine_branch."):]: v for k, v in sd.items() if k.startswith("refiner_fine_branch.")}, fp)
    try:
        ref = ref_shim.build_reference_patchrefiner(cfg, cp, fp)
    finally:
        os.chdir(cwd)
        for f in (cp, fp):
            os.unlink(f)
        os.rmdir(d)
    res = ref.load_state_dict(sd, strict=False)
    assert not res.missing_keys and not res.unexpected_keys, (res.missing_keys[:3], res.unexpected_keys[:3])
    return ref.to(device).eval()


def reference_whole_frames(cfg, sd, hr, cai_mode, process_num, frames, warmup, device="cpu", log=lambda *a: None, exact_fp32=False):
    """The UNMODIFIED reference's own ``PatchRefiner.forward(mode='infer')`` on whole frames of this workload (host cores or,
    with device='cuda', eager PyTorch on this GPU with the reference's own per-patch host RunningAverageMap).
    Returns (frames/s, seconds per frame list, cores)."""
    import torch
    from oracle import pr_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    tf32_was = torch.backends.cudnn.allow_tf32
    if exact_fp32:
        torch.backends.cudnn.allow_tf32 = False        # checker runs: strict fp32 convolutions (PyTorch's default lets cuDNN use TF32)
    ref = _build_reference(cfg, sd, device)
    hr_d = hr.to(device)
    lr_d = O.resizer(tuple(cfg["patch_process_shape"]), hr).to(device)
    times = []
    with torch.no_grad():
        for i in range(warmup + frames):
            random.seed(1)
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            depth, _ = ref(mode="infer", image_lr=lr_d, image_hr=hr_d, cai_mode=cai_mode, process_num=process_num, tile_cfg=None)
            if device != "cpu":
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            log(f"reference whole frame {i} on {device}: {dt:.2f}s")
            if i >= warmup:
                times.append(dt)
    del ref
    torch.backends.cudnn.allow_tf32 = tf32_was
    if device != "cpu":
        torch.cuda.empty_cache()
    return 1.0 / statistics.mean(times), times, torch.get_num_threads(), depth


def reference_bounded_sample(cfg, sd, hr, cai_mode, process_num, n_regular, n_random, tiles=3, log=lambda *a: None):
    """cpu_baseline of the main arm: the reference's OWN code on a bounded sample of the workload (~20-30 s of host work) --
    its ``coarse_forward`` once, then `tiles` calls of its ``random_tile`` (each: process_num crops, roi_align, fine branch,
    FusionUnet, nearest upsample and process_num RunningAverageMap updates at raw resolution; baseline_pretrain.py:149-231),
    extrapolated to the frame's patch count; a regular-stage patch is charged the same network time with the (cheaper)
    canvas-resolution update measured separately."""
    import torch
    from oracle import pr_oracle as O
    from oracle import ref_shim
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    ref = _build_reference(cfg, sd, "cpu")
    from estimator.models.utils import RunningAverageMap, generatemask
    lr = O.resizer(tuple(cfg["patch_process_shape"]), hr)
    tc = ref.tile_cfg
    rh, rw = tc["patch_raw_shape"]
    H, W = tc["image_raw_shape"]
    Hc, Wc = tc["patch_reensemble_shape"]
    with torch.no_grad():
        t0 = time.perf_counter()
        feats, coarse = ref.coarse_forward(lr)
        t_coarse = time.perf_counter() - t0
        tt = {"coarse_prediction": coarse, "coarse_features": feats}
        blur = torch.tensor(generatemask((rh, rw), border=0.15) + 1e-3)
        avg = RunningAverageMap(torch.rand(H, W), torch.rand(H, W))
        random.seed(1)
        times = []
        for i in range(tiles):
            t0 = time.perf_counter()
            avg = ref.random_tile(image_hr=hr[0], tile_temp=tt, blur_mask=blur, avg_depth_map=avg, tile_cfg=tc, process_num=process_num)
            times.append((time.perf_counter() - t0) / process_num)
            log(f"reference random_tile {i}: {times[-1]:.2f}s/patch")
        t_rand = statistics.mean(times)
        # RunningAverageMap.update at raw vs canvas resolution (estimator/models/utils.py:31-36)
        def upd(h, w, ph_, pw_):
            a = RunningAverageMap(torch.rand(h, w), torch.rand(h, w))
            p, c = torch.zeros(h, w), torch.zeros(h, w)
            c[10:10 + ph_, 20:20 + pw_] = 1.0
            t0 = time.perf_counter()
            for _ in range(3):
                a.update(p, c)
            return (time.perf_counter() - t0) / 3
        t_up_raw, t_up_canvas = upd(H, W, rh, rw), upd(Hc, Wc, *cfg["patch_process_shape"])
    t_reg = max(t_rand - t_up_raw + t_up_canvas, 0.0)
    frame_s = t_coarse + n_regular * t_reg + n_random * t_rand
    desc = (f"reference's own code (oracle/_ref via the import shim, torch {torch.__version__}, {cores} threads): coarse_forward {t_coarse:.2f}s + "
            f"{tiles} random_tile calls of {process_num} patches ({t_rand:.2f}s/patch incl. the raw-resolution RunningAverageMap update "
            f"{t_up_raw * 1e3:.0f}ms; canvas-resolution update {t_up_canvas * 1e3:.0f}ms), extrapolated to {n_regular} regular + {n_random} random patches")
    return 1.0 / frame_s, cores, desc


def blend_launch_times(pshape, raw, split, cai_mode, process_num, dev, n=20):
    """CUDA-event durations of the two CAI blend kernels on this workload's geometry (random predictions: the kernels'
    work does not depend on the values).  Inside the frame loop the two ~20-60 us launches sit behind host-side work, so an
    event pair there also measures host latency; here every launch is queued behind a 256 MB L2-flush write and bracketed
    by its own event pair, so the delta is the kernel (+ launch gap) with a cold L2."""
    import numpy as np
    import torch
    from patchrefinerv2_b200 import masks, ops, tiling
    ph, pw = pshape
    tc = tiling.prepare_tile_cfg(pshape, raw, split)
    stages = tiling.schedule(tc, pshape, cai_mode, process_num, random.Random(1))
    bb = np.concatenate([s.bboxs for s in stages])
    grid, first = [], 0
    for s in stages:
        if s.kind == "regular":
            grid.append((s.off_process[0], s.off_process[1], s.grid[0], s.grid[1], first))
            first += s.bboxs.shape[0]
    preds = torch.rand(bb.shape[0], ph, pw, device=dev) * 10
    mask = torch.from_numpy(masks.generatemask(pshape, 0.15).copy()).to(dev)
    rh, rw = tc["patch_raw_shape"]
    H, W = tc["image_raw_shape"]
    Hc, Wc = tc["patch_reensemble_shape"]
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

    def run(fn):
        evs = []
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        for _ in range(n):
            flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        torch.cuda.synchronize(dev)
        return statistics.median(a.elapsed_time(b) for a, b in evs) * 1e-3

    out = {"canvas": {"bytes": 4.0 * (first * ph * pw + ph * pw + 2 * Hc * Wc),
                      "s": run(lambda: ops.blend_canvas(preds[:first], mask, grid, Hc, Wc))}}
    if cai_mode[0] == "r" and bb.shape[0] > first:
        avg_c, cnt_c = ops.blend_canvas(preds[:first], mask, grid, Hc, Wc)
        rmask = torch.from_numpy(masks.random_patch_mask((rh, rw), 0.15).copy()).to(dev)
        starts = torch.from_numpy(np.ascontiguousarray(bb[first:, [1, 0]])).to(dev)
        rprep = ops.blend_raw_prepare(rmask, pw)                # once per geometry, as the model does
        out["raw"] = {"bytes": 4.0 * (2 * Hc * Wc + (bb.shape[0] - first) * ph * pw + rh * rw + 2 * H * W),
                      "s": run(lambda: ops.blend_raw(avg_c, cnt_c, preds[first:], starts, rmask, ph, pw, rh, rw, H, W, prep=rprep))}
    # what an event pair costs by itself in this arrangement: the same pair around a one-element fill (a ~2 us kernel)
    one = flush[:1]
    out["_event_pair_floor_s"] = run(lambda: one.fill_(0.0))
    return out


def kernel_tables(log, steps, ms_total, peaks):
    """Per-kernel / per-layer-group aggregates of a profiled pass (CUDA-event pairs recorded on the launching stream)."""
    agg, layers = {}, {}
    for name, unit, amount, a, b, tag in log:
        dt = a.elapsed_time(b)
        g = agg.setdefault(name, {"unit": unit, "work": 0.0, "ms": 0.0, "launches": 0})
        g["work"] += amount
        g["ms"] += dt
        g["launches"] += 1
        if name == "prv2_umma_gemm":
            grp = tag.split(".")[0] if tag.startswith("vit") or tag.startswith("dpt") else ".".join(tag.split(".")[:2])
            for key in (grp, tag):
                l = layers.setdefault(key, {"flop": 0.0, "ms": 0.0, "launches": 0})
                l["flop"] += amount
                l["ms"] += dt
                l["launches"] += 1
    kern = {}
    for name, g in agg.items():
        rate = g["work"] / (g["ms"] * 1e-3) if g["ms"] > 0 else 0.0
        if g["unit"] == "flop":
            kern[name] = {"bound": "tensor", "achieved": rate / 1e12, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": rate / 1e12 / peaks["tf_sustained"]}
        else:
            kern[name] = {"bound": "hbm", "achieved": rate / 1e9, "peak": peaks["hbm"], "unit": "GB/s", "frac": rate / 1e9 / peaks["hbm"]}
        kern[name].update(launches_per_step=g["launches"] / steps, ms_per_step=g["ms"] / steps,
                          share_of_step=g["ms"] / ms_total, avg_launch_us=1e3 * g["ms"] / g["launches"])
    gemm_layers = {k: {"tflops": v["flop"] / (v["ms"] * 1e-3) / 1e12, "frac": v["flop"] / (v["ms"] * 1e-3) / 1e12 / peaks["tf_sustained"],
                       "ms_per_step": v["ms"] / steps, "launches_per_step": v["launches"] / steps} for k, v in sorted(layers.items())}
    return kern, gemm_layers


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dav2_vitl_2160x3840_4x4_r32", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--patch-batch", type=int, default=0, help="patches per network launch (0 = balance automatically, <= 27)")
    ap.add_argument("--frames-per-step", type=int, default=0, help="frames in one model call (0 = one per GPU: the batch-of-frames work list, weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fp32-mode", action="store_true", help="skip the second (fp32-class) model timed beside the bf16 headline")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--reference-frames", type=int, default=1, help="--impl reference: whole frames of the reference's own forward to time")
    ap.add_argument("--profile-run", action="store_true", help="run under ncu: fewer warm-up steps allowed; the printed number is NOT a bench value")
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference" or args.profile_run, "timing rules: at least 3 warm-up steps"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    enc, pshape, raw, split, cai_mode, process_num = WORKLOADS[args.workload]
    global METRIC
    if args.workload != "dav2_vitl_2160x3840_4x4_r32":                  # the headline metric name belongs to the headline workload only
        METRIC = f"frames_per_sec_{raw[0]}x{raw[1]}_cai_{cai_mode}"
    log = lambda *a: print(*a, file=sys.stderr, flush=True)

    import torch
    from patchrefinerv2_b200 import build_model, tiling
    cfg = make_config(enc, pshape, raw, split)
    tc = tiling.prepare_tile_cfg(pshape, raw, split)
    sched = tiling.schedule(tc, pshape, cai_mode, process_num, random.Random(0))
    n_patches = sum(s.bboxs.shape[0] for s in sched)
    n_regular = sum(s.bboxs.shape[0] for s in sched if s.kind == "regular")
    config = {"workload": args.workload, "image_raw_shape": list(raw), "patch_process_shape": list(pshape), "patch_split_num": list(split),
              "cai_mode": cai_mode, "process_num": process_num, "patches_per_frame": n_patches, "precision": args.precision,
              "parallelism": f"patch-shard x{world} + 1 NCCL sum-reduce" if world > 1 else "single GPU",
              "weights": "random-init (seeded), reference state-dict layout",
              "l2": "no explicit flush: every step streams several GB of activations (>> 126 MB L2) between reuses of any input"}

    if args.impl == "reference":
        # The reference's own CPU implementation of the path: the UNMODIFIED PatchRefiner.forward (oracle/_ref through the import
        # shim) on WHOLE frames of this workload, all host threads.  A ViT-L r32 frame is ~2.5 min of host work, so the run is
        # `--reference-frames` frames (default 1, no warm-up frame) whatever --steps says; `steps` in the line is what was run.
        if rank != 0:
            return
        spec = {k: tuple(v.shape) for k, v in build_model(dict(type="PatchRefiner", config=cfg)).state_dict().items()}
        sd = random_state_dict(spec, 0)
        config["precision"] = "fp32 (the reference's own arithmetic)"
        config["parallelism"] = "host cores"
        try:
            with stdout_to_stderr():
                fps, times, cores, _ = reference_whole_frames(cfg, sd, synthetic_frame(raw, 1), cai_mode, process_num, max(1, args.reference_frames), 0, "cpu", log)
            kind, frames = "reference", len(times)
            desc = (f"the reference's own PatchRefiner.forward(mode='infer', cai_mode={cai_mode!r}) from oracle/_ref (import shim, torch {torch.__version__}, "
                    f"{cores} threads) on {frames} WHOLE frame(s) of this workload, no extrapolation: {', '.join(f'{t:.1f}s' for t in times)}")
        except FileNotFoundError as e:
            log(f"reference tree unavailable ({e}); timing the oracle port instead")
            from oracle import pr_oracle as O
            torch.set_num_threads(os.cpu_count() or 1)
            cores = torch.get_num_threads()
            hr = synthetic_frame(raw, 1)
            t0 = time.perf_counter()
            random.seed(1)
            O.PatchRefinerOracle(cfg, sd).infer(O.resizer(pshape, hr), hr, None, cai_mode, process_num)
            fps, frames, kind = 1.0 / (time.perf_counter() - t0), 1, "port"
            desc = f"oracle port (oracle/pr_oracle.py) on 1 WHOLE frame, {cores} threads"
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": frames,
                "steps_requested": args.steps, "warmup": 0, "warmup_requested": args.warmup, "ms_per_step": 1000.0 / fps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "patches_per_sec": fps * n_patches,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": desc},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    from patchrefinerv2_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout at the first collective; keep stdout for the one JSON line
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
    # A step = ONE model call on a batch of F frames (default F = number of GPUs): the F x 81 patches form one work list that is
    # split into contiguous blocks over the ranks (round-robin for a single frame), every frame keeps its own canvases, ONE sum-reduce combines the batch.  Per-GPU work
    # is one frame's worth at every N (weak scaling); at N = 1 this is the single-frame call of BASELINE config 4.
    F_ = args.frames_per_step or world
    n_local = -(-(n_patches * F_) // world)
    pb = args.patch_batch or -(-n_local // (-(-n_local // 27)))
    config["patch_batch"] = pb
    config["frames_per_step"] = F_
    config["parallelism"] = (f"batch of {F_} frames = {F_ * n_patches} patches split into {world} contiguous blocks (one per rank; host frames are uploaded only to the ranks that cut patches from them) + 1 NCCL sum-reduce per batch"
                             if world > 1 else "single GPU")
    import torch as _t
    hr = _t.cat([synthetic_frame(raw, 1 + f) for f in range(F_)])
    hr_dev = hr.to(dev)
    shard = world > 1
    out_shape = (F_, 1) + ((raw[0], raw[1]) if cai_mode[0] == "r" else tuple(tc["patch_reensemble_shape"]))
    host_out = torch.empty(out_shape, dtype=torch.float32).pin_memory()
    peaks = measured_peaks()

    def make_model(precision):
        m = build_model(dict(type="PatchRefiner", config=cfg, precision=precision, patch_batch=pb, output_device="cuda"))
        sd_ = random_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 0)
        res = m.load_dict(sd_)
        assert not res.missing_keys and not res.unexpected_keys
        return m.cuda().eval(), sd_

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        barrier()
        if profile:
            _lib.profile_log = []
        l0 = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        plog, _lib.profile_log = _lib.profile_log, None
        t = torch.tensor([ms], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item()), _lib.launch_count - l0, plog

    def measure(model, lr_dev, lr_pin, hr_pin, steps, warmup, want_e2e):
        """(value pass WITHOUT per-launch instrumentation, e2e pass through host buffers, separate profiled pass)."""
        def step_resident():
            random.seed(1)
            d, _ = model(mode="infer", image_lr=lr_dev, image_hr=hr_dev, cai_mode=cai_mode, process_num=process_num, shard=shard)
            return d

        def step_e2e():
            # host buffers in, host buffer out: the model uploads image_lr, then image_hr on its copy stream under the coarse pass;
            # rank 0 alone reads the blended frame back (every rank holds it after the reduce)
            random.seed(1)
            d, _ = model(mode="infer", image_lr=lr_pin, image_hr=hr_pin, cai_mode=cai_mode, process_num=process_num, shard=shard)
            if rank == 0:
                host_out.copy_(d, non_blocking=True)
            torch.cuda.current_stream().synchronize()          # the caller holds the depth map on the host when the step ends
            return host_out

        ms, launches, _ = timed(step_resident, steps, warmup)
        res = {"ms": ms, "steps": steps, "launches": launches, "fps": F_ * steps / (ms / 1e3), "step_resident": step_resident}
        if want_e2e:
            ms_e, _, _ = timed(step_e2e, steps, 2)
            res["e2e"] = {"value": F_ * steps / (ms_e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": lr_pin.numel() * 4 + hr_pin.numel() * 4,
                          "d2h_bytes_per_step": host_out.numel() * 4, "ms_per_step": ms_e / steps,
                          "note": "pinned host frames in, pinned host depth out; each rank uploads only the frames its patches are cut from (bytes = the whole job's), under the coarse pass on a copy stream; D2H on rank 0"}
        psteps = min(steps, 3)
        ms_p, _, plog = timed(step_resident, psteps, 1, profile=True)
        res["kernels"], res["gemm_layers"] = kernel_tables(plog, psteps, ms_p, peaks)
        res["profiled_pass"] = {"steps": psteps, "ms_per_step": ms_p / psteps,
                                "note": "separate pass with an event pair around every launch (slower than `value`, which carries no instrumentation)"}
        return res

    def roofline_of(res, traffic):
        gemm = res["kernels"].get("prv2_umma_gemm", {})
        return {"kernel": "umma_gemm_kernel (prv2_umma_gemm)", "bound": "tensor", "achieved": gemm.get("achieved"), "peak": peaks["tf_sustained"],
                "unit": "TFLOP/s", "frac": gemm.get("frac"), "traffic": traffic.get("umma_gemm_kernel", {}).get("dram_bytes_per_launch"),
                "traffic_source": traffic.get("source"), "peak_source": f"{peaks['source']} bf16 sustained",
                "share_of_step": gemm.get("share_of_step"), "launches_per_step": gemm.get("launches_per_step"),
                "note": "aggregate over all launches of the kernel in the profiled pass: sum(algorithmic FLOPs) / sum(CUDA-event durations)"}

    model, sd = make_model(args.precision)
    lr_dev = model.resizer(hr_dev)                         # image_lr as the dataset builds it (general_dataset.py:218), on the device
    lr = lr_dev.cpu()
    lr_pin, hr_pin = lr.pin_memory(), hr.pin_memory()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    main_res = measure(model, lr_dev, lr_pin, hr_pin, args.steps, args.warmup, not args.no_e2e)
    clocks = sampler.stop() if sampler else None
    ms, fps = main_res["ms"], main_res["fps"]

    # multi-GPU parity evidence (untimed): the sharded frame against the same frame refined by this rank alone
    shard_check = None
    if world > 1:
        d_sh = main_res["step_resident"]().clone()
        random.seed(1)
        d_one, _ = model(mode="infer", image_lr=lr_dev, image_hr=hr_dev, cai_mode=cai_mode, process_num=process_num, shard=False)
        shard_check = {"max_rel_diff_vs_unsharded": float(((d_sh - d_one).abs() / d_one.abs().clamp_min(1e-3)).max().item()),
                       "what": f"all {F_} frames of the sharded batch (sharded + all-gathered coarse passes, one block of patches per rank, one sum-reduce) against "
                               "the same batch refined by this rank alone"}
        torch.distributed.barrier()

    traffic = profiled_traffic() or {}
    roofline = roofline_of(main_res, traffic) if rank == 0 else None
    flops_frame = workspace_gb = None
    if rank == 0:
        eng = model._engine
        flops_frame = eng["coarse"].flops(1, *pshape) + n_patches * (eng["fine"].flops(1, *pshape) +
                      eng["fusion"].flops(1, [(f.H, f.W) for f in eng["coarse"].forward(lr_dev[:1])[1]][::-1]))
        workspace_gb = sum(w.nbytes() for e in (eng["coarse"], eng["fine"], eng["fusion"]) for w in e.ws.values()) / 1e9

    # the CAI blend (north_star: HBM-bound target): algorithmic bytes / CUDA-event launch duration, cold L2 (see blend_launch_times)
    if rank == 0:
        try:
            torch.cuda.synchronize(dev)
            time.sleep(2.0)                      # let the power-capped clocks of the frame loop recover: these two kernels are timed alone
            bt = blend_launch_times(pshape, raw, split, cai_mode, process_num, dev)
            floor_s = bt.pop("_event_pair_floor_s", 0.0)
            b_bytes, b_s = sum(v["bytes"] for v in bt.values()), sum(v["s"] for v in bt.values())
            roofline["blend"] = {"kernel": "blend_canvas_fast_kernel + blend_raw_seg_kernel", "bound": "hbm", "achieved": b_bytes / b_s / 1e9, "peak": peaks["hbm"],
                                 "unit": "GB/s", "frac": b_bytes / b_s / 1e9 / peaks["hbm"], "us_per_frame": b_s * 1e6,
                                 "stages": {k: {"algorithmic_bytes": v["bytes"], "us": v["s"] * 1e6, "GBps": v["bytes"] / v["s"] / 1e9,
                                                "frac": v["bytes"] / v["s"] / 1e9 / peaks["hbm"]} for k, v in bt.items()},
                                 "traffic": {k: traffic.get(k, {}).get("dram_bytes_per_launch") for k in ("blend_canvas_fast_kernel", "blend_raw_seg_kernel", "blend_raw_tab_kernel")},
                                 "event_pair_floor_us": floor_s * 1e6,
                                 "note": "median of 20 launches each, 256 MB L2 flush before every launch, own event pair per launch, timed alone (burst HBM peak applies). "
                                         "`us` and `frac` include what the event pair itself costs: event_pair_floor_us is the same pair around a one-element fill; "
                                         "ncu's gpu__time_duration of the same launches is in profiles/ (canvas ~14.4 us, rN ~40 us at the 4K r32 frame)"}
        except Exception as e:
            roofline["blend"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        att = main_res["kernels"].get("prv2_attention")
        if att:
            roofline["attention"] = dict(att, kernel="attention kernel (prv2_attention)")

    # the other precision mode on the same workload, same run: fp32-class (3-pass bf16 split, the mode that owns the 1e-3 parity bar)
    other = "fp32" if args.precision == "bf16" else "bf16"
    other_line, models = None, {args.precision: model}
    if not args.no_fp32_mode and world == 1:          # N = 1 only: the scaling runs time the headline mode
        try:
            m2, _ = make_model(other)
            o_steps = max(2, args.steps // 4) if other == "fp32" else args.steps
            r2 = measure(m2, lr_dev, lr_pin, hr_pin, o_steps, 3, not args.no_e2e)
            models[other] = m2
            if rank == 0:
                other_line = {"precision": other, "dtype": "f16x3 ((hi, lo) FP16 planes, three tensor-core passes per product, fp32 accumulate)" if other == "fp32" else "bf16", "value": r2["fps"], "unit": "frames/s",
                              "ms_per_step": r2["ms"] / r2["steps"], "steps": r2["steps"], "warmup": 3, "e2e": r2.get("e2e"), "gpu_launches": r2["launches"],
                              "patches_per_sec": r2["fps"] * n_patches, "roofline": roofline_of(r2, traffic),
                              "kernels": {k: {kk: v[kk] for kk in ("achieved", "unit", "frac", "ms_per_step", "launches_per_step")} for k, v in r2["kernels"].items()},
                              "note": "roofline.achieved counts ALGORITHMIC FLOPs (one product per multiply); the tensor pipe executes 3 FP16 passes per product in this mode"}
        except Exception as e:
            other_line = {"precision": other, "unavailable": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    parity = None
    if not args.no_parity and world == 1:
        try:
            parity = parity_block(models, cfg, sd, hr, lr_dev, hr_dev, log)
        except Exception as e:
            parity = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    cpu_baseline = eager_gpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            with stdout_to_stderr():
                v, cores, desc = reference_bounded_sample(cfg, sd, hr, cai_mode, process_num, n_regular, n_patches - n_regular, 3, log)
            cpu_baseline = {"value": v, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": desc}
        except Exception as e:
            cpu_baseline = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        try:
            # the honest "beat this" number: the reference's own forward as eager PyTorch fp32 on THIS GPU, whole frames
            for m in models.values():
                m._engine = None
            torch.cuda.empty_cache()
            with stdout_to_stderr():
                v, times, _, _ = reference_whole_frames(cfg, sd, hr, cai_mode, process_num, 2, 1, "cuda", log)
            eager_gpu = {"value": v, "unit": "frames/s", "kind": "reference", "sample": "the reference's own PatchRefiner.forward (oracle/_ref, eager PyTorch, "
                         "PyTorch default math modes: fp32 matmul, cuDNN convolutions may use TF32; library kernels + its per-patch host RunningAverageMap) "
                         f"on this GPU: 1 warm-up + {len(times)} whole frames, {', '.join(f'{t:.2f}s' for t in times)}"}
            if parity is not None and "unavailable" not in parity:
                # whole-frame parity against the reference's own output computed on this GPU in strict fp32 (cuDNN TF32 off)
                with stdout_to_stderr():
                    _, _, _, d_ref = reference_whole_frames(cfg, sd, hr, cai_mode, process_num, 1, 0, "cuda", log, exact_fp32=True)
                for name, m in models.items():
                    random.seed(1)
                    d, _ = m(mode="infer", image_lr=lr_dev, image_hr=hr_dev, cai_mode=cai_mode, process_num=process_num)
                    dr = d_ref.cpu()
                    rel = ((d.cpu() - dr).abs() / dr.clamp_min(0.1))[dr >= 0.1]
                    parity[f"{name}_frame_max_rel_vs_reference_on_gpu"] = float(rel.max())
                    parity[f"{name}_frame_p9999_rel_vs_reference_on_gpu"] = float(rel.flatten().kthvalue(max(1, int(rel.numel() * 0.9999))).values)
                    parity[f"{name}_frame_mean_rel_vs_reference_on_gpu"] = float(rel.mean())
        except Exception as e:                                     # a baseline must never take the bench line down
            eager_gpu = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if F_ == world else "strong", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f16x3 ((hi, lo) FP16 planes, fp32-class)", "data": "synthetic", "config": config,
            "patches_per_sec": fps * n_patches, "frames_per_step": F_, "algorithmic_tflop_per_frame": flops_frame / 1e12,
            "model_tflops_per_gpu": flops_frame * fps / 1e12 / world,
            "l2_policy": "working set per step (activations, several GB) far exceeds the 126 MB L2; no explicit flush",
            "clocks": clocks, "e2e": main_res.get("e2e"), "gpu_launches": main_res["launches"], "parity": parity, f"{other}_mode": other_line,
            "shard_check": shard_check, "roofline": roofline, "profiled_pass": main_res["profiled_pass"], "kernels": main_res["kernels"],
            "gemm_layers": main_res["gemm_layers"], "cpu_baseline": cpu_baseline, "eager_gpu_baseline": eager_gpu, "workspace_gb": workspace_gb}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

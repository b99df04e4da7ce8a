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
    "dav2_vitl_2160x3840_4x4_r32": ("vitl", (448, 448), (2160, 3840), (4, 4), "r32", 4),
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


def cpu_reference_sample(cfg, sd, hr, cai_mode, process_num, n_patches_frame, steps, warmup, log=lambda *a: None):
    """Times the oracle port of the reference's CPU path on the host cores: one coarse pass, `steps`
    single-patch refine passes (crop -> roi_align -> ViT + DPT -> FusionUnet) and the running-average
    blend update of one patch; extrapolates to a full frame.  Returns (frames/s, description)."""
    import torch
    from oracle import pr_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    orc = O.PatchRefinerOracle(cfg, sd)
    lr = O.resizer(tuple(cfg["patch_process_shape"]), hr)
    tc = orc.tile_cfg
    ph, pw = orc.patch_process_shape
    rh, rw = tc["patch_raw_shape"]
    H, W = tc["image_raw_shape"]
    with torch.no_grad():
        t0 = time.perf_counter()
        feats, coarse = orc.coarse_forward(lr)
        t_coarse = time.perf_counter() - t0
        tt = {"coarse_prediction": coarse, "coarse_features": feats}
        times = []
        for i in range(warmup + steps):
            bb = O.make_bboxs([(137 * i) % (H - rh)], [(211 * i) % (W - rw)], rh, rw)
            t0 = time.perf_counter()
            orc._predict(hr[0], bb, tc, tt, 1, None)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            log(f"cpu reference patch {i}: {dt:.2f}s")
        t_patch = statistics.mean(times)
        # blend: one RunningAverageMap.update at raw resolution (estimator/models/utils.py:31-36)
        avg = O.RunningAverageMap(torch.rand(H, W), torch.rand(H, W))
        pred, cnt = torch.zeros(H, W), torch.zeros(H, W)
        cnt[100:100 + rh, 200:200 + rw] = 1.0
        t0 = time.perf_counter()
        for _ in range(3):
            avg.update(pred, cnt)
        t_blend = (time.perf_counter() - t0) / 3
    frame_s = t_coarse + n_patches_frame * (t_patch + t_blend)
    desc = (f"oracle port of the reference CPU path (torch {torch.__version__}, {cores} threads): 1 coarse pass {t_coarse:.2f}s + "
            f"{steps} single-patch refine passes (mean {t_patch:.2f}s) + 1 blend update {t_blend * 1e3:.0f}ms, extrapolated to "
            f"{n_patches_frame} patches/frame")
    return 1.0 / frame_s, cores, desc


def eager_gpu_sample(cfg, sd, hr, process_num, n_patches_frame, dev, steps=3, warmup=1):
    """The same oracle port run as plain eager PyTorch (fp32, library kernels) on THIS GPU -- what the reference's own code
    path does on a CUDA device: network on the GPU, canvases and RunningAverageMap on the host after a D2H copy per patch
    (baseline_pretrain.py:340-373).  Bounded sample like the CPU one: 1 coarse pass + `steps` chunks of `process_num`
    patches + 1 host blend update, extrapolated to a frame.  A reported baseline, never part of the measured path."""
    import torch
    from oracle import pr_oracle as O
    orc = O.PatchRefinerOracle(cfg, {k: v.to(dev) for k, v in sd.items()})
    hr = hr.to(dev)
    lr = O.resizer(tuple(cfg["patch_process_shape"]), hr)
    tc = orc.tile_cfg
    rh, rw = tc["patch_raw_shape"]
    H, W = tc["image_raw_shape"]
    with torch.no_grad():
        feats, coarse = orc.coarse_forward(lr)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        feats, coarse = orc.coarse_forward(lr)
        torch.cuda.synchronize(dev)
        t_coarse = time.perf_counter() - t0
        tt = {"coarse_prediction": coarse, "coarse_features": feats}
        times = []
        for i in range(warmup + steps):
            bb = O.make_bboxs([(137 * (i * process_num + j)) % (H - rh) for j in range(process_num)], [(211 * i) % (W - rw)], rh, rw)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            p = orc._predict(hr[0], bb, tc, tt, process_num, None)
            p = p.cpu()                                            # the reference moves every prediction to the host canvas
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt / process_num)
        t_patch = statistics.mean(times)
        avg = O.RunningAverageMap(torch.rand(H, W), torch.rand(H, W))
        pred, cnt = torch.zeros(H, W), torch.zeros(H, W)
        cnt[100:100 + rh, 200:200 + rw] = 1.0
        t0 = time.perf_counter()
        for _ in range(3):
            avg.update(pred, cnt)
        t_blend = (time.perf_counter() - t0) / 3
    frame_s = t_coarse + n_patches_frame * (t_patch + t_blend)
    del orc
    torch.cuda.empty_cache()
    return {"value": 1.0 / frame_s, "unit": "frames/s", "kind": "oracle port, eager PyTorch fp32 on this GPU + host RunningAverageMap",
            "sample": f"1 coarse pass {t_coarse * 1e3:.0f}ms + {steps} chunks of {process_num} patches ({t_patch * 1e3:.0f}ms/patch incl. D2H) + "
                      f"host blend update {t_blend * 1e3:.0f}ms/patch on {torch.get_num_threads()} threads, extrapolated to {n_patches_frame} patches/frame",
            "network_only_fps": 1.0 / (t_coarse + n_patches_frame * t_patch)}


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
        out["raw"] = {"bytes": 4.0 * (2 * Hc * Wc + (bb.shape[0] - first) * ph * pw + rh * rw + 2 * H * W),
                      "s": run(lambda: ops.blend_raw(avg_c, cnt_c, preds[first:], starts, rmask, ph, pw, rh, rw, H, W))}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dav2_vitl_2160x3840_4x4_r32", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--patch-batch", type=int, default=0, help="patches per network launch (0 = balance automatically, <= 27)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-run", action="store_true", help="run under ncu: fewer warm-up steps allowed; the printed number is NOT a bench value")
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference" or args.profile_run, "timing rules: at least 3 warm-up steps"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    enc, pshape, raw, split, cai_mode, process_num = WORKLOADS[args.workload]

    import torch
    from patchrefinerv2_b200 import build_model, tiling
    cfg = make_config(enc, pshape, raw, split)
    tc = tiling.prepare_tile_cfg(pshape, raw, split)
    n_patches = sum(s.bboxs.shape[0] for s in tiling.schedule(tc, pshape, cai_mode, process_num, random.Random(0)))
    config = {"workload": args.workload, "image_raw_shape": list(raw), "patch_process_shape": list(pshape), "patch_split_num": list(split),
              "cai_mode": cai_mode, "process_num": process_num, "patches_per_frame": n_patches, "precision": args.precision,
              "parallelism": f"patch-shard x{world} + 1 NCCL sum-reduce" if world > 1 else "single GPU",
              "weights": "random-init (seeded), reference state-dict layout",
              "l2": "no explicit flush: every step streams several GB of activations (>> 126 MB L2) between reuses of any input"}

    if args.impl == "reference":
        if rank != 0:
            return
        spec = {k: tuple(v.shape) for k, v in build_model(dict(type="PatchRefiner", config=cfg)).state_dict().items()}
        sd = random_state_dict(spec, 0)
        fps, cores, desc = cpu_reference_sample(cfg, sd, synthetic_frame(raw, 1), cai_mode, process_num, n_patches, max(1, args.steps), max(0, args.warmup))
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config, "patches_per_sec": fps * n_patches,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    from patchrefinerv2_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout at the first collective; keep stdout for the one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    n_local = -(-n_patches // world)
    pb = args.patch_batch or -(-n_local // (-(-n_local // 27)))
    config["patch_batch"] = pb
    model = build_model(dict(type="PatchRefiner", config=cfg, precision=args.precision, patch_batch=pb, output_device="cuda"))
    sd = random_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 0)
    res = model.load_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    model = model.cuda().eval()
    hr = synthetic_frame(raw, 1)
    hr_dev = hr.to(dev)
    lr_dev = model.resizer(hr_dev)                         # image_lr as the dataset builds it (general_dataset.py:218), on the device
    lr = lr_dev.cpu()
    lr_pin, hr_pin = lr.pin_memory(), hr.pin_memory()
    shard = world > 1

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step_resident():
        random.seed(1)
        d, _ = model(mode="infer", image_lr=lr_dev, image_hr=hr_dev, cai_mode=cai_mode, process_num=process_num, shard=shard)
        return d

    host_out = torch.empty((1, 1) + ((raw[0], raw[1]) if cai_mode[0] == "r" else tc["patch_reensemble_shape"]), dtype=torch.float32).pin_memory()

    def step_e2e():
        random.seed(1)
        a = lr_pin.to(dev, non_blocking=True)
        b = hr_pin.to(dev, non_blocking=True)
        d, _ = model(mode="infer", image_lr=a, image_hr=b, cai_mode=cai_mode, process_num=process_num, shard=shard)
        host_out.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller holds the depth map on the host when the step ends
        return host_out

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
        log, _lib.profile_log = _lib.profile_log, None
        t = torch.tensor([ms], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item()), _lib.launch_count - l0, log

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches, log = timed(step_resident, args.steps, args.warmup, profile=True)
    clocks = sampler.stop() if sampler else None
    fps = args.steps / (ms / 1e3)

    e2e = None
    if not args.no_e2e:
        ms_e, _, _ = timed(step_e2e, args.steps, 1)
        h2d = lr.numel() * 4 + hr.numel() * 4
        e2e = {"value": args.steps / (ms_e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": host_out.numel() * 4,
               "ms_per_step": ms_e / args.steps}

    # multi-GPU parity evidence (untimed): the sharded frame against the same frame refined by this rank alone
    shard_check = None
    if world > 1:
        d_sh = step_resident().clone()
        random.seed(1)
        d_one, _ = model(mode="infer", image_lr=lr_dev, image_hr=hr_dev, cai_mode=cai_mode, process_num=process_num, shard=False)
        shard_check = {"max_rel_diff_vs_unsharded": float(((d_sh - d_one).abs() / d_one.abs().clamp_min(1e-3)).max().item())}
        torch.distributed.barrier()

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # per-kernel CUDA-event durations from the timed region (launching stream)
    peaks = measured_peaks()
    agg = {}
    layers = {}
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
        kern[name].update(launches_per_step=g["launches"] / args.steps, ms_per_step=g["ms"] / args.steps,
                          share_of_step=g["ms"] / ms, avg_launch_us=1e3 * g["ms"] / g["launches"])
    gemm_layers = {k: {"tflops": v["flop"] / (v["ms"] * 1e-3) / 1e12, "frac": v["flop"] / (v["ms"] * 1e-3) / 1e12 / peaks["tf_sustained"],
                       "ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps} for k, v in sorted(layers.items())}
    gemm = kern.get("prv2_umma_gemm", {})
    traffic = profiled_traffic() or {}
    roofline = {"kernel": "umma_gemm_kernel (prv2_umma_gemm)", "bound": "tensor", "achieved": gemm.get("achieved"), "peak": peaks["tf_sustained"],
                "unit": "TFLOP/s", "frac": gemm.get("frac"), "traffic": traffic.get("umma_gemm_kernel", {}).get("dram_bytes_per_launch"),
                "traffic_source": traffic.get("source"), "peak_source": f"{peaks['source']} bf16 sustained",
                "share_of_step": gemm.get("share_of_step"), "launches_per_step": gemm.get("launches_per_step"),
                "note": "aggregate over all launches of the kernel in the timed region: sum(algorithmic FLOPs) / sum(CUDA-event durations)"}

    # the CAI blend (north_star: HBM-bound target): algorithmic bytes / CUDA-event launch duration, cold L2 (see blend_launch_times)
    roofline_blend = None
    try:
        torch.cuda.synchronize(dev)
        time.sleep(2.0)                      # let the power-capped clocks of the frame loop recover: these two kernels are timed alone
        bt = blend_launch_times(pshape, raw, split, cai_mode, process_num, dev)
        b_bytes, b_s = sum(v["bytes"] for v in bt.values()), sum(v["s"] for v in bt.values())
        roofline_blend = {"kernel": "blend_canvas_fast_kernel + blend_raw_tab_kernel", "bound": "hbm", "achieved": b_bytes / b_s / 1e9, "peak": peaks["hbm"],
                          "unit": "GB/s", "frac": b_bytes / b_s / 1e9 / peaks["hbm"], "us_per_frame": b_s * 1e6,
                          "stages": {k: {"algorithmic_bytes": v["bytes"], "us": v["s"] * 1e6, "GBps": v["bytes"] / v["s"] / 1e9,
                                         "frac": v["bytes"] / v["s"] / 1e9 / peaks["hbm"]} for k, v in bt.items()},
                          "traffic": {k: traffic.get(k, {}).get("dram_bytes_per_launch") for k in ("blend_canvas_fast_kernel", "blend_raw_tab_kernel")},
                          "note": "median of 20 launches each, 256 MB L2 flush before every launch, own event pair per launch"}
    except Exception as e:
        roofline_blend = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    cpu_baseline = eager_gpu = None
    if not args.no_cpu_baseline and world == 1:
        v, cores, desc = cpu_reference_sample(cfg, sd, hr, cai_mode, process_num, n_patches, 6, 1)
        cpu_baseline = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc}
        try:
            eager_gpu = eager_gpu_sample(cfg, sd, hr, process_num, n_patches, dev)
        except Exception as e:                                     # a baseline must never take the bench line down
            eager_gpu = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    flops_frame = model._engine["coarse"].flops(1, *pshape) + n_patches * (model._engine["fine"].flops(1, *pshape) +
                  model._engine["fusion"].flops(1, [(f.H, f.W) for f in model._engine["coarse"].forward(lr_dev)[1]][::-1]))
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (fp32-class split)", "data": "synthetic", "config": config,
            "patches_per_sec": fps * n_patches, "algorithmic_tflop_per_frame": flops_frame / 1e12,
            "model_tflops_per_gpu": flops_frame * fps / 1e12 / world,
            "l2_policy": "working set per step (activations, several GB) far exceeds the 126 MB L2; no explicit flush",
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "shard_check": shard_check, "roofline": roofline, "roofline_blend": roofline_blend, "kernels": kern, "gemm_layers": gemm_layers, "cpu_baseline": cpu_baseline, "eager_gpu_baseline": eager_gpu,
            "workspace_gb": sum(w.nbytes() for eng in (model._engine["coarse"], model._engine["fine"], model._engine["fusion"]) for w in eng.ws.values()) / 1e9}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Frame rate of the V2 family on one GPU: PatchRefinerPlus shaped like configs/patchrefinerv2_dav2/plus_mobile_u4k_*: DAv2 ViT-L coarse
branch + MobileNetV4-conv-small refiner encoder (mnv4.py) + BiDirectionalFusion (coarse-gated), 2160x3840, 4x4 patches, CAI r32.
Random-init weights of the model's own state-dict spec, synthetic frame, CUDA events, >= 3 warm-up frames.  Prints one JSON line.
    python scripts/bench_plus.py [--precision bf16|fp32] [--steps 4]"""
import argparse
import json
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (weight / frame generators)
from patchrefinerv2_b200 import _lib, build_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--patch-batch", type=int, default=27)
    a = ap.parse_args()
    raw, pshape, split = (2160, 3840), (448, 448), (4, 4)
    enc = "mobilenetv4_conv_small.e2400_r224_in1k"
    cfg = dict(image_raw_shape=list(raw), patch_process_shape=list(pshape), patch_split_num=list(split), fusion_feat_level=6, min_depth=1e-3, max_depth=80.0,
               strategy_refiner_target="offset_coarse", pretrain_stage=False, e2e_training=False, hack_strategy=None,
               coarse_branch=dict(type="DA2", pretrained=None, model_cfg=dict(encoder="vitl", features=256, out_channels=[256, 512, 1024, 1024])),
               refiner=dict(fine_branch=dict(type="LightWeightRefiner", coarse_condition=True, with_decoder=False, encoder_name=enc),
                            fusion_model=dict(type="BiDirectionalFusion", encoder_name=enc, coarse2fine=True, coarse2fine_type="coarse-gated",
                                              coarse_chl=[128, 256, 256, 256, 256, 256], fine_chl=[32, 32, 64, 96, 960],
                                              fine_chl_after_coarse2fine=[128, 256, 256, 256, 256, 256], temp_chl=[32, 64, 64, 128, 256, 512],
                                              dec_chl=[512, 256, 128, 64, 32])),
               pre_norm_bbox=True, pretrain_coarse_model=None, pretrained=None, whole_pretrained=None)
    m = build_model(dict(type="PatchRefinerPlus", config=cfg, precision=a.precision, patch_batch=a.patch_batch, output_device="cuda"))
    spec = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = bench.random_state_dict({k: v for k, v in spec.items() if len(v)}, 0)
    g = torch.Generator().manual_seed(5)
    for k in spec:                                                  # BatchNorm statistics of the encoder: positive variances, small means
        if k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(spec[k], generator=g)
        elif k.endswith("running_mean"):
            sd[k] = 0.1 * torch.randn(spec[k], generator=g)
        elif k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros(spec[k])
    res = m.load_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    m = m.cuda().eval()
    hr = bench.synthetic_frame(raw, 1).cuda()
    lr = m.resizer(hr)

    def step():
        random.seed(1)
        d, _ = m(mode="infer", image_lr=lr, image_hr=hr, cai_mode="r32", process_num=4)
        return d

    for _ in range(3):
        d = step()
    torch.cuda.synchronize()
    _lib.launch_count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        d = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"workload": "plus_mobile_vitl_2160x3840_4x4_r32", "precision": a.precision, "frames_per_sec": 1000.0 / ms, "ms_per_frame": ms,
                      "patches_per_frame": 81, "gpu_launches_per_frame": _lib.launch_count / a.steps, "depth_range": [float(d.min()), float(d.max())],
                      "finite": bool(torch.isfinite(d).all()), "encoder": "MobileNetV4-conv-small on prv2_umma_gemm / prv2_dwconv (mnv4.py)"}))


if __name__ == "__main__":
    main()

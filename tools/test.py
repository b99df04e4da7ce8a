#!/usr/bin/env python
"""``tools/test.py`` of the reference (absent from its snapshot; documented in README.md:57-77 and
docs/user_infer.md:113-130) for the B200 implementation: the ``--test-type general`` path.

    python tools/test.py CONFIG --ckp-path CKPT --cai-mode {m1,m2,rN} \
        --cfg-option general_dataloader.dataset.rgb_image_dir='<dir>' [--save] --work-dir OUT \
        --test-type general [--gray-scale] --image-raw-shape H W --patch-split-num h w

CONFIG is one of the reference's own python configs (loaded unmodified by ``patchrefinerv2_b200.config``);
the checkpoint is a ``torch.save`` dict with ``model_state_dict`` (estimator/trainer/trainer.py:276-294) or a
bare state dict; the loop is ``Tester.run`` (estimator/tester/tester.py:52-106).  Launched with ``torchrun --nproc-per-node N``
the patches of every frame are sharded over the N GPUs (one NCCL sum-reduce per frame) instead of the frames over ranks.
"""
from __future__ import annotations

import argparse
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="PatchRefinerV2 tiled inference on B200 (drop-in for the reference's tools/test.py)")
    ap.add_argument("config")
    ap.add_argument("--ckp-path", default=None, help="checkpoint ({'model_state_dict': ...} or a bare state dict)")
    ap.add_argument("--cai-mode", default="m1", help="m1 | m2 | rN")
    ap.add_argument("--process-num", type=int, default=4, help="patches per reference forward (also fixes the rN bbox stream)")
    ap.add_argument("--cfg-option", nargs="+", default=[], help="dotted key=value overrides of the config")
    ap.add_argument("--save", action="store_true")
    ap.add_argument("--work-dir", default="./work_dir/predictions")
    ap.add_argument("--test-type", default="general", choices=["general", "benchmark", "normal"],
                    help="general: image directory -> depth maps (Tester.run); benchmark: Tester.benchmark on the same images (frames/s, benchmark.txt)")
    ap.add_argument("--repeat-times", type=int, default=10, help="--test-type benchmark: number of runs (tester.py:326)")
    ap.add_argument("--gray-scale", action="store_true")
    ap.add_argument("--image-raw-shape", nargs=2, type=int, default=[2160, 3840])
    ap.add_argument("--patch-split-num", nargs=2, type=int, default=[4, 4])
    ap.add_argument("--seed", type=int, default=621, help="fix_random_seed (estimator/utils/misc.py:16-26)")
    ap.add_argument("--precision", default="fp32", choices=["bf16", "fp32"],
                    help="fp32 (default): the fp32-class mode that meets the reference within 1e-3 relative; bf16: the one-pass mode, ~2.7x faster, its own tolerance (5e-2)")
    ap.add_argument("--patch-batch", type=int, default=27)
    ap.add_argument("--frames-per-call", type=int, default=0,
                    help="frames handed to one model call (0 = one per GPU under torchrun, else 1).  A batch of F frames is ONE work list of F x P patches "
                         "(split over the ranks, one NCCL sum-reduce per batch); the depth maps are the ones F successive single-frame calls give")
    return ap.parse_args(argv)


def build(args):
    """Config -> model (no CUDA needed up to here)."""
    from patchrefinerv2_b200 import build_model
    from patchrefinerv2_b200.config import Config, parse_cfg_options
    cfg = Config.fromfile(args.config)
    cfg.merge_from_dict(parse_cfg_options(args.cfg_option))
    if args.test_type not in ("general", "benchmark"):
        raise NotImplementedError("--test-type general (image directory) and benchmark are implemented; dataset evaluation is out of scope")
    mcfg = cfg.model.to_dict()
    # weights come from --ckp-path: the per-branch 'pretrained' / pretrain_* files named by the training configs are optional here
    for br in (mcfg["config"].get("coarse_branch", {}), mcfg["config"].get("refiner", {}).get("fine_branch", {})):
        if br.get("pretrained") and not os.path.exists(br["pretrained"]):
            br["pretrained"] = None
    for k in ("pretrain_coarse_model", "pretrain_fine_model"):
        if mcfg["config"].get(k) and not os.path.exists(mcfg["config"][k]):
            mcfg["config"][k] = None
    model = build_model(dict(type=mcfg["type"], config=mcfg["config"], precision=args.precision, patch_batch=args.patch_batch))
    if args.ckp_path:
        import torch
        sd = torch.load(args.ckp_path, map_location="cpu")
        sd = sd.get("model_state_dict", sd)
        print(model.load_dict(sd))
    return cfg, model


def main(argv=None):
    args = parse_args(argv)
    cfg, model = build(args)
    import numpy as np
    import torch
    from patchrefinerv2_b200 import frames
    if not torch.cuda.is_available():
        raise SystemExit("tools/test.py needs a CUDA device (sm_100a); there is no CPU path")
    # under torchrun (one rank per GPU) every frame's patches are sharded over the ranks and combined with one NCCL sum-reduce
    # (DESIGN.md section 6); all ranks walk the same files and draw the same random patches, rank 0 writes the outputs
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    shard = world > 1
    if shard:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    random.seed(args.seed); np.random.seed(args.seed); torch.manual_seed(args.seed)
    model = model.cuda().eval()
    img_dir = cfg.general_dataloader.dataset.rgb_image_dir
    if args.test_type == "benchmark":                                      # Tester.benchmark (tester.py:325-404)
        from patchrefinerv2_b200 import metrics
        batches = []
        for name, image_hr in frames.iter_frames(img_dir, args.image_raw_shape):
            hr = image_hr.cuda().unsqueeze(0)
            batches.append({"image_lr": model.resizer(hr), "image_hr": hr})
            if len(batches) >= 50:
                break
        res = metrics.benchmark(model, batches, args.cai_mode, args.process_num, args.image_raw_shape, args.patch_split_num, repeat_times=args.repeat_times,
                                work_dir=args.work_dir if rank == 0 else None, shard=shard, log=print if rank == 0 else (lambda *a: None))
        if shard:
            torch.distributed.destroy_process_group()
        if rank == 0:
            print(res)
        return
    n, t0 = 0, time.perf_counter()
    per_call = args.frames_per_call if args.frames_per_call > 0 else (world if shard else 1)
    tile_cfg = {"image_raw_shape": list(args.image_raw_shape), "patch_split_num": list(args.patch_split_num)}

    def run_batch(names, hrs):
        hr = torch.stack(hrs).cuda()
        lr = model.resizer(hr)                                           # general_dataset.py:218 (on the device)
        result, log = model(mode="infer", cai_mode=args.cai_mode, process_num=args.process_num, tile_cfg=tile_cfg, image_lr=lr, image_hr=hr, shard=shard)
        if args.save and rank == 0:
            for f, name in enumerate(names):                             # frame by frame, as Tester.run writes them (tester.py:73-122)
                print(torch.max(result[f:f + 1]))
                frames.save_prediction(result[f:f + 1], args.work_dir, name, args.gray_scale, log["coarse_prediction"][f:f + 1], args.image_raw_shape)

    names, hrs = [], []
    for name, image_hr in frames.iter_frames(img_dir, args.image_raw_shape):
        names.append(name); hrs.append(image_hr)
        n += 1
        if len(names) == per_call:
            run_batch(names, hrs)
            names, hrs = [], []
    if names:
        run_batch(names, hrs)
    dt = time.perf_counter() - t0
    if shard:
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    print(f"{n} frame(s) in {dt:.2f}s ({n / dt if dt > 0 else 0:.2f} img / s incl. file I/O) -> {args.work_dir if args.save else '(not saved)'}")


if __name__ == "__main__":
    main()

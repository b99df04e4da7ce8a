#!/usr/bin/env python
"""Multi-GPU check of tools/test.py (under gpurun --gpus 2): writes a tiny config / checkpoint / three frames, runs the CLI once on one
GPU frame by frame and once under torchrun with one frame per GPU per model call, and compares the uint16 files (sharded partial sums are
combined by one NCCL sum-reduce, so the depth differs from the sequential blend by rounding only: <= 1e-3 relative + one quantisation step).
    python scripts/tools_torchrun_check.py [nproc]"""
import os
import subprocess
import sys
import tempfile

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pr_oracle as O  # noqa: E402  (seeded tiny weights only; the CLI itself never imports oracle/)

nproc = int(sys.argv[1]) if len(sys.argv) > 1 else 2
d = tempfile.mkdtemp()
os.makedirs(f"{d}/cfg"); os.makedirs(f"{d}/imgs")
cfg = O.make_config("vits", (224, 224), (432, 768), (2, 2))
open(f"{d}/cfg/base_dataset.py", "w").write("general_dataloader = dict(batch_size=1, num_workers=0, dataset=dict(type='ImageDataset', rgb_image_dir='', dataset_name=''))\n")
open(f"{d}/cfg/tiny.py", "w").write("_base_ = ['./base_dataset.py']\nmodel = dict(type='PatchRefiner', config=%r)\n" % (cfg,))
torch.save({"model_state_dict": O.init_patchrefiner_state_dict(cfg, 0)}, f"{d}/ckpt.pth")
rng = np.random.default_rng(5)
for name in ("f0.png", "f1.png", "f2.png"):
    cv2.imwrite(f"{d}/imgs/{name}", rng.integers(0, 256, (216, 384, 3), dtype=np.uint8))
base = [f"{d}/cfg/tiny.py", "--ckp-path", f"{d}/ckpt.pth", "--cai-mode", "r2", "--process-num", "2", "--cfg-option",
        f"general_dataloader.dataset.rgb_image_dir={d}/imgs", "--save", "--test-type", "general", "--image-raw-shape", "432", "768",
        "--patch-split-num", "2", "2", "--precision", "fp32", "--patch-batch", "4", "--seed", "7"]
subprocess.check_call([sys.executable, f"{ROOT}/tools/test.py"] + base + ["--work-dir", f"{d}/one"])
subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1", "--master-port", "29577",
                       f"{ROOT}/tools/test.py"] + base + ["--work-dir", f"{d}/many"])
worst = 0.0
for name in ("f0", "f1", "f2"):
    a = cv2.imread(f"{d}/one/{name}_uint16.png", cv2.IMREAD_UNCHANGED).astype(np.float64)
    b = cv2.imread(f"{d}/many/{name}_uint16.png", cv2.IMREAD_UNCHANGED).astype(np.float64)
    assert a.shape == b.shape == (432, 768)
    err = np.abs(a - b) - 1.0
    worst = max(worst, float((err / np.maximum(a, 1.0)).max()))
    assert np.all(np.abs(a - b) <= 1e-3 * a + 1.0), name
print(f"TOOLS_TORCHRUN_OK nproc={nproc} worst relative difference beyond one quantisation step: {max(worst, 0.0):.2e}")

"""tools/test.py end to end on the GPU (the drop-in CLI of SURVEY.md 8(f) row 1): config file -> registry model -> checkpoint ->
image directory -> uint16 PNGs, compared with the model called directly and with the CPU oracle."""
import importlib.util
import os
import random

import cv2
import numpy as np
import pytest
import torch

from oracle import pr_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CFG = """
_base_ = ['./base_dataset.py']
model = dict(type='PatchRefiner', config=%r)
"""
BASE = "general_dataloader = dict(batch_size=1, num_workers=0, dataset=dict(type='ImageDataset', rgb_image_dir='', dataset_name=''))\n"


def _tools():
    spec = importlib.util.spec_from_file_location("prv2_tools_test_gpu", os.path.join(ROOT, "tools", "test.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_cli_general_mode_writes_reference_format_predictions(tmp_path, tiny_setup):
    cfg, sd, lr, hr = tiny_setup
    (tmp_path / "cfg").mkdir(); (tmp_path / "imgs").mkdir()
    (tmp_path / "cfg" / "base_dataset.py").write_text(BASE)
    (tmp_path / "cfg" / "tiny.py").write_text(CFG % (cfg,))
    torch.save({"model_state_dict": sd}, tmp_path / "ckpt.pth")
    rng = np.random.default_rng(3)
    for name in ("frame_b.png", "frame_a.png"):
        cv2.imwrite(str(tmp_path / "imgs" / name), rng.integers(0, 256, (216, 384, 3), dtype=np.uint8))
    t = _tools()
    argv = [str(tmp_path / "cfg" / "tiny.py"), "--ckp-path", str(tmp_path / "ckpt.pth"), "--cai-mode", "r2", "--process-num", "2",
            "--cfg-option", f"general_dataloader.dataset.rgb_image_dir={tmp_path / 'imgs'}", "--save", "--work-dir", str(tmp_path / "out"),
            "--test-type", "general", "--image-raw-shape", "432", "768", "--patch-split-num", "2", "2", "--precision", "fp32", "--patch-batch", "4", "--seed", "7"]
    t.main(argv)
    outs = sorted(os.listdir(tmp_path / "out"))
    assert outs == ["frame_a.png", "frame_a_coarse.png", "frame_a_uint16.png", "frame_b.png", "frame_b_coarse.png", "frame_b_uint16.png"]
    # the same frames through the oracle (reference arithmetic, CPU): uint16 files within 1e-3 relative (+1 quantisation step)
    from patchrefinerv2_b200 import frames
    random.seed(7)
    orc = O.PatchRefinerOracle(cfg, sd)
    for name, image_hr in frames.iter_frames(str(tmp_path / "imgs"), (432, 768)):          # sorted order, as the CLI walks it
        hr1 = image_hr.unsqueeze(0)
        want, _, _ = orc.infer(O.resizer(cfg["patch_process_shape"], hr1), hr1, None, "r2", 2)
        got = cv2.imread(str(tmp_path / "out" / f"{name}_uint16.png"), cv2.IMREAD_UNCHANGED)
        assert got.dtype == np.uint16 and got.shape == (432, 768)
        w16 = want[0, 0].numpy() * 256
        assert np.all(np.abs(got.astype(np.float64) - w16) <= 1e-3 * np.abs(w16) + 1.0)


def test_cli_frames_per_call_gives_the_same_files(tmp_path, tiny_setup):
    """--frames-per-call F: F frames are one work list (the form that scales under torchrun).  The depth maps must be the very files
    the frame-by-frame run writes -- same schedule draws (one `random` stream consumed frame by frame), same per-patch results."""
    cfg, sd, lr, hr = tiny_setup
    (tmp_path / "cfg").mkdir(); (tmp_path / "imgs").mkdir()
    (tmp_path / "cfg" / "base_dataset.py").write_text(BASE)
    (tmp_path / "cfg" / "tiny.py").write_text(CFG % (cfg,))
    torch.save({"model_state_dict": sd}, tmp_path / "ckpt.pth")
    rng = np.random.default_rng(5)
    for name in ("f0.png", "f1.png", "f2.png"):
        cv2.imwrite(str(tmp_path / "imgs" / name), rng.integers(0, 256, (216, 384, 3), dtype=np.uint8))
    t = _tools()
    base = [str(tmp_path / "cfg" / "tiny.py"), "--ckp-path", str(tmp_path / "ckpt.pth"), "--cai-mode", "r2", "--process-num", "2",
            "--cfg-option", f"general_dataloader.dataset.rgb_image_dir={tmp_path / 'imgs'}", "--save", "--test-type", "general",
            "--image-raw-shape", "432", "768", "--patch-split-num", "2", "2", "--precision", "bf16", "--patch-batch", "4", "--seed", "7"]
    t.main(base + ["--work-dir", str(tmp_path / "one"), "--frames-per-call", "1"])
    t.main(base + ["--work-dir", str(tmp_path / "two"), "--frames-per-call", "2"])          # 3 frames: a batch of 2 and a batch of 1
    assert sorted(os.listdir(tmp_path / "one")) == sorted(os.listdir(tmp_path / "two")) and len(os.listdir(tmp_path / "one")) == 9
    for name in ("f0", "f1", "f2"):
        a = cv2.imread(str(tmp_path / "one" / f"{name}_uint16.png"), cv2.IMREAD_UNCHANGED)
        b = cv2.imread(str(tmp_path / "two" / f"{name}_uint16.png"), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(a, b), name

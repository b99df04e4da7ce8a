"""PatchRefinerPlus (V2 family) end to end: oracle vs the reference (CPU, with the toy encoder injected for timm) and vs the
reference-made goldens; the sm_100a model vs the goldens on the GPU."""
import os
import random
import tempfile

import numpy as np
import pytest
import torch

from oracle import pr_oracle as O
from oracle import ref_shim
from oracle.make_golden import PLUS_MODES, sd_digest


@pytest.fixture(scope="module")
def plus_setup():
    cfg = O.make_plus_config()
    sd = O.init_patchrefinerplus_state_dict(cfg, 0)
    lr, hr = O.synthetic_frame(cfg, 1)
    return cfg, sd, lr, hr


@pytest.mark.parametrize("mode,pn", PLUS_MODES[:1])
def test_plus_oracle_matches_reference_golden(golden_dir, plus_setup, mode, pn):
    cfg, sd, lr, hr = plus_setup
    g = np.load(os.path.join(golden_dir, f"plus_{mode}.npz"))
    assert str(g["sd_sha"]) == sd_digest(sd) and str(g["frame_sha"]) == O.sha256_f32(hr.numpy()), "generators drifted from the golden run"
    assert sorted(sd.keys()) == list(g["keys"])
    random.seed(1)
    depth, coarse, _ = O.PatchRefinerPlusOracle(cfg, sd, O.ToyFineEncoder(4)).infer(lr, hr, None, mode, pn)
    np.testing.assert_allclose(coarse.numpy(), g["coarse"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(depth.numpy(), g["depth"], rtol=1e-3, atol=1e-3)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
def test_plus_oracle_is_bit_identical_to_reference(plus_setup):
    cfg, sd, lr, hr = plus_setup
    d = tempfile.mkdtemp()
    cp = os.path.join(d, "c.pth")
    torch.save({k[len("coarse_branch."):]: v for k, v in sd.items() if k.startswith("coarse_branch.")}, cp)
    cwd = os.getcwd()
    try:
        ref = ref_shim.build_reference_patchrefinerplus(cfg, cp, lambda: O.ToyFineEncoder(3))
        res = ref.load_state_dict(sd, strict=False)
        assert not res.missing_keys and not res.unexpected_keys
        random.seed(1)
        with torch.no_grad():
            dref, log = ref(mode="infer", image_lr=lr, image_hr=hr, cai_mode="r2", process_num=2, tile_cfg=None)
        random.seed(1)
        dor, coarse, _ = O.PatchRefinerPlusOracle(cfg, sd, O.ToyFineEncoder(4)).infer(lr, hr, None, "r2", 2)
        assert torch.equal(dref, dor) and torch.equal(log["coarse_prediction"], coarse)
        from patchrefinerv2_b200 import build_model
        mine = build_model(dict(type="PatchRefinerPlus", config=cfg, fine_encoder=O.ToyFineEncoder(4)))
        assert set(mine.state_dict().keys()) == set(ref.state_dict().keys())
        assert set(mine.get_save_dict().keys()) == set(ref.get_save_dict().keys())
    finally:
        os.chdir(cwd)


def test_plus_surface_without_gpu(plus_setup, monkeypatch):
    import sys
    monkeypatch.setitem(sys.modules, "timm", None)                # no timm (also hides the reference shim's stub when it ran first)
    from patchrefinerv2_b200 import PatchRefinerPlus, build_model
    cfg, sd, lr, hr = plus_setup
    m = build_model(dict(type="PatchRefinerPlus", config=cfg, fine_encoder=O.ToyFineEncoder(4)))
    assert isinstance(m, PatchRefinerPlus)
    res = m.load_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    assert m.tile_cfg == O.prepare_tile_cfg(cfg["patch_process_shape"], cfg["image_raw_shape"], cfg["patch_split_num"])
    eff = dict(cfg, refiner=dict(cfg["refiner"], fine_branch=dict(cfg["refiner"]["fine_branch"], encoder_name="tf_efficientnet_b5_ap")))
    with pytest.raises(NotImplementedError):                      # a timm-only encoder and no timm here: loud, with the way out
        build_model(dict(type="PatchRefinerPlus", config=eff))
    with pytest.raises(RuntimeError):                             # no CPU path
        m(mode="infer", image_lr=lr, image_hr=hr, cai_mode="m1", process_num=2)
    bad = dict(cfg); bad["refiner"] = dict(cfg["refiner"]); bad["refiner"]["fusion_model"] = dict(cfg["refiner"]["fusion_model"], coarse2fine_type="no-such-type")
    with pytest.raises(NotImplementedError):                      # unknown fusion variants fail loudly
        build_model(dict(type="PatchRefinerPlus", config=bad, fine_encoder=O.ToyFineEncoder(4)))


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp32", 1e-3), ("bf16", 5e-2)])
@pytest.mark.parametrize("mode,pn", PLUS_MODES)
def test_b200_plus_matches_reference_golden(golden_dir, plus_setup, mode, pn, prec, tol):
    """fp32 mode: depth within 1e-3 relative PER PIXEL of the reference PatchRefinerPlus output (north-star bar);
    bf16 mode: relative to the depth range, stated separately."""
    from patchrefinerv2_b200 import build_model
    cfg, sd, lr, hr = plus_setup
    g = np.load(os.path.join(golden_dir, f"plus_{mode}.npz"))
    m = build_model(dict(type="PatchRefinerPlus", config=cfg, precision=prec, patch_batch=3, fine_encoder=O.ToyFineEncoder(4)))
    m.load_dict(sd)
    m = m.cuda().eval()
    random.seed(1)
    depth, log = m(mode="infer", image_lr=lr.cuda(), image_hr=hr.cuda(), cai_mode=mode, process_num=pn)
    want = torch.from_numpy(g["depth"])
    assert depth.device.type == "cpu" and depth.dtype == torch.float32 and depth.shape == want.shape
    if prec == "fp32":
        rel = ((depth - want).abs() / want.abs().clamp_min(1e-2)).max().item()
        assert rel < tol, rel
    else:
        assert ((depth - want).abs().max() / want.abs().max()).item() < tol
    assert ((log["coarse_prediction"].cpu() - torch.from_numpy(g["coarse"])).abs().max() / float(g["coarse"].max())) < tol


# ---- convnext-named encoders: four encoder maps + LightWeightRefiner's own upsample_convx stage (lightweight_refiner.py:276-283, 307-314) ----
@pytest.fixture(scope="module")
def convx_setup():
    cfg = O.make_plus_config(convnext=True)
    sd = O.init_patchrefinerplus_state_dict(cfg, 0)
    lr, hr = O.synthetic_frame(cfg, 1)
    return cfg, sd, lr, hr


def test_plus_convx_oracle_matches_reference_golden(golden_dir, convx_setup):
    cfg, sd, lr, hr = convx_setup
    g = np.load(os.path.join(golden_dir, "plus_convx_r2.npz"))
    assert str(g["sd_sha"]) == sd_digest(sd) and str(g["frame_sha"]) == O.sha256_f32(hr.numpy()), "generators drifted from the golden run"
    assert sorted(sd.keys()) == list(g["keys"])
    random.seed(1)
    depth, coarse, _ = O.PatchRefinerPlusOracle(cfg, sd, O.ToyConvNeXtEncoder(4)).infer(lr, hr, None, "r2", 2)
    np.testing.assert_allclose(depth.numpy(), g["depth"], rtol=1e-3, atol=1e-3)
    from patchrefinerv2_b200 import build_model
    mine = build_model(dict(type="PatchRefinerPlus", config=cfg, fine_encoder=O.ToyConvNeXtEncoder(4)))
    assert sorted(mine.state_dict().keys()) == list(g["keys"])                       # incl. refiner_fine_branch.upsample_convx.0.{weight,bias}
    res = mine.load_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
def test_plus_convx_oracle_is_bit_identical_to_reference(convx_setup):
    cfg, sd, lr, hr = convx_setup
    d = tempfile.mkdtemp()
    cp = os.path.join(d, "c.pth")
    torch.save({k[len("coarse_branch."):]: v for k, v in sd.items() if k.startswith("coarse_branch.")}, cp)
    cwd = os.getcwd()
    try:
        ref = ref_shim.build_reference_patchrefinerplus(cfg, cp, lambda: O.ToyConvNeXtEncoder(3))
        res = ref.load_state_dict(sd, strict=False)
        assert not res.missing_keys and not res.unexpected_keys
        random.seed(1)
        with torch.no_grad():
            dref, log = ref(mode="infer", image_lr=lr, image_hr=hr, cai_mode="m1", process_num=2, tile_cfg=None)
        random.seed(1)
        dor, coarse, _ = O.PatchRefinerPlusOracle(cfg, sd, O.ToyConvNeXtEncoder(4)).infer(lr, hr, None, "m1", 2)
        assert torch.equal(dref, dor) and torch.equal(log["coarse_prediction"], coarse)
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp32", 1e-3), ("bf16", 5e-2)])
def test_b200_plus_convx_matches_reference_golden(golden_dir, convx_setup, prec, tol):
    from patchrefinerv2_b200 import build_model
    cfg, sd, lr, hr = convx_setup
    g = np.load(os.path.join(golden_dir, "plus_convx_r2.npz"))
    m = build_model(dict(type="PatchRefinerPlus", config=cfg, precision=prec, patch_batch=3, fine_encoder=O.ToyConvNeXtEncoder(4)))
    m.load_dict(sd)
    m = m.cuda().eval()
    random.seed(1)
    depth, log = m(mode="infer", image_lr=lr.cuda(), image_hr=hr.cuda(), cai_mode="r2", process_num=2)
    want = torch.from_numpy(g["depth"])
    assert depth.shape == want.shape
    if prec == "fp32":
        rel = ((depth - want).abs() / want.abs().clamp_min(1e-2)).max().item()
        assert rel < tol, rel
    else:
        assert ((depth - want).abs().max() / want.abs().max()).item() < tol

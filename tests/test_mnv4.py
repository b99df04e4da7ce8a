"""MobileNetV4-conv-small refiner encoder on the kernels (patchrefinerv2_b200/mnv4.py) against this repository's PyTorch restatement
(oracle/mnv4_oracle.py).  timm is not available offline: the restatement follows the published architecture, parity with timm itself
is UNPINNED -- what these tests pin is that the CUDA path computes the restated network, layer names and shapes included."""
import random

import pytest
import torch

from oracle import pr_oracle as O
from oracle.mnv4_oracle import MobileNetV4ConvSmallFeatures, init_healthy

DEV = "cuda:0"
ENC = "refiner_fine_branch.refiner_encoder."


def test_state_dict_layout_matches_the_restated_timm_module():
    from patchrefinerv2_b200.mnv4 import OUT_CHANNELS, mnv4_conv_small_spec
    for in_chans in (3, 4):
        sd = MobileNetV4ConvSmallFeatures(in_chans).state_dict()
        spec = mnv4_conv_small_spec(in_chans)
        assert list(sd.keys()) == list(spec.keys())
        assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in sd)
    with torch.no_grad():
        feats = MobileNetV4ConvSmallFeatures(4).eval()(torch.rand(1, 4, 64, 96))
    assert [f.shape[1] for f in feats] == list(OUT_CHANNELS) and [f.shape[2] for f in feats] == [32, 16, 8, 4, 2]


def test_plus_model_builds_the_native_encoder_without_timm(monkeypatch):
    import sys
    monkeypatch.setitem(sys.modules, "timm", None)
    from patchrefinerv2_b200 import build_model
    cfg = O.make_plus_config()
    assert cfg["refiner"]["fine_branch"]["encoder_name"].startswith("mobilenetv4_conv_small")
    m = build_model(dict(type="PatchRefinerPlus", config=cfg))                    # no fine_encoder, no timm
    enc_sd = {ENC + k: v for k, v in init_healthy(MobileNetV4ConvSmallFeatures(4), 3).state_dict().items()}
    keys = set(m.state_dict().keys())
    assert set(enc_sd) <= keys
    sd = {k: v for k, v in O.init_patchrefinerplus_state_dict(cfg, 0).items() if not k.startswith(ENC)}
    sd.update(enc_sd)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    got = m.state_dict()
    assert all(torch.equal(got[k], v.float()) for k, v in enc_sd.items())
    other = dict(cfg, refiner=dict(cfg["refiner"], fine_branch=dict(cfg["refiner"]["fine_branch"], encoder_name="tf_efficientnet_b5_ap")))
    with pytest.raises(NotImplementedError):                                      # other timm encoders: loud, with the way out
        build_model(dict(type="PatchRefinerPlus", config=other))


@pytest.mark.gpu
@pytest.mark.parametrize("k,stride,relu", [(3, 1, True), (3, 2, False), (5, 1, False), (5, 2, True)])
@pytest.mark.parametrize("x3,tol", [(False, 1e-2), (True, 2e-5)])
def test_dwconv_vs_torch(x3, tol, k, stride, relu):
    from patchrefinerv2_b200.mnv4 import _DwLayer
    from patchrefinerv2_b200.nn import Act
    g = torch.Generator().manual_seed(k * 10 + stride)
    N, Cc, H, W = 2, 96, 29, 38
    x = torch.randn(N, Cc, H, W, generator=g)
    w = torch.randn(Cc, 1, k, k, generator=g) / k
    b = torch.randn(Cc, generator=g) * 0.1
    want = torch.nn.functional.conv2d(x, w, b, stride=stride, padding=k // 2, groups=Cc)
    want = torch.relu(want) if relu else want
    a = Act.from_nchw(x.to(DEV), x3)
    out = Act.empty(N, want.shape[2], want.shape[3], Cc, x3, DEV)
    _DwLayer(w[:, 0], b, k, stride, relu, DEV)(a, out)
    got = out.to_nchw().cpu()
    assert got.shape == want.shape
    assert ((got - want).abs().max() / want.abs().max()).item() < tol


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp32", 1e-3), ("bf16", 5e-2)])
def test_encoder_features_vs_restatement(prec, tol):
    from patchrefinerv2_b200.mnv4 import MobileNetV4ConvSmallB200
    from patchrefinerv2_b200.nn import Workspace
    m = init_healthy(MobileNetV4ConvSmallFeatures(4), 5)
    g = torch.Generator().manual_seed(7)
    crops = torch.rand(2, 3, 224, 224, generator=g)
    depth = torch.rand(2, 1, 224, 224, generator=g) * 10
    mean = torch.tensor(m.default_cfg["mean"]).view(1, 3, 1, 1)
    std = torch.tensor(m.default_cfg["std"]).view(1, 3, 1, 1)
    with torch.no_grad():
        want = m(torch.cat([(crops - mean) / std, depth], dim=1))
    x3 = prec == "fp32"
    enc = MobileNetV4ConvSmallB200({ENC + k: v for k, v in m.state_dict().items()}, ENC, 4, x3, DEV)
    feats = enc.forward(crops.to(DEV), depth.to(DEV), Workspace(DEV, x3))
    assert len(feats) == 5
    for f, w in zip(feats, want):
        got = f.to_nchw().cpu()
        assert got.shape == w.shape
        assert ((got - w).abs().max() / w.abs().max()).item() < tol


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp32", 1e-3), ("bf16", 5e-2)])
def test_plus_with_native_encoder_vs_oracle(prec, tol):
    """PatchRefinerPlus end to end (r2: regular + random patches) with the MobileNetV4 encoder on the kernels, against the oracle's
    PatchRefinerPlus loop (bit-identical to the reference's, tests/test_plus.py) running the restated encoder in PyTorch."""
    from patchrefinerv2_b200 import build_model
    cfg = O.make_plus_config()
    sd = {k: v for k, v in O.init_patchrefinerplus_state_dict(cfg, 0).items() if not k.startswith(ENC)}
    sd.update({ENC + k: v for k, v in init_healthy(MobileNetV4ConvSmallFeatures(4), 3).state_dict().items()})
    lr, hr = O.synthetic_frame(cfg, 1)
    random.seed(1)
    want, coarse, _ = O.PatchRefinerPlusOracle(cfg, sd, MobileNetV4ConvSmallFeatures(4)).infer(lr, hr, None, "r2", 2)
    m = build_model(dict(type="PatchRefinerPlus", config=cfg, precision=prec, patch_batch=3))
    m.load_dict(sd)
    m = m.cuda().eval()
    random.seed(1)
    depth, log = m(mode="infer", image_lr=lr.cuda(), image_hr=hr.cuda(), cai_mode="r2", process_num=2)
    assert depth.shape == want.shape
    if prec == "fp32":
        assert ((depth - want).abs() / want.abs().clamp_min(1e-2)).max().item() < tol
    else:
        assert ((depth - want).abs().max() / want.abs().max()).item() < tol

"""CPU oracle: a plain PyTorch/NumPy restatement of PatchRefinerV2's tiled high-resolution
inference path (``mode='infer'``, CAI modes m1 / m2 / rN) for the DAv2 + FusionUnet family
(``PatchRefiner``) and the V2 family (``PatchRefinerPlus`` = DAv2 coarse branch + LightWeightRefiner +
``BiDirectionalFusion``).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product package
``patchrefinerv2_b200`` never does (the product fails loudly if its CUDA library is missing).

Parity status: PINNED.  ``tests/test_oracle_vs_reference.py``, ``tests/test_bifusion.py`` and
``tests/test_plus.py`` run the reference's own classes (imported read-only from /root/reference
through ``oracle/ref_shim.py``) against this file on identical weights/frames (bit-identical), and
``tests/golden/*.npz`` (made by ``oracle/make_golden.py`` from the reference itself) are checked on
every run, with or without /root/reference.  ONE piece is UNPINNED: the timm CNN inside
``LightWeightRefiner`` (timm is not installable offline; pinned ``timm==0.9.2`` in the reference's
environment.yml:27).  ``ToyFineEncoder`` stands in for it on both sides -- injected into the
reference as ``timm.create_model``'s result -- so everything around the encoder is still pinned.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Third-party arithmetic the reference itself calls is called here too, not restated:
``torch.nn.functional`` (conv2d / linear / interpolate / layer_norm / gelu),
``torchvision.ops.roi_align`` and ``cv2.GaussianBlur``.  The explicit index-math restatements
at the bottom (``np_*``) are the bit-exact contracts for the CUDA gather / blend kernels.
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional, Sequence, Tuple

import cv2
import numpy as np
import torch
import torch.nn.functional as F
from torchvision.ops import roi_align as tv_roi_align

Tensor = torch.Tensor

# ---------------------------------------------------------------------------------------------
# model hyper-parameters
# ---------------------------------------------------------------------------------------------

#: external/depth_anything_v2/dinov2.py:340-390 (embed_dim, depth, heads) and
#: external/depth_anything_v2/dpt.py:165-170 (intermediate layer indices)
VIT_CFG = {
    "vits": dict(dim=384, depth=12, heads=6, taps=(2, 5, 8, 11)),
    "vitb": dict(dim=768, depth=12, heads=12, taps=(2, 5, 8, 11)),
    "vitl": dict(dim=1024, depth=24, heads=16, taps=(4, 11, 17, 23)),
}
PATCH = 14                    # dinov2.py:405 (patch_size=14)
POS_GRID = 37                 # dinov2.py:404 (img_size=518 -> 37x37 position table)
INTERP_OFFSET = 0.1           # dinov2.py:410
LN_EPS = 1e-6                 # dinov2.py:95 ; estimator/models/blocks/convs.py:11
PIXEL_MEAN = (0.485, 0.456, 0.406)   # external/depth_anything_v2/dpt.py:179
PIXEL_STD = (0.229, 0.224, 0.225)    # external/depth_anything_v2/dpt.py:180


# ---------------------------------------------------------------------------------------------
# deterministic random-init weights (state-dict keys == the reference's)
# ---------------------------------------------------------------------------------------------

def _conv_w(g, co, ci, k, gain=1.0):
    bound = gain / math.sqrt(ci * k * k)
    return (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound


def _vec(g, n, scale):
    return (torch.rand(n, generator=g) * 2 - 1) * scale


def init_dav2_state_dict(encoder: str, features: int, out_channels: Sequence[int], seed: int) -> Dict[str, Tensor]:
    """Random DepthAnythingV2 weights with the reference's key names/shapes
    (external/depth_anything_v2/dpt.py:153-181, dinov2.py:44-170, util/blocks.py:4-146).
    Distributions are chosen for test signal (non-trivial biases / LayerScale), not to mimic
    the reference's init; both sides always load the same tensors."""
    g = torch.Generator().manual_seed(seed)
    c = VIT_CFG[encoder]
    D, depth = c["dim"], c["depth"]
    sd: Dict[str, Tensor] = {}
    p = "pretrained."
    sd[p + "cls_token"] = torch.randn(1, 1, D, generator=g) * 0.02
    sd[p + "pos_embed"] = torch.randn(1, POS_GRID * POS_GRID + 1, D, generator=g) * 0.02
    sd[p + "mask_token"] = torch.zeros(1, D)
    sd[p + "patch_embed.proj.weight"] = _conv_w(g, D, 3, PATCH)
    sd[p + "patch_embed.proj.bias"] = _vec(g, D, 0.05)
    for i in range(depth):
        b = f"{p}blocks.{i}."
        sd[b + "norm1.weight"] = 1.0 + _vec(g, D, 0.1)
        sd[b + "norm1.bias"] = _vec(g, D, 0.05)
        sd[b + "attn.qkv.weight"] = torch.randn(3 * D, D, generator=g) * (1.0 / math.sqrt(D))
        sd[b + "attn.qkv.bias"] = _vec(g, 3 * D, 0.05)
        sd[b + "attn.proj.weight"] = torch.randn(D, D, generator=g) * (1.0 / math.sqrt(D))
        sd[b + "attn.proj.bias"] = _vec(g, D, 0.05)
        sd[b + "ls1.gamma"] = 0.2 + _vec(g, D, 0.1)
        sd[b + "norm2.weight"] = 1.0 + _vec(g, D, 0.1)
        sd[b + "norm2.bias"] = _vec(g, D, 0.05)
        sd[b + "mlp.fc1.weight"] = torch.randn(4 * D, D, generator=g) * (1.0 / math.sqrt(D))
        sd[b + "mlp.fc1.bias"] = _vec(g, 4 * D, 0.05)
        sd[b + "mlp.fc2.weight"] = torch.randn(D, 4 * D, generator=g) * (1.0 / math.sqrt(4 * D))
        sd[b + "mlp.fc2.bias"] = _vec(g, D, 0.05)
        sd[b + "ls2.gamma"] = 0.2 + _vec(g, D, 0.1)
    sd[p + "norm.weight"] = 1.0 + _vec(g, D, 0.1)
    sd[p + "norm.bias"] = _vec(g, D, 0.05)

    h = "depth_head."
    oc = list(out_channels)
    for i, o in enumerate(oc):
        sd[f"{h}projects.{i}.weight"] = _conv_w(g, o, D, 1)
        sd[f"{h}projects.{i}.bias"] = _vec(g, o, 0.05)
    # ConvTranspose2d weights are [Cin, Cout, k, k]  (dpt.py:62-73)
    sd[h + "resize_layers.0.weight"] = (torch.rand(oc[0], oc[0], 4, 4, generator=g) * 2 - 1) / math.sqrt(oc[0])
    sd[h + "resize_layers.0.bias"] = _vec(g, oc[0], 0.05)
    sd[h + "resize_layers.1.weight"] = (torch.rand(oc[1], oc[1], 2, 2, generator=g) * 2 - 1) / math.sqrt(oc[1])
    sd[h + "resize_layers.1.bias"] = _vec(g, oc[1], 0.05)
    sd[h + "resize_layers.3.weight"] = _conv_w(g, oc[3], oc[3], 3)
    sd[h + "resize_layers.3.bias"] = _vec(g, oc[3], 0.05)
    for i, o in enumerate(oc):
        sd[f"{h}scratch.layer{i + 1}_rn.weight"] = _conv_w(g, features, o, 3, gain=1.7)
    for r in (1, 2, 3, 4):
        q = f"{h}scratch.refinenet{r}."
        sd[q + "out_conv.weight"] = _conv_w(g, features, features, 1, gain=1.7)
        sd[q + "out_conv.bias"] = _vec(g, features, 0.05)
        for u in (1, 2):
            for cv in (1, 2):
                sd[f"{q}resConfUnit{u}.conv{cv}.weight"] = _conv_w(g, features, features, 3, gain=1.7)
                sd[f"{q}resConfUnit{u}.conv{cv}.bias"] = _vec(g, features, 0.05)
    sd[h + "scratch.output_conv1.weight"] = _conv_w(g, features // 2, features, 3, gain=1.7)
    sd[h + "scratch.output_conv1.bias"] = _vec(g, features // 2, 0.05)
    sd[h + "scratch.output_conv2.0.weight"] = _conv_w(g, 32, features // 2, 3, gain=1.7)
    sd[h + "scratch.output_conv2.0.bias"] = _vec(g, 32, 0.05)
    sd[h + "scratch.output_conv2.2.weight"] = _conv_w(g, 1, 32, 1, gain=3.0)
    sd[h + "scratch.output_conv2.2.bias"] = _vec(g, 1, 0.05)
    return sd


def init_fusion_unet_state_dict(input_chl, temp_chl, dec_chl, seed: int) -> Dict[str, Tensor]:
    """Random FusionUnet weights (estimator/models/blocks/fusion_model.py:52-82; convs.py:5-75)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for idx, (ic, tc) in enumerate(zip(input_chl, temp_chl)):
        sd[f"encoder_layers_1.{idx}.single_conv.0.weight"] = _conv_w(g, tc, ic, 3, gain=1.7)
        sd[f"encoder_layers_1.{idx}.single_conv.1.weight"] = 1.0 + _vec(g, tc, 0.1)
        sd[f"encoder_layers_1.{idx}.single_conv.1.bias"] = _vec(g, tc, 0.05)
        sd[f"encoder_layers_2.{idx}.single_conv.0.weight"] = _conv_w(g, tc, tc + 2, 3, gain=1.7)
        sd[f"encoder_layers_2.{idx}.single_conv.1.weight"] = 1.0 + _vec(g, tc, 0.1)
        sd[f"encoder_layers_2.{idx}.single_conv.1.bias"] = _vec(g, tc, 0.05)
    rev = list(temp_chl)[::-1]
    _chl = rev[0]
    for i, (tc, dc) in enumerate(zip(rev[1:], dec_chl)):
        cin = tc + _chl + 2
        sd[f"decoder_layers.{i}.conv.double_conv.0.weight"] = _conv_w(g, cin, cin, 3, gain=1.7)
        sd[f"decoder_layers.{i}.conv.double_conv.2.weight"] = _conv_w(g, dc, cin, 3, gain=1.7)
        _chl = dc
    last = dec_chl[-1] if len(dec_chl) else _chl
    sd["final_conv.weight"] = _conv_w(g, 1, last, 3, gain=6.0)
    return sd


def init_patchrefiner_state_dict(cfg: dict, seed: int = 0) -> Dict[str, Tensor]:
    """Full PatchRefiner state dict with the reference's prefixes
    (estimator/models/patchrefiner.py:94,118,135: coarse_branch. / refiner_fine_branch. /
    refiner_fusion_model.)."""
    cb = cfg["coarse_branch"]["model_cfg"]
    fb = cfg["refiner"]["fine_branch"]["model_cfg"]
    fu = cfg["refiner"]["fusion_model"]
    sd: Dict[str, Tensor] = {}
    for k, v in init_dav2_state_dict(cb["encoder"], cb["features"], cb["out_channels"], seed * 3 + 11).items():
        sd["coarse_branch." + k] = v
    for k, v in init_dav2_state_dict(fb["encoder"], fb["features"], fb["out_channels"], seed * 3 + 12).items():
        sd["refiner_fine_branch." + k] = v
    for k, v in init_fusion_unet_state_dict(fu["input_chl"], fu["temp_chl"], fu["dec_chl"], seed * 3 + 13).items():
        sd["refiner_fusion_model." + k] = v
    return sd


def make_config(encoder="vits", patch_process_shape=(448, 448), image_raw_shape=(2160, 3840),
                patch_split_num=(4, 4), max_depth=80.0) -> dict:
    """A config dict shaped like configs/patchrefiner_dav2/pr_u4k.py:10-53 for either encoder size."""
    if encoder == "vitl":
        feats, oc = 256, [256, 512, 1024, 1024]
    elif encoder == "vitb":
        feats, oc = 128, [96, 192, 384, 768]
    else:
        feats, oc = 64, [48, 96, 192, 384]
    half = feats // 2
    return dict(
        image_raw_shape=list(image_raw_shape), patch_process_shape=list(patch_process_shape),
        patch_split_num=list(patch_split_num), fusion_feat_level=6, min_depth=1e-3, max_depth=max_depth,
        strategy_refiner_target="offset_coarse",
        coarse_branch=dict(type="DA2", pretrained=None, model_cfg=dict(encoder=encoder, features=feats, out_channels=oc)),
        refiner=dict(
            fine_branch=dict(type="DA2", pretrained=None, model_cfg=dict(encoder=encoder, features=feats, out_channels=oc)),
            fusion_model=dict(type="FusionUnet", input_chl=[half * 2] + [feats * 2] * 5,
                              temp_chl=[half] + [feats] * 5, dec_chl=[feats] * 4 + [half])),
        sigloss=dict(type="SILogLoss"), pretrained=None, pre_norm_bbox=True,
        pretrain_coarse_model=None, pretrain_fine_model=None)


# ---------------------------------------------------------------------------------------------
# DINOv2 ViT  (external/depth_anything_v2/dinov2.py, dinov2_layers/*)
# ---------------------------------------------------------------------------------------------

def interpolate_pos_embed(pos_embed: Tensor, h: int, w: int) -> Tensor:
    """dinov2.py:179-210.  ``pos_embed`` [1, 37*37+1, D] -> [1, (h//14)*(w//14)+1, D].  The
    reference passes (w, h) = x.shape[-2:] *swapped in name only*; we keep its arithmetic:
    first spatial axis gets scale (H//14+0.1)/37, second (W//14+0.1)/37."""
    N = pos_embed.shape[1] - 1
    n0, n1 = h // PATCH, w // PATCH
    if n0 * n1 == N and h == w:
        return pos_embed
    cls_pos = pos_embed[:, 0]
    patch_pos = pos_embed[:, 1:]
    dim = pos_embed.shape[-1]
    sqrt_n = math.sqrt(N)
    s0, s1 = float(n0 + INTERP_OFFSET) / sqrt_n, float(n1 + INTERP_OFFSET) / sqrt_n
    patch_pos = F.interpolate(
        patch_pos.reshape(1, int(sqrt_n), int(sqrt_n), dim).permute(0, 3, 1, 2),
        scale_factor=(s0, s1), mode="bicubic", antialias=False)
    assert patch_pos.shape[-2] == n0 and patch_pos.shape[-1] == n1
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((cls_pos.unsqueeze(0), patch_pos), dim=1)


def vit_intermediate(sd: Dict[str, Tensor], pre: str, x: Tensor, encoder: str,
                     trace: Optional[dict] = None) -> List[Tensor]:
    """DinoVisionTransformer.get_intermediate_layers(x, taps, return_class_token=True) with the
    class token dropped (dinov2.py:212-231, 269-321; block.py:82-107; attention.py:49-62;
    mlp.py:35-41; layer_scale.py:27-28).  x: normalised image [B,3,H,W].  Returns 4 tensors
    [B, (H//14)*(W//14), D] (final LayerNorm applied, dinov2.py:309-310)."""
    c = VIT_CFG[encoder]
    D, heads = c["dim"], c["heads"]
    B, _, H, W = x.shape
    t = F.conv2d(x, sd[pre + "patch_embed.proj.weight"], sd[pre + "patch_embed.proj.bias"], stride=PATCH)
    t = t.flatten(2).transpose(1, 2)                                   # patch_embed.py:76-78
    t = torch.cat((sd[pre + "cls_token"].expand(B, -1, -1), t), dim=1)  # dinov2.py:218
    t = t + interpolate_pos_embed(sd[pre + "pos_embed"].float(), H, W)  # dinov2.py:219
    if trace is not None:
        trace["tokens0"] = t.clone()
    outs = []
    N = t.shape[1]
    hd = D // heads
    scale = hd ** -0.5
    for i in range(c["depth"]):
        b = f"{pre}blocks.{i}."
        y = F.layer_norm(t, (D,), sd[b + "norm1.weight"], sd[b + "norm1.bias"], LN_EPS)
        qkv = F.linear(y, sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"])
        qkv = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * scale, qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        y = (attn @ v).transpose(1, 2).reshape(B, N, D)
        y = F.linear(y, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])
        t = t + y * sd[b + "ls1.gamma"]
        y = F.layer_norm(t, (D,), sd[b + "norm2.weight"], sd[b + "norm2.bias"], LN_EPS)
        y = F.gelu(F.linear(y, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"]))
        y = F.linear(y, sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])
        t = t + y * sd[b + "ls2.gamma"]
        if trace is not None and i == 0:
            trace["block0"] = t.clone()
        if i in c["taps"]:
            outs.append(F.layer_norm(t, (D,), sd[pre + "norm.weight"], sd[pre + "norm.bias"], LN_EPS)[:, 1:])
    return outs


# ---------------------------------------------------------------------------------------------
# DPT head  (external/depth_anything_v2/dpt.py:38-150 ; util/blocks.py:29-148)
# ---------------------------------------------------------------------------------------------

def _rcu(sd, pre, x):
    """ResidualConvUnit.forward (util/blocks.py:57-80): relu -> conv1 -> relu -> conv2 -> + x."""
    out = F.relu(x)
    out = F.conv2d(out, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    out = F.relu(out)
    out = F.conv2d(out, sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    return out + x


def _ffb(sd, pre, x0, x1, size):
    """FeatureFusionBlock.forward (util/blocks.py:123-148)."""
    out = x0
    if x1 is not None:
        out = out + _rcu(sd, pre + "resConfUnit1.", x1)
    out = _rcu(sd, pre + "resConfUnit2.", out)
    if size is None:
        out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    else:
        out = F.interpolate(out, size=size, mode="bilinear", align_corners=True)
    return F.conv2d(out, sd[pre + "out_conv.weight"], sd[pre + "out_conv.bias"])


def dpt_head(sd: Dict[str, Tensor], pre: str, taps: List[Tensor], ph: int, pw: int):
    """DPTHead.forward (dpt.py:116-150), use_clstoken=False.  Returns (sigmoid map, 6 features)."""
    outs = []
    for i, x in enumerate(taps):
        B, _, D = x.shape
        x = x.permute(0, 2, 1).reshape(B, D, ph, pw)
        x = F.conv2d(x, sd[f"{pre}projects.{i}.weight"], sd[f"{pre}projects.{i}.bias"])
        if i == 0:
            x = F.conv_transpose2d(x, sd[pre + "resize_layers.0.weight"], sd[pre + "resize_layers.0.bias"], stride=4)
        elif i == 1:
            x = F.conv_transpose2d(x, sd[pre + "resize_layers.1.weight"], sd[pre + "resize_layers.1.bias"], stride=2)
        elif i == 3:
            x = F.conv2d(x, sd[pre + "resize_layers.3.weight"], sd[pre + "resize_layers.3.bias"], stride=2, padding=1)
        outs.append(x)
    l1, l2, l3, l4 = [F.conv2d(o, sd[f"{pre}scratch.layer{i + 1}_rn.weight"], padding=1) for i, o in enumerate(outs)]
    s = pre + "scratch."
    path_4 = _ffb(sd, s + "refinenet4.", l4, None, l3.shape[2:])
    path_3 = _ffb(sd, s + "refinenet3.", path_4, l3, l2.shape[2:])
    path_2 = _ffb(sd, s + "refinenet2.", path_3, l2, l1.shape[2:])
    path_1 = _ffb(sd, s + "refinenet1.", path_2, l1, None)
    out = F.conv2d(path_1, sd[s + "output_conv1.weight"], sd[s + "output_conv1.bias"], padding=1)
    out_feat = F.interpolate(out, (int(ph * PATCH), int(pw * PATCH)), mode="bilinear", align_corners=True)
    out = F.conv2d(out_feat, sd[s + "output_conv2.0.weight"], sd[s + "output_conv2.0.bias"], padding=1)
    out = F.relu(out)
    out = F.conv2d(out, sd[s + "output_conv2.2.weight"], sd[s + "output_conv2.2.bias"])
    out = torch.sigmoid(out)
    return out, [l4, path_4, path_3, path_2, path_1, out_feat]


def depth_anything_v2(sd: Dict[str, Tensor], pre: str, x: Tensor, encoder: str, max_depth: float,
                      trace: Optional[dict] = None):
    """DepthAnythingV2.forward (dpt.py:182-203).  x: RGB in [0,1], [B,3,H,W].
    Returns (metric_depth [B,1,H,W], [x_d0, x_blocks_feat_0..3, midas_final_feat])."""
    mean = torch.tensor(PIXEL_MEAN, device=x.device).view(-1, 1, 1)      # device-agnostic: bench.py also runs this port on the GPU as the eager baseline
    std = torch.tensor(PIXEL_STD, device=x.device).view(-1, 1, 1)
    x = (x - mean) / std
    ph, pw = x.shape[-2] // PATCH, x.shape[-1] // PATCH
    taps = vit_intermediate(sd, pre + "pretrained.", x, encoder, trace)
    if trace is not None:
        trace["taps"] = [t.clone() for t in taps]
    depth, feats = dpt_head(sd, pre + "depth_head.", taps, ph, pw)
    return depth * max_depth, feats


# ---------------------------------------------------------------------------------------------
# FusionUnet  (estimator/models/blocks/fusion_model.py:84-122 ; convs.py)
# ---------------------------------------------------------------------------------------------

def _ln_cf(x, w, b):
    """channels-first LayerNorm (convs.py:21-29)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + LN_EPS)
    return w[:, None, None] * x + b[:, None, None]


def _single_conv_ln(sd, pre, x):
    """SingleConvCNNLN (convs.py:64-75): conv3x3(no bias) -> LN over C -> GELU."""
    x = F.conv2d(x, sd[pre + "single_conv.0.weight"], padding=1)
    return F.gelu(_ln_cf(x, sd[pre + "single_conv.1.weight"], sd[pre + "single_conv.1.bias"]))


def _bil(x, size):
    return F.interpolate(x, size, mode="bilinear", align_corners=True)


def fusion_unet(sd: Dict[str, Tensor], pre: str, c_feat: List[Tensor], f_feat: List[Tensor],
                pred1: Tensor, pred2: Tensor, update_base: Optional[Tensor], trace: Optional[dict] = None) -> Tensor:
    """FusionUnet.forward (fusion_model.py:84-122) with UpSample.forward_hardcode (:15-24).
    c_feat / f_feat are ordered finest-first (patchrefiner.py:250-251 reverses them)."""
    temp = []
    for idx, (c, f) in enumerate(zip(c_feat, f_feat)):
        f = _single_conv_ln(sd, f"{pre}encoder_layers_1.{idx}.", torch.cat([c, f], dim=1))
        p1 = _bil(pred1, f.shape[-2:])
        p2 = _bil(pred2, f.shape[-2:])
        f = _single_conv_ln(sd, f"{pre}encoder_layers_2.{idx}.", torch.cat([f, p1, p2], dim=1))
        temp.append(f)
    if trace is not None:
        trace["fusion_enc"] = [t.clone() for t in temp]
    dec = temp[0]
    temp = temp[::-1]
    _feat = temp[0]
    for i, feat in enumerate(temp[1:]):
        x1 = _bil(_feat, feat.shape[-2:])
        p1 = _bil(pred1, feat.shape[-2:])
        p2 = _bil(pred2, feat.shape[-2:])
        x = torch.cat([x1, feat, p1, p2], dim=1)
        x = F.gelu(F.conv2d(x, sd[f"{pre}decoder_layers.{i}.conv.double_conv.0.weight"], padding=1))
        x = F.gelu(F.conv2d(x, sd[f"{pre}decoder_layers.{i}.conv.double_conv.2.weight"], padding=1))
        dec = x
        _feat = x
    if trace is not None:
        trace["fusion_dec"] = dec.clone()
    off = F.conv2d(dec, sd[pre + "final_conv.weight"], padding=1)
    if trace is not None:
        trace["offset"] = off.clone()
    if update_base is not None:
        return torch.clamp(update_base + off, min=0)
    return off


# ---------------------------------------------------------------------------------------------
# BiDirectionalFusion -- the V2 ("PatchRefinerPlus") fusion model
# (estimator/models/blocks/bi_directional_fusion_model.py ; SURVEY.md row a9')
# ---------------------------------------------------------------------------------------------

#: coarse2fine_type -> (C2FModule fusion, gate) (bi_directional_fusion_model.py:366-374)
C2F_TYPES = {"self-agg": (False, False), "coarse-gated": (True, True), "coarse-fusion": (True, False),
             "only-gate": (True, False)}     # only-gate = C2FNOENCModule(fusion=True, gate=False) (:372-373): oracle only, the product does not build it
C2F_FEATURES = 256            # C2FModule default `features` (:149); no config overrides it


def init_bidirectional_fusion_state_dict(coarse_chl, fine_chl, fine_chl_after_coarse2fine, temp_chl, dec_chl, seed: int,
                                         coarse2fine_type: str = "coarse-gated", features: int = C2F_FEATURES, heavy: bool = False) -> Dict[str, Tensor]:
    """Random BiDirectionalFusion weights with the reference's state-dict keys (glb_att=False, coarse2fine=True;
    bi_directional_fusion_model.py:285-374, C2FModule :148-184, C2FNOENCModule :211-251, GatedFusionBlock :84-116, GatedConvUnit
    :24-54).  ``heavy``: BiDirectionalFusionHeavy (:517-561) -- SingleConvCNNLNHeavy (:448-463) and DoubleConvHeavy (:465-485)."""
    fusion, _gate = C2F_TYPES[coarse2fine_type]
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def single(pre, tc, cin):
        sd[pre + "single_conv.0.weight"] = _conv_w(g, tc, cin, 3, gain=1.7)
        sd[pre + "single_conv.1.weight"] = 1.0 + _vec(g, tc, 0.1)
        sd[pre + "single_conv.1.bias"] = _vec(g, tc, 0.05)
        if heavy:
            sd[pre + "single_conv.2.weight"] = _conv_w(g, tc, tc, 3, gain=1.7)
            sd[pre + "single_conv.3.weight"] = 1.0 + _vec(g, tc, 0.1)
            sd[pre + "single_conv.3.bias"] = _vec(g, tc, 0.05)
            sd[pre + "single_conv.4.weight"] = _conv_w(g, tc, tc, 3, gain=1.7)

    for idx, (cc, fc, tc) in enumerate(zip(coarse_chl, fine_chl_after_coarse2fine, temp_chl)):      # (draw order is part of the golden fixtures)
        single(f"fusion_layers_1.{idx}.", tc, cc + fc)
        single(f"fusion_layers_2.{idx}.", tc, tc + 2)
    rev = list(temp_chl)[::-1]
    _chl = rev[0]
    for i, (tc, dc) in enumerate(zip(rev[1:], dec_chl)):
        cin = tc + _chl + 2
        sd[f"f2r_agg.{i}.conv.double_conv.0.weight"] = _conv_w(g, cin, cin, 3, gain=1.7)
        if heavy:
            for k in (2, 4, 6):
                sd[f"f2r_agg.{i}.conv.double_conv.{k}.weight"] = _conv_w(g, cin, cin, 3, gain=1.7)
            sd[f"f2r_agg.{i}.conv.double_conv.8.weight"] = _conv_w(g, dc, cin, 3, gain=1.7)
        else:
            sd[f"f2r_agg.{i}.conv.double_conv.2.weight"] = _conv_w(g, dc, cin, 3, gain=1.7)
        _chl = dc
    sd["final_conv.weight"] = _conv_w(g, 1, dec_chl[-1] if len(dec_chl) else _chl, 3, gain=6.0)

    def unit(pre, feat):
        sd[pre + "conv.weight"] = _conv_w(g, feat, feat, 3)
        sd[pre + "conv.bias"] = _vec(g, feat, 0.05)
        if fusion:
            sd[pre + "fusion_conv.0.weight"] = _conv_w(g, feat, 2 * feat, 3, gain=1.7)
            sd[pre + "fusion_conv.0.bias"] = _vec(g, feat, 0.05)
            sd[pre + "fusion_conv.1.weight"] = 1.0 + _vec(g, feat, 0.1)
            sd[pre + "fusion_conv.1.bias"] = _vec(g, feat, 0.05)
            sd[pre + "fusion_conv.3.weight"] = _conv_w(g, feat, feat, 1, gain=2.0)

    def block(pre, feat):
        sd[pre + "out_conv.weight"] = _conv_w(g, feat, feat, 1)
        sd[pre + "out_conv.bias"] = _vec(g, feat, 0.05)
        unit(pre + "GateresConfUnit1.", feat)
        unit(pre + "GateresConfUnit2.", feat)

    s = "c2f.scratch."
    for i, fc in enumerate(fine_chl):
        sd[f"{s}layer{i + 1}_rn.weight"] = _conv_w(g, features, fc, 3)
    if coarse2fine_type == "only-gate":                                  # C2FNOENCModule (:211-251)
        for lvl in range(1, 6):
            unit(f"{s}layer{lvl}_gate1.", features)
            unit(f"{s}layer{lvl}_gate2.", features)
        sd[s + "upsample_conv.0.weight"] = (torch.rand(fine_chl[0], 32, 2, 2, generator=g) * 2 - 1) / math.sqrt(fine_chl[0])   # ConvTranspose2d [Cin, Cout, k, k]
        sd[s + "upsample_conv.0.bias"] = _vec(g, 32, 0.05)
        sd[s + "upsample_conv.2.weight"] = _conv_w(g, 32, 32, 3, gain=1.7)
        unit(s + "layer6_gate1.", 32)
        unit(s + "layer6_gate2.", 32)
        sd[s + "output_conv.weight"] = _conv_w(g, 1, 32, 3, gain=3.0)
        sd[s + "output_conv.bias"] = _vec(g, 1, 0.05)
        return sd
    for i in range(1, 6):
        block(f"{s}refinenet{i}.", features)
    h2 = coarse_chl[0]
    sd[s + "output_conv1.weight"] = _conv_w(g, features // 2, features, 3)
    sd[s + "output_conv1.bias"] = _vec(g, features // 2, 0.05)
    sd[s + "output_conv2.0.weight"] = _conv_w(g, h2, features // 2, 3, gain=1.7)
    sd[s + "output_conv2.0.bias"] = _vec(g, h2, 0.05)
    block(s + "output_conv2_fusion.", h2)
    sd[s + "output_conv3.0.weight"] = 1.0 + _conv_w(g, 1, h2, 1)       # nn.init.normal_(mean=1.0) (:181)
    sd[s + "output_conv3.0.bias"] = torch.zeros(1)
    return sd


def _gated_conv_unit(sd, pre, x, c_feat, fusion: bool, gate: bool):
    """GatedConvUnit.forward (bi_directional_fusion_model.py:56-82)."""
    out = F.relu(x)
    out = F.conv2d(out, sd[pre + "conv.weight"], sd[pre + "conv.bias"], padding=1)
    out = out + x
    if fusion:
        fused = torch.cat([out, c_feat], dim=1)
        fused = F.conv2d(fused, sd[pre + "fusion_conv.0.weight"], sd[pre + "fusion_conv.0.bias"], padding=1)
        fused = F.relu(_ln_cf(fused, sd[pre + "fusion_conv.1.weight"], sd[pre + "fusion_conv.1.bias"]))
        fused = F.conv2d(fused, sd[pre + "fusion_conv.3.weight"])
        out = out * torch.sigmoid(fused) if gate else fused
    return out


def _gated_fusion_block(sd, pre, xs, size, coarse_feat, fusion: bool, gate: bool, upscale: bool = True):
    """GatedFusionBlock.forward (bi_directional_fusion_model.py:116-146)."""
    out = xs[0]
    if len(xs) == 2:
        out = out + _gated_conv_unit(sd, pre + "GateresConfUnit1.", xs[1], coarse_feat, fusion, gate)
    out = _gated_conv_unit(sd, pre + "GateresConfUnit2.", out, coarse_feat, fusion, gate)
    if upscale:
        if size is None:
            out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
        else:
            out = F.interpolate(out, size=tuple(size), mode="bilinear", align_corners=True)
    return F.conv2d(out, sd[pre + "out_conv.weight"], sd[pre + "out_conv.bias"])


def c2f_module(sd, pre, fine_features: List[Tensor], coarse_features: List[Tensor], fusion: bool, gate: bool):
    """C2FModule.forward (bi_directional_fusion_model.py:184-208).  fine_features: 5 maps, finest first;
    coarse_features: 6 maps, finest first.  Returns ([layer_5_rn, path_5, path_4, path_3, path_2, last_feat], depth)."""
    s = pre + "scratch."
    rn = [F.conv2d(f, sd[f"{s}layer{i + 1}_rn.weight"], padding=1) for i, f in enumerate(fine_features)]
    fb = lambda i, xs, size, cf: _gated_fusion_block(sd, f"{s}refinenet{i}.", xs, size, cf, fusion, gate)
    path_5 = fb(5, [rn[4]], rn[3].shape[2:], coarse_features[5])
    path_4 = fb(4, [path_5, rn[3]], rn[2].shape[2:], coarse_features[4])
    path_3 = fb(3, [path_4, rn[2]], rn[1].shape[2:], coarse_features[3])
    path_2 = fb(2, [path_3, rn[1]], rn[0].shape[2:], coarse_features[2])
    path_1 = fb(1, [path_2, rn[0]], None, coarse_features[1])
    out = F.conv2d(path_1, sd[s + "output_conv1.weight"], sd[s + "output_conv1.bias"], padding=1)
    last = F.relu(F.conv2d(out, sd[s + "output_conv2.0.weight"], sd[s + "output_conv2.0.bias"], padding=1))
    last = _gated_fusion_block(sd, s + "output_conv2_fusion.", [last], None, coarse_features[0], fusion, gate, upscale=False)
    out = F.conv2d(last, sd[s + "output_conv3.0.weight"], sd[s + "output_conv3.0.bias"])
    return [rn[4], path_5, path_4, path_3, path_2, last], out


def c2f_noenc_module(sd, pre, fine_features: List[Tensor], coarse_features: List[Tensor], fusion: bool, gate: bool):
    """C2FNOENCModule.forward (bi_directional_fusion_model.py:253-282): no top-down path; every level is two gated units
    against its coarse map, plus a transposed-conv level 0.  Returns ([path_5 .. path_0], depth)."""
    s = pre + "scratch."
    rn = [F.conv2d(f, sd[f"{s}layer{i + 1}_rn.weight"], padding=1) for i, f in enumerate(fine_features)]
    l0 = F.conv_transpose2d(fine_features[0], sd[s + "upsample_conv.0.weight"], sd[s + "upsample_conv.0.bias"], stride=2)
    l0 = F.conv2d(F.relu(l0), sd[s + "upsample_conv.2.weight"], padding=1)
    paths = []
    for lvl, (x, c) in enumerate(zip(rn[::-1], coarse_features[::-1][:5]), start=1):      # layer1_* works on layer_5_rn with c[5], ...
        x = _gated_conv_unit(sd, f"{s}layer{lvl}_gate1.", x, c, fusion, gate)
        paths.append(_gated_conv_unit(sd, f"{s}layer{lvl}_gate2.", x, c, fusion, gate))
    p0 = _gated_conv_unit(sd, s + "layer6_gate1.", l0, coarse_features[0], fusion, gate)
    p0 = _gated_conv_unit(sd, s + "layer6_gate2.", p0, coarse_features[0], fusion, gate)
    out = F.conv2d(p0, sd[s + "output_conv.weight"], sd[s + "output_conv.bias"], padding=1)
    return paths + [p0], out


def _single_conv_ln_heavy(sd, pre, x):
    """SingleConvCNNLNHeavy (bi_directional_fusion_model.py:448-463): conv -> LN -> conv -> LN -> conv -> GELU (no activation between)."""
    x = _ln_cf(F.conv2d(x, sd[pre + "single_conv.0.weight"], padding=1), sd[pre + "single_conv.1.weight"], sd[pre + "single_conv.1.bias"])
    x = _ln_cf(F.conv2d(x, sd[pre + "single_conv.2.weight"], padding=1), sd[pre + "single_conv.3.weight"], sd[pre + "single_conv.3.bias"])
    return F.gelu(F.conv2d(x, sd[pre + "single_conv.4.weight"], padding=1))


def bidirectional_fusion(sd: Dict[str, Tensor], pre: str, c_feat: List[Tensor], f_feat: List[Tensor], pred1: Tensor, pred2: Tensor,
                         update_base: Optional[Tensor], coarse2fine_type: str = "coarse-gated", trace: Optional[dict] = None,
                         heavy: bool = False) -> Tensor:
    """BiDirectionalFusion.forward (bi_directional_fusion_model.py:379-446) with glb_att=False, coarse2fine=True.
    c_feat / f_feat: 6 maps each, finest first (patchrefinerplus.py:318-326 reverses them); ``pred2`` is replaced by the
    C2F module's depth (:409-414), exactly as the reference does."""
    fusion, gate = C2F_TYPES[coarse2fine_type]
    c_feat, f_feat = list(c_feat), list(f_feat)
    if tuple(c_feat[-1].shape[-2:]) != tuple(f_feat[-1].shape[-2:]):
        c_feat = [_bil(c, f.shape[-2:]) for c, f in zip(c_feat, f_feat)]                       # :392-395
    c2f = c2f_noenc_module if coarse2fine_type == "only-gate" else c2f_module
    feats, out_depth = c2f(sd, pre + "c2f.", f_feat[1:], c_feat, fusion, gate)
    f_feat, pred2 = feats[::-1], out_depth
    if trace is not None:
        trace["c2f_feats"] = [t.clone() for t in f_feat]
        trace["c2f_depth"] = out_depth.clone()
    temp = []
    single = _single_conv_ln_heavy if heavy else _single_conv_ln
    for idx, (c, f) in enumerate(zip(c_feat, f_feat)):
        f = single(sd, f"{pre}fusion_layers_1.{idx}.", torch.cat([c, f], dim=1))
        p1 = _bil(pred1, f.shape[-2:])
        p2 = _bil(pred2, f.shape[-2:])
        temp.append(single(sd, f"{pre}fusion_layers_2.{idx}.", torch.cat([f, p1, p2], dim=1)))
    if trace is not None:
        trace["fusion_enc"] = [t.clone() for t in temp]
    dec = temp[0]
    temp = temp[::-1]
    _feat = temp[0]
    for i, feat in enumerate(temp[1:]):
        x = torch.cat([_bil(_feat, feat.shape[-2:]), feat, _bil(pred1, feat.shape[-2:]), _bil(pred2, feat.shape[-2:])], dim=1)
        for k in ((0, 2, 4, 6, 8) if heavy else (0, 2)):                       # DoubleConv (convs.py:31-45) / DoubleConvHeavy (:465-485)
            x = F.gelu(F.conv2d(x, sd[f"{pre}f2r_agg.{i}.conv.double_conv.{k}.weight"], padding=1))
        dec = _feat = x
    off = F.conv2d(dec, sd[pre + "final_conv.weight"], padding=1)
    if update_base is not None:
        return torch.clamp(update_base + off, min=0)
    return off


def synthetic_fusion_inputs(coarse_chl, fine_chl, sizes_c, sizes_f, B: int, seed: int):
    """Seeded stand-ins for the tensors BiDirectionalFusion consumes: ROI-cropped coarse features, the light-weight
    encoder's features (its timm arithmetic is not available offline: parity of the ENCODER is unpinned, the fusion
    model is pinned on these inputs), the ROI coarse depth and the (ignored, see :409-414) refiner depth."""
    g = torch.Generator().manual_seed(seed)
    c = [torch.randn(B, ch, *sz, generator=g) for ch, sz in zip(coarse_chl, sizes_c)]
    f = [torch.relu(torch.randn(B, ch, *sz, generator=g)) for ch, sz in zip([fine_chl[0]] + list(fine_chl), sizes_f)]
    p1 = torch.rand(B, 1, *sizes_f[0], generator=g) * 10
    p2 = torch.zeros(B, 1, *sizes_f[0])
    return c, f, p1, p2


# ---------------------------------------------------------------------------------------------
# ZoeDepth metric-bins head  (external/zoedepth/models/zoedepth/zoedepth_v1.py:173-233)
# Oracle only so far: the product does not build this head yet (its BEiT-L core is unavailable offline).
# ---------------------------------------------------------------------------------------------

#: configs/patchrefiner_zoedepth/pr_u4k.py:10-66 (the values every ZoeDepth branch of the reference uses)
ZOE_HEAD_CFG = dict(n_bins=64, bin_embedding_dim=128, n_attractors=(16, 8, 4, 1), attractor_alpha=1000, attractor_gamma=2,
                    attractor_kind="mean", attractor_type="inv", min_temp=0.0212, max_temp=50.0, bin_centers_type="softplus")


def init_zoe_head_state_dict(output_channels: Sequence[int], seed: int, cfg: dict = ZOE_HEAD_CFG, n_midas_out: int = 32) -> Dict[str, Tensor]:
    """Random weights with the reference's keys for everything in ZoeDepth except ``core`` (zoedepth_v1.py:84-123)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    bt, outs, E, nb = output_channels[0], list(output_channels[1:]), cfg["bin_embedding_dim"], cfg["n_bins"]

    def conv(name, co, ci, gain=1.0):
        sd[name + ".weight"] = _conv_w(g, co, ci, 1, gain)
        sd[name + ".bias"] = _vec(g, co, 0.05)

    conv("conv2", bt, bt)
    conv("seed_bin_regressor._net.0", 256, bt, 1.7); conv("seed_bin_regressor._net.2", nb, 256)
    conv("seed_projector._net.0", 128, bt, 1.7); conv("seed_projector._net.2", E, 128)
    for i, c in enumerate(outs):
        conv(f"projectors.{i}._net.0", 128, c, 1.7); conv(f"projectors.{i}._net.2", E, 128)
        conv(f"attractors.{i}._net.0", 128, E, 1.7); conv(f"attractors.{i}._net.2", cfg["n_attractors"][i], 128)
    last_in = n_midas_out + 1
    conv("conditional_log_binomial.mlp.0", (last_in + E) // 2, last_in + E, 1.7)
    conv("conditional_log_binomial.mlp.2", 4, (last_in + E) // 2)
    return sd


def _mlp1x1(sd, pre, x, act_last=None):
    x = F.relu(F.conv2d(x, sd[pre + "_net.0.weight"], sd[pre + "_net.0.bias"]))
    x = F.conv2d(x, sd[pre + "_net.2.weight"], sd[pre + "_net.2.bias"])
    return act_last(x) if act_last is not None else x


def zoe_bins_head(sd: Dict[str, Tensor], pre: str, rel_depth: Tensor, btlnck: Tensor, x_blocks: List[Tensor], outconv: Tensor,
                  cfg: dict = ZOE_HEAD_CFG, trace: Optional[dict] = None):
    """ZoeDepth.forward after the core (zoedepth_v1.py:173-233) for bin_centers_type='softplus', inverse_midas=False:
    seed bins (layers/localbins_layers.py:71-96), projectors (:99-118), AttractorLayerUnnormed (layers/attractor.py:139-208,
    inv_attractor :45-57), ConditionalLogBinomial + LogBinomial (layers/dist_layers.py:29-122).
    rel_depth [B,H,W]; btlnck / x_blocks / outconv as the core returns them.  Returns (metric_depth [B,1,h,w], temp_features)."""
    assert cfg["bin_centers_type"] == "softplus" and cfg["attractor_type"] == "inv"
    x_d0 = F.conv2d(btlnck, sd[pre + "conv2.weight"], sd[pre + "conv2.bias"])                     # :173
    b_prev = _mlp1x1(sd, pre + "seed_bin_regressor.", x_d0, F.softplus)                             # :176, unnormed: centres themselves
    prev_emb = _mlp1x1(sd, pre + "seed_projector.", x_d0)                                           # :184
    # NB: AttractorLayerUnnormed.forward calls ``dist(A.unsqueeze(2) - b_centers.unsqueeze(1))`` WITHOUT alpha / gamma
    # (attractor.py:194-197), so the jit function's defaults alpha=300, gamma=2 apply whatever the config says (1000 in
    # every shipped config).  The oracle follows the code, not the config.
    alpha, gamma, K = 300.0, 2, int(cfg["n_bins"])
    b_centers = b_prev
    emb = prev_emb
    for i, x in enumerate(x_blocks):                                                               # :188-195
        emb = _mlp1x1(sd, f"{pre}projectors.{i}.", x)
        xa = emb + F.interpolate(prev_emb, x.shape[-2:], mode="bilinear", align_corners=True)      # attractor.py:176-180
        A = _mlp1x1(sd, f"{pre}attractors.{i}.", xa, F.softplus)
        b_c = F.interpolate(b_prev, A.shape[-2:], mode="bilinear", align_corners=True)
        dx = A.unsqueeze(2) - b_c.unsqueeze(1)
        dc = dx.div(1 + alpha * dx.pow(gamma))                                                      # inv_attractor
        delta = dc.mean(dim=1) if cfg["attractor_kind"] == "mean" else dc.sum(dim=1)
        b_centers = b_c + delta
        b_prev, prev_emb = b_centers, emb
    rel_cond = F.interpolate(rel_depth.unsqueeze(1), size=outconv.shape[2:], mode="bilinear", align_corners=True)   # :207-210
    last = torch.cat([outconv, rel_cond], dim=1)
    cond = F.interpolate(emb, last.shape[-2:], mode="bilinear", align_corners=True)
    q = pre + "conditional_log_binomial.mlp."
    pt = F.conv2d(torch.cat((last, cond), dim=1), sd[q + "0.weight"], sd[q + "0.bias"])
    pt = F.softplus(F.conv2d(F.gelu(pt), sd[q + "2.weight"], sd[q + "2.bias"]))                   # dist_layers.py:95-103
    p_eps = 1e-4
    pp = pt[:, :2] + p_eps
    prob = pp[:, 0] / (pp[:, 0] + pp[:, 1])
    tt = pt[:, 2:] + p_eps
    t = (tt[:, 0] / (tt[:, 0] + tt[:, 1])).unsqueeze(1)
    t = (float(cfg["max_temp"]) - float(cfg["min_temp"])) * t + float(cfg["min_temp"])
    xk = prob.unsqueeze(1)                                                                          # LogBinomial.forward (:52-69)
    eps = 1e-4
    one_minus = torch.clamp(1 - xk, eps, 1)
    xk = torch.clamp(xk, eps, 1)
    k_idx = torch.arange(0, K, device=xk.device).view(1, -1, 1, 1)
    Km1 = torch.tensor([float(K - 1)], device=xk.device).view(1, -1, 1, 1)

    def log_binom(n, k, e=1e-7):                                                                    # :29-33 (Stirling)
        n = n + e
        k = k + e
        return n * torch.log(n) - k * torch.log(k) - (n - k) * torch.log(n - k + e)

    y = log_binom(Km1, k_idx) + k_idx * torch.log(xk) + (K - 1 - k_idx) * torch.log(one_minus)
    probs = torch.softmax(y / t, dim=1)
    centers = F.interpolate(b_centers, probs.shape[-2:], mode="bilinear", align_corners=True)      # :217-218
    depth = torch.sum(probs * centers, dim=1, keepdim=True)
    feats = {"x_d0": x_d0, "midas_final_feat": outconv}
    for i, x in enumerate(x_blocks):
        feats[f"x_blocks_feat_{i}"] = x
    if trace is not None:
        trace.update(probs=probs, bin_centers=centers)
    return depth, feats


# ---------------------------------------------------------------------------------------------
# tiling geometry, masks, running average  (baseline_pretrain.py ; models/utils.py)
# ---------------------------------------------------------------------------------------------

def prepare_tile_cfg(patch_process_shape, image_raw_shape, patch_split_num) -> dict:
    """BaselinePretrain.prepare_tile_cfg (baseline_pretrain.py:96-124)."""
    ph, pw = patch_process_shape
    sh, sw = patch_split_num
    raw = (image_raw_shape[0] // sh, image_raw_shape[1] // sw)
    return {
        "patch_split_num": patch_split_num,
        "patch_reensemble_shape": (ph * sh, pw * sw),
        "patch_raw_shape": raw,
        "image_raw_shape": image_raw_shape,
        "raw_h_split_point": [int(raw[0] * i) for i in range(sh)],
        "raw_w_split_point": [int(raw[1] * i) for i in range(sw)],
    }


def generatemask(size, border=0.1) -> np.ndarray:
    """estimator/models/utils.py:51-60 (OpenCV Gaussian of a box of ones, min-max normalised)."""
    mask = np.zeros(size, dtype=np.float32)
    sigma = int(size[0] / 16)
    k_size = int(2 * np.ceil(2 * int(size[0] / 16)) + 1)
    mask[int(border * size[0]):size[0] - int(border * size[0]), int(border * size[1]):size[1] - int(border * size[1])] = 1
    mask = cv2.GaussianBlur(mask, (int(k_size), int(k_size)), sigma)
    mask = (mask - mask.min()) / (mask.max() - mask.min())
    return mask.astype(np.float32)


def resizer_size(patch_process_shape, in_h, in_w) -> Tuple[int, int]:
    """external/depth_anything/transform.py:43-127 for (keep_aspect_ratio=False,
    ensure_multiple_of=14, resize_method='minimal'): round(target/14)*14 per axis."""
    ph, pw = patch_process_shape
    new_h = int(np.round((ph / in_h) * in_h / PATCH) * PATCH)
    new_w = int(np.round((pw / in_w) * in_w / PATCH) * PATCH)
    return new_h, new_w


def resizer(patch_process_shape, x: Tensor) -> Tensor:
    """Resize.__call__ (transform.py:127-129): bilinear, align_corners=True."""
    nh, nw = resizer_size(patch_process_shape, x.shape[-2], x.shape[-1])
    return F.interpolate(x, (nh, nw), mode="bilinear", align_corners=True)


def bbox_feat_factor(image_raw_shape, patch_process_shape) -> Tensor:
    """baseline_pretrain.py:289-293: Python-double ``1 / W * pw`` rounded to float32."""
    H, W = image_raw_shape
    ph, pw = patch_process_shape
    return torch.tensor([1 / W * pw, 1 / H * ph, 1 / W * pw, 1 / H * ph]).unsqueeze(0)


def regular_bboxes(tile_cfg, offset) -> Tuple[List[int], List[int]]:
    """baseline_pretrain.py:249-257: row-major start lists of the (shifted) regular grid."""
    rh, rw = tile_cfg["patch_raw_shape"]
    oh, ow = offset
    assert ow >= 0 and oh >= 0
    nh = (tile_cfg["image_raw_shape"][0] - oh) // rh
    nw = (tile_cfg["image_raw_shape"][1] - ow) // rw
    return [rh * i + oh for i in range(nh)], [rw * j + ow for j in range(nw)]


def make_bboxs(h_starts, w_starts, rh, rw) -> Tensor:
    """bbox = [w0, h0, w0+rw, h0+rh] .int(), patch order = for h: for w (baseline_pretrain.py:272-287)."""
    rows = [[w0, h0, w0 + rw, h0 + rh] for h0 in h_starts for w0 in w_starts]
    return torch.tensor(rows).int().reshape(-1, 4)


def bboxs_to_feat(bboxs: Tensor, image_raw_shape, patch_process_shape) -> Tensor:
    """baseline_pretrain.py:289-296: [P,4] int32 -> [P,5] float32 (index column first)."""
    bf = bboxs * bbox_feat_factor(image_raw_shape, patch_process_shape)
    inds = torch.arange(bboxs.shape[0]).unsqueeze(-1)
    return torch.cat((inds, bf), dim=-1)


def crop_resize(image_hr: Tensor, bboxs: Tensor, patch_process_shape) -> Tensor:
    """baseline_pretrain.py:272-280: slice + resizer per patch, stacked.  image_hr [3,H,W]."""
    crops = []
    for w0, h0, w1, h1 in bboxs.tolist():
        crops.append(resizer(patch_process_shape, image_hr[:, h0:h1, w0:w1].unsqueeze(0)).squeeze(0))
    return torch.stack(crops, dim=0)


def coarse_postprocess_test(coarse_prediction: Tensor, coarse_features: List[Tensor], bboxs_feat: Tensor, ph: int):
    """PatchRefiner.coarse_postprocess_test (patchrefiner.py:199-217).  ``repeat`` is replaced by
    batch index 0 for every ROI (identical result: every repeated copy equals the original)."""
    rois = bboxs_feat.clone().to(coarse_prediction.device)
    rois[:, 0] = 0
    feats = []
    for feat in coarse_features:
        h, w = feat.shape[-2:]
        feats.append(tv_roi_align(feat, rois, (h, w), h / ph, aligned=True))
    h, w = coarse_prediction.shape[-2:]
    depth = tv_roi_align(coarse_prediction, rois, (h, w), h / ph, aligned=True)
    return depth, feats


class RunningAverageMap:
    """estimator/models/utils.py:22-49, same tensor ops in the same order."""

    def __init__(self, average_map, count_map):
        self.count_map = count_map
        self.average_map_init = average_map
        self.average_map = average_map
        self.update_flag = False

    def update(self, pred_map, ct_map):
        self.update_flag = True
        mask = ct_map > 0
        self.average_map[mask] = (pred_map[mask] * ct_map[mask] + self.count_map[mask] * self.average_map[mask]) / (self.count_map[mask] + ct_map[mask])
        self.count_map[mask] = self.count_map[mask] + ct_map[mask]

    def resize(self, resolution):
        a = self.average_map.unsqueeze(0).unsqueeze(0)
        c = self.count_map.unsqueeze(0).unsqueeze(0)
        self.average_map = F.interpolate(a, size=resolution).squeeze()
        self.count_map = F.interpolate(c, size=resolution, mode="bilinear", align_corners=True).squeeze()

    def get_avg_map(self):
        return self.average_map if self.update_flag else self.average_map_init


# ---------------------------------------------------------------------------------------------
# the whole path
# ---------------------------------------------------------------------------------------------

class PatchRefinerOracle:
    """PatchRefiner.forward(mode='infer') (estimator/models/patchrefiner.py:341-401) together with
    BaselinePretrain.regular_tile / random_tile (baseline_pretrain.py:149-375), restated on CPU."""

    def __init__(self, cfg: dict, state_dict: Dict[str, Tensor]):
        self.cfg = cfg
        self.sd = {k: v.float() for k, v in state_dict.items()}
        self.patch_process_shape = tuple(cfg["patch_process_shape"])
        self.max_depth = float(cfg["max_depth"])
        self.enc_c = cfg["coarse_branch"]["model_cfg"]["encoder"]
        self.enc_f = cfg["refiner"]["fine_branch"]["model_cfg"]["encoder"]
        self.level = cfg["fusion_feat_level"]
        self.target = cfg["strategy_refiner_target"]
        self.tile_cfg = prepare_tile_cfg(self.patch_process_shape, cfg["image_raw_shape"], cfg["patch_split_num"])
        self.trace: Optional[dict] = None       # set to {} to collect intermediates

    # -- pieces -------------------------------------------------------------------------------
    def coarse_forward(self, image_lr):
        """patchrefiner.py:168-185"""
        depth, feats = depth_anything_v2(self.sd, "coarse_branch.", image_lr, self.enc_c, self.max_depth)
        return feats, depth

    def infer_forward(self, imgs_crop, coarse_depth_roi, coarse_feats_roi, trace=None):
        """patchrefiner.py:258-283 (+ refiner_fine_forward :219-232, refiner_fusion_forward :234-256)"""
        r_depth, r_feats = depth_anything_v2(self.sd, "refiner_fine_branch.", imgs_crop, self.enc_f, self.max_depth, trace)
        if trace is not None:
            trace["fine_depth"] = r_depth.clone()
            trace["fine_feats"] = [t.clone() for t in r_feats]
        if self.target == "offset_fine":
            base = r_depth
        elif self.target == "offset_coarse":
            base = coarse_depth_roi
        else:
            base = None
        c_list = list(coarse_feats_roi[-self.level:])[::-1]
        r_list = list(r_feats[-self.level:])[::-1]
        pred = fusion_unet(self.sd, "refiner_fusion_model.", c_list, r_list, coarse_depth_roi, r_depth, base, trace)
        if self.target == "direct":
            pred = torch.sigmoid(pred) * self.max_depth
        return pred

    def _predict(self, image_hr, bboxs, tile_cfg, tile_temp, process_num, record):
        """shared front half of regular_tile / random_tile: crop, roi_align, chunked forward."""
        imgs_crop = crop_resize(image_hr, bboxs, self.patch_process_shape)
        bf = bboxs_to_feat(bboxs, tile_cfg["image_raw_shape"], self.patch_process_shape)
        d_roi, f_roi = coarse_postprocess_test(tile_temp["coarse_prediction"], tile_temp["coarse_features"], bf, self.patch_process_shape[0])
        preds = []
        for s in range(0, imgs_crop.shape[0], process_num):
            sl = slice(s, s + process_num)
            tr = None
            if self.trace is not None and "first_patch" not in self.trace:
                tr = self.trace["first_patch"] = {}
            preds.append(self.infer_forward(imgs_crop[sl], d_roi[sl], [f[sl] for f in f_roi], tr))
        preds = torch.cat(preds, dim=0)
        if record is not None:
            record["bboxs"].append(bboxs.clone())
            record["bboxs_feat"].append(bf.clone())
            record["preds"].append(preds.clone())
            if "roi_first" not in record:
                record["roi_first"] = dict(crop=imgs_crop[:2].clone(), depth=d_roi[:2].clone(), feats=[f[:2].clone() for f in f_roi])
        return preds

    def regular_tile(self, offset, offset_process, image_hr, init_flag, tile_temp, blur_mask, avg, tile_cfg, process_num, record):
        """baseline_pretrain.py:235-375"""
        rh, rw = tile_cfg["patch_raw_shape"]
        hs, ws = regular_bboxes(tile_cfg, offset)
        ph, pw = self.patch_process_shape
        oph, opw = offset_process
        nhp = (tile_cfg["patch_reensemble_shape"][0] - oph) // ph
        nwp = (tile_cfg["patch_reensemble_shape"][1] - opw) // pw
        hsp = [ph * i + oph for i in range(nhp)]
        wsp = [pw * j + opw for j in range(nwp)]
        bboxs = make_bboxs(hs, ws, rh, rw)
        preds = self._predict(image_hr, bboxs, tile_cfg, tile_temp, process_num, record)
        count_map = torch.zeros(tile_cfg["patch_reensemble_shape"])
        pred_depth = torch.zeros(tile_cfg["patch_reensemble_shape"])
        idx = 0
        for h0 in hsp:
            for w0 in wsp:
                d = preds[idx]
                if init_flag:
                    count_map[h0:h0 + ph, w0:w0 + pw] = blur_mask
                    pred_depth[h0:h0 + ph, w0:w0 + pw] = d
                else:
                    count_map = torch.zeros(tile_cfg["patch_reensemble_shape"])
                    pred_depth = torch.zeros(tile_cfg["patch_reensemble_shape"])
                    count_map[h0:h0 + ph, w0:w0 + pw] = blur_mask
                    pred_depth[h0:h0 + ph, w0:w0 + pw] = d
                    avg.update(pred_depth, count_map)
                idx += 1
        if init_flag:
            avg = RunningAverageMap(pred_depth, count_map)
        return avg

    def random_tile(self, image_hr, tile_temp, blur_mask, avg, tile_cfg, process_num, record):
        """baseline_pretrain.py:149-231 (global ``random`` stream: process_num rows then ONE column)."""
        rh, rw = tile_cfg["patch_raw_shape"]
        H, W = tile_cfg["image_raw_shape"]
        hs = [random.randint(0, H - rh - 1) for _ in range(process_num)]
        ws = [random.randint(0, W - rw - 1)]
        bboxs = make_bboxs(hs, ws, rh, rw)
        preds = self._predict(image_hr, bboxs, tile_cfg, tile_temp, process_num, record)
        preds = F.interpolate(preds, tile_cfg["patch_raw_shape"])          # nearest (:210)
        idx = 0
        for h0 in hs:
            for w0 in ws:
                count_map = torch.zeros(tile_cfg["image_raw_shape"])
                pred_depth = torch.zeros(tile_cfg["image_raw_shape"])
                count_map[h0:h0 + rh, w0:w0 + rw] = blur_mask
                pred_depth[h0:h0 + rh, w0:w0 + rw] = preds[idx]
                avg.update(pred_depth, count_map)
                idx += 1
        return avg

    # -- entry --------------------------------------------------------------------------------
    @torch.no_grad()
    def infer(self, image_lr: Tensor, image_hr: Tensor, tile_cfg: Optional[dict] = None, cai_mode: str = "m1",
              process_num: int = 4, record: Optional[dict] = None):
        """patchrefiner.py:341-401.  Returns (depth [1,1,h,w] fp32, coarse_prediction, avg map object)."""
        if tile_cfg is None:
            tile_cfg = self.tile_cfg
        else:
            tile_cfg = prepare_tile_cfg(self.patch_process_shape, tile_cfg["image_raw_shape"], tile_cfg["patch_split_num"])
        assert image_hr.shape[0] == 1
        if record is not None:
            for k in ("bboxs", "bboxs_feat", "preds"):
                record.setdefault(k, [])
        feats, coarse = self.coarse_forward(image_lr)
        if record is not None:
            record["coarse_prediction"] = coarse.clone()
            record["coarse_features"] = [f.clone() for f in feats]
        tt = {"coarse_prediction": coarse, "coarse_features": feats}
        ph, pw = self.patch_process_shape
        rh, rw = tile_cfg["patch_raw_shape"]
        blur = torch.tensor(generatemask((ph, pw), border=0.15))
        hr = image_hr[0]
        avg = self.regular_tile([0, 0], [0, 0], hr, True, tt, blur, None, tile_cfg, process_num, record)
        if cai_mode == "m2" or cai_mode[0] == "r":
            avg = self.regular_tile([0, rw // 2], [0, pw // 2], hr, False, tt, blur, avg, tile_cfg, process_num, record)
            avg = self.regular_tile([rh // 2, 0], [ph // 2, 0], hr, False, tt, blur, avg, tile_cfg, process_num, record)
            avg = self.regular_tile([rh // 2, rw // 2], [ph // 2, pw // 2], hr, False, tt, blur, avg, tile_cfg, process_num, record)
        if cai_mode[0] == "r":
            blur_r = torch.tensor(generatemask((rh, rw), border=0.15) + 1e-3)
            avg.resize(tile_cfg["image_raw_shape"])
            if record is not None:
                record["avg_resized"] = avg.average_map.clone()
                record["count_resized"] = avg.count_map.clone()
            for _ in range(int(cai_mode[1:]) // process_num):
                avg = self.random_tile(hr, tt, blur_r, avg, tile_cfg, process_num, record)
        depth = avg.get_avg_map().unsqueeze(0).unsqueeze(0)
        return depth, coarse, avg


# ---------------------------------------------------------------------------------------------
# PatchRefinerPlus (V2 family) tiled inference  (estimator/models/patchrefinerplus.py)
# ---------------------------------------------------------------------------------------------

TOY_ENCODER_NAME = "mobilenetv4_conv_small.e2400_r224_in1k"     # the name whose 4-channel stem surgery the reference knows (patchrefinerplus.py:158-164)
TOY_ENCODER_CHL = (32, 32, 64, 96, 960)                         # == fine_chl of configs/patchrefinerv2_*/ *mobile* configs


class ToyFineEncoder(torch.nn.Module):
    """Stand-in for the timm ``features_only`` CNN inside ``LightWeightRefiner`` (blocks/lightweight_refiner.py:259-262).
    timm is not installable offline, so the encoder's own arithmetic is UNPINNED; everything around
    it -- pixel normalisation, coarse-depth conditioning, feature ordering, BiDirectionalFusion, tiling, blending --
    is pinned by running the reference ``PatchRefinerPlus`` with this module injected as ``timm.create_model``'s result.
    Five stride-2 stages with the channel counts / strides of mobilenetv4_conv_small; ``conv_stem`` and ``default_cfg``
    are the attributes the reference touches (patchrefinerplus.py:158-164, lightweight_refiner.py:264-265)."""

    default_cfg = {"mean": (0.485, 0.456, 0.406), "std": (0.229, 0.224, 0.225)}

    def __init__(self, in_chans: int = 3):
        super().__init__()
        nn = torch.nn
        c = TOY_ENCODER_CHL
        self.conv_stem = nn.Conv2d(in_chans, c[0], 3, stride=2, padding=1, bias=False)
        self.stages = nn.ModuleList([nn.Conv2d(c[i], c[i + 1], 3, stride=2, padding=1, bias=True) for i in range(4)])

    def forward(self, x):
        feats = [F.relu(self.conv_stem(x))]
        for st in self.stages:
            feats.append(F.relu(st(feats[-1])))
        return feats


TOY_CONVX_NAME = "convnext_large"                               # the name whose 4-channel ``stem_0`` surgery the reference knows (patchrefinerplus.py:194-200)
TOY_CONVX_CHL = (32, 192, 32, 48, 64)                           # encoder_channels: [up-sampled level, stem (192 is hard-coded by the surgery), 3 stages]


class ToyConvNeXtEncoder(torch.nn.Module):
    """Stand-in for timm's ``convnext_large`` ``features_only`` model inside ``LightWeightRefiner``: FOUR maps at strides
    4 / 8 / 16 / 32 (the reference adds the stride-2 and stride-1 levels itself with ``upsample_convx`` and a bilinear resize,
    lightweight_refiner.py:276-283, 307-314).  ``stem_0`` (4x4 stride-4 conv with a bias) and ``default_cfg`` are the attributes
    the reference touches.  The arithmetic of the real encoder is UNPINNED (timm absent); the stage around it is pinned."""

    default_cfg = {"mean": (0.485, 0.456, 0.406), "std": (0.229, 0.224, 0.225)}

    def __init__(self, in_chans: int = 3):
        super().__init__()
        nn = torch.nn
        c = TOY_CONVX_CHL
        self.stem_0 = nn.Conv2d(in_chans, c[1], 4, stride=4)
        self.stages = nn.ModuleList([nn.Conv2d(c[i], c[i + 1], 3, stride=2, padding=1, bias=True) for i in range(1, 4)])

    def forward(self, x):
        feats = [F.relu(self.stem_0(x))]
        for st in self.stages:
            feats.append(F.relu(st(feats[-1])))
        return feats


def init_toy_convx_state_dict(seed: int, in_chans: int = 4) -> Dict[str, Tensor]:
    """Encoder weights plus the reference's own ``upsample_convx`` (ConvTranspose2d(192 -> 32, k = s = 2), keys relative to
    ``refiner_fine_branch.``)."""
    g = torch.Generator().manual_seed(seed)
    c = TOY_CONVX_CHL
    sd = {"refiner_encoder.stem_0.weight": _conv_w(g, c[1], in_chans, 4, gain=1.7), "refiner_encoder.stem_0.bias": _vec(g, c[1], 0.05)}
    for i in range(3):
        sd[f"refiner_encoder.stages.{i}.weight"] = _conv_w(g, c[i + 2], c[i + 1], 3, gain=1.7)
        sd[f"refiner_encoder.stages.{i}.bias"] = _vec(g, c[i + 2], 0.05)
    sd["upsample_convx.0.weight"] = _conv_w(g, c[1], c[0], 2, gain=1.7)              # ConvTranspose2d weight: [in, out, 2, 2]
    sd["upsample_convx.0.bias"] = _vec(g, c[0], 0.05)
    return sd


def init_toy_encoder_state_dict(seed: int, in_chans: int = 4) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    c = TOY_ENCODER_CHL
    sd = {"conv_stem.weight": _conv_w(g, c[0], in_chans, 3, gain=1.7)}
    for i in range(4):
        sd[f"stages.{i}.weight"] = _conv_w(g, c[i + 1], c[i], 3, gain=1.7)
        sd[f"stages.{i}.bias"] = _vec(g, c[i + 1], 0.05)
    return sd


def make_plus_config(encoder="vits", features=256, out_channels=(48, 96, 192, 384), patch_process_shape=(224, 224), image_raw_shape=(432, 768),
                     patch_split_num=(2, 2), coarse2fine_type="coarse-gated", max_depth=80.0, convnext=False) -> dict:
    """A config dict shaped like configs/patchrefinerv2_dav2/plus_mobile_u4k_base_coarse_e2e_c2f_pretrain.py (DA2 coarse branch
    with 256 decoder features, LightWeightRefiner fine branch, BiDirectionalFusion); ``convnext=True``: like plus_convx_u4k_* (four
    encoder maps + the ``upsample_convx`` stage)."""
    if convnext:
        cfg = make_plus_config(encoder, features, out_channels, patch_process_shape, image_raw_shape, patch_split_num, coarse2fine_type, max_depth)
        cfg["refiner"]["fine_branch"].update(encoder_name=TOY_CONVX_NAME, encoder_channels=list(TOY_CONVX_CHL))
        cfg["refiner"]["fusion_model"].update(encoder_name=TOY_CONVX_NAME, fine_chl=list(TOY_CONVX_CHL))
        return cfg
    return dict(
        image_raw_shape=list(image_raw_shape), patch_process_shape=list(patch_process_shape), patch_split_num=list(patch_split_num),
        fusion_feat_level=6, min_depth=1e-3, max_depth=max_depth, strategy_refiner_target="offset_coarse",
        pretrain_stage=False, e2e_training=False, hack_strategy=None,
        coarse_branch=dict(type="DA2", pretrained=None, model_cfg=dict(encoder=encoder, features=features, out_channels=list(out_channels))),
        refiner=dict(
            fine_branch=dict(type="LightWeightRefiner", coarse_condition=True, with_decoder=False, encoder_name=TOY_ENCODER_NAME),
            fusion_model=dict(type="BiDirectionalFusion", encoder_name=TOY_ENCODER_NAME, coarse2fine=True, coarse2fine_type=coarse2fine_type,
                              coarse_chl=[features // 2] + [features] * 5, fine_chl=list(TOY_ENCODER_CHL),
                              fine_chl_after_coarse2fine=[features // 2] + [features] * 5,
                              temp_chl=[32, 64, 64, 128, 256, 512], dec_chl=[512, 256, 128, 64, 32])),
        sigloss=dict(type="SILogLoss"), gmloss=dict(type="GradMatchLoss"), sigweight=1, pre_norm_bbox=True,
        pretrain_coarse_model=None, pretrained=None, whole_pretrained=None)


def init_patchrefinerplus_state_dict(cfg: dict, seed: int = 0) -> Dict[str, Tensor]:
    """State dict with the reference's PatchRefinerPlus prefixes: coarse_branch.* (DepthAnythingV2),
    refiner_fine_branch.refiner_encoder.* (the encoder), refiner_fusion_model.* (BiDirectionalFusion)."""
    cb = cfg["coarse_branch"]["model_cfg"]
    fu = cfg["refiner"]["fusion_model"]
    sd: Dict[str, Tensor] = {}
    for k, v in init_dav2_state_dict(cb["encoder"], cb["features"], cb["out_channels"], seed * 3 + 21).items():
        sd["coarse_branch." + k] = v
    if "convnext" in cfg["refiner"]["fine_branch"]["encoder_name"]:
        for k, v in init_toy_convx_state_dict(seed * 3 + 22).items():
            sd["refiner_fine_branch." + k] = v
    else:
        for k, v in init_toy_encoder_state_dict(seed * 3 + 22).items():
            sd["refiner_fine_branch.refiner_encoder." + k] = v
    for k, v in init_bidirectional_fusion_state_dict(fu["coarse_chl"], fu["fine_chl"], fu["fine_chl_after_coarse2fine"], fu["temp_chl"], fu["dec_chl"],
                                                     seed * 3 + 23, fu["coarse2fine_type"]).items():
        sd["refiner_fusion_model." + k] = v
    return sd


class PatchRefinerPlusOracle(PatchRefinerOracle):
    """PatchRefinerPlus.forward(mode='infer') (patchrefinerplus.py:367-533): the tiling / blending is BaselinePretrain's
    (shared with PatchRefiner); only ``infer_forward`` differs (:330-365): LightWeightRefiner (lightweight_refiner.py:285-322,
    with_decoder=False, coarse_condition=True) then BiDirectionalFusion."""

    def __init__(self, cfg: dict, state_dict: Dict[str, Tensor], encoder: torch.nn.Module):
        self.cfg = cfg
        self.sd = {k: v.float() for k, v in state_dict.items()}
        self.patch_process_shape = tuple(cfg["patch_process_shape"])
        self.max_depth = float(cfg["max_depth"])
        self.enc_c = cfg["coarse_branch"]["model_cfg"]["encoder"]
        self.level = cfg["fusion_feat_level"]
        self.target = cfg["strategy_refiner_target"]
        self.c2f_type = cfg["refiner"]["fusion_model"]["coarse2fine_type"]
        self.tile_cfg = prepare_tile_cfg(self.patch_process_shape, cfg["image_raw_shape"], cfg["patch_split_num"])
        self.trace = None
        pre = "refiner_fine_branch.refiner_encoder."
        encoder.load_state_dict({k[len(pre):]: v for k, v in self.sd.items() if k.startswith(pre)}, strict=True)
        self.encoder = encoder.eval()

    def infer_forward(self, imgs_crop, coarse_depth_roi, coarse_feats_roi, trace=None):
        mean = torch.tensor(self.encoder.default_cfg["mean"], device=imgs_crop.device).view(-1, 1, 1)
        std = torch.tensor(self.encoder.default_cfg["std"], device=imgs_crop.device).view(-1, 1, 1)
        x = (imgs_crop - mean) / std                                                      # lightweight_refiner.py:293
        feats = list(self.encoder(torch.cat([x, coarse_depth_roi], dim=1)))                # :296 (coarse_condition)
        if "convnext" in self.cfg["refiner"]["fine_branch"]["encoder_name"]:              # :307-314: ConvTranspose2d(k = s = 2) + ReLU, then bilinear x2
            q = "refiner_fine_branch.upsample_convx.0."
            feats.insert(0, F.relu(F.conv_transpose2d(feats[0], self.sd[q + "weight"], self.sd[q + "bias"], stride=2)))
        feats.insert(0, F.interpolate(feats[0], scale_factor=2, mode="bilinear", align_corners=True))   # :316-318 / :312-313
        r_feats = feats[::-1]                                                             # :320
        r_depth = torch.zeros_like(x[:, :1])                                              # :321
        if trace is not None:
            trace["fine_feats"] = [t.clone() for t in r_feats]
        if self.target == "offset_fine":
            base = r_depth
        elif self.target == "offset_coarse":
            base = coarse_depth_roi
        else:
            base = None
        c_list = list(coarse_feats_roi[-self.level:])[::-1]
        r_list = list(r_feats[-self.level:])[::-1]
        pred = bidirectional_fusion(self.sd, "refiner_fusion_model.", c_list, r_list, coarse_depth_roi, r_depth, base, self.c2f_type, trace)
        if self.target == "direct":
            pred = torch.sigmoid(pred) * self.max_depth
        return pred


def synthetic_frame(cfg: dict, seed: int = 1, image_raw_shape=None) -> Tuple[Tensor, Tensor]:
    """Synthetic frame: smooth random structure + noise in [0,1]; image_lr = resizer(image_hr)
    (estimator/datasets/general_dataset.py:203-234 output contract)."""
    H, W = image_raw_shape if image_raw_shape is not None else cfg["image_raw_shape"]
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(1, 3, 9, 16, generator=g)
    smooth = F.interpolate(low, (H, W), mode="bicubic", align_corners=True).clamp(0, 1)
    noise = torch.rand(1, 3, H, W, generator=g)
    image_hr = (0.7 * smooth + 0.3 * noise).contiguous()
    image_lr = resizer(tuple(cfg["patch_process_shape"]), image_hr)
    return image_lr, image_hr


# ---------------------------------------------------------------------------------------------
# explicit index-math restatements (bit-exact contracts for the CUDA gather / blend kernels)
# ---------------------------------------------------------------------------------------------

def np_bilinear_ac_axis(n_in: int, n_out: int):
    """ATen upsample_bilinear2d, align_corners=True, fp32 (aten/src/ATen/native/UpSample.h
    area_pixel_compute_scale / compute_source_index): scale=(in-1)/(out-1) in float;
    src=scale*dst; i0=int(src); i1=i0+(i0<in-1); l1=src-i0; l0=1-l1."""
    scale = np.float32(0.0) if n_out <= 1 else np.float32(np.float32(n_in - 1) / np.float32(n_out - 1))
    dst = np.arange(n_out, dtype=np.float32)
    src = (scale * dst).astype(np.float32)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1.0) - l1).astype(np.float32)
    return i0, i1, l0, l1


def _fma32(a, b, c):
    """float32 fused multiply-add emulated exactly in float64 (a*b is exact in f64; one rounding)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def np_bilinear_ac(src: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Bit-exact restatement of F.interpolate(mode='bilinear', align_corners=True) on this
    container's CPU (AVX2/AVX-512 ATen build): r0=fma(lx0,a,lx1*b); r1=fma(lx0,c,lx1*d);
    out=fma(ly0,r0,ly1*r1)  (SURVEY.md 8(a) recipe table).  src [..., H, W] float32."""
    H, W = src.shape[-2:]
    y0, y1, ly0, ly1 = np_bilinear_ac_axis(H, out_h)
    x0, x1, lx0, lx1 = np_bilinear_ac_axis(W, out_w)
    a = src[..., y0[:, None], x0[None, :]]
    b = src[..., y0[:, None], x1[None, :]]
    c = src[..., y1[:, None], x0[None, :]]
    d = src[..., y1[:, None], x1[None, :]]
    LX0 = np.broadcast_to(lx0[None, :], a.shape)
    LX1 = np.broadcast_to(lx1[None, :], a.shape)
    LY0 = np.broadcast_to(ly0[:, None], a.shape)
    LY1 = np.broadcast_to(ly1[:, None], a.shape)
    r0 = _fma32(LX0, a, (LX1 * b).astype(np.float32))
    r1 = _fma32(LX0, c, (LX1 * d).astype(np.float32))
    return _fma32(LY0, r0, (LY1 * r1).astype(np.float32))


def np_nearest_index(n_in: int, n_out: int) -> np.ndarray:
    """ATen nearest (legacy 'nearest' mode): src=min(int(floorf(dst*(float)in/out)), in-1), scale
    computed in float32 (UpSample.h nearest_neighbor_compute_source_index)."""
    scale = np.float32(np.float32(n_in) / np.float32(n_out))
    dst = np.arange(n_out, dtype=np.float32)
    return np.minimum(np.floor((dst * scale).astype(np.float32)).astype(np.int64), n_in - 1)


def np_roi_align_1s(feat: np.ndarray, roi: np.ndarray, spatial_scale: float, out_h: int, out_w: int) -> np.ndarray:
    """torchvision roi_align CPU kernel (torchvision/csrc/ops/cpu/roi_align_kernel.cpp +
    roi_align_common.h pre_calc_for_bilinear_interpolate), aligned=True, sampling_ratio=-1, for
    the case ceil(roi/out)==1 (one sample per bin; true for every config on this path).
    feat [C,H,W] float32; roi = [x1,y1,x2,y2] float32 (already scaled by bbox_feat_factor)."""
    C, H, W = feat.shape
    f = np.float32
    s = f(spatial_scale)
    x1, y1, x2, y2 = [f(f(v) * s) - f(0.5) for v in roi]
    rw, rh = f(x2 - x1), f(y2 - y1)
    bh, bw = f(rh / f(out_h)), f(rw / f(out_w))
    gh, gw = int(np.ceil(rh / f(out_h))), int(np.ceil(rw / f(out_w)))
    assert gh == 1 and gw == 1, "restatement covers the one-sample-per-bin case only"

    def axis(start, b, n_out, n_in):
        p = np.arange(n_out, dtype=np.float32)
        c = ((start + (p * b).astype(np.float32)).astype(np.float32) + f(f(0.5) * b / f(1))).astype(np.float32)
        valid = ~((c < -1.0) | (c > n_in))
        c = np.maximum(c, f(0))
        lo = c.astype(np.int64)
        edge = lo >= n_in - 1
        lo = np.where(edge, n_in - 1, lo)
        hi = np.where(edge, n_in - 1, lo + 1)
        c = np.where(edge, lo.astype(np.float32), c)
        l = (c - lo.astype(np.float32)).astype(np.float32)
        h = (f(1) - l).astype(np.float32)
        return lo, hi, l, h, valid

    ylo, yhi, ly, hy, vy = axis(y1, bh, out_h, H)
    xlo, xhi, lx, hx, vx = axis(x1, bw, out_w, W)
    w1 = (hy[:, None] * hx[None, :]).astype(np.float32)
    w2 = (hy[:, None] * lx[None, :]).astype(np.float32)
    w3 = (ly[:, None] * hx[None, :]).astype(np.float32)
    w4 = (ly[:, None] * lx[None, :]).astype(np.float32)
    v1 = feat[:, ylo[:, None], xlo[None, :]]
    v2 = feat[:, ylo[:, None], xhi[None, :]]
    v3 = feat[:, yhi[:, None], xlo[None, :]]
    v4 = feat[:, yhi[:, None], xhi[None, :]]
    out = (w1 * v1).astype(np.float32)
    out = (out + (w2 * v2).astype(np.float32)).astype(np.float32)
    out = (out + (w3 * v3).astype(np.float32)).astype(np.float32)
    out = (out + (w4 * v4).astype(np.float32)).astype(np.float32)
    out = np.where((vy[:, None] & vx[None, :])[None], out, f(0))
    return out


# ---------------------------------------------------------------------------------------------
# helpers shared by the golden-fixture generator and the tests
# ---------------------------------------------------------------------------------------------

def fake_prediction(bbox_row: Sequence[int], ph: int, pw: int) -> Tensor:
    """Deterministic stand-in for a patch prediction, keyed by the raw bbox; lets the blend be
    exercised at full 4K / r32 size without running a network (geometry goldens)."""
    w0, h0 = int(bbox_row[0]), int(bbox_row[1])
    g = torch.Generator().manual_seed(w0 * 7919 + h0 * 13 + 5)
    base = 5.0 + 70.0 * torch.rand(1, generator=g)
    return (base + 4.0 * torch.rand(1, ph, pw, generator=g)).float()


def sha256_f32(x) -> str:
    import hashlib
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    return hashlib.sha256(a.tobytes()).hexdigest()


class GeometryOracle(PatchRefinerOracle):
    """PatchRefinerOracle with the networks replaced by ``fake_prediction`` (tiling, crop,
    bbox scaling and blend stay the reference's)."""

    def __init__(self, patch_process_shape, image_raw_shape, patch_split_num):
        self.patch_process_shape = tuple(patch_process_shape)
        self.tile_cfg = prepare_tile_cfg(self.patch_process_shape, image_raw_shape, patch_split_num)
        self.trace = None

    def coarse_forward(self, image_lr):
        return [torch.zeros(1, 1, 8, 8)], torch.zeros(1, 1, *self.patch_process_shape)

    def _predict(self, image_hr, bboxs, tile_cfg, tile_temp, process_num, record):
        bf = bboxs_to_feat(bboxs, tile_cfg["image_raw_shape"], self.patch_process_shape)
        ph, pw = self.patch_process_shape
        preds = torch.stack([fake_prediction(b, ph, pw) for b in bboxs.tolist()], dim=0)
        if record is not None:
            record["bboxs"].append(bboxs.clone())
            record["bboxs_feat"].append(bf.clone())
        return preds

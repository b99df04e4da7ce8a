"""MobileNetV4-conv-small feature encoder (the V2 family's light-weight refiner encoder) on the B200 kernels.

The reference builds this network through ``timm.create_model('mobilenetv4_conv_small.e2400_r224_in1k', features_only=True)``
(estimator/models/blocks/lightweight_refiner.py:259-262), swaps the stem for a 4-channel one
(estimator/models/patchrefinerplus.py:159-165) and feeds it ``cat[(crop - mean) / std, coarse depth]``
(lightweight_refiner.py:293-298).  timm is not part of the reference tree and not installable offline, so the arithmetic below is
restated from the published architecture (MobileNetV4 paper, table "MNv4-Conv-S"; timm's ``_gen_mobilenet_v4`` block strings quoted
next to each stage) -- parity of THIS encoder is therefore pinned only against the PyTorch restatement the test suite carries
(tests/test_mnv4.py), NOT against timm ("parity unpinned" in DESIGN.md).  State-dict key names follow timm's ``features_only``
model (``conv_stem``, ``bn1``, ``blocks.<stage>.<block>.{conv,bn1}`` for ConvBnAct, ``.{dw_start,pw_exp,dw_mid,pw_proj}.{conv,bn}``
for UniversalInvertedResidual), so a checkpoint written by the reference loads by name.

Execution: BatchNorm (eval) is folded into the convolution weights at load time; 1x1 convolutions and the 3x3 stride-2
convolutions (2x2 phase split) run on ``prv2_umma_gemm`` with the ReLU / residual epilogues, depthwise convolutions on
``prv2_dwconv``, the input normalisation + depth concatenation on ``prv2_encoder_input``.  The five feature maps come back as
channels-last acts at strides 2, 4, 8, 16, 32 with 32, 32, 64, 96, 960 channels (``fine_chl`` of configs/patchrefinerv2_dav2/plus_mobile_*)."""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

from . import _lib, ops
from .nn import Act, GemmLayer, Workspace, ptr, stream_ptr

BN_EPS = 1e-5
STEM = 32
# (kind, ...): ("cn", kernel, stride, out)  |  ("uir", dw_start kernel, dw_mid kernel, stride, expansion, out)
ARCH = [
    [("cn", 3, 2, 32), ("cn", 1, 1, 32)],                                  # cn_r1_k3_s2_e1_c32, cn_r1_k1_s1_e1_c32
    [("cn", 3, 2, 96), ("cn", 1, 1, 64)],                                  # cn_r1_k3_s2_e1_c96, cn_r1_k1_s1_e1_c64
    [("uir", 5, 5, 2, 3.0, 96)] + [("uir", 0, 3, 1, 2.0, 96)] * 4 + [("uir", 3, 0, 1, 4.0, 96)],      # uir_r1_a5_k5_s2_e3_c96, uir_r4_a0_k3_s1_e2_c96, uir_r1_a3_k0_s1_e4_c96
    [("uir", 3, 3, 2, 6.0, 128), ("uir", 5, 5, 1, 4.0, 128), ("uir", 0, 5, 1, 4.0, 128), ("uir", 0, 5, 1, 3.0, 128),
     ("uir", 0, 3, 1, 4.0, 128), ("uir", 0, 3, 1, 4.0, 128)],              # uir_r1_a3_k3_s2_e6_c128, a5_k5_e4, a0_k5_e4, a0_k5_e3, 2 x a0_k3_e4
    [("cn", 1, 1, 960)],                                                    # cn_r1_k1_s1_e1_c960
]
FEATURE_STAGES = (0, 1, 2, 4)          # after the stem: the last block of these stages (strides 4, 8, 16, 32)
OUT_CHANNELS = (32, 32, 64, 96, 960)
DEFAULT_MEAN, DEFAULT_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def make_divisible(v: float, divisor: int = 8, round_limit: float = 0.9) -> int:
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < round_limit * v:
        new_v += divisor
    return new_v


def _bn_spec(spec, name, c):
    spec[name + ".weight"] = (c,)
    spec[name + ".bias"] = (c,)
    spec[name + ".running_mean"] = (c,)
    spec[name + ".running_var"] = (c,)
    spec[name + ".num_batches_tracked"] = ()


def mnv4_conv_small_spec(in_chans: int = 4) -> "OrderedDict[str, tuple]":
    """State-dict layout (name -> shape) of the features_only encoder with an ``in_chans``-channel stem."""
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    spec["conv_stem.weight"] = (STEM, in_chans, 3, 3)
    _bn_spec(spec, "bn1", STEM)
    cin = STEM
    for si, stage in enumerate(ARCH):
        for bi, blk in enumerate(stage):
            p = f"blocks.{si}.{bi}."
            if blk[0] == "cn":
                _, k, _, cout = blk
                spec[p + "conv.weight"] = (cout, cin, k, k)
                _bn_spec(spec, p + "bn1", cout)
            else:
                _, ks, km, _, e, cout = blk
                mid = make_divisible(cin * e)
                if ks:
                    spec[p + "dw_start.conv.weight"] = (cin, 1, ks, ks)
                    _bn_spec(spec, p + "dw_start.bn", cin)
                spec[p + "pw_exp.conv.weight"] = (mid, cin, 1, 1)
                _bn_spec(spec, p + "pw_exp.bn", mid)
                if km:
                    spec[p + "dw_mid.conv.weight"] = (mid, 1, km, km)
                    _bn_spec(spec, p + "dw_mid.bn", mid)
                spec[p + "pw_proj.conv.weight"] = (cout, mid, 1, 1)
                _bn_spec(spec, p + "pw_proj.bn", cout)
            cin = cout
    return spec


def _fold_bn(sd, conv_key: str, bn_key: str):
    """(w', b') of conv followed by eval-mode BatchNorm: w' = w * g / sqrt(var + eps), b' = beta - mean * g / sqrt(var + eps) (fp64 fold)."""
    w = sd[conv_key + ".weight"].detach().double()
    g, b = sd[bn_key + ".weight"].detach().double(), sd[bn_key + ".bias"].detach().double()
    m, v = sd[bn_key + ".running_mean"].detach().double(), sd[bn_key + ".running_var"].detach().double()
    s = g / torch.sqrt(v + BN_EPS)
    return (w * s.view(-1, 1, 1, 1)).float(), (b - m * s).float()


class _DwLayer:
    def __init__(self, w: torch.Tensor, b: torch.Tensor, k: int, stride: int, relu: bool, device):
        c = w.shape[0]
        self.k, self.stride, self.relu, self.C = k, stride, relu, c
        self.w = w.reshape(c, k * k).t().contiguous().to(device)       # [k*k, C] tap-major (include/prv2_b200.h)
        self.b = b.contiguous().to(device)

    def __call__(self, a: Act, out: Act) -> Act:
        assert a.C == self.C and out.C == self.C
        planes = 2 if out.lo is not None else 1
        _lib.call("prv2_dwconv", ptr(a.hi), ptr(a.lo), a.N, a.H, a.W, a.C, a.cs, ptr(self.w), ptr(self.b), self.k, self.stride, 1 if self.relu else 0,
                  ptr(out.hi), ptr(out.lo), out.cs, stream_ptr(), work=("byte", 2.0 * planes * a.N * a.C * (a.H * a.W + out.H * out.W)))
        return out


def _s2_segments(w: torch.Tensor):
    """3x3 stride-2 pad-1 conv on 2x2 phase-split sources: in(2y+r-1, 2x+s-1) = phase[(r-1)&1][(s-1)&1] at offset (r==0 ? -1 : 0)."""
    segs = []
    for r in range(3):
        for s in range(3):
            py, px = (r - 1) & 1, (s - 1) & 1
            segs.append((py * 2 + px, -1 if r == 0 else 0, -1 if s == 0 else 0, w[:, :, r, s]))
    return segs


class MobileNetV4ConvSmallB200:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, in_chans: int, x3: bool, device, mean=DEFAULT_MEAN, std=DEFAULT_STD):
        self.x3, self.device, self.in_chans = x3, device, in_chans
        assert in_chans in (3, 4)
        self.mean = (C.c_float * 3)(*mean)
        self.std = (C.c_float * 3)(*std)
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
        mk = lambda segs, n_src, cout, **kw: GemmLayer(segs, n_src, cout, x3, device, **kw)
        relu = _lib.ACT_RELU

        def conv_layer(conv_key, bn_key, k, stride, act, name):
            w, b = _fold_bn(sd, conv_key, bn_key)
            if k == 1:
                return mk([(0, 0, 0, w[:, :, 0, 0])], 1, w.shape[0], act=act, bias=b, name=name)
            assert k == 3 and stride == 2
            if w.shape[1] % 8:                                            # the 4-channel stem reads an 8-channel (zero padded) act
                w = torch.cat([w, torch.zeros(w.shape[0], 8 - w.shape[1] % 8, 3, 3)], dim=1)
            return mk(_s2_segments(w), 4, w.shape[0], act=act, bias=b, name=name)

        self.stem = conv_layer("conv_stem", "bn1", 3, 2, relu, "enc.stem")
        self.blocks: List[List[dict]] = []
        cin = STEM
        for si, stage in enumerate(ARCH):
            row = []
            for bi, blk in enumerate(stage):
                p = f"blocks.{si}.{bi}."
                if blk[0] == "cn":
                    _, k, stride, cout = blk
                    row.append(dict(kind="cn", k=k, stride=stride, cout=cout, conv=conv_layer(p + "conv", p + "bn1", k, stride, relu, f"enc.s{si}.cn")))
                else:
                    _, ks, km, stride, e, cout = blk
                    mid = make_divisible(cin * e)
                    d = dict(kind="uir", ks=ks, km=km, stride=stride, mid=mid, cout=cout, skip=(cin == cout and stride == 1))
                    if ks:
                        w, b = _fold_bn(sd, p + "dw_start.conv", p + "dw_start.bn")
                        d["dw_start"] = _DwLayer(w[:, 0], b, ks, stride if not km else 1, False, device)          # no activation after dw_start
                    d["pw_exp"] = conv_layer(p + "pw_exp.conv", p + "pw_exp.bn", 1, 1, relu, f"enc.s{si}.pw_exp")
                    if km:
                        w, b = _fold_bn(sd, p + "dw_mid.conv", p + "dw_mid.bn")
                        d["dw_mid"] = _DwLayer(w[:, 0], b, km, stride, True, device)
                    d["pw_proj"] = conv_layer(p + "pw_proj.conv", p + "pw_proj.bn", 1, 1, _lib.ACT_NONE, f"enc.s{si}.pw_proj")
                    row.append(d)
                cin = cout
            self.blocks.append(row)

    @staticmethod
    def _half(n: int, k: int, stride: int) -> int:
        return (n + 2 * (k // 2) - k) // stride + 1

    def _conv_s2(self, ws: Workspace, name: str, a: Act, layer: GemmLayer, cout: int) -> Act:
        c8 = (a.C + 7) // 8 * 8
        src = a.view_channels(c8)
        h2, w2 = (a.H + 1) // 2, (a.W + 1) // 2
        ph = ws.act(name + "_phase", 4 * a.N, h2, w2, c8)
        ops.phase_split(src, ph)
        out = ws.act(name, a.N, h2, w2, cout)
        layer([ph.batch_slice(slice(k * a.N, (k + 1) * a.N)) for k in range(4)], out=out)
        return out

    def forward(self, crops: torch.Tensor, depth: Optional[torch.Tensor], ws: Workspace) -> List[Act]:
        """crops [N,3,H,W] fp32 in [0,1]; depth [N,1,H,W] fp32 (coarse_condition) or None -> the 5 features, finest first."""
        N, _, H, W = crops.shape
        assert (depth is not None) == (self.in_chans == 4)
        x = ws.act("enc_in", N, H, W, 8)                                   # channels 0..2 colour, 3 depth (or zero), 4..7 zero
        _lib.call("prv2_encoder_input", ptr(crops.contiguous()), ptr(None if depth is None else depth.contiguous()), N, H, W, self.mean, self.std,
                  ptr(x.hi), ptr(x.lo), x.cs, stream_ptr(), work=("byte", N * H * W * (16.0 + 16.0 * (2 if x.lo is not None else 1))))
        a = self._conv_s2(ws, "enc_stem", x, self.stem, STEM)
        feats = [a]
        for si, row in enumerate(self.blocks):
            for bi, d in enumerate(row):
                tag = f"enc_s{si}b{bi}"
                if d["kind"] == "cn":
                    if d["k"] == 3:
                        a = self._conv_s2(ws, tag, a, d["conv"], d["cout"])
                    else:
                        o = ws.act(tag, a.N, a.H, a.W, d["cout"])
                        d["conv"]([a], out=o)
                        a = o
                    continue
                x_in = a
                t = a
                if d["ks"]:
                    s = d["dw_start"].stride
                    o = ws.act(tag + "_dws", t.N, self._half(t.H, d["ks"], s), self._half(t.W, d["ks"], s), t.C)
                    t = d["dw_start"](t, o)
                o = ws.act(tag + "_exp", t.N, t.H, t.W, d["mid"])
                d["pw_exp"]([t], out=o)
                t = o
                if d["km"]:
                    s = d["stride"]
                    o = ws.act(tag + "_dwm", t.N, self._half(t.H, d["km"], s), self._half(t.W, d["km"], s), t.C)
                    t = d["dw_mid"](t, o)
                o = ws.act(tag, t.N, t.H, t.W, d["cout"])
                d["pw_proj"]([t], out=o, res=x_in if d["skip"] else None)
                a = o
            if si in FEATURE_STAGES:
                feats.append(a)
        return feats

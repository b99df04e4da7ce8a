"""BiDirectionalFusion -- the fusion model of the V2 family (``PatchRefinerPlus``) -- on the B200 kernels
(estimator/models/blocks/bi_directional_fusion_model.py:285-446; SURVEY.md row a9').

Two halves, both dense 3x3 / 1x1 convolutions and therefore tcgen05 implicit GEMMs here:

* coarse -> fine (``C2FModule`` :148-208): a DPT-style top-down decoder over the light-weight encoder's features in
  which every residual unit is a ``GatedConvUnit`` (:24-82): ``a = conv(relu(x)) + x``, then
  ``a * sigmoid(conv1x1(relu(LN(conv3x3(cat[a, coarse_roi])))))``.  Per unit that is three launches of
  ``prv2_umma_gemm``: STORE (+bias, +residual), LN epilogue with bias and ReLU over the virtual concat, and the
  ``SIGMOID_GATE`` epilogue that multiplies the gate into ``a`` (and adds the block's skip, and emits the ReLU copy
  the next unit consumes).
* fine -> coarse (:416-446): the same U-Net as ``FusionUnet`` with different module names and an uneven
  (coarse, fine) channel split, so ``FusionUnetB200`` runs it.

``BiDirectionalFusion`` (registered in ``MODELS`` under the reference's name) is the operator-level drop-in: same
constructor keywords, same state-dict keys, ``forward(c_feat, f_feat, pred1, pred2, update_base)`` on CUDA NCHW
tensors.  The light-weight encoder that produces ``f_feat`` in the reference is a timm model whose arithmetic is not
available offline (parity unpinned); callers run it themselves and hand its features in.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib, ops
from .fusion import LN_EPS, FusionUnetB200
from .nn import Act, GemmLayer, Workspace, conv_segments
from .registry import MODELS

#: coarse2fine_type -> (fusion, gate) of the C2FModule (bi_directional_fusion_model.py:366-374)
C2F_TYPES = {"self-agg": (False, False), "coarse-gated": (True, True), "coarse-fusion": (True, False),
             "only-gate": (True, False)}       # only-gate = C2FNOENCModule(fusion=True, gate=False) (:211-283, :372-373)
C2F_FEATURES = 256                 # C2FModule(features=256) (:149)


def bifusion_weight_spec(coarse_chl, fine_chl, fine_chl_after_coarse2fine, temp_chl, dec_chl, coarse2fine_type="coarse-gated",
                         features: int = C2F_FEATURES, heavy: bool = False) -> "OrderedDict[str, tuple]":
    """State-dict names and shapes of BiDirectionalFusion / BiDirectionalFusionHeavy (glb_att=False, coarse2fine=True)."""
    fusion, _ = C2F_TYPES[coarse2fine_type]
    s: "OrderedDict[str, tuple]" = OrderedDict()
    for name, extra in (("fusion_layers_1", None), ("fusion_layers_2", 2)):
        for idx, (cc, fc, tc) in enumerate(zip(coarse_chl, fine_chl_after_coarse2fine, temp_chl)):
            cin = cc + fc if extra is None else tc + extra
            s[f"{name}.{idx}.single_conv.0.weight"] = (tc, cin, 3, 3)
            s[f"{name}.{idx}.single_conv.1.weight"] = (tc,)
            s[f"{name}.{idx}.single_conv.1.bias"] = (tc,)
            if heavy:                                                    # SingleConvCNNLNHeavy (:448-463)
                s[f"{name}.{idx}.single_conv.2.weight"] = (tc, tc, 3, 3)
                s[f"{name}.{idx}.single_conv.3.weight"] = (tc,)
                s[f"{name}.{idx}.single_conv.3.bias"] = (tc,)
                s[f"{name}.{idx}.single_conv.4.weight"] = (tc, tc, 3, 3)
    rev = list(temp_chl)[::-1]
    chl = rev[0]
    for i, (tc, dc) in enumerate(zip(rev[1:], dec_chl)):
        cin = tc + chl + 2
        s[f"f2r_agg.{i}.conv.double_conv.0.weight"] = (cin, cin, 3, 3)
        if heavy:                                                        # DoubleConvHeavy (:465-485)
            for k in (2, 4, 6):
                s[f"f2r_agg.{i}.conv.double_conv.{k}.weight"] = (cin, cin, 3, 3)
            s[f"f2r_agg.{i}.conv.double_conv.8.weight"] = (dc, cin, 3, 3)
        else:
            s[f"f2r_agg.{i}.conv.double_conv.2.weight"] = (dc, cin, 3, 3)
        chl = dc
    s["final_conv.weight"] = (1, dec_chl[-1] if len(dec_chl) else chl, 3, 3)

    def unit(pre, f):
        s[pre + "conv.weight"] = (f, f, 3, 3)
        s[pre + "conv.bias"] = (f,)
        if fusion:
            s[pre + "fusion_conv.0.weight"] = (f, 2 * f, 3, 3)
            s[pre + "fusion_conv.0.bias"] = (f,)
            s[pre + "fusion_conv.1.weight"] = (f,)
            s[pre + "fusion_conv.1.bias"] = (f,)
            s[pre + "fusion_conv.3.weight"] = (f, f, 1, 1)

    def block(pre, f):
        s[pre + "out_conv.weight"] = (f, f, 1, 1)
        s[pre + "out_conv.bias"] = (f,)
        unit(pre + "GateresConfUnit1.", f)
        unit(pre + "GateresConfUnit2.", f)

    q = "c2f.scratch."
    for i, fc in enumerate(fine_chl):
        s[f"{q}layer{i + 1}_rn.weight"] = (features, fc, 3, 3)
    if coarse2fine_type == "only-gate":                                  # C2FNOENCModule (:211-251)
        for lvl in range(1, 6):
            unit(f"{q}layer{lvl}_gate1.", features)
            unit(f"{q}layer{lvl}_gate2.", features)
        s[q + "upsample_conv.0.weight"] = (fine_chl[0], 32, 2, 2)        # ConvTranspose2d [Cin, Cout, k, k]
        s[q + "upsample_conv.0.bias"] = (32,)
        s[q + "upsample_conv.2.weight"] = (32, 32, 3, 3)
        unit(q + "layer6_gate1.", 32)
        unit(q + "layer6_gate2.", 32)
        s[q + "output_conv.weight"] = (1, 32, 3, 3)
        s[q + "output_conv.bias"] = (1,)
        return s
    for i in range(1, 6):
        block(f"{q}refinenet{i}.", features)
    h2 = coarse_chl[0]
    s[q + "output_conv1.weight"] = (features // 2, features, 3, 3)
    s[q + "output_conv1.bias"] = (features // 2,)
    s[q + "output_conv2.0.weight"] = (h2, features // 2, 3, 3)
    s[q + "output_conv2.0.bias"] = (h2,)
    block(q + "output_conv2_fusion.", h2)
    s[q + "output_conv3.0.weight"] = (1, h2, 1, 1)
    s[q + "output_conv3.0.bias"] = (1,)
    return s


class _GatedUnit:
    """GatedConvUnit (bi_directional_fusion_model.py:24-82) as three GEMM launches."""

    def __init__(self, g, pre: str, feat: int, fusion: bool, gate: bool, x3: bool, device, name: str):
        mk = lambda segs, n_src, cout, **kw: GemmLayer(segs, n_src, cout, x3, device, **kw)
        self.feat, self.fusion, self.gate = feat, fusion, gate
        self.conv = mk(conv_segments(g(pre + "conv.weight"), [feat]), 1, feat, bias=g(pre + "conv.bias"), name=name + ".conv")
        if fusion:
            self.fuse = mk(conv_segments(g(pre + "fusion_conv.0.weight"), [feat, feat]), 2, feat, epi=_lib.EPI_LN_GELU, act=_lib.ACT_RELU,
                           bias=g(pre + "fusion_conv.0.bias"), gamma=g(pre + "fusion_conv.1.weight"), beta=g(pre + "fusion_conv.1.bias"),
                           eps=LN_EPS, name=name + ".fuse")
            self.mix = mk([(0, 0, 0, g(pre + "fusion_conv.3.weight")[:, :, 0, 0])], 1, feat,
                          act=_lib.ACT_SIGMOID_GATE if gate else _lib.ACT_NONE, name=name + ".gate")

    def flops_per_pixel(self) -> float:
        f = self.conv.flops_per_pixel
        if self.fusion:
            f += self.fuse.flops_per_pixel + self.mix.flops_per_pixel
        return float(f)

    def __call__(self, A, tag: str, x: Act, x_relu: Act, c: Optional[Act], skip: Optional[Act], out: Act, out_relu: Optional[Act]) -> None:
        """out = unit(x, c) [+ skip]; out_relu = relu(out) when requested."""
        if not self.fusion:
            self.conv([x_relu], out=out, relu_out=out_relu, res=x, res2=skip)
            return
        a = A(tag + "_a", x.N, x.H, x.W, self.feat)
        self.conv([x_relu], out=a, res=x)                                              # :62-68
        f = A(tag + "_f", x.N, x.H, x.W, self.feat)
        self.fuse([a, c], out=f)                                                       # :71-72 conv3x3(cat) + b -> LN -> ReLU
        if self.gate:
            self.mix([f], out=out, relu_out=out_relu, res=a, res2=skip)                # :75-77 a * sigmoid(conv1x1(f)) [+ skip]
        else:
            self.mix([f], out=out, relu_out=out_relu, res=skip)                        # :78-79 out = fused_feat [+ skip]


class _GatedBlock:
    """GatedFusionBlock (bi_directional_fusion_model.py:84-146)."""

    def __init__(self, g, pre: str, feat: int, fusion: bool, gate: bool, x3: bool, device, name: str, two_inputs: bool):
        self.feat = feat
        self.u1 = _GatedUnit(g, pre + "GateresConfUnit1.", feat, fusion, gate, x3, device, name + ".u1") if two_inputs else None
        self.u2 = _GatedUnit(g, pre + "GateresConfUnit2.", feat, fusion, gate, x3, device, name + ".u2")
        self.out_conv = GemmLayer([(0, 0, 0, g(pre + "out_conv.weight")[:, :, 0, 0])], 1, feat, x3, device, bias=g(pre + "out_conv.bias"),
                                  name=name + ".out_conv")

    def __call__(self, A, tag: str, x0: Act, x0_relu: Optional[Act], skip_in: Optional[Act], skip_in_relu: Optional[Act], c: Optional[Act],
                 size, out: Act) -> Act:
        """x0 = xs[0]; (skip_in, skip_in_relu) = xs[1] and its ReLU copy when the block has two inputs."""
        if self.u1 is not None:
            s = A(tag + "_sum", x0.N, x0.H, x0.W, self.feat)
            s_relu = A(tag + "_sum_relu", x0.N, x0.H, x0.W, self.feat)
            self.u1(A, tag + "_u1", skip_in, skip_in_relu, c, x0, s, s_relu)            # output = xs[0] + unit1(xs[1])  (:125-127)
            x0, x0_relu = s, s_relu
        y = A(tag + "_y", x0.N, x0.H, x0.W, self.feat)
        self.u2(A, tag + "_u2", x0, x0_relu, c, None, y, None)                         # :129
        if size is not None and tuple(size) != (y.H, y.W):
            y = ops.resize_bilinear(y, A(tag + "_up", y.N, size[0], size[1], self.feat))   # :131-140 bilinear, align_corners=True
        self.out_conv([y], out=out)                                                    # :142
        return out


class BiDirectionalFusionB200:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, coarse_chl: Sequence[int], fine_chl: Sequence[int],
                 fine_chl_after_coarse2fine: Sequence[int], temp_chl: Sequence[int], dec_chl: Sequence[int], coarse2fine_type: str,
                 x3: bool, device, features: int = C2F_FEATURES, heavy: bool = False):
        if coarse2fine_type not in C2F_TYPES:
            raise NotImplementedError(f"coarse2fine_type={coarse2fine_type!r}: implemented {sorted(C2F_TYPES)}")
        fusion, gate = C2F_TYPES[coarse2fine_type]
        self.x3, self.device, self.features = x3, device, features
        self.only_gate = coarse2fine_type == "only-gate"
        self.coarse_chl, self.fine_chl = list(coarse_chl), list(fine_chl)
        assert len(self.coarse_chl) == 6 and len(self.fine_chl) == 5
        if fusion:
            assert all(c == features for c in self.coarse_chl[1:]), "gated C2F concatenates coarse maps with `features`-channel paths"
        assert list(fine_chl_after_coarse2fine) == [self.coarse_chl[0]] + [features] * 5, "fine_chl_after_coarse2fine is what C2FModule emits"
        g = lambda k: sd[prefix + k].detach().float()
        mk = lambda segs, n_src, cout, **kw: GemmLayer(segs, n_src, cout, x3, device, **kw)
        q = "c2f.scratch."
        self.layer_rn = [mk(conv_segments(g(f"{q}layer{i + 1}_rn.weight"), [fc]), 1, features, name=f"c2f.layer{i + 1}_rn")
                         for i, fc in enumerate(self.fine_chl)]
        after = list(fine_chl_after_coarse2fine)
        self.f2c = FusionUnetB200(sd, prefix, [c + f for c, f in zip(self.coarse_chl, after)], temp_chl, dec_chl, x3, device,
                                  in_splits=list(zip(self.coarse_chl, after)), names=("fusion_layers_1", "fusion_layers_2", "f2r_agg"), heavy=heavy)
        self.ws: Dict[tuple, Workspace] = {}
        if self.only_gate:
            # C2FNOENCModule (:211-283): no top-down path -- two gated units per level against that level's coarse map, and a
            # transposed-conv level 0 (deconv k = stride = 2 -> ReLU -> 3x3 conv) with two 32-channel units
            assert self.coarse_chl[0] == 32, "C2FNOENCModule hard-codes a 32-channel level 0 (:238-247)"
            self.gates = {lvl: (_GatedUnit(g, f"{q}layer{lvl}_gate1.", features, fusion, gate, x3, device, f"c2f.layer{lvl}_gate1"),
                                _GatedUnit(g, f"{q}layer{lvl}_gate2.", features, fusion, gate, x3, device, f"c2f.layer{lvl}_gate2")) for lvl in range(1, 6)}
            self.gates[6] = (_GatedUnit(g, q + "layer6_gate1.", 32, fusion, gate, x3, device, "c2f.layer6_gate1"),
                             _GatedUnit(g, q + "layer6_gate2.", 32, fusion, gate, x3, device, "c2f.layer6_gate2"))
            wt = g(q + "upsample_conv.0.weight")                             # [Cin, 32, 2, 2]
            self.up_deconv = mk([(0, 0, 0, wt.permute(2, 3, 1, 0).reshape(4 * 32, wt.shape[0]))], 1, 4 * 32, epi=_lib.EPI_SHUFFLE, act=_lib.ACT_RELU,
                                bias=g(q + "upsample_conv.0.bias"), shuffle_k=2, name="c2f.upsample_deconv")
            self.up_conv = mk(conv_segments(g(q + "upsample_conv.2.weight"), [32]), 1, 32, name="c2f.upsample_conv")
            self.out_ng = mk(conv_segments(g(q + "output_conv.weight"), [32]), 1, 1, epi=_lib.EPI_F32, bias=g(q + "output_conv.bias"), name="c2f.output_conv")
            return
        self.refine = {i: _GatedBlock(g, f"{q}refinenet{i}.", features, fusion, gate, x3, device, f"c2f.refinenet{i}", two_inputs=i != 5)
                       for i in range(1, 6)}
        h2 = self.coarse_chl[0]
        self.h2 = h2
        self.out1 = mk(conv_segments(g(q + "output_conv1.weight"), [features]), 1, features // 2, bias=g(q + "output_conv1.bias"), name="c2f.output_conv1")
        self.out2 = mk(conv_segments(g(q + "output_conv2.0.weight"), [features // 2]), 1, h2, act=_lib.ACT_RELU, bias=g(q + "output_conv2.0.bias"),
                       name="c2f.output_conv2")
        self.out2_fusion = _GatedBlock(g, q + "output_conv2_fusion.", h2, fusion, gate, x3, device, "c2f.output_conv2_fusion", two_inputs=False)
        self.out3 = mk([(0, 0, 0, g(q + "output_conv3.0.weight")[:, :, 0, 0])], 1, 1, epi=_lib.EPI_F32, bias=g(q + "output_conv3.0.bias"),
                       name="c2f.output_conv3")

    def flops(self, B: int, sizes) -> float:
        """sizes: (h, w) of the six fine levels, finest first (level 0 = twice level 1)."""
        F_ = self.features
        px = [B * h * w for h, w in sizes]
        f = sum(2.0 * px[i + 1] * 9 * fc * F_ for i, fc in enumerate(self.fine_chl))
        if self.only_gate:
            for lvl in range(1, 6):                                  # layer<lvl>_gate* work on level 6 - lvl
                f += px[6 - lvl] * (self.gates[lvl][0].flops_per_pixel() + self.gates[lvl][1].flops_per_pixel())
            f += px[1] * 2.0 * self.fine_chl[0] * 128 + px[0] * (2.0 * 9 * 32 * 32 + self.gates[6][0].flops_per_pixel() + self.gates[6][1].flops_per_pixel() + 2.0 * 9 * 32)
            return f + self.f2c.flops(B, sizes)
        for r in range(1, 6):                                    # refinenet r works at level r, its out_conv at level r-1
            blk = self.refine[r]
            units = (blk.u1.flops_per_pixel() if blk.u1 else 0.0) + blk.u2.flops_per_pixel()
            f += px[r] * units + px[r - 1] * 2.0 * F_ * F_
        f += px[0] * (2.0 * 9 * F_ * (F_ // 2) + 2.0 * 9 * (F_ // 2) * self.h2 + self.out2_fusion.u2.flops_per_pixel() + 2.0 * self.h2 * self.h2 + 2.0 * self.h2)
        return f + self.f2c.flops(B, sizes)

    def forward(self, c_feat: List[Act], f_feat: List[Act], pred1: torch.Tensor, pred2: Optional[torch.Tensor],
                update_base: Optional[torch.Tensor], trace: Optional[dict] = None) -> torch.Tensor:
        """bi_directional_fusion_model.py:379-446.  c_feat / f_feat: six acts each, finest first; pred1 / update_base fp32
        [B,1,H,W].  ``pred2`` is accepted for signature parity and ignored: the reference overwrites it with the C2F
        module's depth (:409-414)."""
        assert len(c_feat) == 6 and len(f_feat) == 6
        B = pred1.shape[0]
        key = (B,) + tuple((f.H, f.W) for f in f_feat)
        ws = self.ws.setdefault(key, Workspace(self.device, self.x3))
        A = ws.act
        if (c_feat[-1].H, c_feat[-1].W) != (f_feat[-1].H, f_feat[-1].W):              # :392-395 (a same-size resize is the identity)
            c_feat = [c if (c.H, c.W) == (f.H, f.W) else ops.resize_bilinear(c, A(f"c_rs{i}", B, f.H, f.W, c.C))
                      for i, (c, f) in enumerate(zip(c_feat, f_feat))]
        Fe = self.features
        fine = f_feat[1:]
        rn, rn_relu = [], []
        for i, src in enumerate(fine):                                                 # :188-192
            o, orl = A(f"l{i + 1}_rn", B, src.H, src.W, Fe), A(f"l{i + 1}_rn_relu", B, src.H, src.W, Fe)
            self.layer_rn[i]([src], out=o, relu_out=orl)
            rn.append(o); rn_relu.append(orl)
        if self.only_gate:
            def two_units(lvl, x, x_relu, c, feat):
                u1, u2 = self.gates[lvl]
                m, m_relu = A(f"g{lvl}_m", B, x.H, x.W, feat), A(f"g{lvl}_m_relu", B, x.H, x.W, feat)
                u1(A, f"g{lvl}_u1", x, x_relu, c, None, m, m_relu)                          # :262-263 (gate1 then gate2 on the same coarse map)
                o = A(f"g{lvl}_o", B, x.H, x.W, feat)
                u2(A, f"g{lvl}_u2", m, m_relu, c, None, o, None)
                return o
            paths = [two_units(lvl, rn[5 - lvl], rn_relu[5 - lvl], c_feat[6 - lvl], Fe) for lvl in range(1, 6)]   # layer1_* : layer_5_rn with coarse[5], ...
            H0, W0 = fine[0].H * 2, fine[0].W * 2
            if (c_feat[0].H, c_feat[0].W) != (H0, W0):
                raise ValueError(f"finest coarse map is {(c_feat[0].H, c_feat[0].W)} but the transposed conv ends at {(H0, W0)}")
            up_relu = A("ng_up_relu", B, H0, W0, 32)
            self.up_deconv([fine[0]], out=up_relu)                                         # :256 ConvTranspose2d(k=2, s=2) + bias -> ReLU
            l0, l0_relu = A("ng_l0", B, H0, W0, 32), A("ng_l0_relu", B, H0, W0, 32)
            self.up_conv([up_relu], out=l0, relu_out=l0_relu)                              # 3x3 conv, no bias
            p0 = two_units(6, l0, l0_relu, c_feat[0], 32)
            depth = ws.f32("c2f_depth", B, 1, H0, W0)
            self.out_ng([p0], out_f32=depth, out_f32_ld=1)                                 # :280 output_conv (3x3, bias)
            feats = [p0] + paths[::-1]                                                     # [::-1] of [path_5 .. path_1, path_0]
            if trace is not None:
                trace["c2f_feats"] = [t.to_nchw() for t in feats]
                trace["c2f_depth"] = depth.clone()
            return self.f2c.forward(c_feat, feats, pred1, depth, update_base, trace)
        size_of = lambda a: (a.H, a.W)
        path_5 = self.refine[5](A, "r5", rn[4], rn_relu[4], None, None, c_feat[5], size_of(rn[3]), A("path_5", B, rn[3].H, rn[3].W, Fe))
        path_4 = self.refine[4](A, "r4", path_5, None, rn[3], rn_relu[3], c_feat[4], size_of(rn[2]), A("path_4", B, rn[2].H, rn[2].W, Fe))
        path_3 = self.refine[3](A, "r3", path_4, None, rn[2], rn_relu[2], c_feat[3], size_of(rn[1]), A("path_3", B, rn[1].H, rn[1].W, Fe))
        path_2 = self.refine[2](A, "r2", path_3, None, rn[1], rn_relu[1], c_feat[2], size_of(rn[0]), A("path_2", B, rn[0].H, rn[0].W, Fe))
        H0, W0 = rn[0].H * 2, rn[0].W * 2                                              # refinenet1: scale_factor=2 (:133-134)
        path_1 = self.refine[1](A, "r1", path_2, None, rn[0], rn_relu[0], c_feat[1], (H0, W0), A("path_1", B, H0, W0, Fe))
        if (c_feat[0].H, c_feat[0].W) != (H0, W0):
            raise ValueError(f"finest coarse map is {(c_feat[0].H, c_feat[0].W)} but the C2F decoder ends at {(H0, W0)}")
        o1 = A("c2f_out1", B, H0, W0, Fe // 2)
        self.out1([path_1], out=o1)                                                    # :201
        last0 = A("c2f_last0", B, H0, W0, self.h2)
        self.out2([o1], out=last0)                                                     # :202 (ReLU: last0 is its own ReLU copy)
        last = self.out2_fusion(A, "o2f", last0, last0, None, None, c_feat[0], None, A("c2f_last", B, H0, W0, self.h2))   # :203
        depth = ws.f32("c2f_depth", B, 1, H0, W0)
        self.out3([last], out_f32=depth, out_f32_ld=1)                                 # :204
        feats = [last, path_2, path_3, path_4, path_5, rn[4]]                          # [::-1] of :207
        if trace is not None:
            trace["c2f_feats"] = [t.to_nchw() for t in feats]
            trace["c2f_depth"] = depth.clone()
        return self.f2c.forward(c_feat, feats, pred1, depth, update_base, trace)


@MODELS.register_module()
class BiDirectionalFusion(nn.Module):
    """Operator-level drop-in for the reference's registered ``BiDirectionalFusion`` (same keywords, same state dict,
    same forward arguments); CUDA only.  ``glb_att=True`` (dead in every shipped config) is not implemented."""
    HEAVY = False

    def __init__(self, encoder_name="", coarse2fine=True, coarse2fine_type="self-agg", fine2coarse=True,
                 coarse_chl=(32, 256, 256, 256, 256, 256), fine_chl=(32, 32, 64, 96, 960),
                 fine_chl_after_coarse2fine=(32, 256, 256, 256, 256, 256), temp_chl=(32, 64, 64, 128, 256, 512),
                 dec_chl=(512, 256, 128, 64, 32), glb_att=False, att_dim=256, select_feat_index=(-1,), pe_type="none",
                 precision: str = "bf16"):
        super().__init__()
        self.heavy = type(self).HEAVY
        if glb_att or not coarse2fine:
            raise NotImplementedError("BiDirectionalFusion: glb_att=True / coarse2fine=False are not implemented")
        if coarse2fine_type not in C2F_TYPES:
            raise NotImplementedError(f"coarse2fine_type={coarse2fine_type!r}: implemented {sorted(C2F_TYPES)}")
        assert precision in ("bf16", "fp32")
        self.encoder_name, self.glb_att, self.coarse2fine_type, self.precision = encoder_name, False, coarse2fine_type, precision
        self.cfg = dict(coarse_chl=list(coarse_chl), fine_chl=list(fine_chl), fine_chl_after_coarse2fine=list(fine_chl_after_coarse2fine),
                        temp_chl=list(temp_chl), dec_chl=list(dec_chl))
        self._weights: "OrderedDict[str, torch.Tensor]" = OrderedDict(
            (k, torch.zeros(shp)) for k, shp in bifusion_weight_spec(coarse2fine_type=coarse2fine_type, heavy=self.heavy, **self.cfg).items())
        self._engine: Optional[BiDirectionalFusionB200] = None

    def state_dict(self, *args, **kwargs):
        return OrderedDict((k, v.clone()) for k, v in self._weights.items())

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        missing = [k for k in self._weights if k not in state_dict]
        unexpected = [k for k in state_dict if k not in self._weights]
        if strict and (missing or unexpected):
            raise RuntimeError(f"missing keys {missing[:4]}..., unexpected keys {unexpected[:4]}...")
        for k, v in state_dict.items():
            if k in self._weights:
                if tuple(v.shape) != tuple(self._weights[k].shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(v.shape)} vs {tuple(self._weights[k].shape)}")
                self._weights[k] = v.detach().float().cpu().clone()
        self._engine = None
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def engine(self, device) -> BiDirectionalFusionB200:
        if device.type != "cuda":
            raise RuntimeError("patchrefinerv2_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if self._engine is None or self._engine.device != device:
            _lib.load()
            self._engine = BiDirectionalFusionB200(self._weights, "", coarse2fine_type=self.coarse2fine_type, x3=self.precision == "fp32",
                                                   device=device, heavy=self.heavy, **self.cfg)
        return self._engine

    @torch.no_grad()
    def forward(self, c_feat, f_feat, pred1, pred2=None, update_base=None, pe_list=None, pe_patch_list=None, trace=None):
        dev = pred1.device
        eng = self.engine(dev)
        x3 = eng.x3
        to_act = lambda t: t if isinstance(t, Act) else Act.from_nchw(t.to(dev), x3)
        c = [to_act(t) for t in c_feat]
        f = [to_act(t) for t in f_feat]
        out = eng.forward(c, f, pred1.float().contiguous(), None, None if update_base is None else update_base.float().contiguous(), trace)
        return out.clone()


@MODELS.register_module()
class BiDirectionalFusionHeavy(BiDirectionalFusion):
    """``BiDirectionalFusionHeavy`` (bi_directional_fusion_model.py:517-677): the same model with three-conv encoder blocks
    (SingleConvCNNLNHeavy) and five-conv decoder blocks (DoubleConvHeavy)."""
    HEAVY = True

"""FusionUnet on the B200 kernels (estimator/models/blocks/fusion_model.py:52-122, convs.py).

Every concatenation is virtual: the conv kernel walks several source tensors in its K loop, and
the two injected depth maps live in an 8-channel slot appended to the feature tensor they are
concatenated with (2 real channels + 6 zeros)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib, ops
from .nn import Act, GemmLayer, Workspace, conv_segments

LN_EPS = 1e-6   # convs.py:11


class FusionUnetB200:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, input_chl, temp_chl, dec_chl, x3: bool, device):
        self.x3, self.device = x3, device
        self.input_chl, self.temp_chl, self.dec_chl = list(input_chl), list(temp_chl), list(dec_chl)
        self.ws: Dict[tuple, Workspace] = {}
        g = lambda k: sd[prefix + k].detach().float()
        mk = lambda segs, n_src, cout, **kw: GemmLayer(segs, n_src, cout, x3, device, **kw)
        self.enc1, self.enc2 = [], []
        for idx, (ic, tc) in enumerate(zip(self.input_chl, self.temp_chl)):
            assert ic % 2 == 0
            q = f"encoder_layers_1.{idx}.single_conv."
            self.enc1.append(mk(conv_segments(g(q + "0.weight"), [ic // 2, ic // 2]), 2, tc, epi=_lib.EPI_LN_GELU,
                                gamma=g(q + "1.weight"), beta=g(q + "1.bias"), eps=LN_EPS, name=f"fusion.enc1.L{idx}"))
            q = f"encoder_layers_2.{idx}.single_conv."
            self.enc2.append(mk(conv_segments(g(q + "0.weight"), [tc + 2]), 1, tc, epi=_lib.EPI_LN_GELU,
                                gamma=g(q + "1.weight"), beta=g(q + "1.bias"), eps=LN_EPS, name=f"fusion.enc2.L{idx}"))
        self.dec = []
        rev = self.temp_chl[::-1]
        chl = rev[0]
        for i, (tc, dc) in enumerate(zip(rev[1:], self.dec_chl)):
            cin = tc + chl + 2
            q = f"decoder_layers.{i}.conv.double_conv."
            c1 = mk(conv_segments(g(q + "0.weight"), [chl, tc + 2]), 2, cin, act=_lib.ACT_GELU, name=f"fusion.dec{i}.conv1")
            c2 = mk(conv_segments(g(q + "2.weight"), [cin]), 1, dc, act=_lib.ACT_GELU, name=f"fusion.dec{i}.conv2")
            self.dec.append((c1, c2, cin, dc))
            chl = dc
        wf = g("final_conv.weight")                       # [1, C, 3, 3] -> 1x1 conv with one output per tap: [9, C]
        self.final_c = wf.shape[1]
        self.final_taps = mk([(0, 0, 0, wf[0].permute(1, 2, 0).reshape(9, self.final_c))], 1, 9, epi=_lib.EPI_F32, name="fusion.final_taps")

    def flops(self, B: int, sizes) -> float:
        """sizes: list of (h, w) per level, finest first."""
        f = 0.0
        for idx, (ic, tc) in enumerate(zip(self.input_chl, self.temp_chl)):
            h, w = sizes[idx]
            f += 2.0 * B * h * w * 9 * (ic * tc + (tc + 2) * tc)
        for i, (c1, c2, cin, dc) in enumerate(self.dec):
            h, w = sizes[len(sizes) - 2 - i]
            f += 2.0 * B * h * w * 9 * (cin * cin + cin * dc)
        h, w = sizes[0]
        f += 2.0 * B * h * w * 9 * self.final_c
        return f

    def forward(self, c_feat: List[Act], f_feat: List[Act], pred1: torch.Tensor, pred2: torch.Tensor,
                update_base: Optional[torch.Tensor], trace: Optional[dict] = None) -> torch.Tensor:
        """fusion_model.py:84-122.  c_feat / f_feat finest first; pred1/pred2/update_base fp32 [B,1,H,W]."""
        B = pred1.shape[0]
        key = (B, pred1.shape[2], pred1.shape[3])
        ws = self.ws.setdefault(key, Workspace(self.device, self.x3))
        A = ws.act
        temp: List[Act] = []
        for idx, (c, f) in enumerate(zip(c_feat, f_feat)):
            tc = self.temp_chl[idx]
            e1 = A(f"e1_{idx}", B, c.H, c.W, tc, cs=tc + 8)
            self.enc1[idx]([c, f], out=e1)
            ops.depth_slots(pred1, pred2, e1, tc)
            t = A(f"t_{idx}", B, c.H, c.W, tc, cs=tc + 8)
            self.enc2[idx]([e1.view_channels(tc + 2)], out=t)
            if idx < len(c_feat) - 1:
                ops.depth_slots(pred1, pred2, t, tc)          # this level is a decoder skip: [feat | pred1 | pred2]
            temp.append(t)
        if trace is not None:
            trace["fusion_enc"] = [t.to_nchw() for t in temp]
        rev = temp[::-1]
        feat = rev[0]
        for i, skip in enumerate(rev[1:]):
            c1, c2, cin, dc = self.dec[i]
            up = ops.resize_bilinear(feat, A(f"d{i}_up", B, skip.H, skip.W, feat.C))
            mid = A(f"d{i}_mid", B, skip.H, skip.W, cin)
            c1([up, skip.view_channels(skip.C + 2)], out=mid)
            o = A(f"d{i}_out", B, skip.H, skip.W, dc)
            c2([mid], out=o)
            feat = o
        if trace is not None:
            trace["fusion_dec"] = feat.to_nchw()
        taps = ws.f32("final_taps", B, feat.H, feat.W, 16)
        self.final_taps([feat], out_f32=taps, out_f32_ld=16)
        out = ws.f32("pred", B, 1, feat.H, feat.W)
        ops.tap_stencil(taps, update_base, out)
        return out

"""FusionUnet on the B200 kernels (estimator/models/blocks/fusion_model.py:52-122, convs.py).

Every concatenation is virtual: the conv kernel walks several source tensors in its K loop.  The
two depth maps the reference concatenates at every level enter as ONE extra 1x1 K segment over an
18-channel im2col tensor of their 3x3 neighbourhoods (``ops.depth_taps``), instead of two extra
channels per tap (which would cost a mostly empty 64-wide K chunk per tap)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib, ops
from .nn import Act, GemmLayer, Workspace, conv_segments

LN_EPS = 1e-6   # convs.py:11


def depth_tap_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, 2, 3, 3] slice of a conv weight (the two concatenated depth channels) -> [Cout, 18] matching the
    channel order of ``ops.depth_taps``: (r*3+s)*2 + d."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], 18)


class FusionUnetB200:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, input_chl, temp_chl, dec_chl, x3: bool, device,
                 in_splits=None, names=("encoder_layers_1", "encoder_layers_2", "decoder_layers"), heavy: bool = False):
        """``in_splits``: per level (coarse channels, fine channels) of the first conv's virtual concat (default: two equal
        halves of ``input_chl``, the FusionUnet case); ``names``: state-dict module names of the two encoder lists and the
        decoder list (BiDirectionalFusion calls them fusion_layers_1 / fusion_layers_2 / f2r_agg, same arithmetic).
        ``heavy``: BiDirectionalFusionHeavy's blocks (bi_directional_fusion_model.py:448-485) -- every encoder block is
        conv -> LN -> conv -> LN -> conv -> GELU, every decoder block five conv -> GELU layers."""
        self.x3, self.device, self.heavy = x3, device, heavy
        self.input_chl, self.temp_chl, self.dec_chl = list(input_chl), list(temp_chl), list(dec_chl)
        self.ws: Dict[tuple, Workspace] = {}
        g = lambda k: sd[prefix + k].detach().float()
        mk = lambda segs, n_src, cout, **kw: GemmLayer(segs, n_src, cout, x3, device, **kw)
        self.enc1, self.enc2 = [], []
        n_e1, n_e2, n_dec = names
        for idx, (ic, tc) in enumerate(zip(self.input_chl, self.temp_chl)):
            if in_splits is None:
                assert ic % 2 == 0
                split = [ic // 2, ic // 2]
            else:
                split = list(in_splits[idx])
                assert sum(split) == ic
            # conv -> LN over channels -> GELU is one kernel while a pixel's channel vector fits one N tile (Cout <= 256);
            # wider levels (BiDirectionalFusion's 512-channel bottom level, a few hundred pixels) store fp32 rows and
            # normalise in prv2_layernorm_gelu
            act0 = _lib.ACT_IDENTITY if heavy else _lib.ACT_NONE          # heavy: the first LayerNorm is NOT followed by an activation
            ln = (lambda q, k=1, act=act0: dict(epi=_lib.EPI_LN_GELU, act=act, gamma=g(q + f"{k}.weight"), beta=g(q + f"{k}.bias"), eps=LN_EPS)) if tc <= 256 else \
                 (lambda q, k=1, act=act0: dict(epi=_lib.EPI_F32))

            def tail(q, name):
                """heavy only: [conv(tc->tc) -> LN] then [conv(tc->tc) -> GELU] behind the first conv -> LN"""
                if not heavy:
                    return None
                c2 = mk(conv_segments(g(q + "2.weight"), [tc]), 1, tc, name=name + ".b", **ln(q, 3))
                c2.ln_split = None if tc <= 256 else (g(q + "3.weight").to(device), g(q + "3.bias").to(device))
                c3 = mk(conv_segments(g(q + "4.weight"), [tc]), 1, tc, act=_lib.ACT_GELU, name=name + ".c")
                return c2, c3
            q = f"{n_e1}.{idx}.single_conv."
            self.enc1.append(mk(conv_segments(g(q + "0.weight"), split), 2, tc, name=f"fusion.enc1.L{idx}", **ln(q)))
            self.enc1[-1].ln_split = None if tc <= 256 else (g(q + "1.weight").to(device), g(q + "1.bias").to(device))
            self.enc1[-1].tail = tail(q, f"fusion.enc1.L{idx}")
            q = f"{n_e2}.{idx}.single_conv."
            w2 = g(q + "0.weight")                     # [tc, tc + 2, 3, 3]: cat[f, pred1, pred2]
            self.enc2.append(mk(conv_segments(w2[:, :tc], [tc]) + [(1, 0, 0, depth_tap_weight(w2[:, tc:tc + 2]))], 2, tc,
                                name=f"fusion.enc2.L{idx}", **ln(q)))
            self.enc2[-1].ln_split = None if tc <= 256 else (g(q + "1.weight").to(device), g(q + "1.bias").to(device))
            self.enc2[-1].tail = tail(q, f"fusion.enc2.L{idx}")
        self.dec = []
        rev = self.temp_chl[::-1]
        chl = rev[0]
        for i, (tc, dc) in enumerate(zip(rev[1:], self.dec_chl)):
            cin = tc + chl + 2
            q = f"{n_dec}.{i}.conv.double_conv."
            w1 = g(q + "0.weight")                     # [cin, cin, 3, 3]: cat[up(chl), skip(tc), pred1, pred2]
            c1 = mk(conv_segments(w1[:, :chl + tc], [chl, tc]) + [(2, 0, 0, depth_tap_weight(w1[:, chl + tc:chl + tc + 2]))], 3, cin,
                    act=_lib.ACT_GELU, name=f"fusion.dec{i}.conv1")
            mids = [mk(conv_segments(g(q + f"{k}.weight"), [cin]), 1, cin, act=_lib.ACT_GELU, name=f"fusion.dec{i}.mid{k}") for k in ((2, 4, 6) if heavy else ())]
            c2 = mk(conv_segments(g(q + ("8.weight" if heavy else "2.weight")), [cin]), 1, dc, act=_lib.ACT_GELU, name=f"fusion.dec{i}.conv2")
            c1.mids = mids
            self.dec.append((c1, c2, cin, dc))
            chl = dc
        wf = g("final_conv.weight")                       # [1, C, 3, 3] -> 1x1 conv with one output per tap: [9, C]
        self.final_c = wf.shape[1]
        self.final_w = wf[0].permute(1, 2, 0).reshape(9, self.final_c).contiguous().to(device)

    def flops(self, B: int, sizes) -> float:
        """sizes: list of (h, w) per level, finest first."""
        f = 0.0
        for idx, (ic, tc) in enumerate(zip(self.input_chl, self.temp_chl)):
            h, w = sizes[idx]
            f += 2.0 * B * h * w * 9 * (ic * tc + (tc + 2) * tc + (4 * tc * tc if self.heavy else 0))
        for i, (c1, c2, cin, dc) in enumerate(self.dec):
            h, w = sizes[len(sizes) - 2 - i]
            f += 2.0 * B * h * w * 9 * (cin * cin * (4 if self.heavy else 1) + cin * dc)
        h, w = sizes[0]
        f += 2.0 * B * h * w * 9 * self.final_c
        return f

    def _conv_ln(self, ws: Workspace, layer: GemmLayer, srcs: List[Act], out: Act) -> None:
        """conv -> channels LayerNorm (-> GELU unless this is a heavy block's inner LayerNorm)."""
        if layer.ln_split is None:
            layer(srcs, out=out)
            return
        rows = out.N * out.H * out.W
        tmp = ws.f32(f"ln_rows_{out.C}", rows, out.C)
        layer(srcs, out_f32=tmp, out_f32_ld=out.C)
        if self.heavy:
            ops.layernorm(tmp, layer.ln_split[0], layer.ln_split[1], LN_EPS, out)
        else:
            ops.layernorm_gelu(tmp, layer.ln_split[0], layer.ln_split[1], LN_EPS, out)

    def _single(self, ws: Workspace, layer: GemmLayer, srcs: List[Act], out: Act, tag: str) -> None:
        """SingleConvCNNLN (convs.py:64-75) or, heavy, SingleConvCNNLNHeavy (bi_directional_fusion_model.py:448-463)."""
        if layer.tail is None:
            self._conv_ln(ws, layer, srcs, out)
            return
        c2, c3 = layer.tail
        a = ws.act(tag + "_a", out.N, out.H, out.W, out.C)
        self._conv_ln(ws, layer, srcs, a)                     # conv -> LN
        b = ws.act(tag + "_b", out.N, out.H, out.W, out.C)
        self._conv_ln(ws, c2, [a], b)                         # conv -> LN
        c3([b], out=out)                                      # conv -> GELU

    def forward(self, c_feat: List[Act], f_feat: List[Act], pred1: torch.Tensor, pred2: torch.Tensor,
                update_base: Optional[torch.Tensor], trace: Optional[dict] = None) -> torch.Tensor:
        """fusion_model.py:84-122.  c_feat / f_feat finest first; pred1/pred2/update_base fp32 [B,1,H,W]."""
        B = pred1.shape[0]
        key = (B, pred1.shape[2], pred1.shape[3])
        ws = self.ws.setdefault(key, Workspace(self.device, self.x3))
        A = ws.act
        temp: List[Act] = []
        dtaps: List[Act] = []
        for idx, (c, f) in enumerate(zip(c_feat, f_feat)):
            tc = self.temp_chl[idx]
            e1 = A(f"e1_{idx}", B, c.H, c.W, tc)
            self._single(ws, self.enc1[idx], [c, f], e1, f"e1h_{idx}")
            d18 = A(f"dtaps_{idx}", B, c.H, c.W, 18, cs=24)
            ops.depth_taps(pred1, pred2, d18)                 # shared by enc2 of this level and the decoder conv that takes it as skip
            t = A(f"t_{idx}", B, c.H, c.W, tc)
            self._single(ws, self.enc2[idx], [e1, d18], t, f"e2h_{idx}")
            temp.append(t)
            dtaps.append(d18)
        if trace is not None:
            trace["fusion_enc"] = [t.to_nchw() for t in temp]
        rev, rev_d = temp[::-1], dtaps[::-1]
        feat = rev[0]
        for i, skip in enumerate(rev[1:]):
            c1, c2, cin, dc = self.dec[i]
            up = ops.resize_bilinear(feat, A(f"d{i}_up", B, skip.H, skip.W, feat.C))
            mid = A(f"d{i}_mid", B, skip.H, skip.W, cin)
            c1([up, skip, rev_d[i + 1]], out=mid)
            for k, lay in enumerate(c1.mids):                 # DoubleConvHeavy: three more cin -> cin conv + GELU layers
                nxt = A(f"d{i}_mid{k}", B, skip.H, skip.W, cin)
                lay([mid], out=nxt)
                mid = nxt
            o = A(f"d{i}_out", B, skip.H, skip.W, dc)
            c2([mid], out=o)
            feat = o
        if trace is not None:
            trace["fusion_dec"] = feat.to_nchw()
        out = ws.f32("pred", B, 1, feat.H, feat.W)
        ops.final_conv3x3(feat, self.final_w, update_base, out)
        return out

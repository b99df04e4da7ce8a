"""ZoeDepth metric-bins head on the B200 kernels (external/zoedepth/models/zoedepth/zoedepth_v1.py:173-233).

Everything ``ZoeDepth.forward`` does AFTER its core: ``conv2``, the seed bin regressor and projector, four projector +
``AttractorLayerUnnormed`` levels, the conditional log-binomial and the expectation over 64 bins.  The 1x1 convolutions are
``prv2_umma_gemm`` launches (the "emb + upsampled previous emb" input of an attractor MLP is two K segments with the same weights, the
``cat[last, cond]`` input of the conditional MLP three), the per-pixel bin arithmetic is two fused fp32 kernels
(``prv2_zoe_attractor``, ``prv2_zoe_logbinomial_depth``).

The reference's core for every shipped ZoeDepth config is MiDaS BEiT-L, whose source is fetched from GitHub at build time and is not
available offline -- so this head is an OPERATOR here (registered as ``ZoeDepthBinsHead``): callers hand in the
core's outputs, exactly what the reference's ``hack_feature`` argument carries (zoedepth_v1.py:164-171).  ``attractor_alpha`` /
``attractor_gamma`` from the configs are NOT used by the reference's unnormed attractor (attractor.py:194-197 calls the jit function
with its defaults alpha=300, gamma=2); this module follows the code.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib, ops
from .nn import Act, GemmLayer, Workspace
from .registry import MODELS

#: configs/patchrefiner_zoedepth/pr_u4k.py:10-66
ZOE_HEAD_CFG = dict(n_bins=64, bin_embedding_dim=128, n_attractors=(16, 8, 4, 1), attractor_alpha=1000, attractor_gamma=2,
                    attractor_kind="mean", attractor_type="inv", min_temp=0.0212, max_temp=50.0, bin_centers_type="softplus")
JIT_ALPHA = 300.0          # inv_attractor's default, the value the reference really runs with (attractor.py:45,194-197)


def zoe_head_weight_spec(output_channels: Sequence[int], cfg: dict = ZOE_HEAD_CFG, n_midas_out: int = 32) -> "OrderedDict[str, tuple]":
    """State-dict names and shapes of ZoeDepth minus ``core`` (zoedepth_v1.py:84-123)."""
    bt, outs, E, nb = output_channels[0], list(output_channels[1:]), cfg["bin_embedding_dim"], cfg["n_bins"]
    s: "OrderedDict[str, tuple]" = OrderedDict()

    def conv(name, co, ci):
        s[name + ".weight"] = (co, ci, 1, 1)
        s[name + ".bias"] = (co,)
    conv("conv2", bt, bt)
    conv("seed_bin_regressor._net.0", 256, bt); conv("seed_bin_regressor._net.2", nb, 256)
    conv("seed_projector._net.0", 128, bt); conv("seed_projector._net.2", E, 128)
    for i, c in enumerate(outs):
        conv(f"projectors.{i}._net.0", 128, c); conv(f"projectors.{i}._net.2", E, 128)
        conv(f"attractors.{i}._net.0", 128, E); conv(f"attractors.{i}._net.2", cfg["n_attractors"][i], 128)
    last_in = n_midas_out + 1
    conv("conditional_log_binomial.mlp.0", (last_in + E) // 2, last_in + E)
    conv("conditional_log_binomial.mlp.2", 4, (last_in + E) // 2)
    return s


class ZoeBinsHeadB200:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, output_channels: Sequence[int], x3: bool, device, cfg: dict = ZOE_HEAD_CFG,
                 n_midas_out: int = 32):
        if cfg["bin_centers_type"] != "softplus" or cfg["attractor_type"] != "inv":
            raise NotImplementedError("ZoeDepth head: bin_centers_type='softplus' with attractor_type='inv' (every shipped config) is implemented")
        self.x3, self.device, self.cfg = x3, device, dict(cfg)
        self.n_bins, self.E, self.n_out = int(cfg["n_bins"]), int(cfg["bin_embedding_dim"]), n_midas_out
        self.ws: Dict[tuple, Workspace] = {}
        g = lambda k: sd[prefix + k].detach().float()
        w1 = lambda k: g(k + ".weight")[:, :, 0, 0]

        def lin(name, srcs=1, **kw):
            w = w1(name)
            cout = w.shape[0]
            if isinstance(srcs, int):
                segs = [(i, 0, 0, w) for i in range(srcs)]           # several sources through the SAME weights = their sum
                n_src = srcs
            else:                                                     # column split of one weight over concatenated sources
                segs, o = [], 0
                for i, c in enumerate(srcs):
                    segs.append((i, 0, 0, w[:, o:o + c]))
                    o += c
                n_src = len(srcs)
            return GemmLayer(segs, n_src, cout, x3, device, bias=g(name + ".bias"), name="zoe." + name, **kw)
        bt, outs = output_channels[0], list(output_channels[1:])
        self.conv2 = lin("conv2")
        self.seed_r0 = lin("seed_bin_regressor._net.0", act=_lib.ACT_RELU)
        self.seed_r2 = lin("seed_bin_regressor._net.2", epi=_lib.EPI_F32)
        self.seed_p0 = lin("seed_projector._net.0", act=_lib.ACT_RELU)
        self.seed_p2 = lin("seed_projector._net.2")
        self.levels = []
        for i, c in enumerate(outs):
            self.levels.append(dict(p0=lin(f"projectors.{i}._net.0", act=_lib.ACT_RELU), p2=lin(f"projectors.{i}._net.2"),
                                    a0=lin(f"attractors.{i}._net.0", srcs=2, act=_lib.ACT_RELU),      # net(x + up(prev_x)) = W x + W up(prev_x) (attractor.py:176-181)
                                    a2=lin(f"attractors.{i}._net.2", epi=_lib.EPI_F32), na=int(cfg["n_attractors"][i])))
        self.mlp0 = lin("conditional_log_binomial.mlp.0", srcs=[n_midas_out, 1, self.E], act=_lib.ACT_GELU)     # cat[(outconv, rel), cond] (dist_layers.py:113)
        self.mlp2 = lin("conditional_log_binomial.mlp.2", epi=_lib.EPI_F32)

    def forward(self, rel_depth: torch.Tensor, btlnck: Act, x_blocks: List[Act], outconv: Act, trace: Optional[dict] = None):
        """rel_depth fp32 [B,H,W]; btlnck / x_blocks (coarse -> fine) / outconv as the core returns them (acts).  Returns
        (metric_depth fp32 [B,1,H,W], temp_features dict of acts) like zoedepth_v1.py:221-233."""
        B, H, W = outconv.N, outconv.H, outconv.W
        if tuple(rel_depth.shape) != (B, H, W):
            raise NotImplementedError("rel_depth must have the core's out_conv resolution (it does for every core of the reference)")
        key = (B, btlnck.H, btlnck.W, H, W)
        ws = self.ws.setdefault(key, Workspace(self.device, self.x3))
        A = ws.act
        E, K = self.E, self.n_bins
        x_d0 = A("x_d0", B, btlnck.H, btlnck.W, btlnck.C)
        self.conv2([btlnck], out=x_d0)                                                          # :173
        t = A("seed_r", B, btlnck.H, btlnck.W, 256)
        self.seed_r0([x_d0], out=t)
        b_prev = ws.f32("seed_raw", B, btlnck.H, btlnck.W, K)
        self.seed_r2([t], out_f32=b_prev, out_f32_ld=K)                                          # :176 (softplus is applied where the centres are read)
        prev_raw = True
        t2 = A("seed_p", B, btlnck.H, btlnck.W, 128)
        self.seed_p0([x_d0], out=t2)
        prev_emb = A("emb_seed", B, btlnck.H, btlnck.W, E)
        self.seed_p2([t2], out=prev_emb)                                                         # :184
        emb = prev_emb
        for i, (x, lv) in enumerate(zip(x_blocks, self.levels)):                                 # :188-195
            t = A(f"proj{i}_h", B, x.H, x.W, 128)
            lv["p0"]([x], out=t)
            emb = A(f"emb{i}", B, x.H, x.W, E)
            lv["p2"]([t], out=emb)
            up = prev_emb if (prev_emb.H, prev_emb.W) == (x.H, x.W) else ops.resize_bilinear(prev_emb, A(f"emb_up{i}", B, x.H, x.W, E))
            t = A(f"att{i}_h", B, x.H, x.W, 128)
            lv["a0"]([emb, up], out=t)
            a_raw = ws.f32(f"att{i}_raw", B, x.H, x.W, max(lv["na"], 4))
            lv["a2"]([t], out_f32=a_raw, out_f32_ld=a_raw.shape[-1])
            b_prev = ops.zoe_attractor(a_raw, lv["na"], b_prev, prev_raw, JIT_ALPHA, self.cfg["attractor_kind"] == "mean")
            prev_raw, prev_emb = False, emb
        rel = Act.from_nchw(rel_depth.reshape(B, 1, H, W), self.x3)                              # :207-210 (same size: the resize is the identity)
        cond = emb if (emb.H, emb.W) == (H, W) else ops.resize_bilinear(emb, A("cond", B, H, W, E))
        h = A("clb_h", B, H, W, self.mlp0.cout)
        self.mlp0([outconv, rel, cond], out=h)
        pt = ws.f32("clb_raw", B, H, W, 4)
        self.mlp2([h], out_f32=pt, out_f32_ld=4)
        depth = ops.zoe_logbinomial_depth(pt, b_prev, float(self.cfg["min_temp"]), float(self.cfg["max_temp"]))
        if trace is not None:
            trace["bin_centers_last"] = b_prev.clone()                                           # [B,h,w,K] at the finest decoder level
        feats = {"x_d0": x_d0, "midas_final_feat": outconv}
        for i, x in enumerate(x_blocks):
            feats[f"x_blocks_feat_{i}"] = x
        return depth, feats


@MODELS.register_module()
class ZoeDepthBinsHead(nn.Module):
    """Operator-level module: ZoeDepth's metric-bins head with the reference's state-dict keys (everything of ``ZoeDepth`` except
    ``core.*``).  ``forward(rel_depth, btlnck, x_blocks, outconv)`` takes the core outputs as CUDA NCHW tensors -- the contents of the
    reference's ``hack_feature=[rel_depth, [btlnck, *x_blocks, outconv]]`` -- and returns ``metric_depth`` [B,1,H,W]."""

    def __init__(self, output_channels=(256, 256, 256, 256, 256), n_midas_out: int = 32, precision: str = "bf16", **cfg):
        super().__init__()
        assert precision in ("bf16", "fp32")
        self.cfg = dict(ZOE_HEAD_CFG)
        self.cfg.update({k: v for k, v in cfg.items() if k in ZOE_HEAD_CFG})
        self.output_channels, self.n_midas_out, self.precision = list(output_channels), int(n_midas_out), precision
        self._weights = OrderedDict((k, torch.zeros(shp)) for k, shp in zoe_head_weight_spec(self.output_channels, self.cfg, self.n_midas_out).items())
        self._engine: Optional[ZoeBinsHeadB200] = None

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

    @torch.no_grad()
    def forward(self, rel_depth, btlnck, x_blocks, outconv, trace=None):
        dev = rel_depth.device
        if dev.type != "cuda":
            raise RuntimeError("patchrefinerv2_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        x3 = self.precision == "fp32"
        if self._engine is None or self._engine.device != dev:
            _lib.load()
            self._engine = ZoeBinsHeadB200(self._weights, "", self.output_channels, x3, dev, self.cfg, self.n_midas_out)
        to_act = lambda t: t if isinstance(t, Act) else Act.from_nchw(t.to(dev), x3)
        depth, _ = self._engine.forward(rel_depth.float().contiguous(), to_act(btlnck), [to_act(x) for x in x_blocks], to_act(outconv), trace)
        return depth

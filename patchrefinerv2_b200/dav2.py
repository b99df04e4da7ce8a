"""DepthAnythingV2 (DINOv2 ViT + DPT head) forward on the B200 kernels.

Follows external/depth_anything_v2/dpt.py:182-203 (forward), dinov2.py:212-231,297-321 (tokens,
blocks, taps), dinov2_layers/{block.py:82-107, attention.py:49-62, mlp.py:35-41,
layer_scale.py:27-28, patch_embed.py:69-82} and dpt.py:116-150 + util/blocks.py:57-80,123-148
(DPT head).  Weights come from a state dict with the reference's key names and are re-packed
once at construction; activations are channels-last bf16 (hi[,lo])."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import _lib, ops
from .nn import Act, GemmLayer, Workspace, ceil_to, conv_segments

VIT_CFG = {   # dinov2.py:340-390 ; dpt.py:165-170
    "vits": dict(dim=384, depth=12, heads=6, taps=(2, 5, 8, 11)),
    "vitb": dict(dim=768, depth=12, heads=12, taps=(2, 5, 8, 11)),
    "vitl": dict(dim=1024, depth=24, heads=16, taps=(4, 11, 17, 23)),
}
PATCH = 14
LN_EPS = 1e-6
KP = 592        # 3*14*14 = 588 im2col columns, pitch padded to a multiple of 8


def interpolate_pos_embed(pos_embed: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """dinov2.py:179-210, evaluated once per input size on the host (weights are constants):
    bicubic resample of the 37x37 table with the reference's +0.1 scale-factor offset."""
    pos_embed = pos_embed.detach().float().cpu()
    N = pos_embed.shape[1] - 1
    n0, n1 = h // PATCH, w // PATCH
    if n0 * n1 == N and h == w:
        return pos_embed[0]
    dim = pos_embed.shape[-1]
    sqrt_n = math.sqrt(N)
    s0, s1 = float(n0 + 0.1) / sqrt_n, float(n1 + 0.1) / sqrt_n
    pp = F.interpolate(pos_embed[:, 1:].reshape(1, int(sqrt_n), int(sqrt_n), dim).permute(0, 3, 1, 2),
                       scale_factor=(s0, s1), mode="bicubic", antialias=False)
    assert pp.shape[-2] == n0 and pp.shape[-1] == n1
    pp = pp.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((pos_embed[:, :1], pp), dim=1)[0].contiguous()


class DepthAnythingV2B200:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, encoder: str, features: int, out_channels, max_depth: float,
                 x3: bool, device):
        c = VIT_CFG[encoder]
        self.D, self.depth, self.heads, self.taps = c["dim"], c["depth"], c["heads"], c["taps"]
        self.features, self.oc, self.max_depth, self.x3, self.device = features, list(out_channels), float(max_depth), x3, device
        self.ws: Dict[int, Workspace] = {}
        self._pos: Dict[Tuple[int, int], torch.Tensor] = {}
        g = lambda k: sd[prefix + k].detach().float()
        dv = lambda t: t.contiguous().to(device)
        D = self.D
        p = "pretrained."
        self.pos_embed = g(p + "pos_embed")
        self.cls = dv(g(p + "cls_token").reshape(D))
        mk = lambda segs, n_src, cout, **kw: GemmLayer(segs, n_src, cout, x3, device, **kw)
        self.patch_embed = mk([(0, 0, 0, g(p + "patch_embed.proj.weight").reshape(D, 588))], 1, D, epi=_lib.EPI_F32,
                              bias=g(p + "patch_embed.proj.bias"), name="vit.patch_embed")
        self.blocks = []
        for i in range(self.depth):
            b = f"{p}blocks.{i}."
            self.blocks.append(dict(
                n1w=dv(g(b + "norm1.weight")), n1b=dv(g(b + "norm1.bias")),
                n2w=dv(g(b + "norm2.weight")), n2b=dv(g(b + "norm2.bias")),
                qkv=mk([(0, 0, 0, g(b + "attn.qkv.weight"))], 1, 3 * D, bias=g(b + "attn.qkv.bias"), name="vit.qkv"),
                proj=mk([(0, 0, 0, g(b + "attn.proj.weight"))], 1, D, epi=_lib.EPI_RESID_F32, bias=g(b + "attn.proj.bias"), gamma=g(b + "ls1.gamma"), name="vit.proj"),
                fc1=mk([(0, 0, 0, g(b + "mlp.fc1.weight"))], 1, 4 * D, act=_lib.ACT_GELU, bias=g(b + "mlp.fc1.bias"), name="vit.fc1"),
                fc2=mk([(0, 0, 0, g(b + "mlp.fc2.weight"))], 1, D, epi=_lib.EPI_RESID_F32, bias=g(b + "mlp.fc2.bias"), gamma=g(b + "ls2.gamma"), name="vit.fc2"),
            ))
        self.norm_w, self.norm_b = dv(g(p + "norm.weight")), dv(g(p + "norm.bias"))

        h = "depth_head."
        oc, Fe = self.oc, features
        self.projects = [mk([(0, 0, 0, g(f"{h}projects.{i}.weight").reshape(oc[i], D))], 1, oc[i], bias=g(f"{h}projects.{i}.bias"), name="dpt.projects") for i in range(4)]

        def deconv(name, cch, k):
            w = g(h + name + ".weight")                      # [Cin, Cout, k, k]  (dpt.py:62-73)
            wg = w.permute(2, 3, 1, 0).reshape(k * k * cch, cch)   # rows n = (ky*k+kx)*Cout + co
            return mk([(0, 0, 0, wg)], 1, k * k * cch, epi=_lib.EPI_SHUFFLE, bias=g(h + name + ".bias"), shuffle_k=k, name="dpt.deconv")
        self.resize0 = deconv("resize_layers.0", oc[0], 4)
        self.resize1 = deconv("resize_layers.1", oc[1], 2)
        # 3x3 stride-2 pad-1 conv on 2x2 phase-split sources: in(2y+r-1, 2x+s-1) = phase[(r-1)&1][(s-1)&1] at offset (r==0 ? -1 : 0)
        w3 = g(h + "resize_layers.3.weight")
        segs = []
        for r in range(3):
            for s in range(3):
                py, px = (r - 1) & 1, (s - 1) & 1
                segs.append((py * 2 + px, -1 if r == 0 else 0, -1 if s == 0 else 0, w3[:, :, r, s]))
        self.resize3 = mk(segs, 4, oc[3], bias=g(h + "resize_layers.3.bias"), name="dpt.conv_s2")
        self.layer_rn = [mk(conv_segments(g(f"{h}scratch.layer{i + 1}_rn.weight"), [oc[i]]), 1, Fe, name=f"dpt.layer{i + 1}_rn") for i in range(4)]
        self.refine = {}
        for r in (1, 2, 3, 4):
            q = f"{h}scratch.refinenet{r}."
            d = {"out_conv": mk([(0, 0, 0, g(q + "out_conv.weight").reshape(Fe, Fe))], 1, Fe, bias=g(q + "out_conv.bias"), name=f"dpt.refine{r}.out_conv")}
            for u in (1, 2):
                d[f"u{u}c1"] = mk(conv_segments(g(f"{q}resConfUnit{u}.conv1.weight"), [Fe]), 1, Fe, act=_lib.ACT_RELU, bias=g(f"{q}resConfUnit{u}.conv1.bias"), name=f"dpt.refine{r}.rcu")
                d[f"u{u}c2"] = mk(conv_segments(g(f"{q}resConfUnit{u}.conv2.weight"), [Fe]), 1, Fe, bias=g(f"{q}resConfUnit{u}.conv2.bias"), name=f"dpt.refine{r}.rcu")
            self.refine[r] = d
        s = h + "scratch."
        self.out1 = mk(conv_segments(g(s + "output_conv1.weight"), [Fe]), 1, Fe // 2, bias=g(s + "output_conv1.bias"), name="dpt.output_conv1")
        self.out2 = mk(conv_segments(g(s + "output_conv2.0.weight"), [Fe // 2]), 1, 32, epi=_lib.EPI_HEAD, bias=g(s + "output_conv2.0.bias"),
                       gamma=g(s + "output_conv2.2.weight").reshape(32), beta=g(s + "output_conv2.2.bias").reshape(1), head_scale=self.max_depth, name="dpt.output_conv2")

    # ------------------------------------------------------------------------------------------
    def _pos_for(self, H, W):
        key = (H, W)
        if key not in self._pos:
            self._pos[key] = interpolate_pos_embed(self.pos_embed, H, W).to(self.device)
        return self._pos[key]

    def flops(self, B: int, H: int, W: int) -> float:
        """Algorithmic FLOPs (2*MACs) of one forward: GEMMs/convs + attention."""
        gh, gw = H // PATCH, W // PATCH
        T = gh * gw
        D = self.D
        M = B * (T + 1)
        f = 2.0 * B * T * 588 * D
        f += self.depth * (2.0 * M * D * (3 * D + D + 8 * D) + 4.0 * B * self.heads * (T + 1) * (T + 1) * 64)
        oc, Fe = self.oc, self.features
        px = lambda s: B * (gh * s) * (gw * s)
        f += sum(2.0 * px(1) * D * oc[i] for i in range(4))
        f += 2.0 * px(1) * oc[0] * oc[0] * 16 + 2.0 * px(1) * oc[1] * oc[1] * 4 + 2.0 * px(0.5) * oc[3] * oc[3] * 9      # (even grids; odd ones round the half level up)
        sizes = [4, 2, 1, 0.5]
        f += sum(2.0 * px(sizes[i]) * oc[i] * Fe * 9 for i in range(4))
        conv = lambda s: 2.0 * px(s) * Fe * Fe * 9
        f += 2 * conv(0.5) + 4 * conv(1) + 4 * conv(2) + 4 * conv(4)            # residual conv units
        f += 2.0 * Fe * Fe * (px(1) + px(2) + px(4) + px(8))                      # out_conv 1x1 after the upsample
        f += 2.0 * px(8) * Fe * (Fe // 2) * 9 + 2.0 * B * H * W * (Fe // 2) * 32 * 9 + 2.0 * B * H * W * 32
        return f

    def forward(self, x: torch.Tensor, trace: Optional[dict] = None):
        """x: RGB in [0,1], fp32 [B,3,H,W] on the device.  Returns (metric depth fp32 [B,1,H,W],
        [x_d0, x_blocks_feat_0..3, midas_final_feat] as channels-last acts)."""
        B, _, H, W = x.shape
        assert H % PATCH == 0 and W % PATCH == 0
        gh, gw = H // PATCH, W // PATCH
        T = gh * gw
        D, x3 = self.D, self.x3
        M = B * (T + 1)
        ws = self.ws.setdefault((B, H, W), Workspace(self.device, x3))
        A = ws.act

        cols = ops.patchify(x.contiguous(), A("cols", 1, 1, B * T, 588, cs=KP))
        emb = ws.f32("emb", B * T, D)
        self.patch_embed([cols], out_f32=emb, out_f32_ld=D)
        xs = ws.f32("x", M, D)
        ops.assemble_tokens(emb, self.cls, self._pos_for(H, W), B, T, D, xs)
        if trace is not None:
            trace["tokens0"] = xs.clone().reshape(B, T + 1, D)
        y = A("y", 1, 1, M, D)
        qkv = A("qkv", 1, 1, M, 3 * D)
        att = A("att", 1, 1, M, D)
        hid = A("hid", 1, 1, M, 4 * D)
        taps: List[Act] = []
        for i, blk in enumerate(self.blocks):
            ops.layernorm(xs, blk["n1w"], blk["n1b"], LN_EPS, y)
            blk["qkv"]([y], out=qkv)
            ops.attention(qkv, B, T + 1, self.heads, att)
            blk["proj"]([att], out_f32=xs, out_f32_ld=D)
            ops.layernorm(xs, blk["n2w"], blk["n2b"], LN_EPS, y)
            blk["fc1"]([y], out=hid)
            blk["fc2"]([hid], out_f32=xs, out_f32_ld=D)
            if trace is not None and i == 0:
                trace["block0"] = xs.clone().reshape(B, T + 1, D)
            if i in self.taps:
                tp = A(f"tap{len(taps)}", B, gh, gw, D)
                ops.layernorm(xs, self.norm_w, self.norm_b, LN_EPS, tp, drop_period=T + 1)
                taps.append(tp)
        if trace is not None:
            trace["taps"] = [t.to_nchw().flatten(2).transpose(1, 2) for t in taps]

        # ---- DPT head (dpt.py:116-150) ----
        oc, Fe = self.oc, self.features
        pr = []
        for i in range(4):
            o = A(f"proj{i}", B, gh, gw, oc[i])
            self.projects[i]([taps[i]], out=o)
            pr.append(o)
        l1_in = A("rs0", B, gh * 4, gw * 4, oc[0])
        self.resize0([pr[0]], out=l1_in)
        l2_in = A("rs1", B, gh * 2, gw * 2, oc[1])
        self.resize1([pr[1]], out=l2_in)
        l3_in = pr[2]
        gh2, gw2 = (gh + 1) // 2, (gw + 1) // 2              # 3x3 stride-2 pad-1 conv (dpt.py:72-80): odd grids round up
        ph4 = A("rs3_phase", 4 * B, gh2, gw2, oc[3])
        ops.phase_split(pr[3], ph4)
        l4_in = A("rs3", B, gh2, gw2, oc[3])
        self.resize3([ph4.batch_slice(slice(k * B, (k + 1) * B)) for k in range(4)], out=l4_in)

        rn, rn_relu = [], []
        for i, src in enumerate((l1_in, l2_in, l3_in, l4_in)):
            o = A(f"l{i + 1}_rn", src.N, src.H, src.W, Fe)
            orl = A(f"l{i + 1}_rn_relu", src.N, src.H, src.W, Fe)
            self.layer_rn[i]([src], out=o, relu_out=orl)
            rn.append(o)
            rn_relu.append(orl)

        def rcu2_up_out(r: int, xin: Act, xin_relu: Act, size) -> Act:
            """resConfUnit2 -> bilinear(align_corners) to `size` -> out_conv (util/blocks.py:137-146)."""
            lay = self.refine[r]
            t = A(f"r{r}_t2", xin.N, xin.H, xin.W, Fe)
            lay["u2c1"]([xin_relu], out=t)                       # relu(conv1(relu(x)) + b)
            yv = A(f"r{r}_y", xin.N, xin.H, xin.W, Fe)
            lay["u2c2"]([t], out=yv, res=xin)                    # conv2(.) + b + x
            up = ops.resize_bilinear(yv, A(f"r{r}_up", xin.N, size[0], size[1], Fe))
            path = A(f"path_{r}", xin.N, size[0], size[1], Fe)
            lay["out_conv"]([up], out=path)
            return path

        def fuse(r: int, x0: Act, skip: Act, skip_relu: Act) -> Tuple[Act, Act]:
            """output = xs[0] + resConfUnit1(xs[1])  (util/blocks.py:131-135) -> (raw, relu copy)."""
            lay = self.refine[r]
            t = A(f"r{r}_t1", skip.N, skip.H, skip.W, Fe)
            lay["u1c1"]([skip_relu], out=t)
            o = A(f"r{r}_sum", skip.N, skip.H, skip.W, Fe)
            orl = A(f"r{r}_sum_relu", skip.N, skip.H, skip.W, Fe)
            lay["u1c2"]([t], out=o, relu_out=orl, res=skip, res2=x0)
            return o, orl

        path_4 = rcu2_up_out(4, rn[3], rn_relu[3], (rn[2].H, rn[2].W))
        s3, s3r = fuse(3, path_4, rn[2], rn_relu[2])
        path_3 = rcu2_up_out(3, s3, s3r, (rn[1].H, rn[1].W))
        s2, s2r = fuse(2, path_3, rn[1], rn_relu[1])
        path_2 = rcu2_up_out(2, s2, s2r, (rn[0].H, rn[0].W))
        s1, s1r = fuse(1, path_2, rn[0], rn_relu[0])
        path_1 = rcu2_up_out(1, s1, s1r, (rn[0].H * 2, rn[0].W * 2))

        o1 = A("out1", B, path_1.H, path_1.W, Fe // 2)
        self.out1([path_1], out=o1)
        out_feat = ops.resize_bilinear(o1, A("out_feat", B, H, W, Fe // 2))
        depth = ws.f32("depth", B, 1, H, W)
        self.out2([out_feat], out_f32=depth, out_f32_ld=1)
        return depth, [rn[3], path_4, path_3, path_2, path_1, out_feat]

"""Single launches of the hot kernels at ViT-L / 12-patch shapes, for ncu captures."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchrefinerv2_b200 import _lib, ops
from patchrefinerv2_b200.nn import Act, GemmLayer, conv_segments

DEV = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
B, T, D, heads = 12, 1025, 1024, 16
M = B * T
reps = 3


def run(fn):
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()


if which in ("all", "qkv"):
    y = Act.empty(1, 1, M, D, False, DEV); y.hi.normal_()
    lay = GemmLayer([(0, 0, 0, torch.randn(3 * D, D) / 32)], 1, 3 * D, False, DEV, bias=torch.randn(3 * D), name="qkv")
    out = Act.empty(1, 1, M, 3 * D, False, DEV)
    run(lambda: lay([y], out=out))
if which in ("all", "proj"):
    y = Act.empty(1, 1, M, D, False, DEV); y.hi.normal_()
    lay = GemmLayer([(0, 0, 0, torch.randn(D, D) / 32)], 1, D, False, DEV, epi=_lib.EPI_RESID_F32, bias=torch.randn(D), gamma=torch.rand(D), name="proj")
    x = torch.zeros(M, D, device=DEV)
    run(lambda: lay([y], out_f32=x, out_f32_ld=D))
if which in ("all", "attn"):
    qkv = Act.empty(1, 1, M, 3 * D, False, DEV); qkv.hi.normal_()
    out = Act.empty(1, 1, M, D, False, DEV)
    run(lambda: ops.attention(qkv, B, T, heads, out))
if which in ("all", "conv"):
    # fusion.dec4.conv1: cat[up(256), skip(128), depth taps] at 448x448 -> 386 channels, GELU (the largest single layer)
    from patchrefinerv2_b200.fusion import depth_tap_weight
    a = Act.empty(4, 448, 448, 256, False, DEV); a.hi.normal_()
    b = Act.empty(4, 448, 448, 128, False, DEV); b.hi.normal_()
    c = Act.empty(4, 448, 448, 18, False, DEV, cs=24); c.hi.normal_()
    w = torch.randn(386, 386, 3, 3) / math.sqrt(386 * 9)
    lay = GemmLayer(conv_segments(w[:, :384], [256, 128]) + [(2, 0, 0, depth_tap_weight(w[:, 384:386]))], 3, 386, False, DEV,
                    act=_lib.ACT_GELU, name="dec4.conv1")
    out = Act.empty(4, 448, 448, 386, False, DEV, cs=392)
    run(lambda: lay([a, b, c], out=out))
if which in ("all", "blend"):
    preds = torch.rand(81, 448, 448, device=DEV)
    mask = torch.rand(448, 448, device=DEV)
    stages = [(0, 0, 4, 4, 0), (0, 224, 4, 3, 16), (224, 0, 3, 4, 28), (224, 224, 3, 3, 40)]
    starts = torch.randint(0, 1600, (32, 2), dtype=torch.int32, device=DEV)
    rmask = torch.rand(540, 960, device=DEV) + 1e-3

    def f():
        avg, cnt = ops.blend_canvas(preds[:49], mask, stages, 1792, 1792)
        ops.blend_raw(avg, cnt, preds[49:], starts, rmask, 448, 448, 540, 960, 2160, 3840)
    run(f)
if which in ("all", "resize"):
    a = Act.empty(12, 256, 256, 256, False, DEV); a.hi.normal_()
    out = Act.empty(12, 448, 448, 256, False, DEV)
    run(lambda: ops.resize_bilinear(a, out))
if which in ("all", "ln"):
    x = torch.randn(M, D, device=DEV)
    w, b = torch.rand(D, device=DEV), torch.rand(D, device=DEV)
    y = Act.empty(1, 1, M, D, False, DEV)
    run(lambda: ops.layernorm(x, w, b, 1e-6, y))
if which in ("all", "roi"):
    f = Act.empty(1, 256, 256, 256, False, DEV); f.hi.normal_()
    rois = torch.tensor([[0.0, 0.0, 112.0, 112.0]] * 12, device=DEV) + torch.arange(12, device=DEV)[:, None] * 20.0
    out = Act.empty(12, 256, 256, 256, False, DEV)
    run(lambda: ops.roi_gather_act(f, rois.contiguous(), 256 / 448, out))
print("done", _lib.launch_count)

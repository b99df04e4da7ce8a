"""Single launches of the hot kernels at ViT-L shapes (PRV2_PROF_B patches per launch, default 27 = the batch bench.py runs), for ncu captures."""
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
B, T, D, heads = int(os.environ.get("PRV2_PROF_B", "27")), 1025, 1024, 16      # 27 = bench.py's patch batch (round 1 profiled 12)
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
if which in ("all", "fc1"):
    y = Act.empty(1, 1, M, D, False, DEV); y.hi.normal_()
    lay = GemmLayer([(0, 0, 0, torch.randn(4 * D, D) / 32)], 1, 4 * D, False, DEV, act=_lib.ACT_GELU, bias=torch.randn(4 * D), name="fc1")
    out = Act.empty(1, 1, M, 4 * D, False, DEV)
    run(lambda: lay([y], out=out))
if which in ("all", "fc2"):
    y = Act.empty(1, 1, M, 4 * D, False, DEV); y.hi.normal_()
    lay = GemmLayer([(0, 0, 0, torch.randn(D, 4 * D) / 64)], 1, D, False, DEV, epi=_lib.EPI_RESID_F32, bias=torch.randn(D), gamma=torch.rand(D), name="fc2")
    x = torch.zeros(M, D, device=DEV)
    run(lambda: lay([y], out_f32=x, out_f32_ld=D))
if which in ("all", "finalconv"):
    f = Act.empty(12, 448, 448, 128, False, DEV); f.hi.normal_()
    w9c = torch.randn(9, 128, device=DEV) / 34
    base = torch.rand(12, 1, 448, 448, device=DEV)
    out = torch.empty(12, 1, 448, 448, device=DEV)
    run(lambda: ops.final_conv3x3(f, w9c, base, out))
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
    # the BASELINE frame's real schedule (2160x3840, 4x4, r32, seed 1) with random predictions
    import random
    import numpy as np
    from patchrefinerv2_b200 import masks, tiling
    tc = tiling.prepare_tile_cfg((448, 448), (2160, 3840), (4, 4))
    st = tiling.schedule(tc, (448, 448), "r32", 4, random.Random(1))
    bb = np.concatenate([s.bboxs for s in st])
    stages, first = [], 0
    for s_ in st:
        if s_.kind == "regular":
            stages.append((s_.off_process[0], s_.off_process[1], s_.grid[0], s_.grid[1], first))
            first += s_.bboxs.shape[0]
    preds = torch.rand(bb.shape[0], 448, 448, device=DEV) * 10
    mask = torch.from_numpy(masks.generatemask((448, 448), 0.15).copy()).to(DEV)
    rmask = torch.from_numpy(masks.random_patch_mask((540, 960), 0.15).copy()).to(DEV)
    starts = torch.from_numpy(np.ascontiguousarray(bb[first:, [1, 0]])).to(DEV)
    flush = torch.empty(64 * 1024 * 1024, device=DEV)
    rprep = ops.blend_raw_prepare(rmask, 448)                 # once per geometry, as the model does

    def f():
        flush.fill_(1.0)                                      # cold L2, as in the frame loop
        avg, cnt = ops.blend_canvas(preds[:first], mask, stages, 1792, 1792)
        ops.blend_raw(avg, cnt, preds[first:], starts, rmask, 448, 448, 540, 960, 2160, 3840, prep=rprep)
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

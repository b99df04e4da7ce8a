"""Micro-benchmark of prv2_umma_gemm at the hot ViT-L / FusionUnet layer shapes (CUDA events on the launching
stream, back-to-back launches).  Env toggles PRV2_GEMM_CG / PRV2_GEMM_STAGES are read by the library.
Usage: python scripts/bench_gemm.py [names...]"""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchrefinerv2_b200 import _lib
from patchrefinerv2_b200.nn import Act, GemmLayer, conv_segments

DEV = "cuda:0"
torch.manual_seed(0)
B, T, D = int(os.environ.get("PRV2_PROF_B", "12")), 1025, 1024
M = B * T
which = sys.argv[1:]
results = []


def timed(name, fn, flops, iters=5):
    if which and name not in which:
        return
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    results.append((name, us, flops / us / 1e6))
    print(f"{name:14s} {us:9.1f} us  {flops / us / 1e6:8.1f} TF/s", flush=True)


def linear(name, K, N, **kw):
    if which and name not in which:
        return
    a = Act.empty(1, 1, M, K, False, DEV); a.hi.normal_()
    lay = GemmLayer([(0, 0, 0, torch.randn(N, K) / math.sqrt(K))], 1, N, False, DEV, bias=torch.randn(N), name=name, **kw)
    if kw.get("epi") == _lib.EPI_RESID_F32:
        x = torch.zeros(M, N, device=DEV)
        fn = lambda: lay([a], out_f32=x, out_f32_ld=N)
    else:
        out = Act.empty(1, 1, M, N, False, DEV)
        fn = lambda: lay([a], out=out)
    timed(name, fn, 2.0 * M * K * N)


def conv(name, Bc, H, W, splits, pitches, Cout, **kw):
    if which and name not in which:
        return
    srcs = []
    for c, cs in zip(splits, pitches):
        a = Act.empty(Bc, H, W, c, False, DEV, cs=cs); a.hi.normal_()
        srcs.append(a)
    w = torch.randn(Cout, sum(splits), 3, 3) / math.sqrt(sum(splits) * 9)
    lay = GemmLayer(conv_segments(w, splits), len(splits), Cout, False, DEV, name=name, **kw)
    out = Act.empty(Bc, H, W, Cout, False, DEV)
    timed(name, lambda: lay(srcs, out=out), 2.0 * Bc * H * W * 9 * sum(splits) * Cout, iters=3)


for kk in (64, 128, 256, 512, 1024, 2048, 4096, 8192):        # K sweep at the qkv shape: separates per-tile overhead from per-chunk cost
    linear(f"k{kk}", kk, 3 * D)
linear("qkv", D, 3 * D)
linear("fc1", D, 4 * D, act=_lib.ACT_GELU)
linear("fc2", 4 * D, D, epi=_lib.EPI_RESID_F32, gamma=torch.rand(D))
linear("proj", D, D, epi=_lib.EPI_RESID_F32, gamma=torch.rand(D))
conv("dec4.conv1", 4, 448, 448, [256, 130], [256, 136], 386, act=_lib.ACT_GELU)
conv("dec4.conv2", 4, 448, 448, [386], [392], 128, act=_lib.ACT_GELU)
conv("dec3.conv1", 4, 256, 256, [256, 258], [256, 264], 514, act=_lib.ACT_GELU)
conv("dec3.conv2", 4, 256, 256, [514], [520], 256, act=_lib.ACT_GELU)
conv("enc1.L0", 4, 448, 448, [128, 128], [128, 128], 128, epi=_lib.EPI_LN_GELU, gamma=torch.rand(128), beta=torch.rand(128))
conv("enc2.L0", 4, 448, 448, [130], [136], 128, epi=_lib.EPI_LN_GELU, gamma=torch.rand(128), beta=torch.rand(128))
conv("enc1.L1", 4, 256, 256, [256, 256], [256, 256], 256, epi=_lib.EPI_LN_GELU, gamma=torch.rand(256), beta=torch.rand(256))

if not which or "attn" in which:
    from patchrefinerv2_b200 import ops
    heads = 16
    qkv = Act.empty(1, 1, M, 3 * D, False, DEV); qkv.hi.normal_()
    ao = Act.empty(1, 1, M, D, False, DEV)
    timed("attn", lambda: ops.attention(qkv, B, T, heads, ao), 4.0 * B * heads * T * T * 64)
print("SUMMARY", os.environ.get("PRV2_GEMM_CG", "auto"), os.environ.get("PRV2_GEMM_STAGES", "-"),
      " ".join(f"{n}={t:.0f}" for n, _, t in results))

"""Diagnostic for the tcgen05 GEMM path: tiny cases with structured operands so that descriptor /
swizzle / layout mistakes show up as recognisable permutations.  Prints, never asserts."""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchrefinerv2_b200 import _lib
from patchrefinerv2_b200.nn import Act, GemmLayer

DEV = "cuda:0"
torch.manual_seed(0)


def run(M, K, N, a, w, x3=False):
    A = Act.from_nchw(a.t().reshape(1, K, 1, M).to(DEV), x3)
    lay = GemmLayer([(0, 0, 0, w)], 1, N, x3, DEV)
    out = Act.empty(1, 1, M, N, x3, DEV)
    out.hi.zero_()
    lay([A], out=out)
    torch.cuda.synchronize()
    return out.to_nchw()[0, :, 0, :].t().cpu()


def report(name, got, want):
    err = (got - want).abs()
    print(f"[{name}] max|err|={err.max().item():.4g} rel={err.max().item() / want.abs().max().item():.4g} "
          f"bad_rows={(err.max(1).values > 0.05 * want.abs().max()).sum().item()}/{got.shape[0]} "
          f"bad_cols={(err.max(0).values > 0.05 * want.abs().max()).sum().item()}/{got.shape[1]}")
    if err.max() > 0.05 * want.abs().max():
        print("  got[0,:8] ", got[0, :8].tolist())
        print("  want[0,:8]", want[0, :8].tolist())
        print("  got[1,:8] ", got[1, :8].tolist())
        print("  want[1,:8]", want[1, :8].tolist())
        print("  got[:8,0] ", got[:8, 0].tolist())
        print("  want[:8,0]", want[:8, 0].tolist())


try:
    # 1. identity weight: out == A (bf16-rounded)
    M, K, N = 128, 64, 64
    a = torch.randn(M, K).bfloat16().float()
    report("identity 128x64x64", run(M, K, N, a, torch.eye(N, K)), a)
    # 2. K = 16 only (single MMA slice)
    a = torch.randn(128, 16).bfloat16().float()
    w = torch.randn(32, 16).bfloat16().float()
    report("k16 128x16x32", run(128, 16, 32, a, w), a @ w.t())
    # 3. multiple chunks and tiles
    for (M, K, N) in [(128, 128, 64), (256, 64, 64), (128, 64, 256), (384, 256, 512), (1025, 384, 1152)]:
        a = torch.randn(M, K).bfloat16().float()
        w = (torch.randn(N, K) / math.sqrt(K)).bfloat16().float()
        report(f"rand {M}x{K}x{N}", run(M, K, N, a, w), a @ w.t())
    a = torch.randn(300, 200)
    w = torch.randn(72, 200) / math.sqrt(200)
    report("x3 300x200x72", run(300, 200, 72, a, w, x3=True), a @ w.t())
except Exception as e:          # noqa: BLE001
    print("EXCEPTION", type(e).__name__, e)
print("launches", _lib.launch_count)

"""Is the fp32-class mode's error floor the tensor core's fp32 accumulation?  x3 GEMM against an fp64 reference, for all-positive
operands (monotone running sum: a round-toward-zero accumulator shows up as a NEGATIVE mean signed error that grows with K) and
for random-sign operands, in both segment orders (main pass first / last).  Run on the GPU box."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchrefinerv2_b200 import _lib
from patchrefinerv2_b200.nn import Act, GemmLayer
DEV = "cuda:0"
torch.manual_seed(0)
for K in (256, 1024, 4096):
    for kind in ("positive", "random"):
        M, N = 256, 256
        a = torch.rand(M, K) + 0.5 if kind == "positive" else torch.randn(M, K)
        w = (torch.rand(N, K) + 0.5 if kind == "positive" else torch.randn(N, K)) / math.sqrt(K)
        want = (a.double() @ w.double().t())
        A = Act.from_nchw(a.t().reshape(1, K, 1, M).to(DEV), True)
        for order in ("main_first", "main_last"):
            os.environ["PRV2_X3_ORDER"] = order
            lay = GemmLayer([(0, 0, 0, w)], 1, N, True, DEV, epi=_lib.EPI_F32)
            out = torch.zeros(M, N, device=DEV)
            lay([A], out_f32=out, out_f32_ld=N)
            err = (out.cpu().double() - want)
            scale = want.abs().mean()
            print(f"K={K:5d} {kind:8s} {order:10s}: mean signed err / mean|y| = {err.mean() / scale:+.3e}   rms err / mean|y| = {err.pow(2).mean().sqrt() / scale:.3e}"
                  f"   (n_mma per output = {3 * K // 16}, n * 2^-24 = {3 * K / 16 * 2 ** -24:.2e})")

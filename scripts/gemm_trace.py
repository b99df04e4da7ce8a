"""Per-tile clock64 timeline of CTA 0 of prv2_umma_gemm (needs a library built with PRV2_EXTRA_NVCC_FLAGS=-DPRV2_GEMM_TRACE_BUILD).
    PRV2_EXTRA_NVCC_FLAGS=-DPRV2_GEMM_TRACE_BUILD python scripts/gemm_trace.py qkv|enc2|k64 [B]"""
import ctypes as C
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchrefinerv2_b200 import _lib, build
from patchrefinerv2_b200.nn import Act, GemmLayer, conv_segments
build.build()
which = sys.argv[1] if len(sys.argv) > 1 else "qkv"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 27
DEV = "cuda:0"
if which in ("qkv", "k64"):
    K = 1024 if which == "qkv" else 64
    M = B * 1025
    a = Act.empty(1, 1, M, K, False, DEV); a.hi.normal_()
    lay = GemmLayer([(0, 0, 0, torch.randn(3072, K) / math.sqrt(K))], 1, 3072, False, DEV, bias=torch.randn(3072))
    out = Act.empty(1, 1, M, 3072, False, DEV)
    fn = lambda: lay([a], out=out)
else:
    from patchrefinerv2_b200.fusion import depth_tap_weight
    Bc = 4
    a = Act.empty(Bc, 448, 448, 128, False, DEV); a.hi.normal_()
    d = Act.empty(Bc, 448, 448, 18, False, DEV, cs=24); d.hi.normal_()
    w = torch.randn(128, 130, 3, 3) / math.sqrt(130 * 9)
    lay = GemmLayer(conv_segments(w[:, :128], [128]) + [(1, 0, 0, depth_tap_weight(w[:, 128:130]))], 2, 128, False, DEV,
                    epi=_lib.EPI_LN_GELU, gamma=torch.rand(128), beta=torch.rand(128))
    out = Act.empty(Bc, 448, 448, 128, False, DEV)
    fn = lambda: lay([a, d], out=out)
for _ in range(3):
    fn()
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * (3 * 64 * 8))()
lib.prv2_debug_gemm_trace.argtypes = [C.c_void_p]
assert lib.prv2_debug_gemm_trace(buf) == 0
t = list(buf)
t0 = min(v for v in t if v)
print(f"{which}: epilogue warp 2 [loop top, params visible, accumulator full, staging free, rows staged, stores issued, warp synced, arrived] | MMA warp [tile top, accumulator free, committed]")
for it in range(24):
    e = [t[(0 * 64 + it) * 8 + k] for k in range(8)]
    m = [t[(1 * 64 + it) * 8 + k] for k in range(3)]
    f = lambda v: f"{v - t0:7d}" if v else "     -1"
    print(f"tile {it:2d} | epi:", " ".join(f(v) for v in e), "| mma:", " ".join(f(v) for v in m))

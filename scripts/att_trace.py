"""PRV2_ATTN_TRACE=1 python scripts/att_trace.py [B] : clock64 timeline of CTA (0,0,0) of the attention kernel (diagnostics)."""
import os, sys
os.environ["PRV2_ATTN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchrefinerv2_b200 import ops
from patchrefinerv2_b200.nn import Act
B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
x3 = "--x3" in sys.argv
T, heads, D = 1025, 16, 1024
qkv = Act.empty(1, 1, B * T, 3 * D, x3, "cuda"); (qkv.hi.view(torch.float16) if x3 else qkv.hi).normal_()
if x3: qkv.lo.view(torch.float16).normal_(std=2.0 ** -11)
out = Act.empty(1, 1, B * T, D, x3, "cuda")
for i in range(2):
    print(f"--- launch {i}", file=sys.stderr)
    ops.attention(qkv, B, T, heads, out)
torch.cuda.synchronize()

import sys, torch
sys.path.insert(0, '/root/repo')
from patchrefinerv2_b200 import ops
from patchrefinerv2_b200.nn import Act
B, T, heads = 1, int(sys.argv[1]), 2
x3 = len(sys.argv) > 2
D = heads * 64
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, T, 3 * D, generator=g)
A = Act.from_nchw(qkv.reshape(B * T, 3 * D).t().reshape(1, 3 * D, 1, B * T).cuda(), x3)
out = Act.empty(1, 1, B * T, D, x3, 'cuda')
ops.attention(A, B, T, heads, out)
torch.cuda.synchronize()
qb = qkv if x3 else qkv.to(torch.bfloat16).float()
q, k, v = qb.reshape(B, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
want = (((q * 64 ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1) @ v).transpose(1, 2).reshape(B, T, D)
got = out.to_nchw()[0, :, 0, :].t().reshape(B, T, D).cpu()
err = (got - want).abs()
print("T", T, "x3", x3, "max err", err.max().item(), "rel", (err.max() / want.abs().max()).item(), "worst row", err.amax(dim=(0, 2)).argmax().item())

"""Diagnostic only (never on the product path): cuBLAS bf16 GEMM rate at the ViT-L linear shapes, to know what the
library reaches on the same problem sizes under the same power cap."""
import torch
import os
M = int(os.environ.get("PRV2_PROF_B", "12")) * 1025
for name, K, N in (("qkv", 1024, 3072), ("fc1", 1024, 4096), ("fc2", 4096, 1024), ("proj", 1024, 1024), ("big", 8192, 8192)):
    m = 8192 if name == "big" else M
    a = torch.randn(m, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, w.t())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"cublas {name:5s} {us:8.1f} us {2.0 * m * K * N / us / 1e6:8.1f} TF/s")

#!/usr/bin/env python
"""Micro-benchmark of prv2_attention at ViT shapes (CUDA events, back-to-back launches, 256 MB L2 flush between them).
    python scripts/bench_attention.py                 # ViT-L: B in {1, 11, 12, 27}, T = 1025, 16 heads
    python scripts/bench_attention.py --x3            # the (hi, lo) fp32-class mode
Prints algorithmic TFLOP/s (4*B*heads*T^2*64 FLOP) and, for orientation, the MUFU bound: B*heads*T^2 exponentials at
16 per clock and SM."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from patchrefinerv2_b200 import ops  # noqa: E402
from patchrefinerv2_b200.nn import Act  # noqa: E402


def main():
    dev = "cuda:0"
    heads, D = (16, 1024) if "--vits" not in sys.argv else (6, 384)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    x3 = "--x3" in sys.argv
    for B, T in ((1, 1025), (11, 1025), (12, 1025), (27, 1025), (27, 257)):
        qkv = Act.empty(1, 1, B * T, 3 * D, x3, dev)
        (qkv.hi.view(torch.float16) if x3 else qkv.hi).normal_()
        if x3:
            qkv.lo.view(torch.float16).normal_(std=2.0 ** -11)
        out = Act.empty(1, 1, B * T, D, x3, dev)
        for _ in range(3):
            ops.attention(qkv, B, T, heads, out)
        torch.cuda.synchronize()
        evs = []
        for _ in range(20):
            flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.attention(qkv, B, T, heads, out); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        t = sorted(x.elapsed_time(y) for x, y in evs)
        us = t[len(t) // 2] * 1e3
        flop = 4.0 * B * heads * T * T * 64
        mufu_us = B * heads * T * T / (16.0 * sm) / 1.965e3          # at the 1.965 GHz boost clock
        print(f"B={B:3d} T={T:5d} heads={heads}: {us:8.1f} us  {flop / us / 1e6:7.1f} TFLOP/s   (MUFU bound {mufu_us:6.1f} us, mode={'fp32-class (hi,lo)' if x3 else 'bf16'})")


if __name__ == "__main__":
    main()

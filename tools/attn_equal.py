#!/usr/bin/env python
"""Bit-equality of two attention implementations of the A/B build: python tools/attn_equal.py 2 112"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("RT_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reptext_b200", "csrc",
                                           "librt_reptext_ab.so"))
import torch
from reptext_b200 import ops

a, b = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(0)
for B, H, S in [(1, 2, 100), (2, 3, 257), (1, 2, 128), (1, 4, 1000), (1, 24, 4608), (1, 3, 9728)]:
    D = H * 128
    qkv = torch.randn(B, S, 3 * D, device="cuda", dtype=torch.bfloat16)
    qkv[..., :D] *= 2.0
    outs = []
    for impl in (a, b):
        out = torch.full((B, S, D), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.attention(qkv, H, 128, 0, D, 2 * D, out=out, impl=impl)
        torch.cuda.synchronize()
        outs.append(out)
    d = (outs[0].float() - outs[1].float()).abs().max().item()
    print(f"B={B} H={H} S={S}: impl{a} == impl{b}: {bool(torch.equal(outs[0], outs[1]))} (max |d| {d:.3e}, nan {int(torch.isnan(outs[1].float()).sum())})", flush=True)

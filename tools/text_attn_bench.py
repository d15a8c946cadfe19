#!/usr/bin/env python
"""Isolated timing of the prompt encoders' attention (head_dim 64): tcgen05 kernel (default), warp-level mma.sync form
(option text_attn_mma), CUDA-core form (option text_attn_simt), at T5-XXL's and CLIP-L's shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reptext_b200 import _lib, ops


def bench(B, S, H, mode, opt, iters=20):
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv = (torch.randn(B, S, 3 * D, device="cuda", generator=g) * 0.5).bfloat16()
    bias = torch.randn(H, 2 * S - 1, device="cuda", generator=g) if mode == "bias" else None
    scale = 1.0 if mode == "bias" else 0.125
    if opt:
        _lib.set_option(opt, 1)
    try:
        for _ in range(3):
            out = ops.text_attention(qkv, H, scale, rel_bias=bias, causal=(mode == "causal"))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = ops.text_attention(qkv, H, scale, rel_bias=bias, causal=(mode == "causal"))
        e1.record()
        torch.cuda.synchronize()
    finally:
        if opt:
            _lib.set_option(opt, 0)
    us = e0.elapsed_time(e1) / iters * 1e3
    fl = 4.0 * B * H * S * S * 64 * (0.5 if mode == "causal" else 1.0)
    return us, fl / us / 1e6


if __name__ == "__main__":
    for (B, S, H, mode) in [(1, 512, 64, "bias"), (1, 512, 64, "plain"), (4, 512, 64, "bias"), (1, 77, 12, "causal"), (8, 77, 12, "causal")]:
        line = f"B={B} S={S} H={H} {mode}:"
        for name, opt in (("tcgen05", None), ("mma.sync", "text_attn_mma"), ("simt", "text_attn_simt")):
            us, tf = bench(B, S, H, mode, opt)
            line += f"  {name} {us:7.1f} us ({tf:6.1f} TFLOP/s)"
        print(line, flush=True)

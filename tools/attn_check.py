#!/usr/bin/env python
"""Correctness of an attention variant against torch fp32 SDPA at ragged and full sizes: python tools/attn_check.py 82"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the A/B kernels live in the -DRT_AB_VARIANTS build (python -m reptext_b200.build --ab)
os.environ.setdefault("RT_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reptext_b200", "csrc",
                                           "librt_reptext_ab.so"))
import torch
import torch.nn.functional as F
from reptext_b200 import ops


def main():
    impls = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["2"])]
    torch.manual_seed(0)
    for B, H, S in [(1, 2, 100), (2, 3, 257), (1, 2, 128), (1, 4, 1000), (1, 24, 4608), (1, 3, 9728)]:
        D = H * 128
        qkv = torch.randn(B, S, 3 * D, device="cuda", dtype=torch.bfloat16)
        qkv[..., :D] *= 2.0  # wider scores: the lazy rescale path is taken
        q, k, v = (qkv[..., i * D:(i + 1) * D].reshape(B, S, H, 128).transpose(1, 2).float() for i in range(3))
        ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, D)
        for impl in impls:
            out = torch.full((B, S, D), float("nan"), device="cuda", dtype=torch.bfloat16)
            ops.attention(qkv, H, 128, 0, D, 2 * D, out=out, impl=impl)
            torch.cuda.synchronize()
            err = float((out.float() - ref).norm() / ref.norm())
            bad = int(torch.isnan(out.float()).sum())
            print(f"B={B} H={H} S={S} impl{impl}: rel-L2 vs fp32 {err:.3e} nan={bad} max|d|={float((out.float()-ref).abs().max()):.3e}", flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Isolated timing of the tcgen05 attention kernel (and its debug variants) at the step's shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the A/B kernels live in the -DRT_AB_VARIANTS build (python -m reptext_b200.build --ab)
os.environ.setdefault("RT_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reptext_b200", "csrc",
                                           "librt_reptext_ab.so"))
import torch
from reptext_b200 import ops

def bench(S, H, impl, iters=10):
    qkv = torch.randn(1, S, 3 * H * 128, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(1, S, H * 128, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): ops.attention(qkv, H, 128, 0, H * 128, 2 * H * 128, out=out, impl=impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.attention(qkv, H, 128, 0, H * 128, 2 * H * 128, out=out, impl=impl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 4.0 * S * S * 128 * H / ms / 1e9

if __name__ == "__main__":
    impls = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["2"])]
    for S, H in [(4608, 24), (9728, 24)]:
        for impl in impls:
            ms, tf = bench(S, H, impl)
            print(f"S={S} H={H} impl{impl} (variant {impl-2}): {ms:.3f} ms {tf:.0f} TF/s", flush=True)

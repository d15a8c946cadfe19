#!/usr/bin/env python
"""One launch of the product attention kernel and of its skeleton (hand-offs + MMAs only) at the step's shape inside a
cudaProfilerStart/Stop window (ncu --profile-from-start off --set full --import-source on)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the A/B kernels live in the -DRT_AB_VARIANTS build (python -m reptext_b200.build --ab)
os.environ.setdefault("RT_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reptext_b200", "csrc",
                                           "librt_reptext_ab.so"))
import torch
from reptext_b200 import ops
S, H = 4608, 24
qkv = torch.randn(1, S, 3 * H * 128, device="cuda", dtype=torch.bfloat16)
out = torch.empty(1, S, H * 128, device="cuda", dtype=torch.bfloat16)
def run():
    for impl in (2, 16):
        ops.attention(qkv, H, 128, 0, H * 128, 2 * H * 128, out=out, impl=impl)
run(); torch.cuda.synchronize()
torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")

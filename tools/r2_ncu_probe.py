#!/usr/bin/env python
"""Round-2 capture set for `ncu --profile-from-start off --set full`: the product attention kernel, its software-pipelined
form (A/B build, impl 112), the LayerNorm-modulate kernel and the step's two largest GEMM shapes, one launch each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("RT_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reptext_b200", "csrc",
                                           "librt_reptext_ab.so"))
import torch
from reptext_b200 import _lib as L, ops

S, H, D = 4608, 24, 3072
dt = torch.bfloat16
qkv = torch.randn(1, S, 3 * D, device="cuda", dtype=dt)
out = torch.empty(1, S, D, device="cuda", dtype=dt)
x = torch.randn(1, S, D, device="cuda", dtype=dt)
mod = torch.randn(1, 4 * D, device="cuda") * 0.3
groups = [(0, 512, mod[:, :D], mod[:, D:2 * D]), (512, S, mod[:, 2 * D:3 * D], mod[:, 3 * D:])]
xn = torch.empty_like(x)
gemms = []
for (M, N, K) in [(4608, 3072, 15360), (4608, 21504, 3072)]:
    A = torch.randn(1, M, K, device="cuda", dtype=dt)
    W = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
    b = torch.randn(N, device="cuda", dtype=dt)
    gemms.append((A, W, b, torch.empty(1, M, N, device="cuda", dtype=dt)))


def run():
    for impl in (2, 112):
        ops.attention(qkv, H, 128, 0, D, 2 * D, out=out, impl=impl)
    ops.layernorm_modulate(x, groups, out=xn)
    for (A, W, b, o) in gemms:
        ops.gemm([ops.Problem(A=A, segs=[ops.Segment(W=W, bias=b, out=o, mode=L.EPI_BIAS)])], 1, dt)


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")

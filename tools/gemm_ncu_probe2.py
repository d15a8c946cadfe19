#!/usr/bin/env python
"""One launch each of our cta_group::2 GEMM with the BIAS and the GELU epilogue, and cuBLAS, on the FF1 shape, inside a
cudaProfilerStart/Stop window (ncu --profile-from-start off --set full --import-source on)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reptext_b200 import ops, _lib as L

dt = torch.bfloat16
M, N, K = 4608, 12288, 3072
A = torch.randn(1, M, K, device="cuda", dtype=dt)
W = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
b = torch.randn(N, device="cuda", dtype=dt)
out = torch.empty(1, M, N, device="cuda", dtype=dt)
def run():
    torch.matmul(A[0], W.t())
    for mode in (L.EPI_BIAS, L.EPI_GELU):
        ops.gemm([ops.Problem(A=A, segs=[ops.Segment(W=W, bias=b, out=out, mode=mode)])], 1, dt, impl=3)
run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")

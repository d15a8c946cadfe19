#!/usr/bin/env python
"""One launch each of cuBLAS (torch.matmul), our cta_group::2 and our cta_group::1 GEMM on two of the step's shapes, inside a
cudaProfilerStart/Stop window, for `ncu --profile-from-start off --set full` (what tile / cluster shape does cuBLAS pick,
what is its tensor-pipe activity and DRAM traffic on the same problem)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reptext_b200 import ops, _lib as L

shapes = [(4608, 3072, 15360), (4608, 21504, 3072)]
dt = torch.bfloat16
data = []
for (M, N, K) in shapes:
    A = torch.randn(1, M, K, device="cuda", dtype=dt)
    W = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
    b = torch.randn(N, device="cuda", dtype=dt)
    out = torch.empty(1, M, N, device="cuda", dtype=dt)
    data.append((A, W, b, out))
# warm-up outside the window (cuBLAS heuristics, TMA descriptor cache, smem attributes)
for (A, W, b, out) in data:
    torch.matmul(A[0], W.t())
    for impl in (3, 2):
        ops.gemm([ops.Problem(A=A, segs=[ops.Segment(W=W, bias=b, out=out, mode=L.EPI_BIAS)])], 1, dt, impl=impl)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for (A, W, b, out) in data:
    torch.matmul(A[0], W.t())
    for impl in (3, 2):
        ops.gemm([ops.Problem(A=A, segs=[ops.Segment(W=W, bias=b, out=out, mode=L.EPI_BIAS)])], 1, dt, impl=impl)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")

#!/usr/bin/env python
"""A/B timing of the tcgen05 GEMM variants on the step's shapes (run under gpurun)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reptext_b200 import ops, _lib as L

def bench(M, N, K, impl, mode=L.EPI_BIAS, iters=20, copies=4):
    dt = torch.bfloat16
    A = [torch.randn(1, M, K, device="cuda", dtype=dt) for _ in range(copies)]
    W = [torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5 for _ in range(copies)]
    b = torch.randn(N, device="cuda", dtype=dt)
    out = torch.empty(1, M, N, device="cuda", dtype=dt)
    def run(i):
        ops.gemm([ops.Problem(A=A[i % copies], segs=[ops.Segment(W=W[i % copies], bias=b, out=out, mode=mode)])], 1, dt, impl=impl)
    for i in range(3): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): run(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 2.0 * M * N * K / ms / 1e9

if __name__ == "__main__":
    shapes = [(4608, 12288, 3072), (4608, 9216, 3072), (4608, 3072, 15360), (4096, 3072, 12288), (4096, 3072, 3072), (4608, 21504, 3072), (512, 12288, 3072)]
    impls = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["2", "3"])]
    for s in shapes:
        line = f"{s}: "
        for impl in impls:
            ms, tf = bench(*s, impl)
            line += f" impl{impl}: {ms:.3f} ms {tf:.0f} TF/s |"
        print(line, flush=True)
    for s in [(4608, 12288, 3072)]:
        for impl in impls:
            ms, tf = bench(*s, impl, mode=L.EPI_GELU)
            print(f"gelu {s} impl{impl}: {ms:.3f} ms {tf:.0f} TF/s")

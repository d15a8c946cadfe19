#!/usr/bin/env python
"""Where does the tcgen05 GEMM lose time?  Times the step's shapes with the epilogue switched off and / or with every
k-block re-reading k = 0 (operands L2-hot) - option "gemm_debug" (results are wrong on purpose; timing only).
Also times cuBLAS (torch.matmul) on the same shapes under the same conditions for reference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reptext_b200 import _lib as L
from tools.gemm_sweep import bench


def cublas(M, N, K, iters=20, copies=4):
    A = [torch.randn(M, K, device="cuda", dtype=torch.bfloat16) for _ in range(copies)]
    W = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) for _ in range(copies)]
    for i in range(3): torch.matmul(A[i % copies], W[i % copies].t())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): torch.matmul(A[i % copies], W[i % copies].t())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 2.0 * M * N * K / ms / 1e9


if __name__ == "__main__":
    shapes = [(4608, 21504, 3072), (4608, 12288, 3072), (4608, 3072, 15360), (4608, 3072, 3072), (1216, 21504, 3072)]
    # warm the clocks
    for _ in range(3): cublas(8192, 8192, 8192, iters=10)
    for s in shapes:
        ms, tf = cublas(*s)
        line = f"{s}: cublas {ms:.3f} ms {tf:.0f} TF/s |"
        for impl in (2, 3):
            for dbg in (0, 1, 2, 3):
                L.set_option("gemm_debug", dbg)
                ms, tf = bench(*s, impl)
                line += f" cg{impl-1}/dbg{dbg}: {tf:.0f}"
            line += " |"
        L.set_option("gemm_debug", 0)
        print(line, flush=True)

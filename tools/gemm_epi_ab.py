#!/usr/bin/env python
"""A/B of the GEMM's epilogue warp count (option gemm_epi_warps: 0 = eight on the CTA-pair kernels, 4 = four) on the
step's shapes (run under gpurun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reptext_b200 import ops, _lib as L


def bench(M, N, K, mode, opt, iters=30, copies=4, debug=0):
    dt = torch.bfloat16
    L.set_option("gemm_epi_warps", opt)
    L.set_option("gemm_debug", debug)
    A = [torch.randn(1, M, K, device="cuda", dtype=dt) for _ in range(copies)]
    W = [torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5 for _ in range(copies)]
    b = torch.randn(N, device="cuda", dtype=dt)
    gate = torch.randn(1, N, device="cuda")
    out = torch.zeros(1, M, N, device="cuda", dtype=dt)

    def run(i):
        ops.gemm([ops.Problem(A=A[i % copies], segs=[ops.Segment(W=W[i % copies], bias=b, out=out, mode=mode)],
                              gate=gate if mode == L.EPI_GATE_RESID else None)], 1, dt)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    L.set_option("gemm_epi_warps", 0)
    L.set_option("gemm_debug", 0)
    ms = e0.elapsed_time(e1) / iters
    return ms * 1e3, 2.0 * M * N * K / ms / 1e9


if __name__ == "__main__":
    shapes = [(4608, 3072, 3072, L.EPI_GATE_RESID), (4608, 3072, 12288, L.EPI_GATE_RESID), (4608, 3072, 15360, L.EPI_GATE_RESID),
              (4608, 12288, 3072, L.EPI_GELU), (4608, 9216, 3072, L.EPI_BIAS), (4608, 21504, 3072, L.EPI_GELU),
              (9728, 3072, 15360, L.EPI_GATE_RESID), (9728, 12288, 3072, L.EPI_GELU)]
    for rep in range(2):
        for (M, N, K, mode) in shapes:
            line = f"M={M:5d} N={N:5d} K={K:5d}:"
            for opt in (4, 0):
                us, tf = bench(M, N, K, mode, opt)
                line += f"  {'four ' if opt == 4 else 'eight'}: {us:7.1f} us {tf:5.0f} TF/s |"
            for dbg in (16, 0, 16, 0):                    # 16: no operand prefetch ahead of the accumulator
                us, tf = bench(M, N, K, mode, 0, debug=dbg)
                line += f"  eight{', no prefetch' if dbg else ''}: {us:7.1f} us |"
            us, tf = bench(M, N, K, mode, 0, debug=1)     # accumulators dropped: what the epilogue still costs
            line += f"  no epilogue: {us:7.1f} us |"
            print(line, flush=True)

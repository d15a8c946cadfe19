#!/usr/bin/env python
"""DRAM traffic and time of the long-K GEMMs with the banded tile order (option gemm_band: 0 auto, -1 off).
Run under: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_tc --csv"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reptext_b200 import ops, _lib as L
dt = torch.bfloat16
for (M, N, K) in [(4608, 3072, 15360), (4096, 3072, 12288)]:
    A = torch.randn(1, M, K, device="cuda", dtype=dt)
    W = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
    b = torch.randn(N, device="cuda", dtype=dt)
    out = torch.zeros(1, M, N, device="cuda", dtype=dt)
    gate = torch.ones(1, N, device="cuda")
    # launches per shape: band off, band auto, then band auto with the L2 hints of gemm_debug 32 (W evict_last) and
    # 64 (W evict_last, A evict_first), twice each
    for band, dbg in ((-1, 0), (0, 0), (0, 32), (0, 64), (-1, 0), (0, 0), (0, 32), (0, 64)):
        L.set_option("gemm_band", band)
        L.set_option("gemm_debug", dbg)
        ops.gemm([ops.Problem(A=A, segs=[ops.Segment(W=W, bias=b, out=out, mode=L.EPI_GATE_RESID)], gate=gate)], 1, dt, impl=3)
    torch.cuda.synchronize()
L.set_option("gemm_band", 0)
L.set_option("gemm_debug", 0)
print("done")

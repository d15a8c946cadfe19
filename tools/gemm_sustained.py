#!/usr/bin/env python
"""Sustained (power-capped) throughput of the tcgen05 GEMM against cuBLAS on the step's shapes: each candidate runs
back to back for ~2.5 s, the last 1.5 s are timed, SM clocks are sampled during the run."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reptext_b200 import ops, _lib as L


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ClockSampler


def sustained(fn, flops, warm_s=1.0, time_s=1.5):
    # calibrate the batch so that one synchronize covers ~100 ms of queued work
    for _ in range(10): fn()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(20): fn()
    torch.cuda.synchronize()
    per = (time.time() - t0) / 20
    batch = max(10, int(0.1 / per))
    t0 = time.time()
    while time.time() - t0 < warm_s:
        for _ in range(batch): fn()
        torch.cuda.synchronize()
    smp = ClockSampler(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    e0.record()
    t0 = time.time()
    while time.time() - t0 < time_s:
        for _ in range(batch): fn()
        n += batch
        e1.record()
        e1.synchronize()
    ms = e0.elapsed_time(e1) / n
    c = smp.stop()
    return flops / ms / 1e9, c.get("sm_mhz") or 0.0, ",".join(c.get("reasons") or [])

if __name__ == "__main__":
    dt = torch.bfloat16
    for (M, N, K) in [(4608, 21504, 3072), (4608, 3072, 15360), (4608, 12288, 3072), (4608, 3072, 3072), (8192, 8192, 8192)]:
        A = [torch.randn(1, M, K, device="cuda", dtype=dt) for _ in range(3)]
        W = [torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5 for _ in range(3)]
        out = torch.empty(1, M, N, device="cuda", dtype=dt)
        b = torch.zeros(N, device="cuda", dtype=dt)
        i = [0]
        def ours(impl):
            def f():
                i[0] += 1
                ops.gemm([ops.Problem(A=A[i[0] % 3], segs=[ops.Segment(W=W[i[0] % 3], bias=b, out=out)])], 1, dt, impl=impl)
            return f
        def cub():
            i[0] += 1
            torch.matmul(A[i[0] % 3][0], W[i[0] % 3].t(), out=out[0])
        fl = 2.0 * M * N * K
        r = [("cuBLAS", sustained(cub, fl)), ("ours cg2", sustained(ours(3), fl)), ("ours cg1", sustained(ours(2), fl))]
        print(f"{(M, N, K)}: " + " | ".join(f"{n}: {v[0]:.0f} TF/s @ {v[1]:.0f} MHz [{v[2]}]" for n, v in r), flush=True)

#!/usr/bin/env python
"""Does a tcgen05.mma instruction carry a fixed cost?  The same GEMM kernel with 256-, 128- and 64-wide column tiles
(= N of the MMA instruction): output widths that are multiples of 256, of 128 only and of 64 only select BN."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gemm_sweep import bench
for impl, name in ((2, "cta_group::1"), (3, "cta_group::2")):
    for N, bn in ((3072, 256), (3200, 128), (3136, 64)):
        if impl == 3 and bn == 64: continue
        ms, tf = bench(4608, N, 3072, impl)
        print(f"{name} N={N} (BN={bn}): {ms:.3f} ms {tf:.0f} TF/s", flush=True)

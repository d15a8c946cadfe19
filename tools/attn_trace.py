#!/usr/bin/env python
"""Hand-off timeline of the attention kernel (library built with RT_AB_VARIANTS; impl 20 = trace variant 18).

CTA 0 records clock64() at every hand-off of every key block (csrc/attn_sm100.cu, kTrace); this prints, per event, the
mean offset in SM cycles from the MMA warp's "P_A first half arrived" event of the same block, over the steady-state
blocks, and the mean block period."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
# the A/B kernels live in the -DRT_AB_VARIANTS build (python -m reptext_b200.build --ab)
os.environ.setdefault("RT_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reptext_b200", "csrc",
                                           "librt_reptext_ab.so"))
import torch
from reptext_b200 import _lib as L, ops

NAMES = {0: "mma: V(j) landed", 1: "mma: P_A half 1 arrived", 2: "mma: P_A complete (+K(j+1))", 3: "mma: issued PV_A b + QK_A(j+1)",
         4: "mma: P_B half 1 arrived", 5: "mma: P_B complete", 6: "mma: issued PV_B b + QK_B(j+1)",
         8: "softmax A: S_A(j) ready", 9: "softmax A: loads + max done", 10: "softmax A: P half arrive", 11: "softmax A: P full arrive",
         20: "PIPE: PV_A a done", 21: "PIPE: PV_A b done", 22: "PIPE: QK_A(j+1) done", 23: "PIPE: PV_B a done",
         24: "PIPE: PV_B b done", 25: "PIPE: QK_B(j+1) done",
         26: "half-row A h0: after max exchange", 27: "half-row A h1: after max exchange",
         28: "half-row B h0: after max exchange", 29: "half-row B h1: after max exchange",
         16: "softmax A: exps done", 17: "softmax B: exps done", 18: "softmax A: P tile free", 19: "softmax B: P tile free",
         30: "mma: S_A free seen (QK_A(j+1) issue)", 31: "mma: S_B free seen (QK_B(j+1) issue)",
         12: "softmax B: S_B(j) ready", 13: "softmax B: loads + max done", 14: "softmax B: P half arrive", 15: "softmax B: P full arrive"}


def main():
    impl = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    S, H = 4608, 24
    qkv = torch.randn(1, S, 3 * H * 128, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(1, S, H * 128, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.attention(qkv, H, 128, 0, H * 128, 2 * H * 128, out=out, impl=impl)
    torch.cuda.synchronize()
    n_kv = (S + 127) // 128
    buf = (C.c_longlong * (128 * 32))()
    lib = L.lib()
    lib.rt_debug_attn_trace.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    L.check(lib.rt_debug_attn_trace(buf, 128 * 32))
    t = np.array(buf[:], dtype=np.int64).reshape(128, 32)[:n_kv]
    lo, hi = 6, n_kv - 4
    period = np.diff(t[lo:hi, 1]).mean()
    print(f"impl {impl}: S={S}, {n_kv} key blocks; mean block period {period:.0f} cycles (MMA work: 2048)")
    base = t[lo:hi, 1]
    rows = []
    for slot, name in NAMES.items():
        if not t[lo:hi, slot].any():
            continue
        d = (t[lo:hi, slot] - base).astype(np.float64)
        rows.append((d.mean(), name, d.std()))
    for m, name, sd in sorted(rows):
        print(f"  {m:8.0f}  (+-{sd:4.0f})  {name}")
    print("first blocks (raw, relative to block 0 event 8):")
    for j in range(0, 4):
        print("  j=%d " % j + " ".join(f"{slot}:{t[j, slot] - t[0, 8]}" for slot in sorted(NAMES)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Cost of the sequence-parallel peer stores, measured in ONE process on a box with >= 2 GPUs: the QKV(+MLP) GEMM and
the attention run on cuda:0 and scatter half of their output into a buffer that lives on cuda:1 (peer access over
NVLink), against the same launch with every destination local.  A/B of the staged (coalesced) and the direct
row-per-thread epilogue stores (option gemm_debug bit 4)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reptext_b200 import ops, _lib as L


def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    assert torch.cuda.device_count() >= 2
    torch.cuda.set_device(0)
    dt = torch.bfloat16
    world, H, hd, K = 2, 24, 128, 3072
    D, Dl = H * hd, H * hd // world
    S_loc = 4864
    Sg = S_loc * world
    # peer access: a cross-device copy makes torch enable it
    probe = torch.zeros(16, device="cuda:1"); probe.to("cuda:0"); torch.zeros(16, device="cuda:0").to("cuda:1")
    A = torch.randn(1, S_loc, K, device="cuda:0", dtype=dt)
    Ws = [torch.randn(D, K, device="cuda:0", dtype=dt) * K ** -0.5 for _ in range(3)]
    Wm = torch.randn(4 * D, K, device="cuda:0", dtype=dt) * K ** -0.5
    nw = torch.ones(hd, device="cuda:0", dtype=dt)
    ids = torch.zeros(S_loc, 3, device="cuda:0")
    rope = ops.rope_table(ids, (16, 56, 56))
    cat = torch.empty(1, S_loc, 5 * D, device="cuda:0", dtype=dt)
    local = [torch.zeros(1, Sg, 3 * Dl, device="cuda:0", dtype=dt) for _ in range(world)]
    remote = [local[0], torch.zeros(1, Sg, 3 * Dl, device="cuda:1", dtype=dt)]
    modes = [L.EPI_QKNORM_ROPE, L.EPI_QKNORM_ROPE, L.EPI_BIAS]

    def gemm(bufs, with_mlp):
        segs = [ops.Segment(W=Ws[i], out=bufs[0], mode=modes[i], out_col0=i * Dl, norm_w=nw if i < 2 else None,
                            scatter=True) for i in range(3)]
        if with_mlp:
            segs.append(ops.Segment(W=Wm, out=cat, mode=L.EPI_GELU, out_col0=D))
        ops.gemm([ops.Problem(A=A, segs=segs)], 1, dt, rope=rope, head_dim=hd, sp_out=bufs, sp_cols=Dl, sp_row0=0)

    for with_mlp in (False, True):
        flops = 2.0 * S_loc * K * (3 * D + (4 * D if with_mlp else 0))
        for dbg, name in ((0, "staged"), (4, "direct")):
            L.set_option("gemm_debug", dbg)
            tl = timeit(lambda: gemm(local, with_mlp))
            tr = timeit(lambda: gemm(remote, with_mlp))
            print(f"QKV{'+MLP' if with_mlp else ''} GEMM {name}: all-local {tl*1e3:.0f} us ({flops/tl/1e9:.0f} TF/s) | "
                  f"half to peer {tr*1e3:.0f} us ({flops/tr/1e9:.0f} TF/s; {S_loc*3*Dl*2/tr/1e6:.0f} GB/s over NVLink)", flush=True)
    L.set_option("gemm_debug", 0)
    # attention: this rank's 12 heads over the whole sequence, rows of the second half go to the peer
    qkv = torch.randn(1, Sg, 3 * Dl, device="cuda:0", dtype=dt)
    out_l = [torch.zeros(1, S_loc, 5 * D, device="cuda:0", dtype=dt) for _ in range(world)]
    out_r = [out_l[0], torch.zeros(1, S_loc, 5 * D, device="cuda:1", dtype=dt)]
    fl = 4.0 * Sg * Sg * hd * (H // world)
    for name, outs in (("all-local", out_l), ("half to peer", out_r)):
        t = timeit(lambda: ops.attention(qkv, H // world, hd, 0, Dl, 2 * Dl, out=outs[0], out_col0=0, sp_out=outs, sp_rows=S_loc))
        print(f"attention {name}: {t*1e3:.0f} us ({fl/t/1e9:.0f} TF/s)", flush=True)
    # correctness of the peer path: remote buffer == local buffer
    gemm(local, False); gemm(remote, False)
    torch.cuda.synchronize()
    print("peer scatter identical:", torch.equal(local[1].cpu(), remote[1].cpu()))


if __name__ == "__main__":
    main()

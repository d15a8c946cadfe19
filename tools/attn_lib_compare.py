#!/usr/bin/env python
"""Same-box comparison of the product attention kernel with the library kernels on this image, at the shapes of the
named configurations (B, H, S, hd): cfg 2 (1, 24, 4608, 128), cfg 4 (2, 24, 4608, 128), cfg 5 on one GPU
(1, 24, 9728, 128) and cfg 5 per rank at 8 GPUs (1, 3, 9728, 128).

    torch SDPA (cuDNN / flash / mem-efficient backends), flash_attn 2.8 (sm_80 kernels), flashinfer single_prefill
    (backends auto / fa2 / fa3 / cutlass) and flashinfer's Blackwell CUTLASS FMHA (fmha_varlen).

Each kernel gets 3 warm-up + 10 timed calls bracketed by CUDA events; between calls a 256 MB buffer is written so that
L2 does not carry q/k/v over (all kernels pay the same flush).  The product kernel ALSO does per-head RMSNorm + RoPE on
q and k in the producing GEMM's epilogue, so the inputs here are already normed / rotated for every contestant.
Writes one JSON line per (shape, kernel) to stdout; a kernel that is missing or fails prints its error instead.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

SHAPES = [("cfg2", 1, 24, 4608), ("cfg4", 2, 24, 4608), ("cfg5_1gpu", 1, 24, 9728), ("cfg5_rank_of_8", 1, 3, 9728)]
HD = 128
ITERS = 10


def timeit(fn, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(ITERS):
        flush.add_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / ITERS


def main():
    from reptext_b200 import ops
    torch.manual_seed(0)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
    rows = []
    for tag, B, H, S in SHAPES:
        D = H * HD
        qkv = torch.randn(B, S, 3 * D, device="cuda", dtype=torch.bfloat16)
        out = torch.empty(B, S, D, device="cuda", dtype=torch.bfloat16)
        flops = 4.0 * S * S * HD * H * B
        q, k, v = (qkv[..., i * D:(i + 1) * D].reshape(B, S, H, HD) for i in range(3))
        qc, kc, vc = (t.contiguous() for t in (q, k, v))                       # [B, S, H, hd]
        qt, kt, vt = (t.transpose(1, 2).contiguous() for t in (qc, kc, vc))   # [B, H, S, hd]
        ref = F.scaled_dot_product_attention(qt.float()[:, :2, :256], kt.float()[:, :2], vt.float()[:, :2])

        def check(o_bshd):
            return float((o_bshd[:, :256, :2].float() - ref.transpose(1, 2)).norm() / ref.norm())

        cands = {}
        cands["product attn_tc_kernel"] = (lambda: ops.attention(qkv, H, HD, 0, D, 2 * D, out=out),
                                           lambda: out.view(B, S, H, HD))
        from torch.nn.attention import SDPBackend, sdpa_kernel
        for nm, be in (("sdpa cudnn", SDPBackend.CUDNN_ATTENTION), ("sdpa flash", SDPBackend.FLASH_ATTENTION),
                       ("sdpa mem-efficient", SDPBackend.EFFICIENT_ATTENTION)):
            def f(be=be):
                with sdpa_kernel(be):
                    return F.scaled_dot_product_attention(qt, kt, vt)
            cands[nm] = (f, lambda f=f: f().transpose(1, 2))
        cands["sdpa default"] = (lambda: F.scaled_dot_product_attention(qt, kt, vt),
                                 lambda: F.scaled_dot_product_attention(qt, kt, vt).transpose(1, 2))
        try:
            from flash_attn import flash_attn_func
            cands["flash_attn 2.8"] = (lambda: flash_attn_func(qc, kc, vc), lambda: flash_attn_func(qc, kc, vc))
        except Exception as e:
            rows.append(dict(shape=tag, kernel="flash_attn 2.8", error=repr(e)[:200]))
        try:
            import flashinfer
            if B == 1:
                for be in ("auto", "fa2", "fa3", "cutlass"):
                    def f(be=be):
                        return flashinfer.single_prefill_with_kv_cache(qc[0], kc[0], vc[0], causal=False, backend=be)
                    cands[f"flashinfer single_prefill[{be}]"] = (f, lambda f=f: f()[None])
            from flashinfer.prefill import fmha_varlen
            offs = (torch.arange(B + 1, device="cuda", dtype=torch.int32) * S)
            q2, k2, v2 = (t.reshape(B * S, H, HD) for t in (qc, kc, vc))

            def fv():
                o = fmha_varlen(q2, k2, v2, offs, offs, max_qo_len=S, causal=False)
                return o[0] if isinstance(o, tuple) else o
            cands["flashinfer fmha_varlen (CUTLASS sm100a)"] = (fv, lambda: fv().reshape(B, S, H, HD))
        except Exception as e:
            rows.append(dict(shape=tag, kernel="flashinfer", error=repr(e)[:200]))

        for nm, (fn, get) in cands.items():
            try:
                fn()
                torch.cuda.synchronize()
                err = check(get())
                ms = timeit(fn, flush)
                rows.append(dict(shape=tag, B=B, H=H, S=S, kernel=nm, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1),
                                 rel_l2_vs_fp32=float(f"{err:.2e}")))
            except Exception as e:
                rows.append(dict(shape=tag, kernel=nm, error=repr(e)[:200]))
            print(json.dumps(rows[-1]), flush=True)
        del qkv, out, qc, kc, vc, qt, kt, vt
        torch.cuda.empty_cache()
    best = {}
    for r in rows:
        if "tflops" in r and r["rel_l2_vs_fp32"] < 2e-2:
            best.setdefault(r["shape"], []).append(r)
    for tag, rs in best.items():
        prod = next(r for r in rs if r["kernel"].startswith("product"))
        lib = max((r for r in rs if not r["kernel"].startswith("product")), key=lambda r: r["tflops"])
        rel = prod["tflops"] / lib["tflops"]
        print(f"VERDICT {tag}: product {prod['tflops']} TF/s vs best library ({lib['kernel']}) {lib['tflops']} TF/s -> "
              f"{'ahead' if rel >= 1 else 'behind'} by {abs(rel - 1) * 100:.1f} %", flush=True)


if __name__ == "__main__":
    main()

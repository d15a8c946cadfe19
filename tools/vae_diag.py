"""VAE path diagnostics + timing on one B200 (not a test: prints every figure, never stops at the first failure).

    python tools/vae_diag.py [--size 1024] [--reps 5]

1. every VAE kernel against fp32 torch (rel-L2), 2. the AutoencoderKL drop-in against the oracle at 128^2 and at --size,
3. device time of encode / decode (CUDA events on the launch stream, after warm-up) for both convolution forms and for the
oracle run by stock torch in bf16 (cuDNN + SDPA, channels_last) - the practical "kernel to beat".
"""
import argparse
import json
import os
import sys
import traceback

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import vae_oracle as V  # noqa: E402
from reptext_b200 import _lib, ops, vae  # noqa: E402
from util import rel_l2  # noqa: E402

BF = torch.bfloat16
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rnd(shape, seed, scale=1.0, dtype=BF):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


def nchw(x, hw):
    B, HW, C = x.shape
    return x.float().view(B, hw[0], hw[1], C).permute(0, 3, 1, 2).contiguous()


def attempt(name, fn):
    try:
        r = fn()
        torch.cuda.synchronize()
        print(f"[ok  ] {name}: {r}", flush=True)
        return r
    except Exception as e:  # noqa: BLE001
        print(f"[FAIL] {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()
        try:
            torch.cuda.synchronize()
        except Exception as e2:  # noqa: BLE001
            print("CUDA context is broken:", e2, flush=True)
            sys.exit(3)
        return None


def conv_case(H, W, C, Co, B, impl):
    x = rnd((B, H * W, C), 1)
    w = rnd((Co, C, 3, 3), 2, (9 * C) ** -0.5)
    b = rnd((Co,), 3, 0.1)
    ref = F.conv2d(nchw(x, (H, W)), w.float(), b.float(), padding=1)
    out = ops.conv3x3(x, (H, W), ops.pack_conv3x3_weight(w), b, impl=impl)
    return f"rel-L2 {rel_l2(nchw(out, (H, W)), ref):.2e}"


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def conv_flops(cfg, size, decode):
    """2 * pixels * Cout * 9 * Cin over the 3x3 convolutions (+ 1x1 shortcuts, attention) of one encode / decode."""
    sd = V.param_shapes(cfg)
    boc = cfg["block_out_channels"]
    nb = len(boc)
    total = 0.0
    lat = size // 8

    def res_of(name):
        if name.startswith("encoder."):
            if "down_blocks" in name:
                i = int(name.split(".")[2])
                r = size >> i
                if "downsamplers" in name:
                    r >>= 1
                return r
            if name.startswith("encoder.conv_in"):
                return size
            return lat
        if "up_blocks" in name:
            i = int(name.split(".")[2])
            r = lat << i
            if "upsamplers" in name:
                r <<= 1
            return r
        if name.startswith("decoder.conv_out") or name.startswith("decoder.conv_norm_out"):
            return size
        return lat

    side = "decoder." if decode else "encoder."
    for k, s in sd.items():
        if not k.startswith(side) or not k.endswith(".weight") or len(s) < 2:
            continue
        r = res_of(k)
        kk = s[2] * s[3] if len(s) == 4 else 1
        total += 2.0 * r * r * s[0] * s[1] * kk
    total += 4.0 * (lat * lat) ** 2 * boc[-1]      # q k^T and P v
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--skip-ops", action="store_true")
    a = ap.parse_args()
    _lib.lib()
    print(torch.cuda.get_device_name(0), flush=True)
    if not a.skip_ops:
        for (H, W, C, Co, B) in [(32, 32, 64, 64, 1), (16, 128, 128, 256, 2), (8, 256, 64, 128, 1), (64, 16, 192, 64, 2),
                                 (128, 128, 512, 512, 1)]:
            for impl in (2, 3):
                attempt(f"conv3x3 {H}x{W} C{C}->{Co} B{B} impl{impl}", lambda: conv_case(H, W, C, Co, B, impl))

        def gn():
            x = rnd((2, 4096, 128), 10, 2.0) + 0.5
            g, b = rnd((128,), 11, 0.2) + 1, rnd((128,), 12, 0.2)
            out = ops.groupnorm_nhwc(x, 32, g, b, silu=True)
            ref = F.silu(F.group_norm(x.float().transpose(1, 2), 32, g.float(), b.float(), eps=1e-6))
            return f"rel-L2 {rel_l2(out.float().transpose(1, 2), ref):.2e}"
        attempt("groupnorm+silu", gn)

        def sm():
            s = rnd((64, 16384), 15, 3.0)
            ref = torch.softmax(s.float(), -1)
            ops.softmax_rows_(s)
            return f"rel-L2 {rel_l2(s, ref):.2e}"
        attempt("softmax_rows", sm)

    results = {}
    for size, boc in ((128, (64, 128, 256, 256)), (a.size, (128, 256, 512, 512))):
        cfg = dict(V.FLUX_VAE_CONFIG, block_out_channels=boc)
        sd = {k: v.to(BF).float() for k, v in V.random_state_dict(cfg, seed=21).items()}
        sdd = {k: v.cuda() for k, v in sd.items()}
        g = torch.Generator().manual_seed(22)
        low = torch.rand(1, 3, size // 16, size // 16, generator=g) * 2 - 1
        img = (F.interpolate(low, size=(size, size), mode="bilinear") * 0.8 +
               0.2 * (torch.rand(1, 3, size, size, generator=g) * 2 - 1)).to(BF).cuda()
        zin = torch.randn(1, 16, size // 8, size // 8, generator=g).to(BF).cuda()
        with torch.no_grad():
            ref_mom = attempt(f"oracle encode {size}", lambda: V.encode_moments(sdd, cfg, img.float()))
            ref_img = attempt(f"oracle decode {size}", lambda: V.decode(sdd, cfg, zin.float()))
        for impl in ("implicit", "im2col"):
            m = attempt(f"build AutoencoderKL {impl}", lambda: vae.AutoencoderKL(cfg, sd, conv_impl=impl))
            if m is None:
                continue
            n0 = _lib.launch_count()
            e = attempt(f"encode {size} {impl} vs oracle",
                        lambda: rel_l2(m.encode(img).latent_dist.parameters, ref_mom))
            n_enc = _lib.launch_count() - n0
            n0 = _lib.launch_count()
            d = attempt(f"decode {size} {impl} vs oracle", lambda: rel_l2(m.decode(zin).sample, ref_img))
            n_dec = _lib.launch_count() - n0
            rec = dict(size=size, conv_impl=impl, encode_rel_l2=e, decode_rel_l2=d, encode_launches=n_enc,
                       decode_launches=n_dec)
            if e is not None and d is not None and size == a.size:
                t_e = timed(lambda: m.encode(img), a.reps)
                t_d = timed(lambda: m.decode(zin), a.reps)
                fe, fd = conv_flops(cfg, size, False), conv_flops(cfg, size, True)
                rec.update(encode_ms=t_e, decode_ms=t_d, encode_tflops=fe / t_e / 1e9, decode_tflops=fd / t_d / 1e9,
                           encode_flop=fe, decode_flop=fd)
            results[f"{size}_{impl}"] = rec
            print(json.dumps(rec), flush=True)
        if size == a.size:
            # stock torch, bf16, channels_last: cuDNN convolutions + SDPA
            sdb = {k: (v.to(BF).contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v.to(BF))
                   for k, v in sdd.items()}
            xi = img.contiguous(memory_format=torch.channels_last)
            zi = zin.contiguous(memory_format=torch.channels_last)
            with torch.no_grad():
                t_e = attempt("torch bf16 encode ms", lambda: timed(lambda: V.encode_moments(sdb, cfg, xi), a.reps))
                t_d = attempt("torch bf16 decode ms", lambda: timed(lambda: V.decode(sdb, cfg, zi), a.reps))
                eb = attempt("torch bf16 encode vs fp32 oracle", lambda: rel_l2(V.encode_moments(sdb, cfg, xi), ref_mom))
                db = attempt("torch bf16 decode vs fp32 oracle", lambda: rel_l2(V.decode(sdb, cfg, zi), ref_img))
            results["torch_bf16"] = dict(size=size, encode_ms=t_e, decode_ms=t_d, encode_rel_l2=eb, decode_rel_l2=db)
            print(json.dumps(results["torch_bf16"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "vae_diag.json"), "w") as fh:
        json.dump(results, fh, indent=1)


if __name__ == "__main__":
    main()

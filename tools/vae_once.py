"""One VAE encode + one decode at 1024 x 1024 (FLUX.1-dev VAE architecture, random weights) after a warm-up pass: the
program ncu profiles for the VAE launch list (profiles/r1_vae_launches.txt).  Prints the per-class device time of the
two calls (CUDA events around every launch, option "profile")."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reptext_b200 import _lib, vae  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = vae.AutoencoderKL.random_init(seed=102)
g = torch.Generator().manual_seed(0)
img = (torch.rand(1, 3, size, size, generator=g) * 2 - 1).to(torch.bfloat16).cuda()
z = torch.randn(1, 16, size // 8, size // 8, generator=g).to(torch.bfloat16).cuda()
if os.environ.get("VAE_WARMUP", "1") == "1":
    m.encode(img), m.decode(z)
torch.cuda.synchronize()
for name, fn in (("encode", lambda: m.encode(img)), ("decode", lambda: m.decode(z))):
    _lib.set_option("profile", 1)
    _lib.profile_reset()
    fn()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.set_option("profile", 0)
    _lib.profile_reset()
    rec = {k: dict(ms=round(v[0], 3), launches=v[2],
                   achieved=round(v[1] / (v[0] / 1e3) / (1e12 if "gemm" in k else 1e9), 1),
                   unit="TFLOP/s" if "gemm" in k else "GB/s") for k, v in prof.items()}
    print(json.dumps({name: rec}), flush=True)

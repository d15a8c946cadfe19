#!/usr/bin/env python
"""Per-step wall / device times of the public T2I __call__ at full size (FLUX.1-dev architecture, synthetic VAE and
prompt encoders), several images per process, with pipe.precompute_modulation on and off (run under gpurun).

    python tools/e2e_step_trace.py [images]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from reptext_b200 import config, models
from reptext_b200.pipeline_flux_controlnet import FluxControlNetPipeline
from reptext_b200.pipeline_utils import SyntheticTextEncoders, SyntheticVAE
from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
from util import box_mask

H = W = 1024
dt, dev = torch.bfloat16, torch.device("cuda")
tr = models.FluxTransformer2DModel.random_init(config.FLUX_DEV, seed=100, dtype=dt, device=dev)
cn = models.FluxControlNetModel.random_init(config.REPTEXT_CONTROLNET, seed=101, dtype=dt, device=dev)
pipe = FluxControlNetPipeline(FlowMatchEulerDiscreteScheduler(), SyntheticVAE(dtype=dt, device=dev),
                              SyntheticTextEncoders(4096, 768, dt, dev), None, None, None, tr, cn)
g = torch.Generator().manual_seed(11)
canny = [torch.rand(1, 3, H, W, generator=g) * 2 - 1]
masks = [box_mask(H, W, (300, 460, 120, 900))]
poss = [(torch.from_numpy(m)[None, None].float() / 255.0) * 2 - 1 for m in masks]
pe = torch.randn(1, 512, 4096, generator=g).to(dt).to(dev)
po = torch.randn(1, 768, generator=g).to(dt).to(dev)
lat = torch.randn(1, 4096, 64, generator=g).to(dt).to(dev)
h_tap = torch.empty(1, 4096, 64, dtype=dt).pin_memory()
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 4
blocking = "--blocking" in sys.argv


def image(on, steps=28):
    pipe.precompute_modulation = on
    stamps = []

    def tap(p, i, t, kw):
        h_tap.copy_(kw["latents"], non_blocking=not blocking)
        stamps.append(time.perf_counter())
        return {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pipe(prompt_embeds=pe, pooled_prompt_embeds=po, height=H, width=W, num_inference_steps=steps, guidance_scale=3.5,
         control_image=canny, control_position=poss, control_mask=masks, controlnet_conditioning_scale=1.0,
         latents=lat.clone(), output_type="latent", callback_on_step_end=tap)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    gaps = [round((b - a) * 1e3, 1) for a, b in zip([t0] + stamps[:-1], stamps)]
    return (t1 - t0) * 1e3, gaps


image(True, 2)
for on in (True, False, True, False):
    for k in range(n_img):
        ms, gaps = image(on)
        worst = max(gaps)
        print(f"table={'on ' if on else 'off'} image {k}: {ms:8.1f} ms; host gap per step max {worst:6.1f} ms at step {gaps.index(worst)}"
              + (f"; gaps {gaps}" if blocking or worst > 150 else ""), flush=True)

#!/usr/bin/env python
"""BASELINE.json configs[2]: a sweep of independent (prompt, seed) samples sharded over the GPUs of one node.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/sweep.py \\
        --config flux-dev --samples 64 --steps 28 --height 1024 --width 1024

Every rank builds the same random-init model pair (no Hub access here), takes samples rank::world
(`reptext_b200.parallel.run_sharded`), runs the T2I pipeline on each - synthetic Arabic glyph lines rendered on the host
like RepText/infer.py does - and the output latents are all-gathered over NCCL in sample order.  Rank 0 prints one JSON
line with images/s for the whole job (wall clock between two barriers, model construction excluded)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORDS = ["مرحبا", "سلام", "القاهرة", "مكتبة", "مقهى", "سوق", "بحر", "نور", "حديقة", "مدرسة", "طريق", "قمر"]


def main(argv=None):
    from reptext_b200 import config, glyphs, models, parallel
    from reptext_b200.pipeline_flux_controlnet import FluxControlNetPipeline
    from reptext_b200.pipeline_utils import SyntheticTextEncoders
    from reptext_b200.vae import AutoencoderKL
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="small", choices=["small", "flux-dev"])
    ap.add_argument("--samples", type=int, default=8)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--width", type=int, default=256)
    a = ap.parse_args(argv)
    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dt = torch.bfloat16
    TR, CN = {"small": (config.SMALL128_TRANSFORMER, config.SMALL128_CONTROLNET),
              "flux-dev": (config.FLUX_DEV, config.REPTEXT_CONTROLNET)}[a.config]
    tr = models.FluxTransformer2DModel.random_init(TR, seed=100, dtype=dt, device=dev)
    cn = models.FluxControlNetModel.random_init(CN, seed=101, dtype=dt, device=dev)
    pipe = FluxControlNetPipeline(FlowMatchEulerDiscreteScheduler(), AutoencoderKL.random_init(seed=102, dtype=dt, device=dev),
                                  SyntheticTextEncoders(TR["joint_attention_dim"], TR["pooled_projection_dim"], dt, dev),
                                  None, None, None, tr, cn)
    font = glyphs.load_font(None, max(a.height // 12, 16))
    samples = [dict(seed=1000 + i, text=f"{WORDS[i % len(WORDS)]} {i}",
                    pos=(a.width // 8 + (i * 37) % (a.width // 3), a.height // 5 + (i * 53) % (a.height // 2)))
               for i in range(a.samples)]

    def denoise(i, s):
        cond = glyphs.build_conditions([s["text"]], [s["pos"]], [(255, 255, 255)], a.width, a.height, font)
        prompt = glyphs.build_prompt("a street sign in city", [s["text"]])
        return pipe(prompt, control_image=cond.control_image, control_position=cond.control_position,
                    control_mask=cond.control_mask, control_glyph=None, controlnet_conditioning_scale=1.0,
                    width=a.width, height=a.height, num_inference_steps=a.steps, guidance_scale=3.5,
                    generator=torch.Generator(device=dev).manual_seed(s["seed"]), output_type="latent").images[0]

    denoise(0, samples[0])                      # warm-up (kernel attributes, workspaces)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    lat = parallel.run_sharded(samples, denoise)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt_s = time.perf_counter() - t0
    ok = bool(torch.isfinite(lat.float()).all().item()) and lat.shape[0] == a.samples
    if rank == 0:
        print(json.dumps(dict(workload=f"cfg3 sweep: {a.samples} (prompt, seed) samples, {a.config}, {a.width}x{a.height}, "
                                       f"{a.steps} steps, host glyph rendering + pipeline + NCCL gather of the latents",
                              n_gpus=world, seconds=dt_s, images_per_s=a.samples / dt_s,
                              steps_per_s=a.samples * a.steps / dt_s, gathered_shape=list(lat.shape), finite=ok)))
    if world > 1:
        dist.destroy_process_group()
    return lat


if __name__ == "__main__":
    main()

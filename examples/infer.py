#!/usr/bin/env python
"""``RepText/infer.py`` on the B200 runtime: the same flow (font -> per-line glyph / position / mask / Canny ->
``pipe(prompt, control_image=..., control_position=..., control_mask=..., control_glyph=...)``), with two differences
forced by the offline box: without checkpoints the modules are random-init and the text encoders are the synthetic
stand-ins of ``reptext_b200.pipeline_utils``; with LOCAL copies of the two repositories the reference's own two loading
lines (``RepText/infer.py:30-33``) run as they are.

    python examples/infer.py --config small --text "مرحبا" --text "RepText" --steps 4 --out results/result.png
    python examples/infer.py --config flux-dev --height 1024 --width 1024 --steps 30          # FLUX.1-dev shapes
    python examples/infer.py --base-model /models/FLUX.1-dev --controlnet-model /models/RepText --height 1024 \
        --width 1024 --steps 30 --font-size 80 --out results/result.png                       # real checkpoints
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main(argv=None):
    from reptext_b200 import config, glyphs, models
    from reptext_b200.pipeline_flux_controlnet import FluxControlNetPipeline
    from reptext_b200.pipeline_utils import SyntheticTextEncoders, SyntheticVAE
    from reptext_b200.vae import AutoencoderKL
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="small", choices=["tiny", "small", "flux-dev"])
    ap.add_argument("--text", action="append", help="one text line (repeatable)")
    ap.add_argument("--font", default=None, help="TrueType font file (default: PIL's bundled face)")
    ap.add_argument("--font-size", type=int, default=40)
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--prompt", default="a street sign in city")
    ap.add_argument("--no-glyph-init", action="store_true", help="control_glyph=None (infer.py:124 'optional')")
    ap.add_argument("--out", default=None)
    ap.add_argument("--output-type", default="pil", choices=["pil", "latent"])
    ap.add_argument("--vae", default="flux", choices=["flux", "synthetic"],
                    help="flux: the AutoencoderKL drop-in (FLUX.1-dev VAE architecture, random weights; fp32 'tiny' uses the "
                         "stand-in); synthetic: the 8x-pooling stand-in")
    ap.add_argument("--base-model", default=None, help="local copy of black-forest-labs/FLUX.1-dev (infer.py:27)")
    ap.add_argument("--controlnet-model", default=None, help="local copy of Shakker-Labs/RepText (infer.py:28)")
    a = ap.parse_args(argv)
    if bool(a.base_model) != bool(a.controlnet_model):
        ap.error("--base-model and --controlnet-model go together")

    TR, CN, dt = {"tiny": (config.TINY_TRANSFORMER, config.TINY_CONTROLNET, torch.float32),
                  "small": (config.SMALL128_TRANSFORMER, config.SMALL128_CONTROLNET, torch.bfloat16),
                  "flux-dev": (config.FLUX_DEV, config.REPTEXT_CONTROLNET, torch.bfloat16)}[a.config]
    dev = torch.device("cuda")
    if a.base_model:
        # RepText/infer.py:30-33, unchanged
        controlnet = models.FluxControlNetModel.from_pretrained(a.controlnet_model, torch_dtype=torch.bfloat16)
        pipe = FluxControlNetPipeline.from_pretrained(
            a.base_model, controlnet=controlnet, torch_dtype=torch.bfloat16
        ).to("cuda")
        return _run(a, pipe, glyphs)
    controlnet = models.FluxControlNetModel.random_init(CN, seed=101, dtype=dt, device=dev)
    transformer = models.FluxTransformer2DModel.random_init(TR, seed=100, dtype=dt, device=dev)
    vae = (AutoencoderKL.random_init(seed=102, dtype=dt, device=dev) if a.vae == "flux" and dt == torch.bfloat16
           else SyntheticVAE(dtype=dt, device=dev))
    pipe = FluxControlNetPipeline(FlowMatchEulerDiscreteScheduler(), vae,
                                  SyntheticTextEncoders(TR["joint_attention_dim"], TR["pooled_projection_dim"], dt, dev),
                                  None, None, None, transformer, controlnet)
    return _run(a, pipe, glyphs)


def _run(a, pipe, glyphs):
    text_list = a.text or ["مرحبا بالعالم", "RepText"]
    font = glyphs.load_font(a.font, a.font_size)
    line_h = int(a.font_size * 1.6)
    text_position_list = [(a.width // 8, a.height // 4 + i * line_h) for i in range(len(text_list))]
    text_color_list = [(255, 255, 255)] * len(text_list)
    cond = glyphs.build_conditions(text_list, text_position_list, text_color_list, a.width, a.height, font)
    prompt = glyphs.build_prompt(a.prompt, text_list, ", filmfotos, film grain, reversal film photography")
    print(prompt)

    generator = torch.Generator(device="cuda").manual_seed(a.seed)
    image = pipe(
        prompt,
        control_image=cond.control_image,        # canny
        control_position=cond.control_position,  # position
        control_mask=cond.control_mask,          # regional mask
        control_glyph=None if a.no_glyph_init else cond.control_glyph,
        controlnet_conditioning_scale=1.0,
        controlnet_conditioning_step=30,
        width=a.width,
        height=a.height,
        num_inference_steps=a.steps,
        guidance_scale=3.5,
        generator=generator,
        output_type=a.output_type,
    ).images
    if a.output_type == "pil":
        image = image[0]
        if a.out:
            os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
            image.save(a.out)
    return image


if __name__ == "__main__":
    main()

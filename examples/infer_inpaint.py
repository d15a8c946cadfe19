#!/usr/bin/env python
"""``RepText/infer_inpaint.py`` on the B200 runtime: the same flow (source photograph -> ``resize_img`` -> font -> per-line
glyph / position / mask / Canny, the position box grown by 5 pixels like the mask -> ``pipe(prompt, control_image=...,
control_position=..., control_mask=..., control_glyph=..., true_guidance_scale=..., control_image_inpaint=...,
control_mask_inpaint=..., controlnet_conditioning_scale_inpaint=...)``).  As in ``examples/infer.py``: without checkpoints
the modules are random-init and the text encoders are the synthetic stand-ins; with LOCAL copies of the three repositories
the reference's own loading lines (``RepText/infer_inpaint.py:57-65``) run as they are.  Without ``--image`` a synthetic
photograph (seeded noise) stands in for ``assets/*.jpg``.

    python examples/infer_inpaint.py --config small --text "RepText" --steps 4 --out results/result_inpaint.png
    python examples/infer_inpaint.py --base-model /models/FLUX.1-dev --controlnet-model /models/RepText \
        --inpaint-model /models/FLUX.1-dev-Controlnet-Inpainting-Beta --image photo.jpg --steps 30 --font-size 70
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402


def main(argv=None):
    from reptext_b200 import config, glyphs, models
    from reptext_b200.pipeline_flux_controlnet_inpaint import FluxControlNetPipeline
    from reptext_b200.pipeline_utils import SyntheticTextEncoders, SyntheticVAE
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from reptext_b200.vae import AutoencoderKL

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="small", choices=["tiny", "small", "flux-dev"])
    ap.add_argument("--image", default=None, help="source photograph (default: a synthetic one of --width x --height)")
    ap.add_argument("--text", action="append", help="one text line (repeatable)")
    ap.add_argument("--font", default=None, help="TrueType font file (default: PIL's bundled face)")
    ap.add_argument("--font-size", type=int, default=40)
    ap.add_argument("--width", type=int, default=256, help="synthetic photograph only; a real one sets the size (resize_img)")
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--prompt", default="a street photo, wall")
    ap.add_argument("--true-guidance-scale", type=float, default=3.5, help="1.0 disables negative guidance")
    ap.add_argument("--out", default=None)
    ap.add_argument("--output-type", default="pil", choices=["pil", "latent"])
    ap.add_argument("--vae", default="flux", choices=["flux", "synthetic"])
    ap.add_argument("--base-model", default=None, help="local copy of black-forest-labs/FLUX.1-dev")
    ap.add_argument("--controlnet-model", default=None, help="local copy of Shakker-Labs/RepText")
    ap.add_argument("--inpaint-model", default=None, help="local copy of alimama-creative/FLUX.1-dev-Controlnet-Inpainting-Beta")
    a = ap.parse_args(argv)
    paths = (a.base_model, a.controlnet_model, a.inpaint_model)
    if any(paths) and not all(paths):
        ap.error("--base-model, --controlnet-model and --inpaint-model go together")

    if a.base_model:
        # RepText/infer_inpaint.py:57-65, unchanged
        controlnet = models.FluxControlNetModel.from_pretrained(a.controlnet_model, torch_dtype=torch.bfloat16)
        controlnet_inpaint = models.FluxControlNetModel.from_pretrained(a.inpaint_model, torch_dtype=torch.bfloat16)
        pipe = FluxControlNetPipeline.from_pretrained(
            a.base_model, controlnet=controlnet, controlnet_inpaint=controlnet_inpaint, torch_dtype=torch.bfloat16
        ).to("cuda")
        return _run(a, pipe, glyphs)
    TR, CN, CNI, dt = {"tiny": (config.TINY_TRANSFORMER, config.TINY_CONTROLNET, config.TINY_INPAINT_CONTROLNET, torch.float32),
                       "small": (config.SMALL128_TRANSFORMER, config.SMALL128_CONTROLNET, config.SMALL128_INPAINT_CONTROLNET,
                                 torch.bfloat16),
                       "flux-dev": (config.FLUX_DEV, config.REPTEXT_CONTROLNET, config.INPAINT_CONTROLNET,
                                    torch.bfloat16)}[a.config]
    dev = torch.device("cuda")
    transformer = models.FluxTransformer2DModel.random_init(TR, seed=100, dtype=dt, device=dev)
    controlnet = models.FluxControlNetModel.random_init(CN, seed=101, dtype=dt, device=dev)
    controlnet_inpaint = models.FluxControlNetModel.random_init(CNI, seed=103, dtype=dt, device=dev)
    vae = (AutoencoderKL.random_init(seed=102, dtype=dt, device=dev) if a.vae == "flux" and dt == torch.bfloat16
           else SyntheticVAE(dtype=dt, device=dev))
    pipe = FluxControlNetPipeline(FlowMatchEulerDiscreteScheduler(), vae,
                                  SyntheticTextEncoders(TR["joint_attention_dim"], TR["pooled_projection_dim"], dt, dev),
                                  None, None, None, transformer, controlnet, controlnet_inpaint)
    return _run(a, pipe, glyphs)


def _run(a, pipe, glyphs):
    if a.image:
        control_image_inpaint = glyphs.resize_img(Image.open(a.image).convert("RGB"))     # infer_inpaint.py:67-68
    else:
        rs = np.random.RandomState(a.seed)
        control_image_inpaint = Image.fromarray(rs.randint(0, 255, (a.height, a.width, 3)).astype(np.uint8))
    width, height = control_image_inpaint.size

    text_list = a.text or ["RepText"]
    font = glyphs.load_font(a.font, a.font_size)
    line_h = int(a.font_size * 1.6)
    text_position_list = [(width // 8, height // 4 + i * line_h) for i in range(len(text_list))]
    text_color_list = [(0, 255, 0)] * len(text_list)
    cond = glyphs.build_conditions(text_list, text_position_list, text_color_list, width, height, font, position_margin=5)
    prompt = glyphs.build_prompt(a.prompt, text_list, ", filmfotos, film grain, reversal film photography")
    print(prompt)

    generator = torch.Generator(device="cuda").manual_seed(a.seed)
    image = pipe(
        prompt,
        true_guidance_scale=a.true_guidance_scale,
        # for text rendering
        control_image=cond.control_image,        # canny
        control_position=cond.control_position,  # position
        control_mask=cond.control_mask,          # regional mask
        control_glyph=cond.control_glyph,        # as init latent
        # for inpainting (the reference passes the LAST line's regional mask, infer_inpaint.py:144)
        control_image_inpaint=control_image_inpaint,
        control_mask_inpaint=cond.control_mask[-1],
        controlnet_conditioning_scale_inpaint=1.0,
        controlnet_conditioning_scale=1.0,
        controlnet_conditioning_step=30,
        width=width,
        height=height,
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

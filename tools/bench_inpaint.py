#!/usr/bin/env python
"""BASELINE.json configs[3] timed end to end: the inpaint pipeline (pipeline_flux_controlnet_inpaint) at 1024x1024 with
FLUX.1-dev-architecture weights (random init), the RepText ControlNet, the inpainting ControlNet and true CFG (effective
batch 2).  Reports ms per denoising step (CUDA events around the loop, first step excluded) and the tensor utilisation
against SURVEY.md 8(d)'s 1.8201e14 FLOPs per step.  Not the headline metric: a parity-config timing for profiles/."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch


def main(steps=10, H=1024, W=1024, T=512):
    from reptext_b200 import config, models
    from reptext_b200.pipeline_flux_controlnet_inpaint import FluxControlNetPipeline
    from reptext_b200.pipeline_utils import SyntheticTextEncoders, SyntheticVAE
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from util import box_mask
    dt, dev = torch.bfloat16, torch.device("cuda")
    TR, CN, CNI = config.FLUX_DEV, config.REPTEXT_CONTROLNET, config.INPAINT_CONTROLNET
    tr = models.FluxTransformer2DModel.random_init(TR, seed=100, dtype=dt, device=dev)
    cn = models.FluxControlNetModel.random_init(CN, seed=101, dtype=dt, device=dev)
    cni = models.FluxControlNetModel.random_init(CNI, seed=103, dtype=dt, device=dev)
    pipe = FluxControlNetPipeline(FlowMatchEulerDiscreteScheduler(), SyntheticVAE(dtype=dt, device=dev),
                                  SyntheticTextEncoders(TR["joint_attention_dim"], TR["pooled_projection_dim"], dt, dev),
                                  None, None, None, tr, cn, cni)
    g = torch.Generator().manual_seed(0)
    N = (H // 16) * (W // 16)
    pe = torch.randn(1, T, TR["joint_attention_dim"], generator=g).to(dt)
    po = torch.randn(1, TR["pooled_projection_dim"], generator=g).to(dt)
    npe = torch.randn(1, T, TR["joint_attention_dim"], generator=g).to(dt)
    npo = torch.randn(1, TR["pooled_projection_dim"], generator=g).to(dt)
    canny = torch.rand(1, 3, H, W, generator=g) * 2 - 1
    mask_img = box_mask(H, W, (H // 3, H // 3 + H // 6, W // 5, W - W // 5))
    pos = (torch.from_numpy(mask_img)[None, None].float() / 255.0) * 2 - 1
    src = (torch.rand(1, 3, H, W, generator=g) * 2 - 1)
    ev = []

    def tap(p, i, t, kw):
        e = torch.cuda.Event(enable_timing=True); e.record(); ev.append(e)
        return {}

    def run(n):
        ev.clear()
        return pipe(prompt_embeds=pe, pooled_prompt_embeds=po, negative_prompt_embeds=npe, negative_pooled_prompt_embeds=npo,
                    height=H, width=W, num_inference_steps=n, guidance_scale=3.5, true_guidance_scale=3.5,
                    control_image=[canny], control_position=[pos], control_mask=[mask_img], control_glyph=src,
                    control_image_inpaint=src, control_mask_inpaint=mask_img, controlnet_conditioning_scale=1.0,
                    controlnet_conditioning_scale_inpaint=1.0, output_type="latent", callback_on_step_end=tap).images

    run(3)
    torch.cuda.synchronize()
    out = run(steps)
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(len(ev) - 1)]
    ms_step = float(np.median(ms))
    flops = 1.8201e14
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0}
    print(json.dumps(dict(workload="cfg4: inpaint pipeline 1024x1024, FLUX.1-dev arch + RepText ControlNet + inpaint ControlNet, true CFG (batch 2), bf16",
                          ms_per_step=ms_step, steps_per_s=1000.0 / ms_step, images_per_s_28_steps=1000.0 / ms_step / 28,
                          flops_per_step=flops, tensor_util_of_sustained=flops / (ms_step / 1e3) / 1e12 / pk["bf16_tflops_sustained"],
                          finite=bool(torch.isfinite(out.float()).all()), per_step_ms=[round(m, 2) for m in ms])))


if __name__ == "__main__":
    main()

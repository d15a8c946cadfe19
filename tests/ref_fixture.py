"""Seeded cases shared by ``tests/golden/make_golden_ref.py`` (runs the REFERENCE's own pipelines over ``tests/ref_shim``),
``tests/test_reference_pin.py`` (CPU: reference == oracle == committed golden) and ``tests/test_reference_gpu.py`` (GPU: the
product pipelines from the same PIL images / prompt strings == the golden).  TEST INFRASTRUCTURE ONLY.

Every case starts where ``RepText/infer.py`` / ``infer_inpaint.py`` hand over to the pipeline: prompt STRINGS, PIL Canny /
position / glyph images, numpy regional masks, a source image + inpaint mask, a seeded ``torch.Generator`` - so prompt
encoding, ``VaeImageProcessor``, ``prepare_image`` / ``prepare_image_with_mask`` / ``prepare_latents_reptext`` (rows a14-a18 of
SURVEY.md section 8) are inside the comparison, not assumed.

Weights: transformer / ControlNets from ``reptext_b200.weights.random_state_dict`` (diffusers names), VAE from
``oracle.vae_oracle.random_state_dict`` (diffusers names; loaded into the BFL autoencoder of the shim through the name map
below), CLIP / T5 from ``oracle.text_oracle.random_state_dict`` (transformers names; loaded into the REAL transformers
modules on the reference side).  The VAE's log-variance head is pinned to -30 (std 3e-7) so that ``latent_dist.sample()``,
which the reference draws from the GLOBAL RNG, is the posterior mean on both sides.
"""
from __future__ import annotations

import numpy as np
import torch

from reptext_b200 import config, weights

PROMPT = "a road sign that reads 'مرحبا' on a desert highway"
PROMPT_2 = "a weathered road sign with the words 'مرحبا بكم' at dusk"
NEGATIVE = "blurry, low quality"

_TINY = ("TINY_TRANSFORMER", "TINY_CONTROLNET", "TINY_INPAINT_CONTROLNET")
_SMALL = ("SMALL128_TRANSFORMER", "SMALL128_CONTROLNET", "SMALL128_INPAINT_CONTROLNET")

CASES = {
    # BASELINE.json configs[0] through infer.py's call shape: fp32, 4 Euler steps, two text lines, ControlNet gated off
    # for the last step
    "ref_tiny_t2i": dict(kind="t2i", dtype="fp32", models=_TINY, H=256, W=192, lines=2, steps=4, cond_step=3, T=64,
                         vae=(32, 64, 64, 64), guidance=3.5, scale=0.9, seed=5),
    # sizes that are multiples of 8 but not of 16: VaeImageProcessor rounds the images down to 256x192 (Lanczos)
    "ref_tiny_t2i_offgrid": dict(kind="t2i", dtype="fp32", models=_TINY, H=264, W=200, lines=1, steps=2, cond_step=30,
                                 T=64, vae=(32, 64, 64, 64), guidance=3.5, scale=1.0, seed=6),
    "ref_tiny_inpaint": dict(kind="inpaint", dtype="fp32", models=_TINY, H=256, W=192, lines=2, steps=3, cond_step=2,
                             T=64, vae=(32, 64, 64, 64), guidance=3.5, scale=0.9, scale_inpaint=0.8, true_cfg=3.0, seed=7),
    # a ControlNet WITH single-stream blocks (row a7): controlnet_single_block_samples reach the transformer
    "ref_tiny_t2i_single": dict(kind="t2i", dtype="fp32", models=_TINY, H=256, W=192, lines=2, steps=2, cond_step=30,
                                T=64, vae=(32, 64, 64, 64), guidance=3.5, scale=0.9, seed=8, cn_single_layers=2),
    # bf16 end to end on both sides (reference: torch CPU bf16), head_dim 128 -> the tcgen05 kernels; product VAE and
    # product prompt encoders
    "ref_small_t2i_bf16": dict(kind="t2i", dtype="bf16", models=_SMALL, H=256, W=256, lines=2, steps=3, cond_step=2,
                               T=128, vae=(64, 128, 128, 128), guidance=3.5, scale=0.9, seed=9),
    "ref_small_inpaint_bf16": dict(kind="inpaint", dtype="bf16", models=_SMALL, H=256, W=256, lines=1, steps=3,
                                   cond_step=30, T=128, vae=(64, 128, 128, 128), guidance=3.5, scale=1.0,
                                   scale_inpaint=0.8, true_cfg=3.0, seed=10),
}


def torch_dtype(case):
    return torch.float32 if case["dtype"] == "fp32" else torch.bfloat16


# --------------------------------------------------------------------------------------------------------- host inputs
def glyph_inputs(H, W, n_lines):
    """Synthetic glyph / Canny / position / regional-mask images shaped like RepText/infer.py:64-104 builds them."""
    from PIL import Image, ImageDraw, ImageFont
    font = ImageFont.load_default(max(H // 10, 12))
    glyph = Image.new("RGB", (W, H), (0, 0, 0))
    cannys, poss, masks = [], [], []
    for li in range(n_lines):
        y = (H // (n_lines + 1)) * (li + 1) - H // 16
        line = Image.new("RGB", (W, H), (0, 0, 0))
        for im in (glyph, line):
            ImageDraw.Draw(im).text((W // 6, y), "مرحبا %d" % li, font=font, fill=(255, 255, 255))
        x0, y0, x1, y1 = ImageDraw.Draw(line).textbbox((W // 6, y), "مرحبا %d" % li, font=font)
        pos = np.zeros((H, W), np.uint8)
        pos[y0:y1, x0:x1] = 255
        msk = np.zeros((H, W), np.uint8)
        msk[max(y0 - 5, 0):y1 + 5, max(x0 - 5, 0):x1 + 5] = 255
        arr = np.array(line.convert("L")).astype(np.int16)
        edge = ((np.abs(np.diff(arr, axis=0, prepend=0)) + np.abs(np.diff(arr, axis=1, prepend=0))) > 50)
        cannys.append(Image.fromarray((255 - edge.astype(np.uint8) * 255)).convert("RGB"))
        poss.append(Image.fromarray(pos))
        masks.append(Image.fromarray(msk))
    return glyph, cannys, poss, masks


def call_kwargs(case, generator_device="cpu"):
    """The keyword arguments of ``pipe(...)`` (RepText/infer.py:117-130, infer_inpaint.py:136-152)."""
    from PIL import Image
    H, W = case["H"], case["W"]
    glyph, cannys, poss, masks = glyph_inputs(H, W, case["lines"])
    kw = dict(prompt=PROMPT, prompt_2=PROMPT_2, height=H, width=W, num_inference_steps=case["steps"],
              guidance_scale=case["guidance"], control_image=cannys, control_position=poss, control_mask=masks,
              control_glyph=glyph, controlnet_conditioning_scale=case["scale"],
              controlnet_conditioning_step=case["cond_step"], max_sequence_length=case["T"],
              generator=torch.Generator(device=generator_device).manual_seed(case["seed"]), output_type="latent")
    if case["kind"] == "inpaint":
        src = np.random.RandomState(case["seed"]).randint(0, 255, (H, W, 3)).astype(np.uint8)
        kw.update(control_image_inpaint=Image.fromarray(src), control_mask_inpaint=masks[0],
                  true_guidance_scale=case["true_cfg"], controlnet_conditioning_scale_inpaint=case["scale_inpaint"],
                  negative_prompt=NEGATIVE)
    return kw


# --------------------------------------------------------------------------------------------------------- weights
def model_configs(case):
    TR, CN, CNI = (getattr(config, n) for n in case["models"])
    if case.get("cn_single_layers"):
        CN = dict(CN, num_single_layers=case["cn_single_layers"])
    return TR, CN, CNI


def _round(sd, case):
    if case["dtype"] == "bf16":
        return {k: v.to(torch.bfloat16).float() for k, v in sd.items()}
    return sd


def state_dicts(case):
    from oracle import text_oracle as TO
    from oracle import vae_oracle as V
    TR, CN, CNI = model_configs(case)
    out = dict(tr=weights.random_state_dict(TR, "transformer", seed=100),
               cn=weights.random_state_dict(CN, "controlnet", seed=101))
    if case["kind"] == "inpaint":
        out["cni"] = weights.random_state_dict(CNI, "controlnet", seed=103)
    vcfg = vae_config(case)
    vsd = V.random_state_dict(vcfg, seed=105)
    lc = vcfg["latent_channels"]
    vsd["encoder.conv_out.weight"][lc:] = 0.0          # log-variance head: constant -30 -> std = exp(-15)
    vsd["encoder.conv_out.bias"][lc:] = -30.0
    out["vae"] = vsd
    tcfg, ccfg = text_configs(case)
    out["t5"] = TO.random_state_dict(TO.t5_param_shapes(tcfg), seed=106)
    out["clip"] = TO.random_state_dict(TO.clip_param_shapes(ccfg), seed=107)
    return {k: _round(v, case) for k, v in out.items()}


def vae_config(case):
    from oracle import vae_oracle as V
    return dict(V.FLUX_VAE_CONFIG, block_out_channels=tuple(case["vae"]))


def text_configs(case):
    from oracle import text_oracle as TO
    TR = model_configs(case)[0]
    d, p = TR["joint_attention_dim"], TR["pooled_projection_dim"]
    tcfg = dict(TO.T5_XXL_CONFIG, vocab_size=1000, d_model=d, d_ff=2 * d, num_layers=2, num_heads=max(d // 64, 1))
    ccfg = dict(TO.CLIP_L_CONFIG, vocab_size=1000, hidden_size=p, intermediate_size=2 * p, num_hidden_layers=2,
                num_attention_heads=max(p // 64, 1))
    return tcfg, ccfg


def tokenizers():
    from reptext_b200.pipeline_utils import SyntheticTokenizer
    return SyntheticTokenizer("clip", 1000, 77), SyntheticTokenizer("t5", 1000, 512)


# --------------------------------------------------------------------------------------------------------- modules
def hf_text_encoders(case, sds, device, dtype):
    """The REAL transformers modules the reference pipelines are written against."""
    import transformers as tf
    tcfg, ccfg = text_configs(case)
    t5 = tf.T5EncoderModel(tf.T5Config(dropout_rate=0.0, **dict(tcfg, feed_forward_proj="gated-gelu"))).eval()
    clip = tf.CLIPTextModel(tf.CLIPTextConfig(attention_dropout=0.0, bos_token_id=998, pad_token_id=999, **ccfg)).eval()
    missing = t5.load_state_dict(sds["t5"], strict=False)
    assert not [k for k in missing.missing_keys if "embed_tokens" not in k], missing
    missing = clip.load_state_dict(sds["clip"], strict=False)
    assert not [k for k in missing.missing_keys if "position_ids" not in k], missing
    return clip.to(device, dtype), t5.to(device, dtype)


def _vae_name_map(ae):
    """(diffusers parameter stem, BFL module, is 1x1 convolution) for the BFL autoencoder of the shim."""
    out = []

    def resnet(name, blk):
        out.extend([(name + "norm1", blk.norm1, False), (name + "conv1", blk.conv1, False),
                    (name + "norm2", blk.norm2, False), (name + "conv2", blk.conv2, False)])
        if blk.in_channels != blk.out_channels:
            out.append((name + "conv_shortcut", blk.nin_shortcut, True))

    def mid(name, mm):
        resnet(name + "mid_block.resnets.0.", mm.block_1)
        resnet(name + "mid_block.resnets.1.", mm.block_2)
        a = name + "mid_block.attentions.0."
        out.append((a + "group_norm", mm.attn_1.norm, False))
        for dn, conv in (("to_q", mm.attn_1.q), ("to_k", mm.attn_1.k), ("to_v", mm.attn_1.v), ("to_out.0", mm.attn_1.proj_out)):
            out.append((a + dn, conv, True))

    e, d = ae.encoder, ae.decoder
    out.append(("encoder.conv_in", e.conv_in, False))
    for i, lvl in enumerate(e.down):
        for j, blk in enumerate(lvl.block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}.", blk)
        if hasattr(lvl, "downsample"):
            out.append((f"encoder.down_blocks.{i}.downsamplers.0.conv", lvl.downsample.conv, False))
    mid("encoder.", e.mid)
    out.extend([("encoder.conv_norm_out", e.norm_out, False), ("encoder.conv_out", e.conv_out, False),
                ("decoder.conv_in", d.conv_in, False)])
    mid("decoder.", d.mid)
    n = len(d.up)
    for i in range(n):                       # diffusers' up_blocks.0 runs first = BFL's up[n - 1]
        lvl = d.up[n - 1 - i]
        for j, blk in enumerate(lvl.block):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}.", blk)
        if hasattr(lvl, "upsample"):
            out.append((f"decoder.up_blocks.{i}.upsamplers.0.conv", lvl.upsample.conv, False))
    out.extend([("decoder.conv_norm_out", d.norm_out, False), ("decoder.conv_out", d.conv_out, False)])
    return out


def shim_vae(case, sds, device, dtype):
    """The shim's AutoencoderKL (BFL autoencoder inside) carrying the diffusers-named VAE state dict."""
    import ref_run
    ref_run.shim()
    from diffusers.models.autoencoders import AutoencoderKL
    cfg = vae_config(case)
    vae = AutoencoderKL(block_out_channels=cfg["block_out_channels"], latent_channels=cfg["latent_channels"],
                        layers_per_block=cfg["layers_per_block"], scaling_factor=cfg["scaling_factor"],
                        shift_factor=cfg["shift_factor"]).eval()
    sd, used = sds["vae"], set()
    with torch.no_grad():
        for stem, mod, one in _vae_name_map(vae.ae):
            w = sd[stem + ".weight"]
            mod.weight.copy_(w[:, :, None, None] if one and w.dim() == 2 else w)
            mod.bias.copy_(sd[stem + ".bias"])
            used.update((stem + ".weight", stem + ".bias"))
    assert used == set(sd), sorted(set(sd) ^ used)[:4]
    return vae.to(device, dtype)


def reference_pipeline(case, sds=None, device="cpu"):
    """The REFERENCE's ``FluxControlNetPipeline`` (by path, unmodified) assembled from shim / transformers modules."""
    import ref_run
    R = ref_run.load()
    from diffusers.models.transformers.transformer_flux import FluxTransformer2DModel
    from diffusers.schedulers import FlowMatchEulerDiscreteScheduler
    sds = sds or state_dicts(case)
    dt = torch_dtype(case)
    TR, CN, CNI = model_configs(case)
    ls = lambda c: {k: (list(v) if isinstance(v, tuple) else v) for k, v in c.items()}
    tr = FluxTransformer2DModel(**ls(TR)).eval()
    tr.load_state_dict(sds["tr"], strict=True)
    cn = R.controlnet_flux.FluxControlNetModel(**ls(CN)).eval()
    cn.load_state_dict(sds["cn"], strict=True)
    clip, t5 = hf_text_encoders(case, sds, device, dt)
    tok, tok2 = tokenizers()
    vae = shim_vae(case, sds, device, dt)
    sch = FlowMatchEulerDiscreteScheduler(**config.SCHEDULER)
    args = dict(scheduler=sch, vae=vae, text_encoder=clip, tokenizer=tok, text_encoder_2=t5, tokenizer_2=tok2,
                transformer=tr.to(device, dt), controlnet=cn.to(device, dt))
    if case["kind"] == "inpaint":
        cni = R.controlnet_flux.FluxControlNetModel(**ls(CNI)).eval()
        cni.load_state_dict(sds["cni"], strict=True)
        return R.inpaint.FluxControlNetPipeline(controlnet_inpaint=cni.to(device, dt), **args)
    return R.t2i.FluxControlNetPipeline(**args)


def product_pipeline(case, sds=None, device="cuda"):
    """The product pipeline on the GPU with the same weights.  fp32 cases: the VAE / prompt encoders are outside the fp32
    hot path (the product's own are bf16 kernels), so the shim VAE and the transformers modules run on the GPU in fp32;
    bf16 cases: the product's own VAE and prompt encoders."""
    from reptext_b200 import models
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
    sds = sds or state_dicts(case)
    dt = torch_dtype(case)
    TR, CN, CNI = model_configs(case)
    tr = models.FluxTransformer2DModel(TR, sds["tr"], dtype=dt, device=device)
    cn = models.FluxControlNetModel(CN, sds["cn"], dtype=dt, device=device)
    tok, tok2 = tokenizers()
    if case["dtype"] == "fp32":
        clip, t5 = hf_text_encoders(case, sds, device, dt)
        vae = shim_vae(case, sds, device, dt)
    else:
        from reptext_b200 import text_encoders as TE
        from reptext_b200.vae import AutoencoderKL
        tcfg, ccfg = text_configs(case)
        clip, t5 = TE.CLIPTextModel(ccfg, sds["clip"], device=device), TE.T5EncoderModel(tcfg, sds["t5"], device=device)
        vae = AutoencoderKL(vae_config(case), sds["vae"], device=device)
    sch = FlowMatchEulerDiscreteScheduler()
    if case["kind"] == "inpaint":
        from reptext_b200.pipeline_flux_controlnet_inpaint import FluxControlNetPipeline
        cni = models.FluxControlNetModel(CNI, sds["cni"], dtype=dt, device=device)
        return FluxControlNetPipeline(sch, vae, clip, tok, t5, tok2, tr, cn, cni)
    from reptext_b200.pipeline_flux_controlnet import FluxControlNetPipeline
    return FluxControlNetPipeline(sch, vae, clip, tok, t5, tok2, tr, cn)


def run_with_taps(pipe, kw):
    """-> (per-step latents [steps, B, N, 64] float32 on the CPU, final output latents)."""
    taps = []

    def tap(p, i, t, k):
        taps.append(k["latents"].detach().float().cpu().clone())
        return {}

    out = pipe(callback_on_step_end=tap, **kw)
    images = out.images if hasattr(out, "images") else out[0]
    return torch.stack(taps), images.detach().float().cpu()


def capture_first_step(pipe):
    """Forward pre-hooks on the reference pipeline's nn.Modules: the keyword arguments of every ``controlnet`` /
    ``controlnet_inpaint`` / ``transformer`` call of step 0, i.e. the tensors the reference's own preparation code built
    (nothing of the reference is modified; the hooks are torch's)."""
    box = dict(controlnet=[], controlnet_inpaint=[], transformer=[])
    handles = []

    def hook(name):
        def fn(mod, args, kwargs):
            box[name].append({k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in kwargs.items()})
        return fn

    def thook(mod, args, kwargs):
        box["transformer"].append({k: (v.detach().clone() if torch.is_tensor(v) else
                                       ([s.detach().clone() for s in v] if isinstance(v, (list, tuple)) else v))
                                   for k, v in kwargs.items()})
        for h in handles:                       # step 0 is complete: stop recording
            h.remove()

    for name in ("controlnet", "controlnet_inpaint"):
        m = getattr(pipe, name, None)
        if m is not None:
            handles.append(m.register_forward_pre_hook(hook(name), with_kwargs=True))
    handles.append(pipe.transformer.register_forward_pre_hook(thook, with_kwargs=True))
    return box


def prepared_from_capture(box):
    """-> dict of float32 CPU tensors: what the reference prepared before / during step 0."""
    f = lambda t: t.detach().float().cpu()
    c0 = box["controlnet"][0]
    out = dict(init_latents=f(c0["hidden_states"]), conds=torch.stack([f(c["controlnet_cond"]) for c in box["controlnet"]]),
               prompt_embeds=f(c0["encoder_hidden_states"]), pooled=f(c0["pooled_projections"]),
               txt_ids=f(c0["txt_ids"]), img_ids=f(c0["img_ids"]), timestep0=f(c0["timestep"]))
    if box["controlnet_inpaint"]:
        out["cond_inpaint"] = f(box["controlnet_inpaint"][0]["controlnet_cond"])
    t = box["transformer"][0]
    if t.get("controlnet_block_samples") is not None:
        out["block_samples0"] = torch.stack([f(s) for s in t["controlnet_block_samples"]])
    if t.get("controlnet_single_block_samples") is not None:
        out["single_block_samples0"] = torch.stack([f(s) for s in t["controlnet_single_block_samples"]])
    return out


def oracle_loop(name, z, dtype, time_dtype=None):
    """oracle.denoise_t2i / denoise_inpaint on the tensors the reference prepared (``z`` = a golden record) in ``dtype``;
    ``time_dtype`` = the dtype the reference held ``timestep`` in (bf16 runs round it: controlnet_flux.py:282).
    -> per-step latents [steps, B, N, 64] float32."""
    from oracle import flux_oracle as O
    case = CASES[name]
    sds = state_dicts(case)
    TR, CN, CNI = model_configs(case)
    c = lambda t: t.to(dtype)
    cs = lambda sd: {k: v.to(dtype) for k, v in sd.items()}
    _, _, _, masks = glyph_inputs(case["H"], case["W"], case["lines"])
    mask_dtype = torch_dtype(case)               # the reference casts the mask to latents.dtype
    taps = []
    args = dict(latents=c(z["init_latents"]), prompt_embeds=c(z["prompt_embeds"]), pooled=c(z["pooled"]),
                control_image_list=[c(x) for x in z["conds"]],
                control_mask_list=[O.regional_mask(np.array(m), mask_dtype).to(dtype) for m in masks], text_ids=c(z["txt_ids"]),
                img_ids=c(z["img_ids"]), timesteps=z["timesteps"], sigmas=z["sigmas"], guidance_scale=case["guidance"],
                conditioning_scale=case["scale"], conditioning_step=case["cond_step"],
                callback=lambda i, t, lat: taps.append(lat.float().clone()), time_dtype=time_dtype)
    with torch.no_grad():
        if case["kind"] == "inpaint":
            O.denoise_inpaint(cs(sds["tr"]), TR, cs(sds["cn"]), CN, cs(sds["cni"]), CNI,
                              control_image_inpaint=c(z["cond_inpaint"]), true_guidance_scale=case["true_cfg"],
                              conditioning_scale_inpaint=case["scale_inpaint"], **args)
        else:
            O.denoise_t2i(cs(sds["tr"]), TR, cs(sds["cn"]), CN, **args)
    return torch.stack(taps)

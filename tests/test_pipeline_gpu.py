"""Pipeline-level parity on the GPU: the reference-shaped ``FluxControlNetPipeline.__call__`` (T2I and inpaint)
against the committed golden vectors (tests/golden/*.npz, fp32 tiny config, BASELINE.json configs[0]) and against
the oracle's denoise loops on the tensors the pipeline itself prepared.

Bars: per-step latent rel-L2 <= 1e-4 in fp32, <= 1e-2 in bf16 (BASELINE.json north_star).
"""
import os
import sys

import numpy as np
import pytest
import torch

from util import box_mask, rel_l2, synth_inputs

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


class LatentPassthroughVAE:
    """Test double: ``encode(x)`` hands back the first 16 channels of ``x`` as the posterior sample (un-doing the
    pipeline's shift/scale), so a test can inject exact packed control latents through ``control_image``."""

    def __init__(self, dtype):
        from reptext_b200.models import FrozenConfig
        self.dtype = dtype
        self.config = FrozenConfig(shift_factor=0.1159, scaling_factor=0.3611, block_out_channels=(1, 1, 1, 1))

    def encode(self, x):
        from reptext_b200.models import FrozenConfig
        z = x[:, :16].double() / self.config.scaling_factor + self.config.shift_factor
        return FrozenConfig(latent_dist=FrozenConfig(sample=lambda generator=None: z))


def _unpack_cond(cond, lh, lw):
    """[B, N, 4C] -> [B, C, lh, lw] (inverse of _pack_latents)."""
    b, n, ch = cond.shape
    x = cond.view(b, lh // 2, lw // 2, ch // 4, 2, 2).permute(0, 3, 1, 4, 2, 5)
    return x.reshape(b, ch // 4, lh, lw)


def _tiny_pipe(dtype, inpaint=False, vae=None, TRname="TINY_TRANSFORMER", CNname="TINY_CONTROLNET",
               CNIname="TINY_INPAINT_CONTROLNET"):
    from reptext_b200 import config, models, weights
    from reptext_b200.pipeline_utils import SyntheticTextEncoders
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
    TR, CN = getattr(config, TRname), getattr(config, CNname)
    rnd = (lambda sd: {k: v.to(dtype).float() for k, v in sd.items()}) if dtype == torch.bfloat16 else (lambda sd: sd)
    tr_sd = rnd(weights.random_state_dict(TR, "transformer", seed=100))
    cn_sd = rnd(weights.random_state_dict(CN, "controlnet", seed=101))
    tr = models.FluxTransformer2DModel(TR, tr_sd, dtype=dtype)
    cn = models.FluxControlNetModel(CN, cn_sd, dtype=dtype)
    enc = SyntheticTextEncoders(TR["joint_attention_dim"], TR["pooled_projection_dim"], dtype=dtype)
    sch = FlowMatchEulerDiscreteScheduler()
    vae = vae or LatentPassthroughVAE(dtype)
    sds = dict(tr=tr_sd, cn=cn_sd)
    if not inpaint:
        from reptext_b200.pipeline_flux_controlnet import FluxControlNetPipeline
        return FluxControlNetPipeline(sch, vae, enc, None, None, None, tr, cn), TR, CN, None, sds
    from reptext_b200.pipeline_flux_controlnet_inpaint import FluxControlNetPipeline
    CNI = getattr(config, CNIname)
    cni_sd = rnd(weights.random_state_dict(CNI, "controlnet", seed=103))
    sds["cni"] = cni_sd
    cni = models.FluxControlNetModel(CNI, cni_sd, dtype=dtype)
    return FluxControlNetPipeline(sch, vae, enc, None, None, None, tr, cn, cni), TR, CN, CNI, sds


@pytest.mark.parametrize("name", ["tiny_t2i", "tiny_t2i_2lines"])
def test_t2i_call_reproduces_golden_latents(name):
    """BASELINE.json configs[0] through the public __call__: 256x256, Euler steps, batch 1, fp32."""
    import make_golden
    spec = make_golden.CASES[name]
    pipe, TR, CN, _, _ = _tiny_pipe(torch.float32)
    H, W, T = spec["H"], spec["W"], spec["T"]
    x = synth_inputs(TR, CN, H, W, T, seed=102, n_lines=spec["lines"])
    lh, lw = 2 * (H // 16), 2 * (W // 16)
    imgs, poss, masks = [], [], []
    for li, cond in enumerate(x["conds"]):
        z = _unpack_cond(cond, lh, lw)               # [1, 32, lh, lw] = canny latents | position latents
        imgs.append(z[:, :16].cuda())
        poss.append(z[:, 16:].cuda())
        y0 = (H // 4) * (li + 1) - H // 8
        masks.append(box_mask(H, W, (y0, y0 + H // 6, W // 5, W - W // 5)))   # same boxes as util.synth_inputs
    taps = []

    def tap(p, i, t, kw):
        taps.append(kw["latents"].float().cpu())
        return {}

    # the position latents go through `.repeat(1, 3, 1, 1)` upstream; the test VAE reads the first 16 channels
    out = pipe(prompt_embeds=x["prompt_embeds"].cuda(), pooled_prompt_embeds=x["pooled"].cuda(), height=H, width=W,
               num_inference_steps=spec["steps"], guidance_scale=3.5, control_image=imgs, control_position=poss,
               control_mask=masks, controlnet_conditioning_scale=1.0,
               controlnet_conditioning_step=spec.get("cond_step", 30), latents=x["latents"].cuda(),
               output_type="latent", callback_on_step_end=tap)
    want = torch.from_numpy(np.load(os.path.join(HERE, "golden", name + ".npz"))["latents_per_step"])
    assert len(taps) == want.shape[0]
    for i, got in enumerate(taps):
        assert rel_l2(got, want[i]) < 1e-4, (name, i, rel_l2(got, want[i]))
    assert torch.equal(out.images.float().cpu(), taps[-1])


def test_inpaint_denoise_reproduces_golden_latents():
    import make_golden
    spec = make_golden.CASES["tiny_inpaint"]
    pipe, TR, CN, CNI, _ = _tiny_pipe(torch.float32, inpaint=True)
    x = synth_inputs(TR, CN, spec["H"], spec["W"], spec["T"], seed=102, n_lines=1)
    g = torch.Generator().manual_seed(104)
    neg_pe = torch.randn(1, spec["T"], TR["joint_attention_dim"], generator=g)
    neg_po = torch.randn(1, TR["pooled_projection_dim"], generator=g)
    cond_inp = torch.randn(1, x["N"], 68, generator=g)
    dev = "cuda"
    pipe._guidance_scale = 3.5
    sc = pipe.scheduler.config
    from reptext_b200._pipeline_common import calculate_shift, retrieve_timesteps
    mu = calculate_shift(x["N"], sc.base_image_seq_len, sc.max_image_seq_len, sc.base_shift, sc.max_shift)
    ts, n = retrieve_timesteps(pipe.scheduler, spec["steps"], dev, None, np.linspace(1.0, 1 / spec["steps"], spec["steps"]), mu=mu)
    taps = []
    lat = pipe._denoise(
        latents=x["latents"].to(dev), latent_image_ids=x["img_ids"].to(dev), text_ids=x["txt_ids"].to(dev),
        prompt_embeds=torch.cat([neg_pe, x["prompt_embeds"]]).to(dev), pooled_prompt_embeds=torch.cat([neg_po, x["pooled"]]).to(dev),
        timesteps=ts, num_inference_steps=n, guidance_scale=3.5,
        control_image_list=[torch.cat([c] * 2).to(dev) for c in x["conds"]], control_mask_list=[m.to(dev) for m in x["masks"]],
        control_mode=None, controlnet_conditioning_scale=1.0, controlnet_conditioning_step=30,
        callback_on_step_end=lambda p, i, t, kw: taps.append(kw["latents"].float().cpu()) or {},
        callback_on_step_end_tensor_inputs=["latents"], control_image_inpaint=torch.cat([cond_inp] * 2).to(dev),
        controlnet_conditioning_scale_inpaint=0.9, true_guidance_scale=3.5)
    want = torch.from_numpy(np.load(os.path.join(HERE, "golden", "tiny_inpaint.npz"))["latents_per_step"])
    assert len(taps) == want.shape[0]
    assert torch.equal(taps[0], x["latents"])          # true-CFG step 0 predicts zero: latents unchanged
    for i, got in enumerate(taps):
        assert rel_l2(got, want[i]) < 1e-4, (i, rel_l2(got, want[i]))
    assert torch.equal(lat.float().cpu(), taps[-1])


def _capture_denoise(pipe):
    box = {}
    inner = pipe._denoise

    def wrapped(**kw):
        box.update({k: (v.clone() if torch.is_tensor(v) else v) for k, v in kw.items()})
        return inner(**kw)

    pipe._denoise = wrapped
    return box


def _glyph_inputs(H, W, n_lines):
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


@pytest.mark.parametrize("dtype,names,tol", [
    (torch.float32, ("TINY_TRANSFORMER", "TINY_CONTROLNET", "TINY_INPAINT_CONTROLNET"), 1e-4),
    (torch.bfloat16, ("SMALL128_TRANSFORMER", "SMALL128_CONTROLNET", "SMALL128_INPAINT_CONTROLNET"), 1e-2)])
@pytest.mark.parametrize("inpaint", [False, True], ids=["t2i", "inpaint"])
def test_call_from_images_matches_oracle_loop(dtype, names, tol, inpaint):
    """Full __call__ from PIL images and a prompt string (synthetic text encoder / VAE stand-ins): the latents
    after every step must match the oracle's loop run on the tensors the pipeline prepared."""
    from oracle import flux_oracle as O
    from reptext_b200.pipeline_utils import SyntheticVAE
    H, W = 256, 192
    pipe, TR, CN, CNI, sds = _tiny_pipe(dtype, inpaint=inpaint, vae=SyntheticVAE(dtype=dtype, posterior_std=0.05),
                                        TRname=names[0], CNname=names[1], CNIname=names[2])
    box = _capture_denoise(pipe)
    glyph, cannys, poss, masks = _glyph_inputs(H, W, 2)
    taps = []
    kw = dict(prompt="a road sign that reads 'مرحبا'", height=H, width=W, num_inference_steps=3,
              guidance_scale=3.5, control_image=cannys, control_position=poss, control_mask=masks, control_glyph=glyph,
              controlnet_conditioning_scale=0.9, controlnet_conditioning_step=2, max_sequence_length=128,
              generator=torch.Generator(device="cuda").manual_seed(5), output_type="latent",
              callback_on_step_end=lambda p, i, t, k: taps.append(k["latents"].float()) or {})
    if inpaint:
        src = np.random.RandomState(0).randint(0, 255, (H, W, 3)).astype(np.uint8)
        from PIL import Image
        kw.update(control_image_inpaint=Image.fromarray(src), control_mask_inpaint=masks[0], true_guidance_scale=3.0,
                  controlnet_conditioning_scale_inpaint=0.8)
    out = pipe(**kw)
    assert out.images.shape == (1, (H // 16) * (W // 16), 64) and len(taps) == 3
    f = lambda v: v.float()
    ts, sg = O.make_sigmas(3, (H // 16) * (W // 16))
    assert torch.allclose(ts, box["timesteps"].float().cpu(), rtol=1e-6)
    dev = "cuda"
    od = lambda sd: {k: v.to(dev) for k, v in sd.items()}
    want = []
    cb = lambda i, t, lat: want.append(lat.clone())
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            args = dict(latents=f(box["latents"]), prompt_embeds=f(box["prompt_embeds"]), pooled=f(box["pooled_prompt_embeds"]),
                        control_image_list=[f(c) for c in box["control_image_list"]],
                        control_mask_list=[f(m) for m in box["control_mask_list"]],
                        text_ids=f(box["text_ids"]), img_ids=f(box["latent_image_ids"]), timesteps=ts.to(dev),
                        sigmas=sg.to(dev), guidance_scale=3.5, conditioning_scale=0.9, conditioning_step=2,
                        callback=cb, time_dtype=dtype)
            if inpaint:
                O.denoise_inpaint(od(sds["tr"]), TR, od(sds["cn"]), CN, od(sds["cni"]), CNI,
                                  control_image_inpaint=f(box["control_image_inpaint"]), true_guidance_scale=3.0,
                                  conditioning_scale_inpaint=0.8, **args)
            else:
                O.denoise_t2i(od(sds["tr"]), TR, od(sds["cn"]), CN, **args)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    for i, (got, w) in enumerate(zip(taps, want)):
        assert rel_l2(got, w) < tol, (i, rel_l2(got, w))
    if inpaint:   # the live glyph-latent init changed the latents inside the glyph mask only
        assert box["control_image_inpaint"].shape[-1] == 68 and box["prompt_embeds"].shape[0] == 2


def test_pipeline_argument_errors():
    pipe, TR, CN, _, _ = _tiny_pipe(torch.float32)
    with pytest.raises(ValueError):
        pipe(prompt="x", height=250, width=256, control_image=[], control_position=[])
    with pytest.raises(ValueError):
        pipe(height=256, width=256, control_image=[], control_position=[])
    with pytest.raises(ValueError):
        pipe(prompt="x", prompt_embeds=torch.zeros(1, 8, 64), height=256, width=256, control_image=[], control_position=[])
    with pytest.raises(ValueError):
        pipe(prompt="x", height=256, width=256, max_sequence_length=1024, control_image=[], control_position=[])


def test_precomputed_modulation_leaves_the_latents_bit_identical():
    """`pipe.precompute_modulation` (the AdaLN vectors of all steps in one pass before the loop,
    models.build_modulation_table) on / off through the public __call__: same latents bit for bit at every step, fewer
    launches."""
    from reptext_b200 import _lib
    pipe, TR, CN, _, _ = _tiny_pipe(torch.bfloat16, TRname="SMALL128_TRANSFORMER", CNname="SMALL128_CONTROLNET")
    H = W = 256
    x = synth_inputs(TR, CN, H, W, 128, seed=43, n_lines=2)
    lh, lw = 2 * (H // 16), 2 * (W // 16)
    imgs, poss, masks = [], [], []
    for li, cond in enumerate(x["conds"]):
        z = _unpack_cond(cond, lh, lw)
        imgs.append(z[:, :16].cuda())
        poss.append(z[:, 16:].cuda())
        y0 = (H // 4) * (li + 1) - H // 8
        masks.append(box_mask(H, W, (y0, y0 + H // 6, W // 5, W - W // 5)))

    def run(on):
        taps = []
        pipe.precompute_modulation = on
        n0 = _lib.launch_count()
        out = pipe(prompt_embeds=x["prompt_embeds"].cuda().bfloat16(), pooled_prompt_embeds=x["pooled"].cuda().bfloat16(),
                   height=H, width=W, num_inference_steps=6, guidance_scale=3.5, control_image=imgs,
                   control_position=poss, control_mask=masks, controlnet_conditioning_scale=1.0,
                   latents=x["latents"].cuda().bfloat16(), output_type="latent",
                   callback_on_step_end=lambda p, i, t, kw: taps.append(kw["latents"].clone()) or {})
        torch.cuda.synchronize()
        return taps, out.images, _lib.launch_count() - n0

    try:
        off, on = run(False), run(True)
    finally:
        pipe.precompute_modulation = True
    assert len(off[0]) == len(on[0]) == 6
    for a, b in zip(off[0], on[0]):
        assert torch.equal(a, b)
    assert torch.equal(off[1], on[1]) and on[2] < off[2], (on[2], off[2])


def test_step_invariant_cache_leaves_the_latents_bit_identical():
    """SURVEY.md 8f.2 through the public __call__ (two text lines: the ControlNet runs twice per step on the same prompt
    tensors): `pipe.cache_step_invariants` on / off give the same latents bit for bit at every step, and the cached
    run launches fewer kernels."""
    from reptext_b200 import _lib
    pipe, TR, CN, _, _ = _tiny_pipe(torch.bfloat16, TRname="SMALL128_TRANSFORMER", CNname="SMALL128_CONTROLNET")
    H = W = 256
    x = synth_inputs(TR, CN, H, W, 128, seed=41, n_lines=2)
    lh, lw = 2 * (H // 16), 2 * (W // 16)
    imgs, poss, masks = [], [], []
    for li, cond in enumerate(x["conds"]):
        z = _unpack_cond(cond, lh, lw)
        imgs.append(z[:, :16].cuda())
        poss.append(z[:, 16:].cuda())
        y0 = (H // 4) * (li + 1) - H // 8
        masks.append(box_mask(H, W, (y0, y0 + H // 6, W // 5, W - W // 5)))

    def run(on):
        taps = []
        pipe.cache_step_invariants = on
        n0 = _lib.launch_count()
        out = pipe(prompt_embeds=x["prompt_embeds"].cuda().bfloat16(), pooled_prompt_embeds=x["pooled"].cuda().bfloat16(),
                   height=H, width=W, num_inference_steps=5, guidance_scale=3.5, control_image=imgs,
                   control_position=poss, control_mask=masks, controlnet_conditioning_scale=1.0,
                   latents=x["latents"].cuda().bfloat16(), output_type="latent",
                   callback_on_step_end=lambda p, i, t, kw: taps.append(kw["latents"].clone()) or {})
        torch.cuda.synchronize()
        return taps, out.images, _lib.launch_count() - n0

    off, on = run(False), run(True)
    assert len(off[0]) == len(on[0]) == 5
    for a, b in zip(off[0], on[0]):
        assert torch.equal(a, b)
    assert torch.equal(off[1], on[1]) and on[2] < off[2], (on[2], off[2])
    assert pipe.transformer._inv_key is None and pipe.controlnet._inv_src is None     # released after the loop

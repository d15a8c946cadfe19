"""Parity PINNED TO THE REFERENCE'S OWN CODE (CPU; no GPU needed).

``/root/reference/RepText/{controlnet_flux,pipeline_flux_controlnet,pipeline_flux_controlnet_inpaint}.py`` are imported by
path, unmodified, over the stand-in ``diffusers`` of ``tests/ref_shim`` (BFL block arithmetic from torchtitan, the real
transformers CLIP / T5) and RUN.  Three links are checked:

  reference run  ==  committed golden (tests/golden/ref_*.npz)        [needs /root/reference; skipped on the GPU box]
  reference ``FluxControlNetModel.forward``  ==  ``oracle.controlnet_forward``   [needs /root/reference]
  oracle (loops AND preparation, from the same PIL images / prompts)  ==  committed golden    [runs anywhere]

The GPU tests (tests/test_reference_gpu.py) then compare the product pipelines with the same golden files.
"""
import os

import numpy as np
import pytest
import torch

import ref_fixture as F
import ref_run
from oracle import flux_oracle as O
from util import rel_l2, synth_inputs

HERE = os.path.dirname(os.path.abspath(__file__))
needs_ref = pytest.mark.skipif(not ref_run.available(), reason="the reference (/root/reference) is not on this box")
needs_shim = pytest.mark.skipif(not ref_run.shim_importable(), reason="torchtitan's BFL modules are not importable")


def golden(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _ls(cfg):
    return {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}


# ------------------------------------------------------------------------------------------ reference == oracle (models)
@needs_ref
@pytest.mark.parametrize("single_layers", [0, 2])
@pytest.mark.parametrize("guidance_embeds", [True, False])
def test_reference_controlnet_forward_equals_oracle(single_layers, guidance_embeds):
    """RepText/controlnet_flux.py:216-413, executed, against oracle.controlnet_forward: fp32, batch 2, <= 1e-6
    (rows a1-a9; ``single_layers=2`` runs the loop at :354-381 and the zero-linears at :390-392 - row a7)."""
    from reptext_b200 import config, weights
    R = ref_run.load()
    TR = config.TINY_TRANSFORMER
    CN = dict(config.TINY_CONTROLNET, num_single_layers=single_layers, guidance_embeds=guidance_embeds)
    m = R.controlnet_flux.FluxControlNetModel(**_ls(CN)).eval()
    sd = weights.random_state_dict(CN, "controlnet", seed=101)
    m.load_state_dict(sd, strict=True)
    x = synth_inputs(TR, CN, 256, 192, 64, seed=3, batch=2)
    t = torch.tensor([0.75, 0.3])
    g = torch.tensor([3.5, 2.0]) if guidance_embeds else None
    with torch.no_grad():
        out = m(hidden_states=x["latents"], controlnet_cond=x["conds"][0], conditioning_scale=0.7,
                encoder_hidden_states=x["prompt_embeds"], pooled_projections=x["pooled"], timestep=t,
                img_ids=x["img_ids"], txt_ids=x["txt_ids"], guidance=g, return_dict=True)
        rb, rs = out.controlnet_block_samples, out.controlnet_single_block_samples
        ob, os_ = O.controlnet_forward(sd, CN, x["latents"], x["conds"][0], 0.7, x["prompt_embeds"], x["pooled"], t,
                                       x["img_ids"], x["txt_ids"], g)
    assert len(rb) == len(ob) == CN["num_layers"]
    for a, b in zip(ob, rb):
        assert float(b.norm()) > 1.0 and rel_l2(a, b) < 1e-6
    if single_layers:
        assert len(rs) == len(os_) == single_layers
        for a, b in zip(os_, rs):
            assert b.shape == x["latents"].shape[:2] + (256,) and rel_l2(a, b) < 1e-6
    else:
        assert rs is None and os_ is None


@needs_shim
def test_shim_transformer_equals_oracle():
    """The stand-in FluxTransformer2DModel (BFL blocks + diffusers 0.36's loop: guidance embedder, residual injection
    ``i // ceil(L / len)`` for both block families) against oracle.transformer_forward."""
    from reptext_b200 import config, weights
    ref_run.shim()
    from diffusers.models.transformers.transformer_flux import FluxTransformer2DModel
    TR, CN = config.TINY_TRANSFORMER, config.TINY_CONTROLNET
    sd = weights.random_state_dict(TR, "transformer", seed=100)
    m = FluxTransformer2DModel(**_ls(TR)).eval()
    m.load_state_dict(sd, strict=True)
    x = synth_inputs(TR, CN, 256, 192, 64, seed=4, batch=2)
    g = torch.Generator().manual_seed(9)
    D = TR["num_attention_heads"] * TR["attention_head_dim"]
    blocks = [0.1 * torch.randn(2, x["N"], D, generator=g) for _ in range(2)]
    singles = [0.1 * torch.randn(2, x["N"], D, generator=g) for _ in range(3)]
    t, gd = torch.tensor([0.6, 0.2]), torch.tensor([3.5, 1.0])
    with torch.no_grad():
        got = m(hidden_states=x["latents"], timestep=t, guidance=gd, pooled_projections=x["pooled"],
                encoder_hidden_states=x["prompt_embeds"], controlnet_block_samples=blocks,
                controlnet_single_block_samples=singles, txt_ids=x["txt_ids"], img_ids=x["img_ids"], return_dict=False)[0]
        want = O.transformer_forward(sd, TR, x["latents"], x["prompt_embeds"], x["pooled"], t, x["img_ids"], x["txt_ids"],
                                     gd, blocks, singles)
    assert rel_l2(want, got) < 1e-6


# ------------------------------------------------------------------------------------------ reference run == golden
@needs_ref
@pytest.mark.parametrize("name", sorted(F.CASES))
def test_reference_run_reproduces_committed_golden(name):
    """The committed tests/golden/ref_*.npz ARE what the reference's own ``__call__`` produces here."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_ref
    torch.set_num_threads(8)
    rec = make_golden_ref.run_case(name)
    want = golden(name)
    assert set(rec) == set(want)
    tol = 1e-6 if F.CASES[name]["dtype"] == "fp32" else 1e-3       # bf16 on the CPU: reduction order may differ by box
    for k in want:
        assert rec[k].shape == tuple(want[k].shape), k
        assert rel_l2(torch.from_numpy(rec[k]), want[k]) <= tol, (k, rel_l2(torch.from_numpy(rec[k]), want[k]))


# ------------------------------------------------------------------------------------------ oracle == golden (anywhere)
@pytest.mark.parametrize("name", sorted(F.CASES))
def test_oracle_loop_equals_reference_run(name):
    """oracle.denoise_t2i / denoise_inpaint on the tensors the reference prepared == the reference's per-step latents
    (pipeline_flux_controlnet.py:1017-1130, pipeline_flux_controlnet_inpaint.py:1140-1295): fp32 <= 1e-5 at every step;
    the bf16 cases run the oracle in bf16 on the CPU like the reference did."""
    torch.set_num_threads(8)
    z = golden(name)
    fp32 = F.CASES[name]["dtype"] == "fp32"
    got = F.oracle_loop(name, z, torch.float32 if fp32 else torch.bfloat16)
    want = z["latents_per_step"]
    assert got.shape == want.shape
    for i in range(want.shape[0]):
        e = rel_l2(got[i], want[i])
        assert e < (1e-5 if fp32 else 1e-2), (name, i, e)
    if F.CASES[name]["kind"] == "inpaint":          # true-CFG step 0 predicts zero: latents unchanged
        assert torch.equal(want[0], z["init_latents"])
    if not fp32:
        # the fp32 truth stored next to the bf16 run (same prepared inputs, timesteps rounded like the bf16 run rounds
        # them) is reproducible, and the reference's OWN bf16 error against it is what the GPU test compares with
        truth = F.oracle_loop(name, z, torch.float32, torch.bfloat16)
        assert rel_l2(truth, z["latents_per_step_fp32"]) < 1e-5
        print(name, "reference bf16 run vs fp32:", [f"{rel_l2(want[i], truth[i]):.1e}" for i in range(want.shape[0])])


@pytest.mark.parametrize("name", ["ref_tiny_t2i", "ref_tiny_t2i_offgrid", "ref_tiny_inpaint"])
def test_oracle_preparation_equals_reference_run(name):
    """Rows a14-a18 restated in the oracle, from the same PIL images and prompt strings: ``prepare_image``
    (pipeline_flux_controlnet.py:663-731), ``prepare_image_with_mask`` (pipeline_flux_controlnet_inpaint.py:761-826),
    ``prepare_latents_reptext`` (:608-660; dead in T2I, live in inpaint), the sigma schedule, ids and the prompt encoders
    == what the reference's own preparation code built (recorded by hooks on its modules during the golden run)."""
    from oracle import text_oracle as TO
    from oracle import vae_oracle as V
    torch.set_num_threads(8)
    case = F.CASES[name]
    z = golden(name)
    sds = F.state_dicts(case)
    vcfg = F.vae_config(case)
    kw = F.call_kwargs(case)
    H, W = case["H"], case["W"]
    inpaint = case["kind"] == "inpaint"
    gen = kw["generator"]

    def encode(x, generator=None):
        with torch.no_grad():
            return V.sample_posterior(V.encode_moments(sds["vae"], vcfg, x), generator=generator)

    # prompt encoders (the reference calls CLIP on `prompt`, T5 on `prompt_2`; negative prompt first when CFG is on)
    tcfg, ccfg = F.text_configs(case)
    tok, tok2 = F.tokenizers()
    prompts = ([F.NEGATIVE] if inpaint else []) + [F.PROMPT]
    prompts2 = ([F.NEGATIVE] if inpaint else []) + [F.PROMPT_2]
    with torch.no_grad():
        pe = torch.cat([TO.t5_encoder(sds["t5"], tcfg, tok2([p], padding="max_length", max_length=case["T"],
                                                             truncation=True).input_ids) for p in prompts2])
        po = torch.cat([TO.clip_text(sds["clip"], ccfg, tok([p], padding="max_length", max_length=77,
                                                            truncation=True).input_ids)[1] for p in prompts])
    assert rel_l2(pe, z["prompt_embeds"]) < 1e-5 and rel_l2(po, z["pooled"]) < 1e-5

    # control conditions
    for li in range(case["lines"]):
        cond = O.prepare_image(encode, kw["control_image"][li], kw["control_position"][li], H, W, 1, torch.float32,
                               do_classifier_free_guidance=inpaint)
        assert cond.shape == z["conds"][li].shape and rel_l2(cond, z["conds"][li]) < 1e-5, (li, rel_l2(cond, z["conds"][li]))
    if inpaint:
        ci = O.prepare_image_with_mask(encode, kw["control_image_inpaint"], kw["control_mask_inpaint"], H, W, 1,
                                       torch.float32, do_classifier_free_guidance=True)
        assert ci.shape == z["cond_inpaint"].shape == (2, (H // 16) * (W // 16), 68)
        assert rel_l2(ci, z["cond_inpaint"]) < 1e-5

    # initial latents, schedule, ids
    lat = O.prepare_latents_reptext(encode, kw["control_glyph"], 1, H, W, torch.float32, gen, live=inpaint)
    # live (inpaint): 0.10 * glyph latents enter inside the mask, VAE oracle vs BFL autoencoder ~1e-5; dead (T2I): the noise
    assert rel_l2(lat, z["init_latents"]) < (5e-5 if inpaint else 1e-7)
    ts, sg = O.make_sigmas(case["steps"], (H // 16) * (W // 16))
    assert torch.allclose(ts, z["timesteps"], rtol=1e-6) and torch.allclose(sg, z["sigmas"], rtol=1e-6, atol=1e-7)
    assert torch.equal(O.prepare_latent_image_ids(2 * (H // 16), 2 * (W // 16)), z["img_ids"])
    assert torch.equal(torch.zeros(case["T"], 3), z["txt_ids"])


# ------------------------------------------------------------------ reference == product host logic (text-to-render span)
@needs_ref
def test_text_to_render_span_equals_reference_run():
    """``encode_prompt(..., get_text_to_render=True)`` (pipeline_flux_controlnet.py:257-280, :423-454): the reference's own
    method, run on its own pipeline object, against the product's host logic on the same tokenizers and the same
    (transformers) encoders - embeddings, pooled vector, text ids and the (start, end) span of the quoted text; and the
    reference's error when the quoted text is not found as a token window."""
    from types import SimpleNamespace
    from reptext_b200._pipeline_common import RepTextPipelineBase as B
    from reptext_b200.pipeline_flux_controlnet import FluxControlNetPipeline as P
    torch.set_num_threads(8)
    case = F.CASES["ref_tiny_t2i"]
    ref = F.reference_pipeline(case)
    mine = SimpleNamespace(tokenizer=ref.tokenizer, tokenizer_2=ref.tokenizer_2, text_encoder=ref.text_encoder,
                           text_encoder_2=ref.text_encoder_2, tokenizer_max_length=ref.tokenizer_max_length,
                           _execution_device=torch.device("cpu"))
    for name in ("_locate_text_to_render", "_get_t5_prompt_embeds", "_get_clip_prompt_embeds", "_encode_text", "_text_ids"):
        setattr(mine, name, getattr(B, name).__get__(mine))
    for prompt in ("a street sign that says ' hello big world ' in the city, film grain",
                   'a shop front with " open all night " written above the door'):
        with torch.no_grad():
            want = ref.encode_prompt(prompt, None, device="cpu", max_sequence_length=case["T"], get_text_to_render=True)
            got = P.encode_prompt(mine, prompt, None, device="cpu", max_sequence_length=case["T"], get_text_to_render=True)
        assert len(want) == len(got) == 5 and tuple(got[3:]) == tuple(want[3:]) and got[4] > got[3] > 0
        for w, g in zip(want[:3], got[:3]):
            assert w.shape == g.shape and w.dtype == g.dtype and torch.equal(w, g)
    with torch.no_grad():
        plain = P.encode_prompt(mine, "no quotes needed here", None, device="cpu", max_sequence_length=case["T"])
    assert len(plain) == 3
    # the quoted text tokenises differently on its own than inside the prompt (no break after the closing quote): both raise
    bad = "a sign that says 'hello big world'now in the city"
    for call in (lambda: ref.encode_prompt(bad, None, device="cpu", max_sequence_length=case["T"], get_text_to_render=True),
                 lambda: P.encode_prompt(mine, bad, None, device="cpu", max_sequence_length=case["T"], get_text_to_render=True)):
        with pytest.raises(ValueError, match="No match found"):
            call()


@needs_ref
def test_get_timesteps_equals_reference_run():
    """``get_timesteps`` (pipeline_flux_controlnet.py:474-484), the reference's own method on its own pipeline object and
    scheduler, against the product's on the product's scheduler: same tail of the schedule, same count, same begin index."""
    from types import SimpleNamespace
    from reptext_b200._pipeline_common import RepTextPipelineBase as B
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
    ref = F.reference_pipeline(F.CASES["ref_tiny_t2i"])
    sig = np.linspace(1.0, 1 / 12, 12)
    mine = SimpleNamespace(scheduler=FlowMatchEulerDiscreteScheduler())
    for strength in (1.0, 0.75, 0.3, 0.05, 0.0):
        ref.scheduler.set_timesteps(sigmas=sig, mu=0.8)
        mine.scheduler.set_timesteps(sigmas=sig, mu=0.8)
        want_ts, want_n = ref.get_timesteps(12, strength, "cpu")
        got_ts, got_n = B.get_timesteps(mine, 12, strength, "cpu")
        assert got_n == want_n and torch.allclose(got_ts.float(), want_ts.float(), rtol=0, atol=1e-4)
        assert mine.scheduler.begin_index == ref.scheduler.begin_index


@needs_ref
@pytest.mark.parametrize("kind", ["ref_tiny_t2i", "ref_tiny_inpaint"])
def test_check_inputs_accepts_and_rejects_like_the_reference_run(kind):
    """``check_inputs`` of both pipelines (pipeline_flux_controlnet.py:486-531, the inpaint file's copy): every argument set
    the reference's own method accepts is accepted, every one it rejects is rejected with the same exception type."""
    from reptext_b200._pipeline_common import RepTextPipelineBase as B
    ref = F.reference_pipeline(F.CASES[kind])
    mine = B.__new__(B)
    e = torch.zeros(1, 4, 8)
    ok = dict(prompt="a", prompt_2=None, height=256, width=256)
    cases = [ok, dict(ok, prompt=["a", "b"], prompt_2=["c", "d"]), dict(ok, prompt_2="b", max_sequence_length=512),
             dict(ok, prompt=None, prompt_embeds=e, pooled_prompt_embeds=e), dict(ok, height=1000, width=760),
             dict(ok, callback_on_step_end_tensor_inputs=["latents", "prompt_embeds"]),
             dict(ok, height=250), dict(ok, width=12), dict(ok, prompt=None), dict(ok, prompt=3), dict(ok, prompt_2=3.5),
             dict(ok, prompt_embeds=e), dict(ok, prompt=None, prompt_2="b", prompt_embeds=e, pooled_prompt_embeds=e),
             dict(ok, prompt=None, prompt_embeds=e), dict(ok, max_sequence_length=513),
             dict(ok, callback_on_step_end_tensor_inputs=["noise_pred"])]
    outcomes = []
    for kw in cases:
        res = []
        for obj in (ref, mine):
            try:
                obj.check_inputs(**kw)
                res.append(None)
            except Exception as ex:          # noqa: BLE001 - the TYPE is what is compared
                res.append(type(ex))
        assert res[0] is res[1], (kw, res)
        outcomes.append(res[0])
    assert outcomes.count(None) == 6 and outcomes.count(ValueError) == 10


@needs_ref
def test_pack_unpack_ids_shift_equal_reference_functions():
    """The once-per-call helpers, reference function against product function on the same CPU tensors, bit for bit:
    ``_pack_latents`` / ``_unpack_latents`` / ``_prepare_latent_image_ids`` (pipeline_flux_controlnet.py:535-570),
    ``calculate_shift`` (:78-88) and ``retrieve_timesteps`` with custom sigmas (:104-160) - for both pipeline files."""
    from reptext_b200 import _pipeline_common as PC
    from reptext_b200.scheduler import FlowMatchEulerDiscreteScheduler
    R = ref_run.load()
    B = PC.RepTextPipelineBase
    g = torch.Generator().manual_seed(5)
    for mod in (R.t2i, R.inpaint):
        RP = mod.FluxControlNetPipeline
        for (b, c, h, w) in ((1, 16, 32, 32), (2, 16, 24, 40), (3, 4, 2, 6)):
            x = torch.randn(b, c, h, w, generator=g)
            p_ref, p_mine = RP._pack_latents(x, b, c, h, w), B._pack_latents(x, b, c, h, w)
            assert torch.equal(p_ref, p_mine)
            assert torch.equal(RP._unpack_latents(p_ref, h * 8, w * 8, 16), B._unpack_latents(p_mine, h * 8, w * 8, 16))
            ids_ref = RP._prepare_latent_image_ids(b, h, w, "cpu", torch.float32)
            ids_mine = B._prepare_latent_image_ids(b, h, w, "cpu", torch.float32)
            assert ids_ref.shape == ids_mine.shape and torch.equal(ids_ref, ids_mine)
        for n in (256, 1024, 4096, 9216):
            assert mod.calculate_shift(n) == PC.calculate_shift(n)
            assert mod.calculate_shift(n, 256, 4096, 0.5, 1.16) == PC.calculate_shift(n, 256, 4096, 0.5, 1.16)
    # the schedule through retrieve_timesteps (custom sigmas + mu), shim scheduler under the reference function
    from diffusers.schedulers import FlowMatchEulerDiscreteScheduler as ShimScheduler
    for steps, n in ((4, 256), (28, 4096), (30, 9216)):
        sig = np.linspace(1.0, 1 / steps, steps)
        mu = PC.calculate_shift(n)
        s_ref, s_mine = ShimScheduler(**F.config.SCHEDULER), FlowMatchEulerDiscreteScheduler()
        t_ref, n_ref = R.t2i.retrieve_timesteps(s_ref, steps, "cpu", None, sig, mu=mu)
        t_mine, n_mine = PC.retrieve_timesteps(s_mine, steps, "cpu", None, sig, mu=mu)
        assert n_ref == n_mine == steps and torch.equal(t_ref, t_mine) and torch.equal(s_ref.sigmas, s_mine.sigmas)
    with pytest.raises(ValueError):
        PC.retrieve_timesteps(FlowMatchEulerDiscreteScheduler(), 4, "cpu", [1, 2], [0.5, 0.1])
    with pytest.raises(ValueError):
        R.t2i.retrieve_timesteps(ShimScheduler(**F.config.SCHEDULER), 4, "cpu", [1, 2], [0.5, 0.1])


@needs_ref
def test_prepare_latents_equals_reference_run():
    """``prepare_latents`` (pipeline_flux_controlnet.py:573-606): the reference's own method on its pipeline object against
    the product's, same seeded generators (one, and a list of one per sample), passed-in latents, and the length check."""
    from reptext_b200._pipeline_common import RepTextPipelineBase as B
    ref = F.reference_pipeline(F.CASES["ref_tiny_t2i"])
    mine = B.__new__(B)
    mine.vae_scale_factor = ref.vae_scale_factor
    gens = lambda seeds: [torch.Generator().manual_seed(s) for s in seeds]
    for (b, h, w, seeds) in ((1, 256, 256, None), (2, 256, 384, None), (2, 200, 136, [3, 4])):
        g_ref = gens(seeds) if seeds else torch.Generator().manual_seed(7)
        g_mine = gens(seeds) if seeds else torch.Generator().manual_seed(7)
        l_ref, i_ref = ref.prepare_latents(b, 16, h, w, torch.float32, "cpu", g_ref)
        l_mine, i_mine = mine.prepare_latents(b, 16, h, w, torch.float32, "cpu", g_mine)
        assert l_ref.shape == l_mine.shape and torch.equal(l_ref, l_mine) and torch.equal(i_ref, i_mine)
        given = torch.randn_like(l_ref)
        k_ref, _ = ref.prepare_latents(b, 16, h, w, torch.float32, "cpu", None, latents=given)
        k_mine, _ = mine.prepare_latents(b, 16, h, w, torch.float32, "cpu", None, latents=given)
        assert torch.equal(k_ref, k_mine) and torch.equal(k_mine, given)
    for obj in (ref, mine):
        with pytest.raises(ValueError):
            obj.prepare_latents(3, 16, 256, 256, torch.float32, "cpu", gens([1, 2]))


@needs_ref
@pytest.mark.parametrize("kind", ["ref_tiny_t2i", "ref_tiny_inpaint"])
@torch.no_grad()
def test_product_condition_preparation_equals_reference_run_on_cpu(kind):
    """Rows a17 / a18 for the PRODUCT's host code (not only the oracle's): ``prepare_image`` (pipeline_flux_controlnet.py
    :663-731), ``prepare_image_with_mask`` (pipeline_flux_controlnet_inpaint.py:761-826) and the T2I ``prepare_latents_reptext``
    (:608-660), product method against the reference's own method, both on the CPU with the SAME VAE module in the ``vae``
    slot, the same PIL images, the same global / explicit RNG state - at the 256 x 256 size and at an off-grid request
    (250 -> resized to 240) that exercises the image processors' rounding and Lanczos resize."""
    from reptext_b200 import _pipeline_common as PC
    torch.set_num_threads(8)
    case = F.CASES[kind]
    ref = F.reference_pipeline(case)
    if kind == "ref_tiny_inpaint":
        from reptext_b200.pipeline_flux_controlnet_inpaint import FluxControlNetPipeline as P
    else:
        from reptext_b200.pipeline_flux_controlnet import FluxControlNetPipeline as P
    mine = P.__new__(P)
    mine.vae = ref.vae
    PC.RepTextPipelineBase._setup(mine)
    if kind == "ref_tiny_inpaint":
        mine.mask_processor = type(mine.image_processor)(vae_scale_factor=mine.vae_scale_factor, do_resize=True,
                                                         do_convert_grayscale=True, do_normalize=False, do_binarize=True)
    assert mine.vae_scale_factor == ref.vae_scale_factor
    kw = F.call_kwargs(case)
    for (H, W) in ((case["H"], case["W"]), (248, 200)):
        pi = dict(image=kw["control_image"][0], width=W, height=H, device="cpu", dtype=torch.float32,
                  image_position=kw["control_position"][0])
        for nb in (1, 2):
            torch.manual_seed(11)
            want, hw, ww = ref.prepare_image(batch_size=nb, num_images_per_prompt=1, **pi)
            torch.manual_seed(11)
            got, hg, wg = mine.prepare_image(batch_size=nb, num_images_per_prompt=1, **pi)
            assert (hw, ww) == (hg, wg) and want.shape == got.shape and rel_l2(got, want) < 1e-6, (H, W, nb)
        # a reference quirk kept: the pipelines pass batch_size * num_images_per_prompt AND num_images_per_prompt, and the
        # pack multiplies the two again - num_images_per_prompt > 1 fails in the reference's view(), and here alike
        for obj in (ref, mine):
            with pytest.raises(RuntimeError):
                obj.prepare_image(batch_size=2, num_images_per_prompt=2, **pi)
        if kind == "ref_tiny_inpaint":
            torch.manual_seed(12)
            want, _, _ = ref.prepare_image_with_mask(image=kw["control_image_inpaint"], mask=kw["control_mask_inpaint"],
                                                     width=W, height=H, batch_size=1, num_images_per_prompt=1, device="cpu",
                                                     dtype=torch.float32, do_classifier_free_guidance=True)
            torch.manual_seed(12)
            got, _, _ = mine.prepare_image_with_mask(image=kw["control_image_inpaint"], mask=kw["control_mask_inpaint"],
                                                     width=W, height=H, batch_size=1, num_images_per_prompt=1, device="cpu",
                                                     dtype=torch.float32, do_classifier_free_guidance=True)
            assert want.shape == got.shape and want.shape[-1] == 68 and rel_l2(got, want) < 1e-6, (H, W)
    if kind == "ref_tiny_t2i":
        glyph_r = ref.image_processor.preprocess(kw["control_glyph"], height=case["H"], width=case["W"])
        glyph_m = mine.image_processor.preprocess(kw["control_glyph"], height=case["H"], width=case["W"])
        assert torch.equal(glyph_r, glyph_m)
        args = (1, 16, case["H"], case["W"], torch.float32, "cpu")
        want, ids_w = ref.prepare_latents_reptext(glyph_r, *args, torch.Generator().manual_seed(3))
        got, ids_g = mine.prepare_latents_reptext(glyph_m, *args, torch.Generator().manual_seed(3))
        assert torch.equal(want, got) and torch.equal(ids_w, ids_g)


# ------------------------------------------------------------------ reference script == product host glyph path (8f row 4)
def _infer_script_pieces():
    """The reference's own ``infer.py`` code, unmodified, as executable pieces: its helper functions (``contains_chinese``,
    ``canny``), the per-text-line ``for`` loop (:71-100), the statements after it up to the prompt (:102-113).  The script
    itself cannot be imported (it loads checkpoints at ``__main__`` and imports diffusers-bound modules at the top)."""
    import ast
    src = open(os.path.join(ref_run.REF, "infer.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef)
            or (isinstance(n, (ast.Import, ast.ImportFrom)) and "controlnet" not in ast.unparse(n))]
    main = [n for n in tree.body if isinstance(n, ast.If)][0].body
    loop_at = [i for i, n in enumerate(main) if isinstance(n, ast.For)][0]
    tail = []
    for n in main[loop_at + 1:]:
        if isinstance(n, ast.Assign) and "Generator" in ast.unparse(n):
            break
        if not (isinstance(n, ast.Expr) and "print" in ast.unparse(n)):
            tail.append(n)
    mod = lambda nodes: compile(ast.fix_missing_locations(ast.Module(body=nodes, type_ignores=[])), "infer.py", "exec")
    return mod(keep), mod([main[loop_at]]), mod(tail)


@needs_ref
def test_glyph_conditions_and_prompt_equal_the_reference_script():
    """``reptext_b200.glyphs.build_conditions`` / ``build_prompt`` / ``canny`` / ``contains_chinese`` against the code of
    ``RepText/infer.py`` itself (:11-22, :71-113), executed from the file with the same font, lines, positions and colours:
    Canny, position, regional-mask and accumulated glyph images pixel for pixel, and the prompt string."""
    cv2 = pytest.importorskip("cv2")
    from PIL import ImageFont
    from reptext_b200 import glyphs
    helpers, loop, tail = _infer_script_pieces()
    for (W, H, size, texts, poss, cols) in (
            (512, 384, 48, ["Shakker Labs", "RepText"], [(60, 80), (60, 200)], [(255, 255, 255), (255, 200, 40)]),
            (256, 256, 32, ["B200"], [(40, 100)], [(255, 255, 255)]),
            (384, 256, 40, ["哩布哩布", "Lovart AI"], [(30, 40), (30, 140)], [(255, 255, 255), (200, 255, 255)])):
        font = ImageFont.load_default(size)
        ns = dict(width=W, height=H, font=font, text_list=texts, text_position_list=poss, text_color_list=cols,
                  control_image_list=[], control_position_list=[], control_mask_list=[],
                  control_glyph_all=np.zeros([H, W, 3], dtype=np.uint8))
        exec(helpers, ns)
        exec(loop, ns)
        exec(tail, ns)
        got = glyphs.build_conditions(texts, poss, cols, W, H, font)
        for name, mine in (("control_image_list", got.control_image), ("control_position_list", got.control_position),
                           ("control_mask_list", got.control_mask)):
            assert len(ns[name]) == len(mine) == len(texts)
            for a, b in zip(ns[name], mine):
                assert a.mode == b.mode and a.size == b.size and np.array_equal(np.array(a), np.array(b)), name
        assert np.array_equal(np.array(ns["control_glyph_all"]), np.array(got.control_glyph))
        assert ns["prompt"] == glyphs.build_prompt("a street sign in city", texts,
                                                   ", filmfotos, film grain, reversal film photography")
        for t in texts + ["مرحبا", "abc 漢"]:
            assert ns["contains_chinese"](t) == glyphs.contains_chinese(t)
        rnd = np.random.RandomState(0).randint(0, 255, (64, 96, 3)).astype(np.uint8)
        assert np.array_equal(ns["canny"](rnd), glyphs.canny(rnd))


@needs_ref
def test_inpaint_script_helpers_equal_the_reference_script():
    """``glyphs.resize_img`` and ``build_conditions(position_margin=5)`` against the code of ``RepText/infer_inpaint.py``
    itself (:25-46, the per-line loop at :84-117 whose POSITION box is grown by 5 pixels like the mask), executed from
    the file."""
    pytest.importorskip("cv2")
    import ast
    from PIL import Image, ImageFont
    from reptext_b200 import glyphs
    tree = ast.parse(open(os.path.join(ref_run.REF, "infer_inpaint.py")).read())
    fns = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name in ("resize_img", "canny", "contains_chinese"))
           or (isinstance(n, (ast.Import, ast.ImportFrom)) and not any(w in ast.unparse(n) for w in ("controlnet", "diffusers")))]
    main = [n for n in tree.body if isinstance(n, ast.If)][0].body
    loop = [n for n in main if isinstance(n, ast.For)][0]
    mod = lambda nodes: compile(ast.fix_missing_locations(ast.Module(body=nodes, type_ignores=[])), "infer_inpaint.py", "exec")
    ns = {}
    exec(mod(fns), ns)
    rs = np.random.RandomState(1)
    for (w, h), kw in (((1600, 1067), {}), ((900, 1400), {}), ((640, 480), dict(size=(512, 384))),
                       ((1500, 1000), dict(pad_to_max_side=True)), ((333, 517), dict(max_side=768, min_side=512, base_pixel_number=16))):
        img = Image.fromarray(rs.randint(0, 255, (h, w, 3)).astype(np.uint8))
        a, b = ns["resize_img"](img, **kw), glyphs.resize_img(img, **kw)
        assert a.size == b.size and np.array_equal(np.array(a), np.array(b)), ((w, h), kw)
    W, H, texts, poss, cols = 512, 384, ["Shakker Labs", "RepText"], [(60, 80), (60, 200)], [(0, 255, 0), (255, 255, 255)]
    font = ImageFont.load_default(48)
    ns.update(width=W, height=H, font=font, text_list=texts, text_position_list=poss, text_color_list=cols,
              control_image_list=[], control_position_list=[], control_mask_list=[],
              control_glyph_all=np.zeros([H, W, 3], dtype=np.uint8))
    exec(mod([loop]), ns)
    got = glyphs.build_conditions(texts, poss, cols, W, H, font, position_margin=5)
    for name, mine in (("control_image_list", got.control_image), ("control_position_list", got.control_position),
                       ("control_mask_list", got.control_mask)):
        for a, b in zip(ns[name], mine):
            assert np.array_equal(np.array(a), np.array(b)), name
    assert np.array_equal(np.array(ns["control_position_list"][0]), np.array(ns["control_mask_list"][0]))
    assert np.array_equal(np.array(ns["control_glyph_all"]), np.array(got.control_glyph))
